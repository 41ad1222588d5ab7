// K1+K2 fused: L2 norm of the region embeddings and the region x text similarity in ONE kernel
// that reads the fp32 NCHW conv output directly.
//
// Replaces, for every level at once, model/heads/text_contrastive.py:134-147 (permute, F.normalize,
// matmul, alpha*s+beta) and model/yolo_clip.py:198-206 (max/argmax over classes, level concat).
//
//   S[b, a, c] = alpha * <x[b,:,a], t^[c,:]> / max(||x[b,:,a]||, 1e-12) + beta
//
// Why fused: the separate normalise kernel writes and the GEMM re-reads a bf16 copy of the
// activations (25.8 MB + 8.6 MB of HBM traffic per image against 17.2 MB of input), and a GEMM
// that streams the A tile once per N tile is bound by the SM's L2->smem ingest, not by the
// tensor pipe.  Here every activation is read from HBM exactly once, as fp32, and the converted
// A tile stays on chip for the whole vocabulary:
//
//   warp 0      TMA producer, text tiles  [128 classes x 64 k] bf16, SWIZZLE_128B, 4-stage ring
//               (4 x 16 KiB: deep enough to cover the L2 latency, shallow enough that the
//               activation loads issued at an anchor-tile boundary do not queue behind 128 KiB
//               of text prefetch)
//   warp 1      MMA issuer   tcgen05.mma.kind::f16, A FROM TENSOR MEMORY, 128 x N x 16, N <= 128
//   warp 2      TMA producer, activations [64 k x 128 anchors] fp32 straight from NCHW
//               (anchors contiguous), 4-stage ring; also owns the TMEM allocation
//   warp 3      (idle; optional L2 prefetch of the next anchor tile, off by default)
//   warps 4-7   converters   thread = anchor row: fp32 smem column -> sum of squares, bf16x2 ->
//               tcgen05.st into the A region of TMEM (256 columns = 128 rows x 512 k)
//   warps 8-11  epilogue     tcgen05.ld -> alpha/||x|| scale, +beta, running max/argmax and/or
//               smem-staged coalesced logit stores
//
// TMEM: columns [0,256) = two 128-column fp32 accumulators (epilogue of tile n overlaps the
// MMAs of tile n+1), columns [256,512) = the A operand.  A block kb of the NEXT anchor tile is
// converted as soon as the last N tile of the current one has consumed block kb (per-block
// mbarriers), so the conversion and the fp32 stream hide behind the MMAs.
//
// CTA pairs (CG = 2, the dim = 512 similarity): the kernel runs as clusters of two CTAs on one
// TPC.  The leader's warp 1 issues tcgen05.mma.cta_group::2 for a 256-anchor x 128-class tile;
// each CTA converts its own 128 anchors into its own tensor memory, stages only HALF of every
// text tile (64 classes) in its shared memory and drains its own 128 accumulator rows.  This
// halves the text bytes every SM pulls from L2 - the bound of the single-CTA kernel (27 GB of
// L2 -> SM traffic per launch, 12.9 TB/s).  Cross-CTA protocol: both CTAs' text loads complete
// on the leader's `b_full` (cp.async.bulk.tensor .cta_group::2); tcgen05.commit multicasts
// `b_empty` / `a_free` / `t_full` to both CTAs; the converter and epilogue warps of the second
// CTA arrive remotely on the leader's `a_ready` / `t_empty`.
#include "common.cuh"
#include "ptx.cuh"
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <type_traits>

namespace ovdet {
namespace {

constexpr int F_BLOCK_M = 128;
constexpr int F_BLOCK_N = 128;
constexpr int F_BLOCK_K = 64;
constexpr int F_MAX_KB = 8;                       // dim <= 512: A fills 256 TMEM columns
constexpr int F_MAX_SLOTS = 12;                   // slots of the A ring (BN64: 12 x 32 columns)
#ifndef OVDET_F_CONV8
#define OVDET_F_CONV8 1                          // eight converter warps, half a k block each (see CONV8 in the kernel)
#endif
#ifndef OVDET_F_BN64
#define OVDET_F_BN64 0                           // CTA-pair cosine kernels: 64-class N tiles, 12-slot A ring (see BN64)
#endif
#ifndef OVDET_F_A_STAGES
#define OVDET_F_A_STAGES 4
#endif
#ifndef OVDET_F_B_STAGES
#define OVDET_F_B_STAGES 4
#endif
#ifndef OVDET_F_AHEAD2
#define OVDET_F_AHEAD2 1                         // converter publishes blocks late (see AHEAD2 in the kernel); 2: at most two
#endif
constexpr int F_A_STAGES = OVDET_F_A_STAGES;      // fp32 activation ring (tuning: -DOVDET_F_A_STAGES=n)
constexpr int F_A_STAGE_BYTES = F_BLOCK_K * F_BLOCK_M * 4;     // 32 KiB fp32 [k][anchor]
constexpr int F_THREADS = 384;
constexpr int F_TMEM_COLS = 512;
constexpr int F_ACC_COL = 0;
constexpr int F_A_COL = 256;
constexpr int F_PITCH = 33;                       // staging pitch (floats) of the scalar logit stores
constexpr int F_VPITCH = 36;                      // staging pitch of the 16-byte logit stores (144-byte rows)
constexpr int F_MAX_LEVELS = 4;

// CG = CTAs per MMA (tcgen05 cta_group): 1 = one CTA per 128-anchor tile; 2 = a CTA PAIR on one
// 256-anchor tile, each CTA staging HALF of every text tile (64 classes x 64 k = 8 KiB), which
// halves the L2 -> SM text traffic that bounds the single-CTA kernel.  The text ring always holds
// 64 KiB (4 x 16 KiB): deep enough to cover the L2 latency, shallow enough that the activation
// loads issued at an anchor-tile boundary do not queue behind it.  With CG = 2 a stage holds TWO
// k blocks (two 8 KiB boxes on one barrier), so the single MMA-issuing thread - which now has
// half the time per k block - synchronises once per 8 MMAs instead of once per 4.
// KPS3 (the projected CTA-pair kernel, K = 4 x 64 + 16): THREE k blocks per stage, three stages of
// 24 KiB.  With two blocks per stage the 16-wide constant block was a stage of its own - one MMA
// (64 tensor cycles) behind a full barrier round trip of the issuing thread: dropping that stage in a
// timing experiment took 13 % off the kernel for 6 % of its MMA work.  Stages are now {0,1,2} {3,const}.
// BN64 (the dim = 512 CTA-pair cosine kernels): N tiles of 64 classes - a text box is [32 rows x 64 k] =
// 4 KiB and a stage holds FOUR k blocks (16 MMAs of 32 cycles per barrier round trip).
template <int CG, bool KPS3 = false, bool BN64 = false>
struct FSmem {
  static constexpr int bn = BN64 ? 64 : F_BLOCK_N;                   // classes per N tile
  static constexpr int kps = KPS3 ? 3 : (BN64 ? 4 : CG);             // k blocks per text stage
  static constexpr int b_sub_bytes = (bn / CG) * F_BLOCK_K * 2;      // one TMA box: [N / CG rows x 64 k]
  static constexpr int b_stages = KPS3 ? 3 : OVDET_F_B_STAGES;
  static constexpr int b_stage_bytes = kps * b_sub_bytes;            // 16 KiB
  static constexpr int b_off = 0;
  static constexpr int a_off = b_off + b_stages * b_stage_bytes;                     // 64 KiB
  static constexpr int epi_off = a_off + F_A_STAGES * F_A_STAGE_BYTES;               // +128 KiB
  static constexpr int epi_warp_bytes = 32 * F_VPITCH * 4;           // 4608: fp32 staging, or two 2 KiB TMA-store buffers
  static constexpr int epi_bytes = 4 * epi_warp_bytes;
  static constexpr int norm_off = epi_off + epi_bytes;
  static constexpr int norm_bytes = 3 * 3 * F_BLOCK_M * 4;           // 3 tiles in flight x (two partial sums of squares, row scale)
  static constexpr int xbuf_off = norm_off + norm_bytes;              // EPI2: group 1's partial rows, two parities
  static constexpr int xbuf_bytes = 2 * 3 * F_BLOCK_M * 4;
  static constexpr int bar_off = xbuf_off + xbuf_bytes;
  // b_full, b_empty, as_full, as_empty, a_ready, a_free, tmem_full, tmem_empty, norm_ready
  static constexpr int num_bars = 2 * b_stages + 2 * F_A_STAGES + 2 * F_MAX_SLOTS + 2 + 2 + 3;
  static constexpr int tmem_ptr_off = bar_off + num_bars * 8;
  static constexpr int total = tmem_ptr_off + 16;
  static constexpr int bytes = total + 1024;
};
static_assert(FSmem<1>::bytes <= 227 * 1024 && FSmem<2>::bytes <= 227 * 1024 && FSmem<2, true>::bytes <= 227 * 1024 &&
              FSmem<2, false, true>::bytes <= 227 * 1024,
              "shared memory budget");

struct LevelMaps { CUtensorMap m[F_MAX_LEVELS]; };   // activations; (projected) one text operand per level

struct FusedParams {
  int levels;
  int batch;
  int hw[F_MAX_LEVELS];
  int mt[F_MAX_LEVELS];            // M tiles per image of the level
  int off[F_MAX_LEVELS];           // anchor offset of the level in the concatenated order
  int tile_start[F_MAX_LEVELS + 1];
  int anchors;                     // per image, all levels
  int classes;
  int kb;                          // k blocks of 64 the MMA walks (3 * kb_in with split3)
  int kb_in;                       // fp32 input blocks per anchor tile: ceil(dim / 64)
  int normalize;                   // 1: scale rows by 1 / max(||x||, 1e-12) (cosine); 0: raw dot product
  int n_tiles;                     // N tiles per anchor tile (projected: ng_tiles of G' first, then classes)
  int proj;                        // 1: projected similarity (hidden features in, quadratic-form norm)
  int ng_tiles;                    // projected: N tiles holding G' (kop rows)
  int cpad;                        // projected: first G' row of the operand (classes padded to 128)
  int kop;                         // projected: operand row length = kb_in * 64 + 16
  int text_batched;
  float alpha, beta;
  void* logits;
  int logits_bf16;
  int logits_tma;                  // bf16 logits leave through TMA stores (cmaps: one [classes, hw, batch] map per level)
  long long ldc;
  float* row_max;
  int* row_arg;
  float* inv_norm;
  // Small launches (fewer anchor tiles than CTA pairs): the class tiles of an anchor tile are split
  // over `nsplit` work items; each writes its partial (max, argmax) and the last one to arrive
  // (per 32-row group, counted with an atomic) merges them.  Scratch is caller memory.
  int nsplit;
  float* part_max;                 // [nsplit, batch * anchors]
  int* part_arg;                   // [nsplit, batch * anchors]
  int* part_count;                 // [tiles * 4], zero on entry, left zero
  // Vocabulary-parallel exchange (vocab_parallel.cu): this launch holds classes
  // [vp_class_offset, vp_class_offset + classes) of a vocabulary sharded over vp_world GPUs.  A row's
  // (score, global class) is packed into one 64-bit key and max-reduced straight into EVERY rank's
  // key array with system-scope atomics over NVLink peer mappings, from the epilogue warp that
  // produced it - no local row_max / row_arg, no separate collective.  The step counter lives in
  // device memory (the key-array parity is its low bit), so a captured graph replays correctly.
  int vp_world;                    // 0: off
  int vp_class_offset;
  long long vp_rows;               // batch * anchors
  const unsigned long long* vp_step;              // this rank's counter: the step being contributed to
  unsigned long long* vp_keys[OVDET_MAX_PEERS];   // keys[2][vp_rows] of every rank (peer-mapped)
  int dbg;
  unsigned long long* trace;       // OVDET_TRACE builds only: per-role clock stamps of CTA 0 (tools/trace_fused.py)
};
#ifdef OVDET_TRACE
#define OVDET_TR(role, tag)                                                                       \
  do {                                                                                            \
    if (p.trace != nullptr && blockIdx.x == 0 && lane == 0 && tr_n < 32768)                       \
      p.trace[(role) * 32768 + tr_n++] = ((unsigned long long)(tag) << 48) | ((unsigned long long)clock64() & 0xFFFFFFFFFFFFull); \
  } while (0)
#else
#define OVDET_TR(role, tag) do { } while (0)
#endif

// With CG == 2 the two CTAs of a pair take tiles 2i and 2i+1; `mt` is then rounded up to an even
// count per (level, image) when the text is per-image, so that a pair never straddles two images
// (both CTAs multiply against the same text tile).  A tile index past the end, or an M tile that
// starts past the level's last anchor, has rows <= 0: it loads zeros (TMA out-of-bounds fill)
// and writes nothing.
struct TileCoord { int b, level, m0, rows; long long out_row0; };

__device__ __forceinline__ TileCoord decode_tile(const FusedParams& p, int tile) {
  const bool past_end = tile >= p.tile_start[p.levels];
  if (past_end) tile = p.tile_start[p.levels] - 1;
  int l = 0;
#pragma unroll
  for (int i = 1; i < F_MAX_LEVELS; ++i)
    if (i < p.levels && tile >= p.tile_start[i]) l = i;
  const int r = tile - p.tile_start[l];
  TileCoord t;
  t.level = l;
  t.b = r / p.mt[l];
  t.m0 = (r - t.b * p.mt[l]) * F_BLOCK_M;
  t.rows = past_end ? 0 : min(F_BLOCK_M, p.hw[l] - t.m0);
  t.out_row0 = (long long)t.b * p.anchors + p.off[l] + t.m0;
  return t;
}

// KB_T  8: dim == 512, every k-block loop unrolled; 0: run-time block count.
// SPLIT3: fp32-accurate product from three bf16 passes (resident for dim <= 128 - the attention
// row -, streaming through an 8-slot ring of the A region for dim <= 512 and classes <= 128).  x = hi + lo with
// hi = bf16(x), lo = bf16(x - hi); the activation blocks are written to tensor memory as
// [hi | hi | lo] and the text operand is laid out [hi | lo | hi] (ovdet_cast_text), so the plain
// block loop accumulates hi*hi + hi*lo + lo*hi (the lo*lo term is < 2^-16 relative).
//
// PROJ ("next" row f-2): the head's 1x1 projection folded into the vocabulary (ops.py
// project_vocabulary).  The activations are the HIDDEN features x (K = hidden + 1 with the
// constant 1 in an extra 16-wide k block that is written to tensor memory once); the operand of
// level l holds the projected classes [W^T t_c | <b, t_c>] and then the rows of
// G' = [[W^T W, W^T b], [b^T W, b^T b]].  The first ng N tiles of every anchor tile multiply
// against G': their epilogue accumulates q = sum_j (x' G')_j x'_j = ||W x + b||^2 with x' read
// back from the A region of tensor memory; the class tiles keep the raw running max/argmax and
// the row is scaled by alpha / sqrt(q) once at the end.  No logits in this mode.
// IN16: the activations arrive as bf16 (the head ran under autocast) instead of fp32: the TMA box
// is [64 k x 128 anchors] bf16 (16 KiB of the 32 KiB stage), the converter widens, accumulates the
// sum of squares in fp32 and re-packs - half the HBM and PCIe bytes of the fp32 input.
// MODE (what the epilogue produces; compile-time so that the scores-only kernel carries none of the
// other modes' code or live registers - with them folded in at run time the projected mode lost 6 %
// and the main mode 1.5 %): 0 = row max / argmax only, 1 = logits (and optionally max / argmax),
// 2 = vocabulary-parallel keys.
// EPI2 (kernels that fit 128 registers; with logits only the TMA-store path, one 2 KiB staging
// buffer per warp): a SECOND epilogue warpgroup (warps 12-15, 512 threads).  Group e drains the N tiles whose accumulator stage is e, so the two stages are
// emptied concurrently; the groups' partial (max, argmax, q) of a row meet in shared memory at the
// end of the anchor tile and group 0 emits.
// F16OP (the "fp16" precision tier, CTA pairs, dim = 512): both operands are fp16 instead of bf16 - the
// same tensor rate with 11-bit significands, |dlogit| ~ 1e-5 (max ~6e-5 over 10^7 logits) against
// ~4e-3 for bf16.  fp16 has a narrow exponent range, so every anchor row is scaled by a power of two s
// chosen from its first 64 channels (their sampled maximum lands in [4, 8)); the sum of squares is
// taken over the scaled values and cos = <s x, t> / ||s x|| is unchanged.  Rows whose later channels
// exceed 8000 x that maximum saturate at +-65504 (finite).  The text operand carries the unit rows
// times 16 (ovdet_l2norm_text, split = 3); the 1/16 goes into the row scale.
template <int KB_T, bool SPLIT3, int CG, bool PROJ, bool IN16 = false, int MODE = 0, bool EPI2 = false,
          bool F16OP = false, bool C8 = false>
__global__ void __launch_bounds__((EPI2 || C8) ? F_THREADS + 128 : F_THREADS, 1)
sim_fused_kernel(const __grid_constant__ LevelMaps amaps, const __grid_constant__ LevelMaps bmaps,
                 const __grid_constant__ LevelMaps cmaps, const FusedParams p) {
  // BN64: N tiles of 64 classes.  Tensor memory then holds two 64-column accumulators and a TWELVE-slot A
  // ring (12 x 32 columns = 1.5 anchor tiles) instead of two 128-column accumulators and exactly one tile:
  // the first four blocks of the next anchor tile are converted and stored while the current tile is still
  // being multiplied, and the other four go into slots the last N tile releases in its FIRST half.  With
  // the 8-slot ring every block of the next tile had to wait for the last N tile's MMAs on the same slot
  // and the converter - ~550 cycles per block against 256 cycles of MMAs - left the tensor pipe idle for
  // ~2800 cycles per anchor tile (tools/trace_fused.py, profiles/r2_trace_*.txt).
  // CONV8 (the dim = 512 cosine kernels without a second epilogue group): EIGHT converter warps - warps
  // 12-15 join warps 4-7, each warp converts HALF of every 64-k block (32 k = 16 TMEM columns of its 32
  // rows) and holds FOUR half blocks in registers.  The boundary between anchor tiles is a serial chain in
  // the converter (tools/trace_fused.py: ~6000 cycles for 8 publishes + 5 conversions of ~600 cycles with
  // four warps at 96 B/clk of shared-memory loads); with eight warps a conversion takes half as long,
  // the shared-memory pipe runs at its 128 B/clk, and four blocks instead of three are ready before
  // the boundary begins.  The sum of squares of a row is the sum of the two halves' partial sums, formed
  // by the epilogue (which also writes inv_norm).
  // It pays where anchor-tile boundaries are frequent - few N tiles per anchor tile (80 prompts, batch 64:
  // 0.220 -> 0.179 ms, 0.78 -> 0.94 of the HBM bandwidth) - and costs 1-3 % at 1203 prompts, where the
  // 128-register cap of a 512-thread block squeezes the epilogue: the launcher picks it for <= 3 N tiles.
  static_assert(!C8 || (!SPLIT3 && !PROJ && KB_T == 8 && !EPI2 && MODE != 2), "eight converter warps: the cosine kernels");
  constexpr bool CONV8 = C8;
  constexpr bool BN64 = CG == 2 && !PROJ && !SPLIT3 && KB_T == 8 && OVDET_F_BN64;
  constexpr int BN = BN64 ? 64 : F_BLOCK_N;
  constexpr int A_COL = BN64 ? 2 * BN : F_A_COL;                   // first column of the A ring
  using FSmem = ovdet::FSmem<CG, (PROJ && CG == 2), BN64>;
  constexpr int F_B_STAGES = FSmem::b_stages;
  constexpr int F_B_STAGE_BYTES = FSmem::b_stage_bytes;
  constexpr int KPS = FSmem::kps;
  constexpr int F_B_SUB_BYTES = FSmem::b_sub_bytes;
  static_assert(CG == 1 || KB_T > 0, "CTA pairs need a compile-time k-block count");
  static_assert(!(PROJ && SPLIT3), "the projected mode is a single bf16 pass");
  static_assert(!EPI2 || MODE != 2, "two epilogue groups: not for the key exchange");
  static_assert(!F16OP || (!SPLIT3 && !PROJ && !IN16), "fp16 operands: the plain cosine similarity from fp32 input");
  // cluster rank: 0 = leader (issues the MMAs, owns the barriers the pair synchronises on)
  const uint32_t rank = CG == 2 ? ptx::cluster_ctarank() : 0u;
  const int pair0 = blockIdx.x / CG, pair_stride = gridDim.x / CG;
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t raw = ptx::smem_u32(smem_dyn);
  const uint32_t base = (raw + 1023u) & ~1023u;
  unsigned char* base_ptr = smem_dyn + (base - raw);

  const uint32_t smem_b = base + FSmem::b_off;
  const uint32_t smem_a = base + FSmem::a_off;
  const float* a_stage_ptr = reinterpret_cast<const float*>(base_ptr + FSmem::a_off);
  float* epi_stage = reinterpret_cast<float*>(base_ptr + FSmem::epi_off);
  float* norm_s = reinterpret_cast<float*>(base_ptr + FSmem::norm_off);
  const uint32_t bars = base + FSmem::bar_off;
  int bi_ = 0;
  const uint32_t b_full0 = bars + 8u * bi_;      bi_ += F_B_STAGES;
  const uint32_t b_empty0 = bars + 8u * bi_;     bi_ += F_B_STAGES;
  const uint32_t as_full0 = bars + 8u * bi_;     bi_ += F_A_STAGES;
  const uint32_t as_empty0 = bars + 8u * bi_;    bi_ += F_A_STAGES;
  const uint32_t a_ready0 = bars + 8u * bi_;     bi_ += F_MAX_SLOTS;
  const uint32_t a_free0 = bars + 8u * bi_;      bi_ += F_MAX_SLOTS;
  const uint32_t t_full0 = bars + 8u * bi_;      bi_ += 2;
  const uint32_t t_empty0 = bars + 8u * bi_;     bi_ += 2;
  const uint32_t n_ready0 = bars + 8u * bi_;
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(base_ptr + FSmem::tmem_ptr_off);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
#ifdef OVDET_TRACE
  int tr_n = 0;
#endif

  if (warp == 0 && lane == 0) {
    for (int l = 0; l < p.levels; ++l) { ptx::prefetch_tmap(&amaps.m[l]); ptx::prefetch_tmap(&bmaps.m[PROJ ? l : 0]); }
    for (int s = 0; s < F_B_STAGES; ++s) { ptx::mbar_init(b_full0 + 8u * s, 1); ptx::mbar_init(b_empty0 + 8u * s, 1); }
    for (int s = 0; s < F_A_STAGES; ++s) { ptx::mbar_init(as_full0 + 8u * s, 1); ptx::mbar_init(as_empty0 + 8u * s, CONV8 ? 8 : 4); }
    // a_ready / t_empty collect the converter / epilogue warps of BOTH CTAs on the leader
    for (int k = 0; k < F_MAX_SLOTS; ++k) { ptx::mbar_init(a_ready0 + 8u * k, (CONV8 ? 8 : 4) * CG);
      // projected: the epilogue warps read x' back from the A region, so they release it too
      ptx::mbar_init(a_free0 + 8u * k, PROJ ? (EPI2 ? 9 : 5) : 1); }
    for (int s = 0; s < 2; ++s) { ptx::mbar_init(t_full0 + 8u * s, 1); ptx::mbar_init(t_empty0 + 8u * s, 4 * CG); }
    for (int s = 0; s < 3; ++s) ptx::mbar_init(n_ready0 + 8u * s, CONV8 ? 8 : 4);
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc_cg<CG>(base + FSmem::tmem_ptr_off, F_TMEM_COLS);
    ptx::tmem_relinquish_cg<CG>();
  }
  ptx::tc_fence_before();
  if constexpr (CG == 2) ptx::cluster_sync_all();      // peer barriers are initialised before any remote arrive
  else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  // programmatic dependent launch: a dependent kernel (the decode of ovdet_head_step) may be
  // scheduled from here on; it waits on this grid's completion itself before reading the scores
  cudaTriggerProgrammaticLaunchCompletion();

  const unsigned lazy_ns = (p.dbg & 16) ? 0u : ((p.dbg & 32) ? 256u : 64u);   // back-off of the non-critical waits
  const int total_tiles = p.tile_start[p.levels];
  const int total_pairs = (total_tiles + CG - 1) / CG;
  const int NSPLIT = p.nsplit;                                   // >= 1
  const int total_work = total_pairs * NSPLIT;
  // barriers of the leader CTA that both CTAs arrive on (shared::cluster addresses)
  const uint32_t lead_b_full0 = CG == 2 ? ptx::map_to_cta(b_full0, 0) : b_full0;
  const uint32_t lead_a_ready0 = CG == 2 ? ptx::map_to_cta(a_ready0, 0) : a_ready0;
  const uint32_t lead_t_empty0 = CG == 2 ? ptx::map_to_cta(t_empty0, 0) : t_empty0;
  // (KB_T counts the blocks the MMA walks: with SPLIT3 three per input block)
  const int KB_IN = KB_T ? (SPLIT3 ? KB_T / 3 : KB_T) : p.kb_in;  // fp32 input blocks per anchor tile
  const int KB = PROJ ? KB_IN + 1 : (KB_T ? KB_T : p.kb);        // k blocks the MMA walks
  const int KB_A = PROJ ? KB_IN : KB;                            // A blocks rewritten per anchor tile
  // The A region is a ring of R slots of 32 columns.  Resident modes: R = KB_A, block i lives in
  // slot i for the whole anchor tile.  Streaming (SPLIT3 with more than 8 blocks, one N tile only):
  // R = 8 and block i of tile lt takes slot (lt * KB_A + i) % 8 - every block is consumed exactly
  // once, so its slot is handed back as soon as its MMAs retire.
  // BN64: R = 12 > KB_A = 8: block i of tile lt takes slot (8 lt + i) % 12 for the tile's lifetime.
  const int R = BN64 ? F_MAX_SLOTS : ((SPLIT3 && KB_A > F_MAX_KB) ? F_MAX_KB : KB_A);
  auto a_slot = [&](uint32_t lt_, int i) { return (int)((lt_ * (uint32_t)KB_A + (uint32_t)i) % (uint32_t)R); };
  auto a_phase = [&](uint32_t lt_, int i) { return ((lt_ * (uint32_t)KB_A + (uint32_t)i) / (uint32_t)R) & 1u; };
  // SPLIT3 visits the blocks input-block-major: i = 3 * j + part, parts (hi, hi, lo) against the
  // text segments (hi, lo, hi) of the operand layout [hi | lo | hi]
  auto b_kblock = [&](int i) { return SPLIT3 ? (i % 3) * KB_IN + i / 3 : i; };
  const int NSTAGE = (KB + KPS - 1) / KPS;                       // text stages per N tile
  const int NT = p.n_tiles;
  const int NG = PROJ ? p.ng_tiles : 0;
  // N tile nt of an anchor tile: first operand row, MMA N (multiple of 16), valid columns
  auto ntile_row0 = [&](int nt) { return nt < NG ? p.cpad + nt * BN : (nt - NG) * BN; };
  auto ntile_valid = [&](int nt) {
    const int n = nt < NG ? p.kop - nt * BN : p.classes - (nt - NG) * BN;
    return n >= BN ? BN : n;
  };
  auto ntile_nsize = [&](int nt) { return (ntile_valid(nt) + 15) & ~15; };

  // Producer / issuer warps run warp-uniform control flow; `issue` is 1 in one elected lane and
  // predicates the single-thread instructions (see ptx::elect_one).
  if (warp == 0) {
    // ================================ text (B) producer ======================================
    const uint32_t issue = ptx::elect_one();
    uint32_t g = 0;                                    // N tiles produced so far
    for (int w = pair0; w < total_work; w += pair_stride) {
      const int tile = (NSPLIT == 1 ? w : w / NSPLIT) * CG + (int)rank;
      const int nt_b = NSPLIT == 1 ? 0 : (w % NSPLIT) * NT / NSPLIT, nt_e = NSPLIT == 1 ? NT : (w % NSPLIT + 1) * NT / NSPLIT;
      (void)nt_b; (void)nt_e;
      const TileCoord tc = decode_tile(p, tile);
      const int tb = p.text_batched ? tc.b : 0;
      const CUtensorMap* bmap = &bmaps.m[PROJ ? tc.level : 0];
      for (int nt = nt_b; nt < nt_e; ++nt, ++g) {
        const int row0 = ntile_row0(nt);
        const int n_half = ntile_nsize(nt) >> 1;
        (void)n_half;
#pragma unroll
        for (int sb = 0; sb < NSTAGE; ++sb) {
          const uint32_t it = g * (uint32_t)NSTAGE + (uint32_t)sb;         // stages produced so far
          const uint32_t s = it % F_B_STAGES, ph = (it / F_B_STAGES) & 1u;
          const int subs = min(KPS, KB - sb * KPS);                         // k blocks in this stage
          ptx::mbar_wait_lazy(b_empty0 + 8u * s, ph ^ 1u, lazy_ns);
          if constexpr (CG == 1) {
            ptx::mbar_arrive_expect_tx_if(issue, b_full0 + 8u * s, F_B_STAGE_BYTES);
            ptx::tma_load_3d_if(issue, smem_b + s * F_B_STAGE_BYTES, bmap, b_full0 + 8u * s,
                                b_kblock(sb) * F_BLOCK_K, row0, tb);
          } else {
            // this CTA's half of the N tile (rows [rank * n/2, (rank + 1) * n/2) of it) lands in its
            // own shared memory; both halves complete on the LEADER's barrier, which expects both
            if (rank == 0) ptx::mbar_arrive_expect_tx_if(issue, b_full0 + 8u * s, 2 * subs * F_B_SUB_BYTES);
#pragma unroll
            for (int j = 0; j < KPS; ++j)
              if (j < subs)
                ptx::tma_load_3d_pair_if(issue, smem_b + s * F_B_STAGE_BYTES + j * F_B_SUB_BYTES, bmap,
                                         lead_b_full0 + 8u * s, b_kblock(sb * KPS + j) * F_BLOCK_K,
                                         row0 + (int)rank * n_half, tb);
          }
        }
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ================================ MMA issuer (leader CTA) ================================
    // broadcast from lane 0 so that the compiler keeps every MMA operand in uniform registers
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t smem_b_u = __shfl_sync(0xffffffffu, smem_b, 0);
    uint32_t g = 0, lt = 0;
    for (int w = pair0; w < total_work; w += pair_stride, ++lt) {
      const int nt_b = NSPLIT == 1 ? 0 : (w % NSPLIT) * NT / NSPLIT, nt_e = NSPLIT == 1 ? NT : (w % NSPLIT + 1) * NT / NSPLIT;
      for (int nt = nt_b; nt < nt_e; ++nt, ++g) {
        const int n_size = ntile_nsize(nt);
        const uint32_t idesc = F16OP ? ptx::umma_idesc_f16_f32(F_BLOCK_M * CG, (uint32_t)n_size)
                                     : ptx::umma_idesc_bf16_f32(F_BLOCK_M * CG, (uint32_t)n_size);
        const uint32_t as = g & 1u;
        const bool first_nt = nt == nt_b, last_nt = nt == nt_e - 1;
        const uint32_t it0 = g * (uint32_t)NSTAGE;
        // peek at the first text stage while waiting for the accumulator to drain
        bool ready = ptx::mbar_try_wait(b_full0 + 8u * (it0 % F_B_STAGES), (it0 / F_B_STAGES) & 1u);
        OVDET_TR(0, 1);
        ptx::mbar_wait(t_empty0 + 8u * as, ((g >> 1) & 1u) ^ 1u);
        OVDET_TR(0, 2);
        const uint32_t d_tmem = tmem_u + (uint32_t)F_ACC_COL + as * (uint32_t)BN;
#pragma unroll
        for (int sb = 0; sb < NSTAGE; ++sb) {
          const uint32_t it = it0 + (uint32_t)sb;
          const uint32_t s = it % F_B_STAGES, ph = (it / F_B_STAGES) & 1u;
          if (first_nt) {                                                // A blocks converted (both CTAs)?
#pragma unroll
            for (int j = 0; j < KPS; ++j)
              if (sb * KPS + j < KB_A)
                ptx::mbar_wait(a_ready0 + 8u * a_slot(lt, sb * KPS + j), a_phase(lt, sb * KPS + j));
            OVDET_TR(0, 3);
          }
          ptx::mbar_wait_if_not(ready, b_full0 + 8u * s, ph);
          OVDET_TR(0, 4);
          ptx::tc_fence_after();
          if (sb + 1 < NSTAGE)                                           // hide the next wait's latency
            ready = ptx::mbar_try_wait(b_full0 + 8u * ((it + 1) % F_B_STAGES), ((it + 1) / F_B_STAGES) & 1u);
          if (ptx::elect_one()) {
#pragma unroll
            for (int j = 0; j < KPS; ++j) {
              const int kb = sb * KPS + j;
              if (kb < KB) {
                const uint64_t b_desc = ptx::umma_desc_k_sw128(smem_b_u + s * F_B_STAGE_BYTES + j * F_B_SUB_BYTES);
                const uint32_t a_tmem = tmem_u + (uint32_t)(A_COL + (kb < KB_A ? a_slot(lt, kb) : kb) * 32);
                // projected: the last block is the 16-wide constant block of x' = [x, 1]
                const int ksteps = (PROJ && kb == KB - 1) ? 1 : F_BLOCK_K / 16;
#pragma unroll
                for (int k = 0; k < F_BLOCK_K / 16; ++k)
                  if (k < ksteps)
                    ptx::umma_bf16_ts_cg<CG>(d_tmem, a_tmem + 8u * k, b_desc + 2u * k, idesc, (kb | k) != 0);
              }
            }
            ptx::umma_commit_cg<CG>(b_empty0 + 8u * s);                  // text stage reusable (both CTAs)
            if (last_nt) {                                               // these A blocks may be overwritten
#pragma unroll
              for (int j = 0; j < KPS; ++j)
                if (sb * KPS + j < KB_A) ptx::umma_commit_cg<CG>(a_free0 + 8u * a_slot(lt, sb * KPS + j));
            }
          }
          __syncwarp();
        }
        if (ptx::elect_one()) ptx::umma_commit_cg<CG>(t_full0 + 8u * as);
        __syncwarp();
        OVDET_TR(0, 5);
      }
    }
  } else if (warp == 2) {
    // ================================ activation (A) producer ================================
    const uint32_t issue = ptx::elect_one();
    uint32_t ia = 0;
    for (int w = pair0; w < total_work; w += pair_stride) {
      const int tile = (NSPLIT == 1 ? w : w / NSPLIT) * CG + (int)rank;
      const int nt_b = NSPLIT == 1 ? 0 : (w % NSPLIT) * NT / NSPLIT, nt_e = NSPLIT == 1 ? NT : (w % NSPLIT + 1) * NT / NSPLIT;
      (void)nt_b; (void)nt_e;
      const TileCoord tc = decode_tile(p, tile);
      const CUtensorMap* map = &amaps.m[tc.level];
#pragma unroll
      for (int kb = 0; kb < KB_IN; ++kb, ++ia) {
        const uint32_t s = ia % F_A_STAGES;
        const uint32_t ph = (ia / F_A_STAGES) & 1u;
        ptx::mbar_wait_lazy(as_empty0 + 8u * s, ph ^ 1u, lazy_ns);
        OVDET_TR(4, 30);
        ptx::mbar_arrive_expect_tx_if(issue, as_full0 + 8u * s, IN16 ? F_A_STAGE_BYTES / 2 : F_A_STAGE_BYTES);
        ptx::tma_load_3d_if(issue, smem_a + s * F_A_STAGE_BYTES, map, as_full0 + 8u * s, tc.m0,
                            kb * F_BLOCK_K, tc.b);
      }
    }
  } else if (warp == 3) {
    // ================================ L2 prefetcher ==========================================
    // Experiment kept behind OVDET_DBG=1: pull the next anchor tile into L2 one tile ahead.  With
    // the 2-stage activation ring of the first version it hid the HBM latency of the loads issued
    // at a tile boundary; with the 4-stage ring it only competes with the text stream for the
    // SM's TMA path (measured 2.25 ms with it vs 2.14 ms without, single-CTA kernel), so it is off.
    const uint32_t issue = ptx::elect_one();
    uint32_t lt = 0;
    for (int w = pair0; w < total_work; w += pair_stride, ++lt) {
      const int tile = (NSPLIT == 1 ? w : w / NSPLIT) * CG + (int)rank;
      const int nt_b = NSPLIT == 1 ? 0 : (w % NSPLIT) * NT / NSPLIT, nt_e = NSPLIT == 1 ? NT : (w % NSPLIT + 1) * NT / NSPLIT;
      (void)nt_b; (void)nt_e;
      const int next = tile + pair_stride * CG;
      if (next >= total_tiles || !(p.dbg & 1)) break;       // off unless OVDET_DBG bit 0 is set
      if (CG == 1 && (p.dbg & 4)) {
        ptx::mbar_wait(a_ready0, lt & 1u);
      } else {                                         // the current tile's first N tile is done
        const uint32_t g0 = lt * (uint32_t)NT;     // (prefetch experiment: only meaningful with nsplit == 1)
        ptx::mbar_wait(t_full0 + 8u * (g0 & 1u), (g0 >> 1) & 1u);
      }
      const TileCoord tc = decode_tile(p, next);
      const CUtensorMap* map = &amaps.m[tc.level];
#pragma unroll
      for (int kb = 0; kb < KB_IN; ++kb)
        ptx::tma_prefetch_l2_3d_if(issue, map, tc.m0, kb * F_BLOCK_K, tc.b);
    }
  } else if (CONV8 && ((warp >= 4 && warp < 8) || warp >= 12)) {
    // ================================ converters, eight warps (CONV8) ========================
    if constexpr (CONV8) {
      const int lg = warp & 3;
      const int half = warp >= 12 ? 1 : 0;             // which 32 k of every 64-k block
      const int arow = lg * 32 + lane;                 // anchor row of the tile == TMEM lane
      constexpr int AH = 4;                            // half blocks held in registers
      uint32_t held[AH][16];
      bool poll_freed = false, poll_landed = false;    // results of the polls issued one block earlier
      uint32_t ia = 0, lt = 0;
      auto publish_half = [&](uint32_t lt_, int i, const uint32_t (&regs)[16], bool peeked) {
        const int t = a_slot(lt_, i);
        if (warp == 4) OVDET_TR(1, 12);
        if (!peeked) ptx::mbar_wait_lazy(a_free0 + 8u * t, a_phase(lt_, i) ^ 1u, lazy_ns >> 1);
        if (warp == 4) OVDET_TR(1, 13);
        ptx::tc_fence_after();
        ptx::tmem_st_32x32_x16(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(A_COL + t * 32 + 16 * half), regs);
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (CG == 2) ptx::mbar_arrive_cluster(lead_a_ready0 + 8u * t);
          else ptx::mbar_arrive(a_ready0 + 8u * t);
        }
        if (warp == 4) OVDET_TR(1, 14);
      };
      for (int w = pair0; w < total_work; w += pair_stride, ++lt) {
        float ss0 = 0.f, ss1 = 0.f, ss2 = 0.f, ss3 = 0.f;
        float row_scale = 1.0f;                        // F16OP: power of two applied before the fp16 rounding
        bool unit_rows = false;                        // F16OP: every row of this warp keeps scale 1 (see below)
#pragma unroll
        for (int kb = 0; kb < KB_T; ++kb, ++ia) {
          const uint32_t s = ia % F_A_STAGES;
          if (warp == 4) OVDET_TR(1, 10);
          // (the CTA's FIRST tile publishes every block as soon as it is converted: its slots have never
          // been used, and holding blocks back would only delay the first MMA - at batch 1, where a CTA pair
          // has one tile, by the conversion time of four blocks)
          if (lt > 0 && kb >= AH) publish_half(lt, kb - AH, held[kb % AH], poll_freed);
          if (!poll_landed) ptx::mbar_wait_lazy(as_full0 + 8u * s, (ia / F_A_STAGES) & 1u, lazy_ns >> 1);
          if (warp == 4) OVDET_TR(1, 11);
          const float* col = a_stage_ptr + s * (F_A_STAGE_BYTES / 4) + arow;
          const __nv_bfloat16* col16 = reinterpret_cast<const __nv_bfloat16*>(a_stage_ptr + s * (F_A_STAGE_BYTES / 4)) + arow;
          auto ldx = [&](int k) -> float {               // k = position in the 64-k block
            if constexpr (IN16) return __bfloat162float(col16[k * F_BLOCK_M]);
            else return col[k * F_BLOCK_M];
          };
          if constexpr (F16OP) {
            if (kb == 0) {                               // both halves take the same eight samples: the same scale
              float m = 0.f;
#pragma unroll
              for (int k = 0; k < 64; k += 8) m = fmaxf(m, fabsf(ldx(k)));
              const uint32_t e = (__float_as_uint(m) >> 23) & 0xffu;
              row_scale = e >= 2u ? __uint_as_float((256u - e) << 23) : 1.0f;
              // Every row of the warp with its sampled maximum in [2^-4, 2^5): fp16 holds such rows as they
              // are (full 11-bit significands down to 6e-5, head-room x2000 above the sample), so the warp
              // skips the 64 multiplies per block - they made the fp16 converter 40 % slower than the bf16
              // one (700 vs 495 cycles per block in the trace), on the boundary's critical chain.
              unit_rows = __all_sync(0xffffffffu, e >= 123u && e <= 131u);
              if (unit_rows) row_scale = 1.0f;
            }
          }
          {                                              // polls for the next iteration / the closing publishes
            const int nb = kb + 1;
            poll_freed = nb >= AH ? ptx::mbar_test_wait(a_free0 + 8u * a_slot(lt, nb - AH), a_phase(lt, nb - AH) ^ 1u) : true;
            poll_landed = ptx::mbar_test_wait(as_full0 + 8u * ((ia + 1u) % F_A_STAGES), ((ia + 1u) / F_A_STAGES) & 1u);
          }
          uint32_t (&packed)[16] = held[kb % AH];
          auto convert = [&](auto scaled) {
#pragma unroll
            for (int k = 0; k < 32; k += 4) {
              float x0 = ldx(32 * half + k + 0), x1 = ldx(32 * half + k + 1);
              float x2 = ldx(32 * half + k + 2), x3 = ldx(32 * half + k + 3);
              if constexpr (decltype(scaled)::value) { x0 *= row_scale; x1 *= row_scale; x2 *= row_scale; x3 *= row_scale; }
              ss0 = fmaf(x0, x0, ss0); ss1 = fmaf(x1, x1, ss1);
              ss2 = fmaf(x2, x2, ss2); ss3 = fmaf(x3, x3, ss3);
              packed[(k >> 1) + 0] = F16OP ? pack_f16x2_sat(x0, x1) : pack_bf16x2(x0, x1);
              packed[(k >> 1) + 1] = F16OP ? pack_f16x2_sat(x2, x3) : pack_bf16x2(x2, x3);
            }
          };
          if (F16OP && !unit_rows) convert(std::true_type{});
          else convert(std::false_type{});
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(as_empty0 + 8u * s);   // staging slot may be refilled (8 arrivals)
          if (lt == 0) publish_half(lt, kb, held[kb % AH], true);
        }
        if (lt > 0) {
          // the closing publishes: all their slots are polled first (the polls' latencies overlap), then stored
          bool freed[AH];
          freed[0] = poll_freed;
#pragma unroll
          for (int i = 1; i < AH; ++i)
            freed[i] = ptx::mbar_test_wait(a_free0 + 8u * a_slot(lt, KB_T - AH + i), a_phase(lt, KB_T - AH + i) ^ 1u);
#pragma unroll
          for (int i = 0; i < AH; ++i) publish_half(lt, KB_T - AH + i, held[(KB_T - AH + i) % AH], freed[i]);
        }
        const int slot = lt % 3;
        norm_s[(slot * 3 + half) * F_BLOCK_M + arow] = (ss0 + ss1) + (ss2 + ss3);
        if (half == 0) norm_s[(slot * 3 + 2) * F_BLOCK_M + arow] = row_scale;
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(n_ready0 + 8u * slot);
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ================================ converters =============================================
    const int lg = warp & 3;
    const int arow = lg * 32 + lane;                 // anchor row of the tile == TMEM lane
    if constexpr (PROJ) {
      // x' = [x, 1]: the constant block (k = KB_IN * 64: 1, then zeros) never changes
      uint32_t one[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) one[i] = 0u;
      one[0] = 0x00003F80u;                           // bf16 pair (1.0, 0.0)
      ptx::tmem_st_32x32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(A_COL + KB_IN * 32), one);
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
    }
    // AHEAD2 (the resident dim = 512 cosine kernels): converted blocks are published TWO blocks late, so
    // that at an anchor-tile boundary - where block kb of the A region is only released by the last N
    // tile's MMAs - two blocks of the next tile are already sitting in registers and the conversion of
    // block kb + 2 overlaps the wait for block kb's release.  The clock-stamp trace (tools/trace_fused.py)
    // showed the MMA warp waiting ~3800 cycles per anchor tile for converted blocks: a block takes the
    // converter ~550 cycles (the 32 KiB fp32 read alone is 256 cycles of shared-memory bandwidth) against
    // 256 cycles of MMAs, and with one block in flight those 550 cycles were serial, eight times over.
    constexpr bool AHEAD2 = !SPLIT3 && !PROJ && KB_T == 8 && OVDET_F_AHEAD2;
    // blocks held in registers: two, or three where the kernel has the registers (384 threads, no second
    // epilogue group): every block held is one ~550-cycle conversion less in the boundary's serial chain
    constexpr int AH = !AHEAD2 ? 1 : ((EPI2 || OVDET_F_AHEAD2 == 2) ? 2 : 3);
    uint32_t held[AH][32];                             // AHEAD2: blocks converted but not yet published
    bool poll_freed = false, poll_landed = false;      // AHEAD2: results of the polls issued one block earlier
    uint32_t ia = 0, lt = 0;
    // block `i` of the A region for anchor tile lt_: wait until the previous tile's MMAs have read it,
    // store, publish
    auto publish_at = [&](uint32_t lt_, int i, const uint32_t (&regs)[32], bool peeked) {
      const int t = a_slot(lt_, i);
      if (warp == 4) OVDET_TR(1, 12);
      if (!peeked) ptx::mbar_wait_lazy(a_free0 + 8u * t, a_phase(lt_, i) ^ 1u, lazy_ns >> 1);
      if (warp == 4) OVDET_TR(1, 13);
      ptx::tc_fence_after();
      ptx::tmem_st_32x32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(A_COL + t * 32), regs);
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CG == 2) ptx::mbar_arrive_cluster(lead_a_ready0 + 8u * t);
        else ptx::mbar_arrive(a_ready0 + 8u * t);
      }
      if (warp == 4) OVDET_TR(1, 14);
    };
    for (int w = pair0; w < total_work; w += pair_stride, ++lt) {
      const int tile = (NSPLIT == 1 ? w : w / NSPLIT) * CG + (int)rank;
      const int nt_b = NSPLIT == 1 ? 0 : (w % NSPLIT) * NT / NSPLIT, nt_e = NSPLIT == 1 ? NT : (w % NSPLIT + 1) * NT / NSPLIT;
      (void)nt_b; (void)nt_e;
      const TileCoord tc = decode_tile(p, tile);
      // the sum of squares in the association of the eight-warp converter above (one partial sum per
      // 32-k half of every block, four interleaved accumulators each): the kernels of one shape family
      // - with and without a second epilogue group - then agree BIT FOR BIT on every row norm
      float ss0 = 0.f, ss1 = 0.f, ss2 = 0.f, ss3 = 0.f;
      float st0 = 0.f, st1 = 0.f, st2 = 0.f, st3 = 0.f;
      float row_scale = 1.0f;                          // F16OP: power of two applied before the fp16 rounding
      bool unit_rows = false;                          // F16OP: every row of this warp keeps scale 1 (as above)
      auto publish = [&](int i, const uint32_t (&regs)[32]) { publish_at(lt, i, regs, false); };
#pragma unroll
      for (int kb = 0; kb < KB_IN; ++kb, ++ia) {
        const uint32_t s = ia % F_A_STAGES;
        if (warp == 4) OVDET_TR(1, 10);
        if constexpr (AHEAD2) {
          // the register buffer this block converts into still holds the block AH positions back.  Both of
          // this iteration's barriers were POLLED DURING THE PREVIOUS CONVERSION (poll_* below): a try_wait
          // on a completed barrier still takes 150-180 cycles in this kernel (tools/trace_fused.py), which
          // in the boundary's serial chain was a third of every block's time.  A poll that came back true
          // is final; one that came back false is followed by the real wait.
          if (lt > 0 && kb >= AH) publish_at(lt, kb - AH, held[kb % AH], poll_freed);   // (first tile: see below)
          if (!poll_landed) ptx::mbar_wait_lazy(as_full0 + 8u * s, (ia / F_A_STAGES) & 1u, lazy_ns >> 1);
        } else {
          ptx::mbar_wait_lazy(as_full0 + 8u * s, (ia / F_A_STAGES) & 1u, lazy_ns >> 1);
        }
        if (warp == 4) OVDET_TR(1, 11);
        const float* col = a_stage_ptr + s * (F_A_STAGE_BYTES / 4) + arow;
        const __nv_bfloat16* col16 = reinterpret_cast<const __nv_bfloat16*>(a_stage_ptr + s * (F_A_STAGE_BYTES / 4)) + arow;
        auto ldx = [&](int k) -> float {
          if constexpr (IN16) return __bfloat162float(col16[k * F_BLOCK_M]);
          else return col[k * F_BLOCK_M];
        };
        uint32_t (&packed)[32] = held[kb % AH];
        uint32_t packed_lo[32];                            // dead (eliminated) unless SPLIT3
        if constexpr (AHEAD2) {
          // polls for the NEXT iteration (or the tile's closing publishes): issued now, consumed after
          // this block's ~400 cycles of shared-memory reads and packing
          const int nb = kb + 1;                           // next block of this tile, or KB_IN = the closing publishes
          poll_freed = nb >= AH ? ptx::mbar_test_wait(a_free0 + 8u * a_slot(lt, nb - AH), a_phase(lt, nb - AH) ^ 1u) : true;
          poll_landed = ptx::mbar_test_wait(as_full0 + 8u * ((ia + 1u) % F_A_STAGES), ((ia + 1u) / F_A_STAGES) & 1u);
        }
        if constexpr (F16OP) {
          if (kb == 0) {
            // the row's power-of-two scale from eight samples of its first block: 2^(2 - exponent of the
            // largest), i.e. that sample lands in [4, 8); zero / denormal rows keep 1
            float m = 0.f;
#pragma unroll
            for (int k = 0; k < 64; k += 8) m = fmaxf(m, fabsf(ldx(k)));
            const uint32_t e = (__float_as_uint(m) >> 23) & 0xffu;
            row_scale = e >= 2u ? __uint_as_float((256u - e) << 23) : 1.0f;
            unit_rows = __all_sync(0xffffffffu, e >= 123u && e <= 131u);
            if (unit_rows) row_scale = 1.0f;
          }
        }
        auto convert = [&](auto scaled) {
#pragma unroll
        for (int k = 0; k < 64; k += 4) {
          float x0 = ldx(k + 0), x1 = ldx(k + 1);
          float x2 = ldx(k + 2), x3 = ldx(k + 3);
          if constexpr (decltype(scaled)::value) { x0 *= row_scale; x1 *= row_scale; x2 *= row_scale; x3 *= row_scale; }
          if (k < 32) {
            ss0 = fmaf(x0, x0, ss0); ss1 = fmaf(x1, x1, ss1);
            ss2 = fmaf(x2, x2, ss2); ss3 = fmaf(x3, x3, ss3);
          } else {
            st0 = fmaf(x0, x0, st0); st1 = fmaf(x1, x1, st1);
            st2 = fmaf(x2, x2, st2); st3 = fmaf(x3, x3, st3);
          }
          packed[(k >> 1) + 0] = F16OP ? pack_f16x2_sat(x0, x1) : pack_bf16x2(x0, x1);
          packed[(k >> 1) + 1] = F16OP ? pack_f16x2_sat(x2, x3) : pack_bf16x2(x2, x3);
          if constexpr (SPLIT3) {
            const uint32_t h01 = packed[(k >> 1) + 0], h23 = packed[(k >> 1) + 1];
            packed_lo[(k >> 1) + 0] = pack_bf16x2(x0 - __uint_as_float(h01 << 16),
                                                  x1 - __uint_as_float(h01 & 0xffff0000u));
            packed_lo[(k >> 1) + 1] = pack_bf16x2(x2 - __uint_as_float(h23 << 16),
                                                  x3 - __uint_as_float(h23 & 0xffff0000u));
          }
        }
        };
        if (F16OP && !unit_rows) convert(std::true_type{});
        else convert(std::false_type{});
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(as_empty0 + 8u * s);   // staging slot may be refilled
        if constexpr (SPLIT3) {
          publish(3 * kb, packed);
          publish(3 * kb + 1, packed);
          publish(3 * kb + 2, packed_lo);
        } else if constexpr (!AHEAD2) {
          publish(kb, packed);
        } else if (lt == 0) {
          publish_at(lt, kb, packed, true);            // the CTA's first tile: its slots have never been used
        }
      }
      if (AHEAD2 && lt > 0) {                          // the tile's last blocks
        bool freed[AH];                                // all closing slots polled first: the latencies overlap
        freed[0] = poll_freed;
#pragma unroll
        for (int i = 1; i < AH; ++i)
          freed[i] = ptx::mbar_test_wait(a_free0 + 8u * a_slot(lt, KB_IN - AH + i), a_phase(lt, KB_IN - AH + i) ^ 1u);
#pragma unroll
        for (int i = 0; i < AH; ++i) publish_at(lt, KB_IN - AH + i, held[(KB_IN - AH + i) % AH], freed[i]);
      }
      const float ssq = ((ss0 + ss1) + (ss2 + ss3)) + ((st0 + st1) + (st2 + st3));
      float inv = 1.0f / fmaxf(sqrtf(ssq), 1e-12f);
      const int slot = lt % 3;
      if constexpr (F16OP) {
        // ||x|| = ||s x|| / s; the accumulators hold <s x, 16 t>: the row factor is 1 / (16 s max(||x||, eps))
        inv = 1.0f / fmaxf(sqrtf(ssq) / row_scale, 1e-12f);
        norm_s[slot * F_BLOCK_M + arow] = inv / (16.0f * row_scale);
      } else {
        norm_s[slot * F_BLOCK_M + arow] = inv;
      }
      if (!PROJ && p.inv_norm != nullptr && arow < tc.rows) p.inv_norm[tc.out_row0 + arow] = inv;
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(n_ready0 + 8u * slot);
    }
  } else if (warp >= 8) {
    // ================================ epilogue ===============================================
    const int lg = warp & 3;
    const int eg = EPI2 ? (warp - 8) >> 2 : 0;                 // epilogue group = accumulator stage it drains
    // staging: 4608 bytes per warp; with two groups 2048 bytes each (one TMA-store buffer).  Both are
    // multiples of 512: the 64-byte swizzle is a function of the absolute shared-memory address
    float* stage = epi_stage + (EPI2 ? (warp - 8) * 512 : lg * (FSmem::epi_warp_bytes / 4));
    const uint32_t stage_u32 = ptx::smem_u32(stage);
    uint32_t tma_chunk = 0;
    constexpr bool vp = MODE == 2;
    void* const logits_ptr = MODE == 1 ? p.logits : nullptr;
    const bool logits_tma = MODE == 1 && (EPI2 || p.logits_tma);      // EPI2 + logits: the host guarantees the TMA path
    const bool want_max = p.row_max != nullptr || vp;
    // a finished row: local (score, class) or, vocabulary-parallel, one max-reduction per rank
    const long long vp_off = vp ? (long long)(__ldcg(p.vp_step) & 1ull) * p.vp_rows : 0;
    auto emit_row = [&](long long row, float v, int idx) {
      if (vp) {
        const unsigned long long key = vp_pack_key(v, p.vp_class_offset + idx);
        for (int g = 0; g < p.vp_world; ++g) atomicMax_system(p.vp_keys[g] + vp_off + row, key);
      } else {
        p.row_max[row] = v;
        if (p.row_arg != nullptr) p.row_arg[row] = idx;
      }
    };
    uint32_t acc_it = 0, lt = 0;
    for (int w = pair0; w < total_work; w += pair_stride, ++lt) {
      const int tile = (NSPLIT == 1 ? w : w / NSPLIT) * CG + (int)rank;
      const int nt_b = NSPLIT == 1 ? 0 : (w % NSPLIT) * NT / NSPLIT, nt_e = NSPLIT == 1 ? NT : (w % NSPLIT + 1) * NT / NSPLIT;
      (void)nt_b; (void)nt_e;
      const TileCoord tc = decode_tile(p, tile);
      const int r_in_tile = lg * 32 + lane;
      const bool row_ok = r_in_tile < tc.rows;
      const long long grow = tc.out_row0 + r_in_tile;
      const int slot = lt % 3;
      if constexpr (!PROJ) ptx::mbar_wait(n_ready0 + 8u * slot, (lt / 3) & 1u);
      float scale;
      if constexpr (CONV8) {
        // the row's sum of squares = the two converter halves' partial sums; F16OP: over the scaled values
        const float ssq = norm_s[(slot * 3 + 0) * F_BLOCK_M + r_in_tile] + norm_s[(slot * 3 + 1) * F_BLOCK_M + r_in_tile];
        float inv, factor;
        if constexpr (F16OP) {
          const float rs = norm_s[(slot * 3 + 2) * F_BLOCK_M + r_in_tile];
          inv = 1.0f / fmaxf(sqrtf(ssq) / rs, 1e-12f);      // ||x|| = ||s x|| / s
          factor = inv / (16.0f * rs);                        // the accumulators hold <s x, 16 t>
        } else {
          inv = 1.0f / fmaxf(sqrtf(ssq), 1e-12f);
          factor = inv;
        }
        if (p.inv_norm != nullptr && row_ok) p.inv_norm[grow] = inv;
        scale = p.normalize ? p.alpha * factor : p.alpha;
      } else {
        scale = PROJ ? 1.0f : (p.normalize ? p.alpha * norm_s[slot * F_BLOCK_M + r_in_tile] : p.alpha);
      }
      float q = 0.f;                                   // projected: ||W x + b||^2
      const float beta = p.beta;
      float bv[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
      int bi[4] = {0, 0, 0, 0};
      // No logits to write and a row scale >= 0: max_c fma(scale, acc_c, beta) ==
      // fma(scale, max_c acc_c, beta) exactly (rounding is monotone), so the running maximum is
      // taken over the RAW accumulators and the affine map is applied once per row.  Scores only
      // (the attention row): one FMNMX3 per two values.  With argmax: compare + select + index
      // per value, no FFMA; the class returned is the argmax of the fp32 accumulators (lowest
      // index among equal accumulators).
      // logits rows start 16-byte aligned when the leading dimension is padded to 16 bytes
      const bool vec_logits = logits_ptr != nullptr && ((uintptr_t)logits_ptr & 15) == 0 &&
                              (p.ldc * (p.logits_bf16 ? 2 : 4)) % 16 == 0;
      const bool raw_mode = PROJ || (want_max && logits_ptr == nullptr && p.alpha >= 0.f && !(p.dbg & 2));
      const bool max_only = raw_mode && p.row_arg == nullptr && !vp;
      float raw_best = -INFINITY;
      // projected + EPI2: every epilogue warp releases the A region once per anchor tile - after the
      // last G' tile it drains, or (none of them its own) once its first tile of this anchor tile
      // has arrived, which proves that the MMA warp has moved on to this anchor tile
      bool a_released = false;
      for (int nt = nt_b; nt < nt_e; ++nt, ++acc_it) {
        if (EPI2 && (int)(acc_it & 1u) != eg) continue;
        const int n0 = (nt - NG) * BN;                 // first class of a class tile
        const int n_valid = ntile_valid(nt);
        const int nchunks = (n_valid + 31) >> 5;
        const int as = acc_it & 1;
        if ((warp & 3) == 0) OVDET_TR(2 + eg, 20);
        ptx::mbar_wait(t_full0 + 8u * as, (acc_it >> 1) & 1u);
        if ((warp & 3) == 0) OVDET_TR(2 + eg, 21);
        ptx::tc_fence_after();
        const uint32_t t_row = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(F_ACC_COL + as * BN);
        if (PROJ && nt < NG) {
          // G' tile: q += sum_j acc_j * x'_j, x' (bf16 pairs) read back from the A region
          const uint32_t a_row = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)A_COL;
          uint32_t acc[32], xa[16];
          for (int c = 0; c < nchunks; ++c) {
            const int j0 = nt * BN + (c << 5);                         // k index of the chunk's first column
            ptx::tmem_ld_32x32(t_row + (uint32_t)(c << 5), acc);
            ptx::tmem_ld_32x32_x16(a_row + (uint32_t)(j0 >> 1), xa);
            ptx::tmem_ld_wait();
            const int valid = n_valid - (c << 5);                      // 32, or 16 in the last tile
            float q0 = 0.f, q1 = 0.f;
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (2 * i < valid) {
                q0 = fmaf(__uint_as_float(acc[2 * i]), __uint_as_float(xa[i] << 16), q0);
                q1 = fmaf(__uint_as_float(acc[2 * i + 1]), __uint_as_float(xa[i] & 0xffff0000u), q1);
              }
            q += q0 + q1;
          }
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (CG == 2) ptx::mbar_arrive_cluster(lead_t_empty0 + 8u * as);
            else ptx::mbar_arrive(t_empty0 + 8u * as);
            if (EPI2 ? (nt + 2 >= NG) : (nt == NG - 1)) {              // done reading x': the A region may be rewritten
              for (int kb = 0; kb < KB_IN; ++kb) ptx::mbar_arrive(a_free0 + 8u * kb);
            }
          }
          if (EPI2 && nt + 2 >= NG) a_released = true;
          continue;
        }
        if (PROJ && EPI2 && !a_released) {                              // this group drained no G' tile
          a_released = true;
          if (lane == 0)
            for (int kb = 0; kb < KB_IN; ++kb) ptx::mbar_arrive(a_free0 + 8u * kb);
        }

        auto consume = [&](uint32_t (&r)[32], int c) {
          const int c0 = c << 5;
          const int valid = n_valid - c0;
          if (max_only) {
            float m0 = raw_best, m1 = -INFINITY;
            if (valid >= 32) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                m0 = fmaxf(m0, fmaxf(__uint_as_float(r[j]), __uint_as_float(r[j + 1])));
                m1 = fmaxf(m1, fmaxf(__uint_as_float(r[j + 2]), __uint_as_float(r[j + 3])));
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (j < valid) m0 = fmaxf(m0, __uint_as_float(r[j]));
            }
            raw_best = fmaxf(m0, m1);
            return;
          }
          if (!raw_mode) {
#pragma unroll
            for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(fmaf(scale, __uint_as_float(r[j]), beta));
          }
          if (want_max) {
            const int col = n0 + c0;
            if (valid >= 32) {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float v = __uint_as_float(r[j]);
                const float old = bv[j & 3];
                bv[j & 3] = fmaxf(old, v);                 // value chain: one FMNMX per element ...
                if (v > old) bi[j & 3] = col + j;           // ... the predicate is off that chain
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float v = __uint_as_float(r[j]);
                const float old = bv[j & 3];
                if (j < valid) {
                  bv[j & 3] = fmaxf(old, v);
                  if (v > old) bi[j & 3] = col + j;
                }
              }
            }
          }
          if (logits_tma) {
            // bf16 logits, 16-byte aligned rows: the warp's 32 x 32 block goes to shared memory (64-byte
            // rows in the 64-byte swizzle pattern: the 16-byte writes by row are bank-conflict free)
            // and leaves as ONE bulk tensor store; rows past the level's last anchor and columns past
            // the last class are clipped by the TMA unit.  Two buffers per warp: the store of chunk
            // c - 1 may still be reading its buffer while chunk c is packed (a third buffer measured no
            // faster: the epilogue's arithmetic, not the store queue, sets the pace).  The per-thread
            // 16-byte global stores this replaces (32 rows per instruction) were LSU-bound.
            const uint32_t buf = EPI2 ? stage_u32 : stage_u32 + (tma_chunk & 1u) * 2048u;
            ++tma_chunk;
            if (lane == 0) {
              if constexpr (EPI2) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
              else asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            }
            __syncwarp();
            const uint32_t rowaddr = buf + lane * 64u;
            const uint32_t sw = (lane >> 1) & 3u;
#pragma unroll
            for (int v = 0; v < 4; ++v)
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::
                  "r"(rowaddr + (((uint32_t)v ^ sw) << 4)),
                  "r"(pack_bf16x2(__uint_as_float(r[8 * v + 0]), __uint_as_float(r[8 * v + 1]))),
                  "r"(pack_bf16x2(__uint_as_float(r[8 * v + 2]), __uint_as_float(r[8 * v + 3]))),
                  "r"(pack_bf16x2(__uint_as_float(r[8 * v + 4]), __uint_as_float(r[8 * v + 5]))),
                  "r"(pack_bf16x2(__uint_as_float(r[8 * v + 6]), __uint_as_float(r[8 * v + 7]))) : "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0 && tc.rows > lg * 32) {
              asm volatile(
                  "cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];"
                  :: "l"(&cmaps.m[tc.level]), "r"(n0 + c0), "r"(tc.m0 + lg * 32), "r"(tc.b), "r"(buf) : "memory");
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
          } else if (!EPI2 && logits_ptr != nullptr && vec_logits && p.logits_bf16) {
            // 16-byte aligned rows (padded leading dimension), bf16: every thread writes its own
            // row's 32 classes (64 bytes) straight from registers with four 16-byte stores; measured
            // faster than turning the block through shared memory (3.15 vs 3.48 ms at batch 256).
            if (row_ok) {
              const int col0 = n0 + c0;
              uint4* dst = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(logits_ptr) + grow * p.ldc + col0);
#pragma unroll
              for (int v = 0; v < 4; ++v)
                if (col0 + 8 * v < (int)p.ldc)
                  dst[v] = make_uint4(pack_bf16x2(__uint_as_float(r[8 * v + 0]), __uint_as_float(r[8 * v + 1])),
                                      pack_bf16x2(__uint_as_float(r[8 * v + 2]), __uint_as_float(r[8 * v + 3])),
                                      pack_bf16x2(__uint_as_float(r[8 * v + 4]), __uint_as_float(r[8 * v + 5])),
                                      pack_bf16x2(__uint_as_float(r[8 * v + 6]), __uint_as_float(r[8 * v + 7])));
            }
          } else if (!EPI2 && logits_ptr != nullptr && vec_logits) {
            // fp32: the warp's 32 x 32 block is turned through shared memory (144-byte pitch: the
            // 128-bit writes by row and the 128-bit reads by quarter-row are both bank-conflict free)
            // so that one store instruction writes 4 rows x 128 contiguous bytes instead of 32 rows
            // x 4 bytes (2.72 -> 1.96 ms at batch 128).  Columns in [classes, ldc) are padding.
            float4* srow = reinterpret_cast<float4*>(stage + lane * F_VPITCH);
#pragma unroll
            for (int q4 = 0; q4 < 8; ++q4)
              srow[q4] = make_float4(__uint_as_float(r[4 * q4]), __uint_as_float(r[4 * q4 + 1]),
                                     __uint_as_float(r[4 * q4 + 2]), __uint_as_float(r[4 * q4 + 3]));
            __syncwarp();
            const int col0 = n0 + c0;
            const int rows_here = min(32, tc.rows - lg * 32);
            const long long out_row0 = tc.out_row0 + lg * 32;
            {
              float* out = static_cast<float*>(logits_ptr);
              const int c4 = (lane & 7) * 4;                             // 8 lanes x 4 classes per row
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int row = 4 * i + (lane >> 3);
                const float4 a = *reinterpret_cast<const float4*>(stage + row * F_VPITCH + c4);
                if (row < rows_here && col0 + c4 < (int)p.ldc)
                  *reinterpret_cast<float4*>(out + (out_row0 + row) * p.ldc + col0 + c4) = a;
              }
            }
            __syncwarp();
          } else if (!EPI2 && logits_ptr != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; ++j) stage[lane * F_PITCH + j] = __uint_as_float(r[j]);
            __syncwarp();
            const int col = n0 + c0 + lane;
            const bool col_ok = lane < valid;
            const int rows_here = min(32, tc.rows - lg * 32);
            const long long out_row0 = tc.out_row0 + lg * 32;
            // unaligned rows (odd class count, exact reference strides): one row per instruction, 32
            // lanes = 32 consecutive classes, 2- or 4-byte stores.  Slow (5.4 ms at batch 256, bf16):
            // callers that can pad the leading dimension to 16 bytes get the paths above.
            if (p.logits_bf16) {
              __nv_bfloat16* out = static_cast<__nv_bfloat16*>(logits_ptr);
              for (int i = 0; i < rows_here; ++i)
                if (col_ok) out[(out_row0 + i) * p.ldc + col] = __float2bfloat16_rn(stage[i * F_PITCH + lane]);
            } else {
              float* out = static_cast<float*>(logits_ptr);
              for (int i = 0; i < rows_here; ++i)
                if (col_ok) out[(out_row0 + i) * p.ldc + col] = stage[i * F_PITCH + lane];
            }
            __syncwarp();
          }
        };

        uint32_t ra[32], rb[32];
        if constexpr (MODE == 1 && EPI2 && !PROJ) {
          if (n_valid == BN) {
            // Hot path of the materialised-logits kernel (full tile, bf16 logits through TMA stores, two
            // epilogue groups): straight-line over the four 32-column chunks, and - as in the scores-only
            // hot path below - the accumulator goes back to the MMA warp as soon as its last 32 columns
            // are in registers; the affine map, the running max / argmax, the packing and the store of
            // that chunk run after the arrive.  The ncu source page of the generic loop showed ~215-260
            // issued instructions per chunk against ~150 of arithmetic, on SM sub-partitions whose only
            // eligible warps are two epilogue warps: this mode is bound by epilogue issue slots.
            auto emit_chunk = [&](uint32_t (&r)[32], int c) {
              const int col = n0 + (c << 5);
#pragma unroll
              for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(fmaf(scale, __uint_as_float(r[j]), beta));
              if (want_max) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                  const float v = __uint_as_float(r[j]);
                  const float old = bv[j & 3];
                  bv[j & 3] = fmaxf(old, v);
                  if (v > old) bi[j & 3] = col + j;
                }
              }
              if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
              __syncwarp();
              const uint32_t rowaddr = stage_u32 + lane * 64u;
              const uint32_t sw = (lane >> 1) & 3u;
#pragma unroll
              for (int v = 0; v < 4; ++v)
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::
                    "r"(rowaddr + (((uint32_t)v ^ sw) << 4)),
                    "r"(pack_bf16x2(__uint_as_float(r[8 * v + 0]), __uint_as_float(r[8 * v + 1]))),
                    "r"(pack_bf16x2(__uint_as_float(r[8 * v + 2]), __uint_as_float(r[8 * v + 3]))),
                    "r"(pack_bf16x2(__uint_as_float(r[8 * v + 4]), __uint_as_float(r[8 * v + 5]))),
                    "r"(pack_bf16x2(__uint_as_float(r[8 * v + 6]), __uint_as_float(r[8 * v + 7]))) : "memory");
              asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
              __syncwarp();
              if (lane == 0 && tc.rows > lg * 32) {
                asm volatile(
                    "cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];"
                    :: "l"(&cmaps.m[tc.level]), "r"(col), "r"(tc.m0 + lg * 32), "r"(tc.b), "r"(stage_u32) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
              }
            };
            ptx::tmem_ld_32x32(t_row, ra);
            ptx::tmem_ld_wait();
            ptx::tmem_ld_32x32(t_row + 32u, rb);
            emit_chunk(ra, 0);
            ptx::tmem_ld_wait();
            if constexpr (BN == 128) {
              ptx::tmem_ld_32x32(t_row + 64u, ra);
              emit_chunk(rb, 1);
              ptx::tmem_ld_wait();
              ptx::tmem_ld_32x32(t_row + 96u, rb);
              emit_chunk(ra, 2);
              ptx::tmem_ld_wait();
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if constexpr (CG == 2) ptx::mbar_arrive_cluster(lead_t_empty0 + 8u * as);
              else ptx::mbar_arrive(t_empty0 + 8u * as);
            }
            emit_chunk(rb, BN / 32 - 1);
            continue;
          }
        }
        if (raw_mode && n_valid == BN) {
          // Hot path (full tile, no logits): straight-line, and the accumulator is handed back to
          // the MMA warp as soon as its last 32 columns are in registers - the compare work on
          // that chunk runs after the arrive, off the MMA -> epilogue -> MMA critical chain.
          auto scan = [&](const uint32_t (&r)[32], int col) {
            if (p.dbg & 64) {                       // experiment: drain only (results are wrong)
              raw_best = fmaxf(raw_best, __uint_as_float(r[0]));
              return;
            }
            if (max_only) {
              float m0 = raw_best, m1 = -INFINITY;
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                m0 = fmaxf(m0, fmaxf(__uint_as_float(r[j]), __uint_as_float(r[j + 1])));
                m1 = fmaxf(m1, fmaxf(__uint_as_float(r[j + 2]), __uint_as_float(r[j + 3])));
              }
              raw_best = fmaxf(m0, m1);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                // running max as FMNMX (4-cycle chain); the compare that decides the index reads the
                // previous value and is not on the chain (an FSETP -> FSEL pair per element was: the
                // epilogue warps sat in fixed-latency waits for 39 % of their samples)
                const float v = __uint_as_float(r[j]);
                const float old = bv[j & 3];
                bv[j & 3] = fmaxf(old, v);
                if (v > old) bi[j & 3] = col + j;
              }
            }
          };
          ptx::tmem_ld_32x32(t_row, ra);
          ptx::tmem_ld_wait();
          ptx::tmem_ld_32x32(t_row + 32u, rb);
          scan(ra, n0);
          ptx::tmem_ld_wait();
          if constexpr (BN == 128) {
            ptx::tmem_ld_32x32(t_row + 64u, ra);
            scan(rb, n0 + 32);
            ptx::tmem_ld_wait();
            ptx::tmem_ld_32x32(t_row + 96u, rb);
            scan(ra, n0 + 64);
            ptx::tmem_ld_wait();
          }
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (CG == 2) ptx::mbar_arrive_cluster(lead_t_empty0 + 8u * as);
            else ptx::mbar_arrive(t_empty0 + 8u * as);
          }
          if ((warp & 3) == 0) OVDET_TR(2 + eg, 22);
          scan(rb, n0 + BN - 32);
          if ((warp & 3) == 0) OVDET_TR(2 + eg, 23);
          continue;
        }
        ptx::tmem_ld_32x32(t_row, ra);
        for (int c = 0; c < nchunks; c += 2) {
          ptx::tmem_ld_wait();
          if (c + 1 < nchunks) ptx::tmem_ld_32x32(t_row + (uint32_t)((c + 1) << 5), rb);
          consume(ra, c);
          if (c + 1 < nchunks) {
            ptx::tmem_ld_wait();
            if (c + 2 < nchunks) ptx::tmem_ld_32x32(t_row + (uint32_t)((c + 2) << 5), ra);
            consume(rb, c + 1);
          }
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (CG == 2) ptx::mbar_arrive_cluster(lead_t_empty0 + 8u * as);
          else ptx::mbar_arrive(t_empty0 + 8u * as);
        }
      }
      if constexpr (EPI2) {
        // the two groups' partial results of this row meet in shared memory (one buffer per anchor-
        // tile parity, one named barrier per anchor tile: group 1's write for tile lt + 2 follows the
        // barrier of lt + 1, which group 0 reaches only after its read for lt); group 0 emits
        float best = bv[0];
        int best_idx = bi[0];
#pragma unroll
        for (int qd = 1; qd < 4; ++qd)
          if (bv[qd] > best || (bv[qd] == best && bi[qd] < best_idx)) { best = bv[qd]; best_idx = bi[qd]; }
        float* xbuf = reinterpret_cast<float*>(base_ptr + FSmem::xbuf_off) + (lt & 1u) * (3 * F_BLOCK_M);
        if (eg == 1) {
          xbuf[r_in_tile] = max_only ? raw_best : best;
          reinterpret_cast<int*>(xbuf)[F_BLOCK_M + r_in_tile] = best_idx;
          xbuf[2 * F_BLOCK_M + r_in_tile] = q;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (eg == 1) continue;
        const float ob = xbuf[r_in_tile];
        const int oi = reinterpret_cast<const int*>(xbuf)[F_BLOCK_M + r_in_tile];
        q += xbuf[2 * F_BLOCK_M + r_in_tile];
        if (max_only) {
          raw_best = fmaxf(raw_best, ob);
        } else if (ob > best || (ob == best && oi < best_idx)) {
          best = ob;
          best_idx = oi;
        }
        bv[0] = best; bi[0] = best_idx;
        bv[1] = bv[2] = bv[3] = -INFINITY;
      }
      if (!PROJ && NSPLIT > 1) {
        // partial result of this class range -> scratch; the last part to arrive merges
        float best = bv[0];
        int best_idx = bi[0];
#pragma unroll
        for (int qd = 1; qd < 4; ++qd)
          if (bv[qd] > best || (bv[qd] == best && bi[qd] < best_idx)) { best = bv[qd]; best_idx = bi[qd]; }
        if (max_only) best = raw_best;
        if (raw_mode) best = fmaf(scale, best, beta);
        const int part = w % NSPLIT;
        const long long rows_total = (long long)p.batch * p.anchors;
        if (row_ok) {
          p.part_max[part * rows_total + grow] = best;
          p.part_arg[part * rows_total + grow] = best_idx;
        }
        __threadfence();
        __syncwarp();
        int last = 0;
        if (lane == 0) last = atomicAdd(p.part_count + tile * 4 + lg, 1) == NSPLIT - 1;
        last = __shfl_sync(0xffffffffu, last, 0);
        if (last) {
          __threadfence();
          if (row_ok) {
            float m = -INFINITY;
            int mi = 0;
            for (int q2 = 0; q2 < NSPLIT; ++q2) {             // class ranges ascend with the part index
              const float v = __ldcg(p.part_max + q2 * rows_total + grow);
              const int vi = __ldcg(p.part_arg + q2 * rows_total + grow);
              if (v > m) { m = v; mi = vi; }
            }
            emit_row(grow, m, mi);
          }
          if (lane == 0) p.part_count[tile * 4 + lg] = 0;      // ready for the next launch
        }
      } else if (PROJ) {
        float best = bv[0];
        int best_idx = bi[0];
#pragma unroll
        for (int qd = 1; qd < 4; ++qd)
          if (bv[qd] > best || (bv[qd] == best && bi[qd] < best_idx)) { best = bv[qd]; best_idx = bi[qd]; }
        if (max_only) best = raw_best;
        const float inv = 1.0f / fmaxf(sqrtf(fmaxf(q, 0.f)), 1e-12f);  // F.normalize's eps
        if (row_ok) {
          p.row_max[grow] = fmaf(p.alpha * inv, best, beta);
          if (p.row_arg != nullptr) p.row_arg[grow] = best_idx;
          if (p.inv_norm != nullptr) p.inv_norm[grow] = inv;
        }
      } else if (max_only) {
        if (row_ok) p.row_max[grow] = fmaf(scale, raw_best, beta);
      } else if (want_max && row_ok) {
        float best = bv[0];
        int best_idx = bi[0];
#pragma unroll
        for (int q = 1; q < 4; ++q)
          if (bv[q] > best || (bv[q] == best && bi[q] < best_idx)) { best = bv[q]; best_idx = bi[q]; }
        if (raw_mode) best = fmaf(scale, best, beta);
        emit_row(grow, best, best_idx);
      }
    }
    // outstanding logit stores read this warp's staging buffers
    if (logits_tma && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  ptx::tc_fence_before();
  if constexpr (CG == 2) ptx::cluster_sync_all();      // the peer may still read this CTA's smem / TMEM
  else __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_cg<CG>(tmem_base, F_TMEM_COLS);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn fused_encode_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      return nullptr;
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}

// Tensor maps are a pure function of the encode arguments, so they are cached by those arguments
// ("an internal cache of CUtensorMaps keyed by (ptr, dims, strides)"): a serving
// loop that reuses its buffers pays cuTensorMapEncodeTiled once per buffer, not four to eight
// times per step.  64 entries, round-robin replacement, one mutex; a stale entry can never be
// returned for different arguments because the whole argument list is the key.
struct MapKey {
  const void* addr;
  cuuint64_t dims[3], strides[2];
  cuuint32_t box[3];
  int dtype, swizzle, l2;
  bool operator==(const MapKey& o) const { return memcmp(this, &o, sizeof(MapKey)) == 0; }
};
struct MapCache {
  static constexpr int kEntries = 64;
  std::mutex mu;
  MapKey keys[kEntries];
  CUtensorMap maps[kEntries];
  int used = 0, next = 0;
};

CUresult encode3(EncodeTiledFn enc, CUtensorMap* out, CUtensorMapDataType dtype, const void* addr,
                 const cuuint64_t (&dims)[3], const cuuint64_t (&strides)[2], const cuuint32_t (&box)[3],
                 CUtensorMapSwizzle swizzle, CUtensorMapL2promotion l2) {
  static MapCache cache;
  MapKey k;
  memset(&k, 0, sizeof(k));                          // padding bytes take part in the comparison
  k.addr = addr;
  for (int i = 0; i < 3; ++i) { k.dims[i] = dims[i]; k.box[i] = box[i]; }
  k.strides[0] = strides[0]; k.strides[1] = strides[1];
  k.dtype = (int)dtype; k.swizzle = (int)swizzle; k.l2 = (int)l2;
  {
    std::lock_guard<std::mutex> lock(cache.mu);
    for (int i = 0; i < cache.used; ++i)
      if (cache.keys[i] == k) { *out = cache.maps[i]; return CUDA_SUCCESS; }
  }
  const cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, dtype, 3, const_cast<void*>(addr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle, l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return r;
  std::lock_guard<std::mutex> lock(cache.mu);
  const int slot = cache.used < MapCache::kEntries ? cache.used++ : (cache.next++ % MapCache::kEntries);
  cache.keys[slot] = k;
  cache.maps[slot] = *out;
  return CUDA_SUCCESS;
}

}  // namespace

// Shared launcher.  `dim` is the real length of an activation vector; it is padded to a multiple
// of 64 by the TMA zero fill on the activation side and by zero columns in the text operand.
// Cosine / raw modes: one operand `text_op` with rows of kop = (split3 ? 3 : 1) * ceil(dim / 64) * 64.
// Projected mode (`level_ops` != NULL): one operand per level, [Cpad + kop, kop] with
// kop = ceil(dim / 64) * 64 + 16 (ops.py project_vocabulary), scores / argmax only.
int fused_launch(const float* const* obj_embeds, const int64_t* hw, const int64_t* stride_b,
                 const int64_t* stride_d, int num_levels, int64_t batch, int64_t dim,
                 const void* text_op, const void* const* level_ops, int64_t classes, int text_batched,
                 int normalize, int split3, float alpha, float beta, void* logits, int logits_dtype,
                 int64_t ldc, float* row_max, int32_t* row_arg, float* inv_norm, void* stream,
                 int in_bf16, void* split_ws, size_t split_ws_bytes, const VpTarget* vp, int f16_operands) {
  const int proj = level_ops != nullptr;
  if (f16_operands && (proj || split3 || in_bf16 || vp || !normalize)) return OVDET_ERR_INVALID_ARG;
  if (vp) {                                         // vocabulary-parallel: keys instead of row_max / row_arg
    if (proj || split3 || logits || row_max || row_arg || alpha < 0.f) return OVDET_ERR_INVALID_ARG;
    if (vp->world < 1 || vp->world > OVDET_MAX_PEERS || vp->class_offset < 0 || !vp->step) return OVDET_ERR_INVALID_ARG;
    for (int g = 0; g < vp->world; ++g)
      if (!vp->keys[g] || ((uintptr_t)vp->keys[g] & 7)) return OVDET_ERR_INVALID_ARG;
  }
  if (in_bf16 && (proj || split3)) return OVDET_ERR_UNSUPPORTED_SHAPE;   // bf16 activations: cosine mode only
  if (batch == 0) return check_device();            // an empty batch is a no-op (its pointers may be null)
  if (!obj_embeds || !hw || !stride_b || !stride_d || (!text_op && !proj) || batch < 0 || classes <= 0 || dim <= 0)
    return OVDET_ERR_INVALID_ARG;
  if (num_levels <= 0) return OVDET_ERR_INVALID_ARG;
  if (!logits && !row_max && !vp) return OVDET_ERR_INVALID_ARG;
  if (row_arg && !row_max) return OVDET_ERR_INVALID_ARG;
  if (logits && (ldc < classes || (logits_dtype != OVDET_F32 && logits_dtype != OVDET_BF16)))
    return OVDET_ERR_INVALID_ARG;
  if (proj && (logits || split3 || alpha < 0.f || !row_max)) return OVDET_ERR_INVALID_ARG;
  const int kb_in = (int)ceil_div<int64_t>(dim, F_BLOCK_K);
  const int kb = proj ? kb_in + 1 : kb_in * (split3 ? 3 : 1);
  // more than 8 k blocks only in the streaming three-pass mode: one N tile (classes <= 128), dim <= 512
  const bool streaming = split3 && kb > F_MAX_KB && kb <= 3 * F_MAX_KB && classes <= F_BLOCK_N;
  if (num_levels > F_MAX_LEVELS || (kb > F_MAX_KB && !streaming) || batch > 65535)
    return OVDET_ERR_UNSUPPORTED_SHAPE;

  if (!proj && ((uintptr_t)text_op & 15)) return OVDET_ERR_INVALID_ARG;
  // CTA pairs (cta_group::2) for the dim = 512 similarity and the hidden = 256 projected one;
  // OVDET_FUSED_CG=1 forces single CTAs
  static const int cg_env = []() { const char* e = getenv("OVDET_FUSED_CG"); return e ? atoi(e) : 2; }();
  // ... and for the attention row's small resident shapes (hidden <= 128: 1-2 k blocks, 3 or 6 with the
  // three-pass recipe; scores only): there the text tile is re-read from L2 once per anchor tile for very
  // little MMA work, which is the L2 -> SM bound CTA pairs halve
  const bool small_pair = !proj && !logits && !vp && !in_bf16 && !f16_operands && row_max && !row_arg &&
                          ((!split3 && (kb == 1 || kb == 2)) || (split3 && (kb == 3 || kb == 6)));
  const int cg = (cg_env == 2 && ((!split3 && ((!proj && kb == 8) || (proj && kb_in == 4))) || small_pair)) ? 2 : 1;
  EncodeTiledFn enc = nullptr;
  FusedParams p{};
  LevelMaps maps, bmaps, cmaps;
  long long anchors = 0, tiles = 0;
  for (int l = 0; l < num_levels; ++l) {
    if (!obj_embeds[l] || hw[l] <= 0) return OVDET_ERR_INVALID_ARG;
    if (proj && (!level_ops[l] || ((uintptr_t)level_ops[l] & 15))) return OVDET_ERR_INVALID_ARG;
    // TMA needs 16-byte aligned base and strides
    const int64_t amask = in_bf16 ? 7 : 3;
    if (((uintptr_t)obj_embeds[l] & 15) || (stride_d[l] & amask) || (stride_b[l] & amask) || stride_d[l] < hw[l])
      return OVDET_ERR_UNSUPPORTED_SHAPE;
    p.hw[l] = (int)hw[l];
    p.mt[l] = (int)ceil_div<int64_t>(hw[l], F_BLOCK_M);
    // a pair multiplies against ONE text tile: it never straddles two images (per-image text) or
    // two levels (per-level operands)
    if (cg == 2 && (text_batched || proj)) p.mt[l] = (p.mt[l] + 1) & ~1;
    p.off[l] = (int)anchors;
    p.tile_start[l] = (int)tiles;
    anchors += hw[l];
    tiles += (long long)batch * p.mt[l];
  }
  if (anchors >= (1ll << 30) || tiles >= (1ll << 31) || classes >= (1 << 30)) return OVDET_ERR_UNSUPPORTED_SHAPE;
  if (int rc = check_device()) return rc;
  if (batch == 0) return OVDET_OK;
  enc = fused_encode_fn();
  if (!enc) return OVDET_ERR_DRIVER;
  for (int l = 0; l < num_levels; ++l) {
    cuuint64_t dims[3] = {(cuuint64_t)hw[l], (cuuint64_t)dim, (cuuint64_t)batch};
    const cuuint64_t esz = in_bf16 ? 2 : 4;
    cuuint64_t strides[2] = {(cuuint64_t)stride_d[l] * esz, (cuuint64_t)stride_b[l] * esz};
    cuuint32_t box[3] = {(cuuint32_t)F_BLOCK_M, (cuuint32_t)F_BLOCK_K, 1};
    if (batch == 1) strides[1] = (cuuint64_t)dim * stride_d[l] * esz;    // unused but must be valid
    CUresult r = encode3(enc, &maps.m[l], in_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                         obj_embeds[l], dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
    if (r != CUDA_SUCCESS) return OVDET_ERR_DRIVER;
  }
  for (int l = num_levels; l < F_MAX_LEVELS; ++l) maps.m[l] = maps.m[0];
  const int64_t kop = proj ? (int64_t)kb_in * F_BLOCK_K + 16 : (int64_t)kb * F_BLOCK_K;
  // classes per N tile of this launch (kernel: BN)
  const int bn = (cg == 2 && !proj && !split3 && OVDET_F_BN64) ? 64 : F_BLOCK_N;
  const int64_t cpad = ceil_div<int64_t>(classes, F_BLOCK_N) * F_BLOCK_N;
  const int64_t op_rows = proj ? cpad + kop : classes;
  for (int l = 0; l < (proj ? num_levels : 1); ++l) {
    const int64_t tb = text_batched ? batch : 1;
    cuuint64_t dims[3] = {(cuuint64_t)kop, (cuuint64_t)op_rows, (cuuint64_t)tb};
    cuuint64_t strides[2] = {(cuuint64_t)kop * 2, (cuuint64_t)op_rows * (cuuint64_t)kop * 2};
    cuuint32_t box[3] = {(cuuint32_t)F_BLOCK_K, (cuuint32_t)(bn / cg), 1};
    const void* op = proj ? level_ops[l] : text_op;
    CUresult r = encode3(enc, &bmaps.m[l], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, op, dims, strides, box,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
    if (r != CUDA_SUCCESS) return OVDET_ERR_DRIVER;
  }
  for (int l = (proj ? num_levels : 1); l < F_MAX_LEVELS; ++l) bmaps.m[l] = bmaps.m[0];
  // bf16 logits with 16-byte aligned rows leave through TMA stores: one [classes, hw, batch] map per
  // level over the level's rows of the [batch, anchors, ldc] array (OVDET_DBG=4: per-thread stores)
  static const int dbg_env0 = []() { const char* e = getenv("OVDET_DBG"); return e ? atoi(e) : 0; }();
  int logits_tma = 0;
  if (logits && logits_dtype == OVDET_BF16 && !((uintptr_t)logits & 15) && (ldc * 2) % 16 == 0 && !(dbg_env0 & 4)) {
    logits_tma = 1;
    long long off = 0;
    for (int l = 0; l < num_levels; ++l) {
      cuuint64_t dims[3] = {(cuuint64_t)classes, (cuuint64_t)hw[l], (cuuint64_t)batch};
      cuuint64_t strides[2] = {(cuuint64_t)ldc * 2, (cuuint64_t)anchors * (cuuint64_t)ldc * 2};
      cuuint32_t box[3] = {32, 32, 1};
      void* base = static_cast<char*>(logits) + (size_t)off * (size_t)ldc * 2;
      CUresult r = encode3(enc, &cmaps.m[l], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, base, dims, strides, box,
                           CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE);
      if (r != CUDA_SUCCESS) { logits_tma = 0; break; }
      off += hw[l];
    }
  }
  for (int l = logits_tma ? num_levels : 0; l < F_MAX_LEVELS; ++l) cmaps.m[l] = maps.m[0];
  p.logits_tma = logits_tma;
  p.levels = num_levels;
  p.batch = (int)batch;
  p.tile_start[num_levels] = (int)tiles;
  p.anchors = (int)anchors;
  p.classes = (int)classes;
  p.kb = kb;
  p.kb_in = kb_in;
  p.normalize = normalize ? 1 : 0;
  p.proj = proj;
  p.ng_tiles = proj ? (int)ceil_div<int64_t>(kop, F_BLOCK_N) : 0;
  p.cpad = (int)cpad;
  p.kop = (int)kop;
  p.n_tiles = (int)ceil_div<int64_t>(classes, bn) + p.ng_tiles;
  p.text_batched = text_batched ? 1 : 0;
  p.alpha = alpha;
  p.beta = beta;
  p.logits = logits;
  p.logits_bf16 = logits_dtype == OVDET_BF16;
  p.ldc = ldc;
  p.row_max = row_max;
  p.row_arg = row_arg;
  p.inv_norm = inv_norm;
  if (vp) {
    p.vp_world = vp->world;
    p.vp_class_offset = vp->class_offset;
    p.vp_rows = vp->rows;
    p.vp_step = vp->step;
    for (int g = 0; g < vp->world; ++g) p.vp_keys[g] = vp->keys[g];
  }
  // small launch: split the class tiles of every anchor tile over the idle CTA pairs
  p.nsplit = 1;
  if (cg == 2 && !proj && !split3 && !logits && (row_max || vp) && split_ws && !((uintptr_t)split_ws & 15)) {
    const long long pairs = (tiles + 1) / 2;
    const long long max_pairs = sm_count() / 2;
    long long ns = pairs > 0 ? max_pairs / pairs : 1;
    if (ns > 4) ns = 4;
    if (ns > p.n_tiles) ns = p.n_tiles;
    const long long rows = (long long)batch * anchors;
    const size_t need = (size_t)ns * rows * 8 + (size_t)(tiles + 2) * 16;
    if (ns >= 2 && split_ws_bytes >= need) {
      p.nsplit = (int)ns;
      p.part_count = static_cast<int*>(split_ws);                       // zero on entry, left zero
      p.part_max = reinterpret_cast<float*>(static_cast<char*>(split_ws) + (size_t)(tiles + 2) * 16);
      p.part_arg = reinterpret_cast<int*>(p.part_max + ns * rows);
    }
  }
  // experiment switches: 1 L2 prefetch warp on, 2 no raw-accumulator
  // epilogue, 16 / 32 back-off of the non-critical waits off / 256 ns
  static const int dbg_env = []() { const char* e = getenv("OVDET_DBG"); return e ? atoi(e) : 0; }();
  p.dbg = dbg_env;
#ifdef OVDET_TRACE
  static const unsigned long long trace_env = []() { const char* e = getenv("OVDET_TRACE_PTR"); return e ? strtoull(e, nullptr, 0) : 0ull; }();
  p.trace = reinterpret_cast<unsigned long long*>(trace_env);
#endif

  // every (shape variant, epilogue mode) instantiation the dispatch below can pick
  const int mode = vp ? 2 : (logits ? 1 : 0);
  if (f16_operands && cg != 2) return OVDET_ERR_UNSUPPORTED_SHAPE;   // fp16 tier: the dim = 512 CTA-pair kernel only
  if (vp && cg != 2) return OVDET_ERR_UNSUPPORTED_SHAPE;      // key exchange: the dim = 512 CTA-pair kernel only
#define OVDET_FOR_EACH_FUSED(X)                                                                        \
  X(8, false, 2, false, false, 0, false, false, false) X(8, false, 2, false, false, 1, false, false, false) X(8, false, 2, false, false, 2, false, false, false) \
  X(8, false, 2, false, true, 0, false, false, false)  X(8, false, 2, false, true, 1, false, false, false)  X(8, false, 2, false, true, 2, false, false, false)  \
  X(4, false, 2, true, false, 0, false, false, false)                                                                \
  X(8, false, 2, false, false, 0, true, false, false)  X(8, false, 2, false, true, 0, true, false, false)  X(4, false, 2, true, false, 0, true, false, false)     \
  X(8, false, 2, false, false, 1, true, false, false)  X(8, false, 2, false, true, 1, true, false, false)                           \
  X(8, false, 1, false, false, 0, false, false, false) X(8, false, 1, false, false, 1, false, false, false)                        \
  X(0, false, 1, false, false, 0, false, false, false) X(0, false, 1, false, false, 1, false, false, false)                        \
  X(0, true, 1, false, false, 0, false, false, false)  X(0, true, 1, false, false, 1, false, false, false)                         \
  X(0, false, 1, true, false, 0, false, false, false)                                                                \
  X(0, false, 1, false, true, 0, false, false, false)  X(0, false, 1, false, true, 1, false, false, false)  \
  X(8, false, 2, false, false, 0, false, true, false) X(8, false, 2, false, false, 1, false, true, false)            \
  X(8, false, 2, false, false, 1, true, true, false)                                                          \
  X(1, false, 2, false, false, 0, false, false, false) X(2, false, 2, false, false, 0, false, false, false)          \
  X(3, true, 2, false, false, 0, false, false, false)  X(6, true, 2, false, false, 0, false, false, false)  \
  X(8, false, 2, false, false, 0, false, false, true) X(8, false, 2, false, true, 0, false, false, true)   \
  X(8, false, 2, false, false, 0, false, true, true)  X(8, false, 2, false, false, 1, false, false, true)  \
  X(8, false, 2, false, false, 1, false, true, true)  X(8, false, 2, false, true, 1, false, false, true)
  if (int rc = once_per_device(1, []() -> int {
#define OVDET_SET_SMEM(KB, S3, CGV, PR, I16, MD, E2, F16, C8V)                                             \
        OVDET_CUDA_TRY(cudaFuncSetAttribute(sim_fused_kernel<KB, S3, CGV, PR, I16, MD, E2, F16, C8V>,   \
                                            cudaFuncAttributeMaxDynamicSharedMemorySize,                   \
                                            FSmem<CGV, (PR && CGV == 2), (CGV == 2 && !PR && !S3 && KB == 8 && OVDET_F_BN64)>::bytes));
        OVDET_FOR_EACH_FUSED(OVDET_SET_SMEM)
#undef OVDET_SET_SMEM
        return OVDET_OK;
      })) return rc;
  // second epilogue warpgroup (CTA-pair scores-only kernels): OVDET_EPI2 bit 0 = projected mode (the
  // epilogue is on its critical path: K = 272 per tile), bit 1 = cosine mode, bit 2 = bf16 logits (TMA stores)
  static const int epi2_env = []() { const char* e = getenv("OVDET_EPI2"); return e ? atoi(e) : 5; }();
  // (measured and dropped: a second group for the attention row's CTA-pair kernels - 0.140 vs 0.136 ms at P3)
  const bool epi2 = cg == 2 && ((mode == 0 && ((proj && (epi2_env & 1)) || (!proj && kb == 8 && (epi2_env & 2)))) ||
                                (mode == 1 && logits_tma && (epi2_env & 4))) &&
                    !(f16_operands && mode == 0);
  // eight converter warps (CONV8 in the kernel): few N tiles per anchor tile, no second epilogue group
  static const int conv8_env = []() { const char* e = getenv("OVDET_CONV8_TILES"); return e ? atoi(e) : 3; }();
  // ... or a launch so small that every CTA pair converts at most two anchor tiles (batch 1: the first
  // tile's conversion chain is fully exposed, there is nothing to hide it behind)
  const long long work_items = (tiles + 1) / 2 * p.nsplit;
  const bool conv8 = OVDET_F_CONV8 && cg == 2 && !proj && !split3 && kb == 8 && !epi2 && mode != 2 &&
                     ((p.n_tiles + p.nsplit - 1) / p.nsplit <= conv8_env ||
                      (conv8_env > 0 && work_items <= 2ll * (sm_count() / 2)));
  // shape variant of this launch
  const int v_kb = cg == 2 ? (proj ? 4 : p.kb) : ((!in_bf16 && !proj && !split3 && p.kb == 8) ? 8 : 0);
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attr[1];
  cfg.blockDim = dim3(F_THREADS);
  cfg.stream = as_stream(stream);
  if (cg == 2) {
    // one CTA per SM, launched as clusters of two (the pair shares a TPC)
    const long long pairs = (tiles + 1) / 2 * p.nsplit;                // work items
    const int max_pairs = sm_count() / 2;
    cfg.gridDim = dim3((unsigned)(2 * (pairs < max_pairs ? pairs : max_pairs)));
    cfg.dynamicSmemBytes = FSmem<2>::bytes;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
  } else {
    cfg.gridDim = dim3((unsigned)(tiles < sm_count() ? tiles : sm_count()));
    cfg.dynamicSmemBytes = FSmem<1>::bytes;
  }
  bool launched = false;
#define OVDET_TRY_LAUNCH(KB, S3, CGV, PR, I16, MD, E2, F16, C8V)                                           \
  if (!launched && v_kb == KB && (split3 != 0) == S3 && cg == CGV && (proj != 0) == PR &&               \
      (in_bf16 != 0) == I16 && mode == MD && epi2 == E2 && (f16_operands != 0) == F16 && conv8 == C8V) { \
    cfg.blockDim = dim3((E2 || C8V) ? F_THREADS + 128 : F_THREADS);                                     \
    cfg.dynamicSmemBytes = FSmem<CGV, (PR && CGV == 2), (CGV == 2 && !PR && !S3 && KB == 8 && OVDET_F_BN64)>::bytes; \
    OVDET_CUDA_TRY(cudaLaunchKernelEx(&cfg, sim_fused_kernel<KB, S3, CGV, PR, I16, MD, E2, F16, C8V>, maps, bmaps, cmaps, p)); \
    launched = true;                                                                                    \
  }
  OVDET_FOR_EACH_FUSED(OVDET_TRY_LAUNCH)
#undef OVDET_TRY_LAUNCH
#undef OVDET_FOR_EACH_FUSED
  if (!launched) return OVDET_ERR_UNSUPPORTED_SHAPE;
  OVDET_LAUNCH_CHECK();
  return OVDET_OK;
}

}  // namespace ovdet

extern "C" int ovdet_similarity_fused(const float* const* obj_embeds, const int64_t* hw,
                                      const int64_t* stride_b, const int64_t* stride_d,
                                      int num_levels, int64_t batch, int64_t dim,
                                      const void* text_op, int64_t classes, int text_batched,
                                      float alpha, float beta, void* logits, int logits_dtype,
                                      int64_t ldc, float* row_max, int32_t* row_arg,
                                      float* inv_norm, void* stream) {
  if (dim % 64 != 0) return OVDET_ERR_UNSUPPORTED_SHAPE;   // the text operand has exactly `dim` columns
  return ovdet::fused_launch(obj_embeds, hw, stride_b, stride_d, num_levels, batch, dim, text_op, nullptr,
                             classes, text_batched, /*normalize=*/1, /*split3=*/0, alpha, beta, logits,
                             logits_dtype, ldc, row_max, row_arg, inv_norm, stream, 0, nullptr, 0);
}

extern "C" size_t ovdet_similarity_split_workspace_bytes(int64_t batch, int64_t anchors) {
  // scratch of the small-launch class split: worth having when batch * anchors fits half the CTA pairs
  if (batch <= 0 || anchors <= 0) return 0;
  const long long rows = (long long)batch * anchors;
  const long long tiles = batch * ((anchors + 127) / 128 + 8);
  if (tiles / 2 > 74 / 2) return 0;
  return (size_t)4 * rows * 8 + (size_t)(tiles + 2) * 16 + 16;
}

extern "C" int ovdet_similarity_fused_ws(const float* const* obj_embeds, const int64_t* hw,
                                         const int64_t* stride_b, const int64_t* stride_d,
                                         int num_levels, int64_t batch, int64_t dim,
                                         const void* text_op, int64_t classes, int text_batched,
                                         float alpha, float beta, float* row_max, int32_t* row_arg,
                                         float* inv_norm, void* workspace, size_t workspace_bytes,
                                         int embed_dtype, void* stream) {
  if (dim % 64 != 0) return OVDET_ERR_UNSUPPORTED_SHAPE;
  return ovdet::fused_launch(obj_embeds, hw, stride_b, stride_d, num_levels, batch, dim, text_op, nullptr,
                             classes, text_batched, 1, 0, alpha, beta, nullptr, OVDET_F32, classes, row_max,
                             row_arg, inv_norm, stream, embed_dtype == OVDET_BF16, workspace, workspace_bytes);
}

extern "C" int ovdet_similarity_fused_bf16in(const void* const* obj_embeds, const int64_t* hw,
                                             const int64_t* stride_b, const int64_t* stride_d,
                                             int num_levels, int64_t batch, int64_t dim,
                                             const void* text_op, int64_t classes, int text_batched,
                                             float alpha, float beta, void* logits, int logits_dtype,
                                             int64_t ldc, float* row_max, int32_t* row_arg,
                                             float* inv_norm, void* stream) {
  if (dim % 64 != 0) return OVDET_ERR_UNSUPPORTED_SHAPE;
  return ovdet::fused_launch(reinterpret_cast<const float* const*>(obj_embeds), hw, stride_b, stride_d,
                             num_levels, batch, dim, text_op, nullptr, classes, text_batched, 1, 0, alpha,
                             beta, logits, logits_dtype, ldc, row_max, row_arg, inv_norm, stream, 1, nullptr, 0);
}

extern "C" int ovdet_similarity_fused_fp32(const float* const* obj_embeds, const int64_t* hw,
                                           const int64_t* stride_b, const int64_t* stride_d,
                                           int num_levels, int64_t batch, int64_t dim,
                                           const void* text_op3, int64_t classes, int text_batched,
                                           float alpha, float beta, void* logits, int logits_dtype,
                                           int64_t ldc, float* row_max, int32_t* row_arg,
                                           float* inv_norm, void* stream) {
  if (dim % 64 != 0) return OVDET_ERR_UNSUPPORTED_SHAPE;
  return ovdet::fused_launch(obj_embeds, hw, stride_b, stride_d, num_levels, batch, dim, text_op3, nullptr,
                             classes, text_batched, /*normalize=*/1, /*split3=*/1, alpha, beta, logits,
                             logits_dtype, ldc, row_max, row_arg, inv_norm, stream, 0, nullptr, 0);
}

extern "C" int ovdet_similarity_projected(const float* const* hidden, const int64_t* hw,
                                          const int64_t* stride_b, const int64_t* stride_d,
                                          int num_levels, int64_t batch, int64_t hidden_dim,
                                          const void* const* level_ops, int64_t classes, int text_batched,
                                          float alpha, float beta, float* row_max, int32_t* row_arg,
                                          float* inv_norm, void* stream) {
  if (!level_ops) return OVDET_ERR_INVALID_ARG;
  return ovdet::fused_launch(hidden, hw, stride_b, stride_d, num_levels, batch, hidden_dim, nullptr, level_ops,
                             classes, text_batched, /*normalize=*/1, /*split3=*/0, alpha, beta, nullptr,
                             OVDET_F32, classes, row_max, row_arg, inv_norm, stream, 0, nullptr, 0);
}

extern "C" int ovdet_similarity_fused_fp16(const float* const* obj_embeds, const int64_t* hw,
                                           const int64_t* stride_b, const int64_t* stride_d,
                                           int num_levels, int64_t batch, int64_t dim,
                                           const void* text_op16, int64_t classes, int text_batched,
                                           float alpha, float beta, void* logits, int logits_dtype,
                                           int64_t ldc, float* row_max, int32_t* row_arg,
                                           float* inv_norm, void* stream) {
  if (dim != 512) return OVDET_ERR_UNSUPPORTED_SHAPE;      // the CTA-pair kernel's shape
  return ovdet::fused_launch(obj_embeds, hw, stride_b, stride_d, num_levels, batch, dim, text_op16, nullptr,
                             classes, text_batched, /*normalize=*/1, /*split3=*/0, alpha, beta, logits,
                             logits_dtype, ldc, row_max, row_arg, inv_norm, stream, 0, nullptr, 0, nullptr, 1);
}
