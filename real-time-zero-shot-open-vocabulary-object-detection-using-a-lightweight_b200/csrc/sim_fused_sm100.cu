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
//   warp 3      L2 prefetch of the next anchor tile (cp.async.bulk.prefetch.tensor)
//   warps 4-7   converters   thread = anchor row: fp32 smem column -> sum of squares, bf16x2 ->
//               tcgen05.st into the A region of TMEM (256 columns = 128 rows x 512 k)
//   warps 8-11  epilogue     tcgen05.ld -> alpha/||x|| scale, +beta, running max/argmax and/or
//               smem-staged coalesced logit stores
//
// TMEM: columns [0,256) = two 128-column fp32 accumulators (epilogue of tile n overlaps the
// MMAs of tile n+1), columns [256,512) = the A operand.  A block kb of the NEXT anchor tile is
// converted as soon as the last N tile of the current one has consumed block kb (per-block
// mbarriers), so the conversion and the fp32 stream hide behind the MMAs.
#include "common.cuh"
#include "ptx.cuh"
#include <cstdlib>

namespace ovdet {
namespace {

constexpr int F_BLOCK_M = 128;
constexpr int F_BLOCK_N = 128;
constexpr int F_BLOCK_K = 64;
constexpr int F_MAX_KB = 8;                       // dim <= 512: A fills 256 TMEM columns
constexpr int F_B_STAGES = 4;                     // text ring; with dim = 512 stage == kb & 3 (static)
constexpr int F_A_STAGES = 4;                     // fp32 activation ring
constexpr int F_B_STAGE_BYTES = F_BLOCK_N * F_BLOCK_K * 2;     // 16 KiB
constexpr int F_A_STAGE_BYTES = F_BLOCK_K * F_BLOCK_M * 4;     // 32 KiB fp32 [k][anchor]
constexpr int F_THREADS = 384;
constexpr int F_TMEM_COLS = 512;
constexpr int F_ACC_COL = 0;
constexpr int F_A_COL = 256;
constexpr int F_PITCH = 33;
constexpr int F_MAX_LEVELS = 4;

struct FSmem {
  static constexpr int b_off = 0;
  static constexpr int a_off = b_off + F_B_STAGES * F_B_STAGE_BYTES;                 // 64 KiB
  static constexpr int epi_off = a_off + F_A_STAGES * F_A_STAGE_BYTES;               // +128 KiB
  static constexpr int epi_bytes = 4 * 32 * F_PITCH * 4;
  static constexpr int norm_off = epi_off + epi_bytes;
  static constexpr int norm_bytes = 3 * F_BLOCK_M * 4;
  static constexpr int bar_off = norm_off + norm_bytes;
  // b_full, b_empty, as_full, as_empty, a_ready, a_free, tmem_full, tmem_empty, norm_ready
  static constexpr int num_bars = 2 * F_B_STAGES + 2 * F_A_STAGES + 2 * F_MAX_KB + 2 + 2 + 3;
  static constexpr int tmem_ptr_off = bar_off + num_bars * 8;
  static constexpr int total = tmem_ptr_off + 16;
};
constexpr int F_SMEM_BYTES = FSmem::total + 1024;
static_assert(F_SMEM_BYTES <= 227 * 1024, "shared memory budget");

struct LevelMaps { CUtensorMap m[F_MAX_LEVELS]; };

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
  int n_tiles;
  int text_batched;
  float alpha, beta;
  void* logits;
  int logits_bf16;
  long long ldc;
  float* row_max;
  int* row_arg;
  float* inv_norm;
  int dbg;
};

struct TileCoord { int b, level, m0, rows; long long out_row0; };

__device__ __forceinline__ TileCoord decode_tile(const FusedParams& p, int tile) {
  int l = 0;
#pragma unroll
  for (int i = 1; i < F_MAX_LEVELS; ++i)
    if (i < p.levels && tile >= p.tile_start[i]) l = i;
  const int r = tile - p.tile_start[l];
  TileCoord t;
  t.level = l;
  t.b = r / p.mt[l];
  t.m0 = (r - t.b * p.mt[l]) * F_BLOCK_M;
  t.rows = min(F_BLOCK_M, p.hw[l] - t.m0);
  t.out_row0 = (long long)t.b * p.anchors + p.off[l] + t.m0;
  return t;
}

// KB_T  8: dim == 512, every k-block loop unrolled; 0: run-time block count.
// SPLIT3 (dim <= 128): fp32-accurate product from three bf16 passes.  x = hi + lo with
// hi = bf16(x), lo = bf16(x - hi); the activation blocks are written to tensor memory as
// [hi | hi | lo] and the text operand is laid out [hi | lo | hi] (ovdet_cast_text), so the plain
// block loop accumulates hi*hi + hi*lo + lo*hi (the lo*lo term is < 2^-16 relative).
template <int KB_T, bool SPLIT3>
__global__ void __launch_bounds__(F_THREADS, 1)
sim_fused_kernel(const __grid_constant__ LevelMaps amaps, const __grid_constant__ CUtensorMap tmap_b,
                 const FusedParams p) {
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
  const uint32_t a_ready0 = bars + 8u * bi_;     bi_ += F_MAX_KB;
  const uint32_t a_free0 = bars + 8u * bi_;      bi_ += F_MAX_KB;
  const uint32_t t_full0 = bars + 8u * bi_;      bi_ += 2;
  const uint32_t t_empty0 = bars + 8u * bi_;     bi_ += 2;
  const uint32_t n_ready0 = bars + 8u * bi_;
  volatile uint32_t* tmem_ptr_smem = reinterpret_cast<volatile uint32_t*>(base_ptr + FSmem::tmem_ptr_off);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_b);
    for (int l = 0; l < p.levels; ++l) ptx::prefetch_tmap(&amaps.m[l]);
    for (int s = 0; s < F_B_STAGES; ++s) { ptx::mbar_init(b_full0 + 8u * s, 1); ptx::mbar_init(b_empty0 + 8u * s, 1); }
    for (int s = 0; s < F_A_STAGES; ++s) { ptx::mbar_init(as_full0 + 8u * s, 1); ptx::mbar_init(as_empty0 + 8u * s, 4); }
    for (int k = 0; k < F_MAX_KB; ++k) { ptx::mbar_init(a_ready0 + 8u * k, 4); ptx::mbar_init(a_free0 + 8u * k, 1); }
    for (int s = 0; s < 2; ++s) { ptx::mbar_init(t_full0 + 8u * s, 1); ptx::mbar_init(t_empty0 + 8u * s, 4); }
    for (int s = 0; s < 3; ++s) ptx::mbar_init(n_ready0 + 8u * s, 4);
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(base + FSmem::tmem_ptr_off, F_TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  const int total_tiles = p.tile_start[p.levels];
  const int KB = KB_T ? KB_T : p.kb;
  const int KB_IN = KB_T ? KB_T : p.kb_in;
  const int NT = p.n_tiles;

  // Producer / issuer warps run warp-uniform control flow; `issue` is 1 in one elected lane and
  // predicates the single-thread instructions (see ptx::elect_one).
  if (warp == 0) {
    // ================================ text (B) producer ======================================
    const uint32_t issue = ptx::elect_one();
    uint32_t g = 0;                                    // N tiles produced so far
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const TileCoord tc = decode_tile(p, tile);
      const int tb = p.text_batched ? tc.b : 0;
      for (int nt = 0; nt < NT; ++nt, ++g) {
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
          const uint32_t it = g * (uint32_t)KB + (uint32_t)kb;       // k blocks produced so far
          const uint32_t s = it % F_B_STAGES, ph = (it / F_B_STAGES) & 1u;
          ptx::mbar_wait(b_empty0 + 8u * s, ph ^ 1u);
          ptx::mbar_arrive_expect_tx_if(issue, b_full0 + 8u * s, F_B_STAGE_BYTES);
          ptx::tma_load_3d_if(issue, smem_b + s * F_B_STAGE_BYTES, &tmap_b, b_full0 + 8u * s,
                              kb * F_BLOCK_K, nt * F_BLOCK_N, tb);
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer =============================================
    // broadcast from lane 0 so that the compiler keeps every MMA operand in uniform registers
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    uint32_t g = 0, lt = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++lt) {
      for (int nt = 0; nt < NT; ++nt, ++g) {
        int n_size = p.classes - nt * F_BLOCK_N;
        n_size = n_size >= F_BLOCK_N ? F_BLOCK_N : ((n_size + 15) & ~15);
        const uint32_t idesc = ptx::umma_idesc_bf16_f32(F_BLOCK_M, (uint32_t)n_size);
        const uint32_t as = g & 1u;
        const bool first_nt = nt == 0, last_nt = nt == NT - 1;
        const uint32_t it0 = g * (uint32_t)KB;
        // peek at the first text stage while waiting for the accumulator to drain
        bool ready = ptx::mbar_try_wait(b_full0 + 8u * (it0 % F_B_STAGES), (it0 / F_B_STAGES) & 1u);
        ptx::mbar_wait(t_empty0 + 8u * as, ((g >> 1) & 1u) ^ 1u);
        const uint32_t d_tmem = tmem_u + (uint32_t)F_ACC_COL + as * (uint32_t)F_BLOCK_N;
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
          const uint32_t it = it0 + (uint32_t)kb;
          const uint32_t s = it % F_B_STAGES, ph = (it / F_B_STAGES) & 1u;
          if (first_nt) ptx::mbar_wait(a_ready0 + 8u * kb, lt & 1u);     // A block converted?
          ptx::mbar_wait_if_not(ready, b_full0 + 8u * s, ph);
          ptx::tc_fence_after();
          if (kb + 1 < KB)                                               // hide the next wait's latency
            ready = ptx::mbar_try_wait(b_full0 + 8u * ((it + 1) % F_B_STAGES), ((it + 1) / F_B_STAGES) & 1u);
          const uint64_t b_desc = ptx::umma_desc_k_sw128(smem_b + s * F_B_STAGE_BYTES);
          const uint32_t a_tmem = tmem_u + (uint32_t)(F_A_COL + kb * 32);
          if (ptx::elect_one()) {
#pragma unroll
            for (int k = 0; k < F_BLOCK_K / 16; ++k)
              ptx::umma_bf16_ts(d_tmem, a_tmem + 8u * k, b_desc + 2u * k, idesc, (kb | k) != 0);
            ptx::umma_commit(b_empty0 + 8u * s);                  // text stage reusable
            if (last_nt) ptx::umma_commit(a_free0 + 8u * kb);     // A block kb may be overwritten
          }
          __syncwarp();
        }
        if (ptx::elect_one()) ptx::umma_commit(t_full0 + 8u * as);
        __syncwarp();
      }
    }
  } else if (warp == 2) {
    // ================================ activation (A) producer ================================
    const uint32_t issue = ptx::elect_one();
    uint32_t ia = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const TileCoord tc = decode_tile(p, tile);
      const CUtensorMap* map = &amaps.m[tc.level];
#pragma unroll
      for (int kb = 0; kb < KB_IN; ++kb, ++ia) {
        const uint32_t s = ia % F_A_STAGES;
        const uint32_t ph = (ia / F_A_STAGES) & 1u;
        ptx::mbar_wait(as_empty0 + 8u * s, ph ^ 1u);
        ptx::mbar_arrive_expect_tx_if(issue, as_full0 + 8u * s, F_A_STAGE_BYTES);
        ptx::tma_load_3d_if(issue, smem_a + s * F_A_STAGE_BYTES, map, as_full0 + 8u * s, tc.m0,
                            kb * F_BLOCK_K, tc.b);
      }
    }
  } else if (warp == 3) {
    // ================================ L2 prefetcher ==========================================
    // Only A_STAGES + 1 blocks of the next anchor tile can be staged before the current tile
    // releases its TMEM blocks, so the rest of the fp32 tile is fetched inside the last N tile.
    // Pull the whole next tile into L2 one tile ahead so that those loads are L2 hits.
    const uint32_t issue = ptx::elect_one();
    uint32_t lt = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++lt) {
      const int next = tile + gridDim.x;
      if (next >= total_tiles || (p.dbg & 1)) break;
      ptx::mbar_wait(a_ready0, lt & 1u);               // conversion of the current tile has begun
      const TileCoord tc = decode_tile(p, next);
      const CUtensorMap* map = &amaps.m[tc.level];
#pragma unroll
      for (int kb = 0; kb < KB_IN; ++kb)
        ptx::tma_prefetch_l2_3d_if(issue, map, tc.m0, kb * F_BLOCK_K, tc.b);
    }
  } else if (warp >= 4 && warp < 8) {
    // ================================ converters =============================================
    const int lg = warp & 3;
    const int arow = lg * 32 + lane;                 // anchor row of the tile == TMEM lane
    uint32_t ia = 0, lt = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++lt) {
      const TileCoord tc = decode_tile(p, tile);
      float ss0 = 0.f, ss1 = 0.f, ss2 = 0.f, ss3 = 0.f;
      // block `t` of the A region: wait until the previous tile's MMAs have read it, store, publish
      auto publish = [&](int t, const uint32_t (&regs)[32]) {
        ptx::mbar_wait(a_free0 + 8u * t, (lt & 1u) ^ 1u);
        ptx::tc_fence_after();
        ptx::tmem_st_32x32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(F_A_COL + t * 32), regs);
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(a_ready0 + 8u * t);
      };
#pragma unroll
      for (int kb = 0; kb < KB_IN; ++kb, ++ia) {
        const uint32_t s = ia % F_A_STAGES;
        ptx::mbar_wait(as_full0 + 8u * s, (ia / F_A_STAGES) & 1u);
        const float* col = a_stage_ptr + s * (F_A_STAGE_BYTES / 4) + arow;
        uint32_t packed[32];
        uint32_t packed_lo[32];                            // dead (eliminated) unless SPLIT3
#pragma unroll
        for (int k = 0; k < 64; k += 4) {
          const float x0 = col[(k + 0) * F_BLOCK_M], x1 = col[(k + 1) * F_BLOCK_M];
          const float x2 = col[(k + 2) * F_BLOCK_M], x3 = col[(k + 3) * F_BLOCK_M];
          ss0 = fmaf(x0, x0, ss0); ss1 = fmaf(x1, x1, ss1);
          ss2 = fmaf(x2, x2, ss2); ss3 = fmaf(x3, x3, ss3);
          packed[(k >> 1) + 0] = pack_bf16x2(x0, x1);
          packed[(k >> 1) + 1] = pack_bf16x2(x2, x3);
          if constexpr (SPLIT3) {
            const uint32_t h01 = packed[(k >> 1) + 0], h23 = packed[(k >> 1) + 1];
            packed_lo[(k >> 1) + 0] = pack_bf16x2(x0 - __uint_as_float(h01 << 16),
                                                  x1 - __uint_as_float(h01 & 0xffff0000u));
            packed_lo[(k >> 1) + 1] = pack_bf16x2(x2 - __uint_as_float(h23 << 16),
                                                  x3 - __uint_as_float(h23 & 0xffff0000u));
          }
        }
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(as_empty0 + 8u * s);   // staging slot may be refilled
        publish(kb, packed);
        if constexpr (SPLIT3) {
          publish(KB_IN + kb, packed);
          publish(2 * KB_IN + kb, packed_lo);
        }
      }
      const float inv = 1.0f / fmaxf(sqrtf((ss0 + ss1) + (ss2 + ss3)), 1e-12f);
      const int slot = lt % 3;
      norm_s[slot * F_BLOCK_M + arow] = inv;
      if (p.inv_norm != nullptr && arow < tc.rows) p.inv_norm[tc.out_row0 + arow] = inv;
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(n_ready0 + 8u * slot);
    }
  } else if (warp >= 8) {
    // ================================ epilogue ===============================================
    const int lg = warp & 3;
    float* stage = epi_stage + lg * 32 * F_PITCH;
    const bool want_max = p.row_max != nullptr;
    uint32_t acc_it = 0, lt = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++lt) {
      const TileCoord tc = decode_tile(p, tile);
      const int r_in_tile = lg * 32 + lane;
      const bool row_ok = r_in_tile < tc.rows;
      const long long grow = tc.out_row0 + r_in_tile;
      const int slot = lt % 3;
      ptx::mbar_wait(n_ready0 + 8u * slot, (lt / 3) & 1u);
      const float scale = p.normalize ? p.alpha * norm_s[slot * F_BLOCK_M + r_in_tile] : p.alpha;
      const float beta = p.beta;
      float bv[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
      int bi[4] = {0, 0, 0, 0};
      // Max without argmax or logits (the attention row): the row scale is >= 0, so
      // max_c fma(scale, acc_c, beta) == fma(scale, max_c acc_c, beta) exactly (rounding is
      // monotone) and the running maximum is taken over the raw accumulators, one FMNMX3 per
      // two values instead of FFMA + compare + two selects per value.
      const bool max_only = want_max && p.row_arg == nullptr && p.logits == nullptr && p.alpha >= 0.f;
      float raw_best = -INFINITY;
      for (int nt = 0; nt < NT; ++nt, ++acc_it) {
        const int n0 = nt * F_BLOCK_N;
        const int n_valid = min(F_BLOCK_N, p.classes - n0);
        const int nchunks = (n_valid + 31) >> 5;
        const int as = acc_it & 1;
        ptx::mbar_wait(t_full0 + 8u * as, (acc_it >> 1) & 1u);
        ptx::tc_fence_after();
        const uint32_t t_row = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(F_ACC_COL + as * F_BLOCK_N);

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
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(fmaf(scale, __uint_as_float(r[j]), beta));
          if (want_max) {
            const int col = n0 + c0;
            if (valid >= 32) {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float v = __uint_as_float(r[j]);
                if (v > bv[j & 3]) { bv[j & 3] = v; bi[j & 3] = col + j; }
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float v = __uint_as_float(r[j]);
                if (j < valid && v > bv[j & 3]) { bv[j & 3] = v; bi[j & 3] = col + j; }
              }
            }
          }
          if (p.logits != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; ++j) stage[lane * F_PITCH + j] = __uint_as_float(r[j]);
            __syncwarp();
            const int col = n0 + c0 + lane;
            const bool col_ok = lane < valid;
            const int rows_here = min(32, tc.rows - lg * 32);
            const long long out_row0 = tc.out_row0 + lg * 32;
            if (p.logits_bf16) {
              __nv_bfloat16* out = static_cast<__nv_bfloat16*>(p.logits);
              for (int i = 0; i < rows_here; ++i)
                if (col_ok) out[(out_row0 + i) * p.ldc + col] = __float2bfloat16_rn(stage[i * F_PITCH + lane]);
            } else {
              float* out = static_cast<float*>(p.logits);
              for (int i = 0; i < rows_here; ++i)
                if (col_ok) out[(out_row0 + i) * p.ldc + col] = stage[i * F_PITCH + lane];
            }
            __syncwarp();
          }
        };

        uint32_t ra[32], rb[32];
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
        if (lane == 0) ptx::mbar_arrive(t_empty0 + 8u * as);
      }
      if (max_only) {
        if (row_ok) p.row_max[grow] = fmaf(scale, raw_best, beta);
      } else if (want_max && row_ok) {
        float best = bv[0];
        int best_idx = bi[0];
#pragma unroll
        for (int q = 1; q < 4; ++q)
          if (bv[q] > best || (bv[q] == best && bi[q] < best_idx)) { best = bv[q]; best_idx = bi[q]; }
        p.row_max[grow] = best;
        if (p.row_arg != nullptr) p.row_arg[grow] = best_idx;
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, F_TMEM_COLS);
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

}  // namespace

// Shared launcher.  `dim` is the real embedding length; it is padded to a multiple of 64 by the
// TMA zero fill on the activation side and by zero columns in the text operand, whose row length
// is kop = (split3 ? 3 : 1) * ceil(dim / 64) * 64.
int fused_launch(const float* const* obj_embeds, const int64_t* hw, const int64_t* stride_b,
                 const int64_t* stride_d, int num_levels, int64_t batch, int64_t dim,
                 const void* text_op, int64_t classes, int text_batched, int normalize, int split3,
                 float alpha, float beta, void* logits, int logits_dtype, int64_t ldc,
                 float* row_max, int32_t* row_arg, float* inv_norm, void* stream) {
  if (!obj_embeds || !hw || !stride_b || !stride_d || !text_op || batch < 0 || classes <= 0 || dim <= 0)
    return OVDET_ERR_INVALID_ARG;
  if (num_levels <= 0) return OVDET_ERR_INVALID_ARG;
  if (!logits && !row_max) return OVDET_ERR_INVALID_ARG;
  if (row_arg && !row_max) return OVDET_ERR_INVALID_ARG;
  if (logits && (ldc < classes || (logits_dtype != OVDET_F32 && logits_dtype != OVDET_BF16)))
    return OVDET_ERR_INVALID_ARG;
  const int kb_in = (int)ceil_div<int64_t>(dim, F_BLOCK_K);
  const int kb = kb_in * (split3 ? 3 : 1);
  if (num_levels > F_MAX_LEVELS || kb > F_MAX_KB || batch > 65535) return OVDET_ERR_UNSUPPORTED_SHAPE;
  if ((uintptr_t)text_op & 15) return OVDET_ERR_INVALID_ARG;
  EncodeTiledFn enc = nullptr;
  FusedParams p{};
  LevelMaps maps;
  long long anchors = 0, tiles = 0;
  for (int l = 0; l < num_levels; ++l) {
    if (!obj_embeds[l] || hw[l] <= 0) return OVDET_ERR_INVALID_ARG;
    // TMA needs 16-byte aligned base and strides
    if (((uintptr_t)obj_embeds[l] & 15) || (stride_d[l] & 3) || (stride_b[l] & 3) || stride_d[l] < hw[l])
      return OVDET_ERR_UNSUPPORTED_SHAPE;
    p.hw[l] = (int)hw[l];
    p.mt[l] = (int)ceil_div<int64_t>(hw[l], F_BLOCK_M);
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
    cuuint64_t strides[2] = {(cuuint64_t)stride_d[l] * 4, (cuuint64_t)stride_b[l] * 4};
    cuuint32_t box[3] = {(cuuint32_t)F_BLOCK_M, (cuuint32_t)F_BLOCK_K, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    if (batch == 1) strides[1] = (cuuint64_t)dim * stride_d[l] * 4;      // unused but must be valid
    CUresult r = enc(&maps.m[l], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(obj_embeds[l]),
                     dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return OVDET_ERR_DRIVER;
  }
  for (int l = num_levels; l < F_MAX_LEVELS; ++l) maps.m[l] = maps.m[0];
  CUtensorMap map_b;
  {
    const int64_t tb = text_batched ? batch : 1;
    const int64_t kop = (int64_t)kb * F_BLOCK_K;
    cuuint64_t dims[3] = {(cuuint64_t)kop, (cuuint64_t)classes, (cuuint64_t)tb};
    cuuint64_t strides[2] = {(cuuint64_t)kop * 2, (cuuint64_t)classes * (cuuint64_t)kop * 2};
    cuuint32_t box[3] = {(cuuint32_t)F_BLOCK_K, (cuuint32_t)F_BLOCK_N, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&map_b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(text_op), dims, strides,
                     box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return OVDET_ERR_DRIVER;
  }
  p.levels = num_levels;
  p.batch = (int)batch;
  p.tile_start[num_levels] = (int)tiles;
  p.anchors = (int)anchors;
  p.classes = (int)classes;
  p.kb = kb;
  p.kb_in = kb_in;
  p.normalize = normalize ? 1 : 0;
  p.n_tiles = (int)ceil_div<int64_t>(classes, F_BLOCK_N);
  p.text_batched = text_batched ? 1 : 0;
  p.alpha = alpha;
  p.beta = beta;
  p.logits = logits;
  p.logits_bf16 = logits_dtype == OVDET_BF16;
  p.ldc = ldc;
  p.row_max = row_max;
  p.row_arg = row_arg;
  p.inv_norm = inv_norm;
  { const char* e = getenv("OVDET_DBG"); p.dbg = e ? atoi(e) : 0; }

  static bool attr_set = false;
  if (!attr_set) {
    OVDET_CUDA_TRY(cudaFuncSetAttribute(sim_fused_kernel<8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, F_SMEM_BYTES));
    OVDET_CUDA_TRY(cudaFuncSetAttribute(sim_fused_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, F_SMEM_BYTES));
    OVDET_CUDA_TRY(cudaFuncSetAttribute(sim_fused_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, F_SMEM_BYTES));
    attr_set = true;
  }
  const int grid = (int)(tiles < sm_count() ? tiles : sm_count());
  if (split3)
    sim_fused_kernel<0, true><<<grid, F_THREADS, F_SMEM_BYTES, as_stream(stream)>>>(maps, map_b, p);
  else if (p.kb == 8)
    sim_fused_kernel<8, false><<<grid, F_THREADS, F_SMEM_BYTES, as_stream(stream)>>>(maps, map_b, p);
  else
    sim_fused_kernel<0, false><<<grid, F_THREADS, F_SMEM_BYTES, as_stream(stream)>>>(maps, map_b, p);
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
  return ovdet::fused_launch(obj_embeds, hw, stride_b, stride_d, num_levels, batch, dim, text_op, classes,
                             text_batched, /*normalize=*/1, /*split3=*/0, alpha, beta, logits, logits_dtype,
                             ldc, row_max, row_arg, inv_norm, stream);
}
