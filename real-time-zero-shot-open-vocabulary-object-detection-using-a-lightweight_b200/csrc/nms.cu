// K4: per-image candidate gather, rescale/clip, total-order sort, optional top-k and greedy NMS.
//
// Replaces the host numpy tail of the reference, inference/detector.py:185-208: the boolean
// mask gather (:185-187), boxes / scale_factor (:193-196), np.clip to the original size
// (:199-202), `_nms` (:225-256: argsort(scores)[::-1], greedy, keep iou <= thr) and
// `_compute_iou` (:258-287: float32, +1e-7 on the union) - for every image of the batch, where
// the reference handles image 0 only.
//
// One CTA per image, four phases separated by block barriers:
//   1. gather   pass-mask words -> 64-bit keys (ordered(score) << 32 | anchor) in the workspace
//   2. sort     bitonic, descending; 4096-key blocks in shared memory, larger merge steps in the
//               (L2-resident) workspace.  Keys are pairwise distinct, so the order is the total
//               order (score desc, anchor index desc) == np.argsort(kind='stable')[::-1]
//   3. select   optional top-k truncation; bitmask of the selected anchors + prefix popcounts so
//               that every survivor knows its index in the thresholded array (what `_nms` returns)
//   (resident path, the common case: no top-k, <= 4096 candidates, <= 65536 anchors: the keys are
//   placed in anchor order straight into the shared-memory sort block from an in-block prefix
//   scan of the mask popcounts - which is also every survivor's index in the thresholded array -
//   sorted there, and never touch the workspace; phases 1-3 then cost three dependent global
//   reads instead of ten.  At batch 1 the kernel went from 37.6 to ~20 us.)
//   4. nms      chunks of 512 sorted candidates: boxes gathered / rescaled / clipped into shared
//               memory, a 512 x 512 suppression bitmask built with one warp ballot per 32 IoUs
//               (only tiles on or above the diagonal), one warp resolves the chunk 32 candidates
//               at a time; later chunks are first tested against the boxes already kept.
//
// The float32 IoU follows numpy operation by operation (no FMA contraction, IEEE division).
#include "common.cuh"

namespace ovdet {
namespace {

constexpr int NMS_THREADS = 512;
constexpr int NMS_WARPS = NMS_THREADS / 32;
constexpr int SORT_CAP = 4096;                 // keys per shared-memory sort block (32 KiB)
constexpr int CHUNK = 512;                     // candidates per NMS chunk
constexpr int CHUNK_WORDS = CHUNK / 32;        // 16
constexpr int FAST_WORDS = 2048;               // resident path: pass-mask words per image (anchors <= 65536)
constexpr int NMS_SMEM_BYTES = SORT_CAP * 8 + CHUNK * 16 + 4 * CHUNK * 4 + NMS_THREADS * 4 + 2 * FAST_WORDS * 4;
static_assert(CHUNK * CHUNK_WORDS * 4 <= SORT_CAP * 8, "mask must fit in the sort block");
static_assert(CHUNK == NMS_THREADS, "one thread per chunk candidate");

struct NmsParams {
  const float* boxes;
  const float* scores;
  const int* classes;
  const uint32_t* pass_mask;
  int pdl;                                     // launched with programmatic stream serialization (ovdet_head_step)
  int use_conf;                                // 1: candidates = scores > conf (computed here, no pass mask)
  float conf;
  int anchors;
  int words;
  const float* scale;
  const float* clip_wh;
  float iou_thr;
  int class_aware;
  int topk;
  int max_det;
  float* out_boxes;
  float* out_scores;
  int* out_classes;
  int* out_anchor;
  int* out_keep;
  int* out_count;
  int* out_candidates;
  unsigned char* ws;
  size_t ws_per_image;
  int pow2_cap;                                // next power of two >= anchors (>= 32)
};

// workspace carve-up of one image (all offsets 16-byte aligned)
struct WsLayout {
  size_t keys, kept_boxes, kept_cls, sel, prefix, total;
  __host__ __device__ WsLayout(int anchors, int words, int pow2_cap) {
    size_t o = 0;
    keys = o;        o += (size_t)pow2_cap * 8;
    kept_boxes = o;  o += (size_t)anchors * 16;
    kept_cls = o;    o += (((size_t)anchors * 4) + 15) & ~(size_t)15;
    sel = o;         o += (((size_t)words * 4) + 15) & ~(size_t)15;
    prefix = o;      o += (((size_t)(words + 1) * 4) + 15) & ~(size_t)15;
    total = o;
  }
};

__device__ __forceinline__ void cmp_swap_desc(unsigned long long& a, unsigned long long& b, bool desc) {
  // desc: larger key first
  if ((a < b) == desc) { const unsigned long long t = a; a = b; b = t; }
}

// bitonic steps j = j_hi, j_hi/2, ..., 1 of merge size k over `len` keys held in shared memory;
// `base` is the global index of sk[0] (direction depends on the global index).
__device__ void smem_bitonic_steps(unsigned long long* sk, int len, int base, int k, int j_hi) {
  for (int j = j_hi; j >= 1; j >>= 1) {
    for (int t = threadIdx.x; t < (len >> 1); t += NMS_THREADS) {
      const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
      const int l = i | j;
      const bool desc = (((base + i) & k) == 0);
      unsigned long long a = sk[i], b = sk[l];
      cmp_swap_desc(a, b, desc);
      sk[i] = a; sk[l] = b;
    }
    __syncthreads();
  }
}

// descending bitonic sort of keys[0, P) (P a power of two) held in the workspace
__device__ void sort_keys_desc(unsigned long long* keys, int P, unsigned long long* sk) {
  const int blk = P < SORT_CAP ? P : SORT_CAP;
  // phase A: every block fully sorted (direction alternates with the global index)
  for (int base = 0; base < P; base += blk) {
    for (int i = threadIdx.x; i < blk; i += NMS_THREADS) sk[i] = keys[base + i];
    __syncthreads();
    for (int k = 2; k <= blk; k <<= 1) smem_bitonic_steps(sk, blk, base, k, k >> 1);
    for (int i = threadIdx.x; i < blk; i += NMS_THREADS) keys[base + i] = sk[i];
    __syncthreads();
  }
  // phase B: merges wider than one block: wide steps in the workspace, the rest in smem
  for (int k = blk << 1; k <= P; k <<= 1) {
    for (int j = k >> 1; j >= blk; j >>= 1) {
      for (int t = threadIdx.x; t < (P >> 1); t += NMS_THREADS) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int l = i | j;
        unsigned long long a = keys[i], b = keys[l];
        const bool desc = ((i & k) == 0);
        if ((a < b) == desc) { keys[i] = b; keys[l] = a; }
      }
      __syncthreads();
    }
    for (int base = 0; base < P; base += blk) {
      for (int i = threadIdx.x; i < blk; i += NMS_THREADS) sk[i] = keys[base + i];
      __syncthreads();
      smem_bitonic_steps(sk, blk, base, k, blk >> 1);
      for (int i = threadIdx.x; i < blk; i += NMS_THREADS) keys[base + i] = sk[i];
      __syncthreads();
    }
  }
}

__device__ __forceinline__ float box_area(const float4 b) {
  return __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
}

// The threshold test `RN(inter / denom) <= thr` (float32 division, as numpy computes the IoU)
// without the division: for denom > 0 the rounded quotient is <= thr exactly when the real
// quotient is below the midpoint m between thr and the next float above it (or equal to it when
// the tie rounds down to thr, i.e. thr's mantissa is even).  m has a 25-bit mantissa and denom a
// 24-bit one, so denom * m is exact in double and the comparison is exact.
//
// The double arithmetic is itself slow on this GPU (a warp-wide FP64 multiply or conversion issues at a small
// fraction of the fp32 rate: with it on every pair the 171-candidate mask of a batch-1 call took 17 k cycles),
// so it only DECIDES THE CLOSE CALLS: an approximate quotient a = inter * rcp(denom) (relative error < 2^-21)
// settles every pair with a < thr (1 - 2^-20) (certainly kept) or a > thr (1 + 2^-20) (certainly suppressed);
// what lies between - about one pair in a million - takes the exact test below.  Same decisions, bit for bit.
struct IouTest {
  double mid;        // (thr + nextafter(thr, +inf)) / 2
  int tie_down;      // a quotient exactly at `mid` rounds to thr
  int exact_ok;      // thr finite and >= 0: the double test applies
  float thr;
  float lo, hi;      // below lo: iou <= thr for certain; above hi: iou > thr for certain (lo > hi: filter off)
};

// the close calls and the out-of-range inputs: exact, out of line (about one pair in a million gets here)
__device__ __noinline__ bool suppresses_exact(const float inter, const float denom, const IouTest& t) {
  if (t.exact_ok && denom > 0.f && inter <= 3.0e38f) {          // (NaN fails both tests)
    const double lhs = (double)inter, rhs = __dmul_rn((double)denom, t.mid);
    const bool keep = lhs < rhs || (t.tie_down && lhs == rhs);
    return !keep;
  }
  const float iou = __fdiv_rn(inter, denom);
  return !(iou <= t.thr);
}

// Branch-free part of the test: intersection and denominator as numpy forms them, then the approximate
// quotient q = inter * rcp.approx(denom) (|relative error| < 2^-21).  `sure_sup`: q above the upper bound -
// suppressed for certain; `unsure`: q between the bounds, not a number, or a denominator beyond the
// approximation's range - the exact test decides.  (Non-positive denominators need no guard: the approximate
// quotient then has the sign / infinity the IEEE quotient has, and the bounds are positive.)
__device__ __forceinline__ void iou_classify(const float4 a, const float area_a, const float4 b, const float area_b,
                                             const IouTest& t, float& inter, float& denom, bool& sure_sup, bool& unsure) {
  const float ix1 = fmaxf(a.x, b.x), iy1 = fmaxf(a.y, b.y);
  const float ix2 = fminf(a.z, b.z), iy2 = fminf(a.w, b.w);
  const float iw = fmaxf(0.f, __fsub_rn(ix2, ix1));
  const float ih = fmaxf(0.f, __fsub_rn(iy2, iy1));
  inter = __fmul_rn(iw, ih);
  const float uni = __fsub_rn(__fadd_rn(area_a, area_b), inter);
  denom = __fadd_rn(uni, 1e-7f);
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(denom));
  const float q = __fmul_rn(inter, r);
  sure_sup = q > t.hi;
  unsure = !((sure_sup || q < t.lo) && denom < 1.0e30f);
}

// true when `b` must be dropped because of the already kept `a`: NOT (iou <= thr)
__device__ __forceinline__ bool suppresses(const float4 a, const float area_a, const float4 b,
                                           const float area_b, const IouTest& t) {
  float inter, denom;
  bool sure_sup, unsure;
  iou_classify(a, area_a, b, area_b, t, inter, denom, sure_sup, unsure);
  if (!unsure) return sure_sup;
  return suppresses_exact(inter, denom, t);
}

__global__ void __launch_bounds__(NMS_THREADS)
nms_batched_kernel(const NmsParams p) {
  // shared memory (dynamic, > 48 KiB in total): the sort block and the suppression mask alias
  // each other (different phases)
  extern __shared__ __align__(16) unsigned char s_raw[];
  unsigned long long* s_big = reinterpret_cast<unsigned long long*>(s_raw);      // 32 KiB
  float4* s_box = reinterpret_cast<float4*>(s_raw + SORT_CAP * 8);                // 8 KiB
  float* s_area = reinterpret_cast<float*>(s_box + CHUNK);
  int* s_cls = reinterpret_cast<int*>(s_area + CHUNK);
  int* s_anchor = s_cls + CHUNK;
  int* s_kept_pos = s_anchor + CHUNK;           // positions (inside the chunk) kept by this chunk
  int* s_scan = s_kept_pos + CHUNK;             // NMS_THREADS entries
  int* s_prefix = s_scan + NMS_THREADS;         // FAST_WORDS entries (resident path)
  uint32_t* s_mask = reinterpret_cast<uint32_t*>(s_prefix + FAST_WORDS);   // FAST_WORDS entries (use_conf)
  __shared__ uint32_t s_removed[CHUNK_WORDS];
  __shared__ int s_count, s_kept_total, s_kept_chunk, s_carry;

#ifdef OVDET_NMS_TRACE
  // timing build only: thread 0 drops the low 32 bits of the SM clock at the phase boundaries into the unused
  // tail of out_keep (tools/batch1_breakdown.py reads them back)
#define NMS_TR(i) do { if (threadIdx.x == 0 && p.out_keep && p.max_det >= 300) p.out_keep[(size_t)blockIdx.x * p.max_det + 280 + (i)] = (int)clock64(); } while (0)
#else
#define NMS_TR(i) do { } while (0)
#endif
  NMS_TR(0);
  if (p.pdl) cudaGridDependencySynchronize();   // the decode kernel (and, through it, the similarity kernel) is done
  NMS_TR(1);
  const int b = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int A = p.anchors, W = p.words;
  const WsLayout L(A, W, p.pow2_cap);
  unsigned char* ws = p.ws + (size_t)b * p.ws_per_image;
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(ws + L.keys);
  float4* kept_boxes = reinterpret_cast<float4*>(ws + L.kept_boxes);
  int* kept_cls = reinterpret_cast<int*>(ws + L.kept_cls);
  uint32_t* sel = reinterpret_cast<uint32_t*>(ws + L.sel);
  int* prefix = reinterpret_cast<int*>(ws + L.prefix);

  const float* scores = p.scores + (size_t)b * A;
  const float4* boxes = reinterpret_cast<const float4*>(p.boxes) + (size_t)b * A;
  const int* classes = p.classes ? p.classes + (size_t)b * A : nullptr;
  const uint32_t* pm = p.pass_mask ? p.pass_mask + (size_t)b * W : nullptr;

  if (tid == 0) { s_count = 0; s_kept_total = 0; s_carry = 0; }
  __syncthreads();

  const uint32_t tail_bits = (A & 31) ? ((1u << (A & 31)) - 1u) : 0xffffffffu;
  // use_conf: the confidence threshold of detector.py:184 evaluated here (strict >, NaN never
  // passes), coalesced over the scores, one ballot per 32 anchors into shared memory - so that the
  // box decode does not have to wait for the scores (it runs beside the similarity kernel)
  const bool own_mask = p.use_conf && W <= FAST_WORDS;
  if (own_mask) {
    for (int a0 = 0; a0 < W * 32; a0 += NMS_THREADS) {
      const int a = a0 + tid;
      const bool pass = a < A && scores[a] > p.conf;
      const uint32_t bits = __ballot_sync(0xffffffffu, pass);
      if (lane == 0 && (a >> 5) < W) s_mask[a >> 5] = bits;
    }
    __syncthreads();
  }
  auto mask_word = [&](int w) -> uint32_t {
    if (own_mask) return s_mask[w];
    uint32_t bits = pm ? pm[w] : 0xffffffffu;
    if (w == W - 1) bits &= tail_bits;
    return bits;
  };
  int N, M;
  bool resident = false;
  // one mask word per thread (anchors <= 16384): the scores of the word's first four candidates are loaded
  // BEFORE the block-wide scan, so that their L2 round trip overlaps it (the key build used to walk the set
  // bits with one dependent load each: ~3 k cycles at batch 1 for clumps of three candidates per word)
  float early_score[4] = {0.f, 0.f, 0.f, 0.f};
  int early_anchor[4] = {0, 0, 0, 0};
  int early_n = 0;
  uint32_t early_rest = 0;
  const unsigned long long* sorted_keys = keys;     // where phase 4 reads the sorted keys from

  if (W <= FAST_WORDS) {
    // ---- 0. exclusive prefix of the mask popcounts (contiguous words per thread) -------------
    const int per = (W + NMS_THREADS - 1) / NMS_THREADS;           // <= 4
    const int w_lo = tid * per;
    int local = 0;
    for (int i = 0; i < per; ++i)
      if (w_lo + i < W) local += __popc(mask_word(w_lo + i));
    if (per == 1 && tid < W) {
      early_rest = mask_word(tid);
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (early_rest) {
          const int a = (tid << 5) + __ffs(early_rest) - 1;
          early_rest &= early_rest - 1;
          early_anchor[q] = a;
          early_score[q] = scores[a];
          early_n = q + 1;
        }
    }
    int incl = local;                                              // inclusive scan over the block
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) s_scan[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int v = (lane < NMS_WARPS) ? s_scan[lane] : 0;
#pragma unroll
      for (int o = 1; o < NMS_WARPS; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += u;
      }
      if (lane < NMS_WARPS) s_scan[lane] = v;                      // inclusive warp totals
    }
    __syncthreads();
    N = s_scan[NMS_WARPS - 1];
    int run = incl - local + (warp ? s_scan[warp - 1] : 0);
    for (int i = 0; i < per; ++i)
      if (w_lo + i < W) { s_prefix[w_lo + i] = run; run += __popc(mask_word(w_lo + i)); }
    resident = N <= SORT_CAP && (p.topk == 0 || p.topk >= N);
    __syncthreads();
    NMS_TR(2);
  } else {
    N = -1;
  }

  if (resident) {
    if (p.out_candidates && tid == 0) p.out_candidates[b] = N;
    if (N == 0) {
      if (tid == 0) p.out_count[b] = 0;
      return;
    }
    // ---- 1r. keys in anchor order, straight into the sort block ------------------------------
    for (int w = tid; w < W; w += NMS_THREADS) {
      uint32_t bits = mask_word(w);
      int slot = s_prefix[w];
      if (W <= NMS_THREADS) {                      // (== the `per == 1` case above: w == tid, one word per thread)
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (q < early_n)
            s_big[slot++] = ((unsigned long long)float_to_ordered(early_score[q]) << 32) | (unsigned)early_anchor[q];
        bits = early_rest;
      }
      while (bits) {
        const int l = __ffs(bits) - 1;
        bits &= bits - 1;
        const int a = (w << 5) + l;
        s_big[slot++] = ((unsigned long long)float_to_ordered(scores[a]) << 32) | (unsigned)a;
      }
    }
    __syncthreads();
    NMS_TR(3);
    // ---- 2r. sort in shared memory -------------------------------------------------------------
    if (N <= CHUNK) {
      // rank sort: thread i counts the keys larger than its own (keys are pairwise distinct, so
      // the ranks are a permutation): one pass of broadcast reads and one barrier instead of
      // the 36 barrier-separated steps a 256-key bitonic network takes.
      // every thread takes a slice of the comparisons of one key (512 / N threads per key, ranks summed
      // with shared-memory atomics): at N = 171 three threads per key instead of 171 busy and 341 idle
      const int per_key = N <= NMS_THREADS / 4 ? 4 : (N <= NMS_THREADS / 2 ? 2 : 1);
      const int key_i = tid / per_key, part = tid - key_i * per_key;
      unsigned long long mine = 0ull;
      if (tid < N) s_scan[tid] = 0;
      __syncthreads();
      if (key_i < N) {
        mine = s_big[key_i];
        const int len = (N + per_key - 1) / per_key;
        const int j_end = min(N, (part + 1) * len);
        int rank = 0;
        for (int j = part * len; j < j_end; ++j) rank += s_big[j] > mine;
        if (per_key > 1) atomicAdd(&s_scan[key_i], rank);
        else s_scan[key_i] = rank;
      }
      __syncthreads();
      if (key_i < N && part == 0) s_big[s_scan[key_i]] = mine;
      __syncthreads();
    } else {
      int P = 32;
      while (P < N) P <<= 1;
      for (int i = N + tid; i < P; i += NMS_THREADS) s_big[i] = 0ull;
      __syncthreads();
      for (int k = 2; k <= P; k <<= 1) smem_bitonic_steps(s_big, P, 0, k, k >> 1);
    }
    M = N;
    if (M > CHUNK) {            // several NMS chunks: the mask will overwrite the sort block
      for (int i = tid; i < M; i += NMS_THREADS) keys[i] = s_big[i];
      __syncthreads();
    } else {
      sorted_keys = s_big;      // single chunk: read into registers before the mask is built
    }
  } else {
  // ---- 1. gather ---------------------------------------------------------------------------
  for (int w = tid; w < W; w += NMS_THREADS) {
    uint32_t bits = mask_word(w);
    const int c = __popc(bits);
    if (c) {
      int slot = atomicAdd(&s_count, c);
      while (bits) {
        const int l = __ffs(bits) - 1;
        bits &= bits - 1;
        const int a = (w << 5) + l;
        keys[slot++] = ((unsigned long long)float_to_ordered(scores[a]) << 32) | (unsigned)a;
      }
    }
  }
  __syncthreads();
  N = s_count;
  if (p.out_candidates && tid == 0) p.out_candidates[b] = N;
  if (N == 0) {
    if (tid == 0) p.out_count[b] = 0;
    return;
  }
  int P = 32;
  while (P < N) P <<= 1;
  for (int i = N + tid; i < P; i += NMS_THREADS) keys[i] = 0ull;
  __syncthreads();

  // ---- 2. sort -----------------------------------------------------------------------------
  sort_keys_desc(keys, P, s_big);

  // ---- 3. select (top-k) + rank of every selected anchor in the thresholded array ------------
  M = (p.topk > 0 && p.topk < N) ? p.topk : N;
  for (int w = tid; w < W; w += NMS_THREADS) sel[w] = 0u;
  __syncthreads();
  for (int i = tid; i < M; i += NMS_THREADS) {
    const unsigned a = (unsigned)(keys[i] & 0xffffffffull);
    atomicOr(&sel[a >> 5], 1u << (a & 31));
  }
  __syncthreads();
  for (int w0 = 0; w0 < W; w0 += NMS_THREADS) {       // exclusive scan of popcounts
    const int w = w0 + tid;
    const int c = (w < W) ? __popc(sel[w]) : 0;
    s_scan[tid] = c;
    __syncthreads();
    for (int off = 1; off < NMS_THREADS; off <<= 1) {
      const int v = (tid >= off) ? s_scan[tid - off] : 0;
      __syncthreads();
      s_scan[tid] += v;
      __syncthreads();
    }
    const int carry = s_carry;
    if (w < W) prefix[w] = carry + s_scan[tid] - c;
    __syncthreads();
    if (tid == NMS_THREADS - 1) s_carry = carry + s_scan[tid];
    __syncthreads();
  }
  }

  NMS_TR(4);
  // ---- 4. greedy NMS over the sorted candidates ----------------------------------------------
  const float scale = p.scale ? p.scale[b] : 1.0f;
  const bool do_scale = p.scale != nullptr;
  const bool do_clip = p.clip_wh != nullptr;
  const float clip_w = do_clip ? p.clip_wh[2 * b] : 0.f;
  const float clip_h = do_clip ? p.clip_wh[2 * b + 1] : 0.f;
  IouTest thr;
  thr.thr = p.iou_thr;
  thr.exact_ok = (p.iou_thr >= 0.f && p.iou_thr < 3.0e38f) ? 1 : 0;
  {
    const float up = __uint_as_float(__float_as_uint(p.iou_thr) + 1u);      // next float above (thr >= 0)
    thr.mid = 0.5 * ((double)p.iou_thr + (double)up);
    thr.tie_down = (__float_as_uint(p.iou_thr) & 1u) == 0u;
    // the filter's certainty bounds (rounded outwards); off for thresholds too small to bracket
    if (thr.exact_ok && p.iou_thr >= 1.0e-30f) {
      thr.lo = __fmul_rd(p.iou_thr, 1.0f - 9.5367431640625e-07f);
      thr.hi = __fmul_ru(p.iou_thr, 1.0f + 9.5367431640625e-07f);
    } else {
      thr.lo = -INFINITY;
      thr.hi = INFINITY;
    }
  }
  const bool aware = p.class_aware && classes != nullptr;
  uint32_t* mask = reinterpret_cast<uint32_t*>(s_big);           // [CHUNK][CHUNK_WORDS]

  for (int c0 = 0; c0 < M; c0 += CHUNK) {
    const int n = min(CHUNK, M - c0);
    const int kept_before = s_kept_total;
    if (kept_before >= p.max_det) break;
    // 4a. load the chunk: gather, rescale, clip
    float4 mine = make_float4(0.f, 0.f, 0.f, 0.f);
    float my_area = 0.f;
    int my_cls = -1;
    if (tid < n) {
      const unsigned a = (unsigned)(sorted_keys[c0 + tid] & 0xffffffffull);
      float4 v = boxes[a];
      if (do_scale) {
        v.x = __fdiv_rn(v.x, scale); v.y = __fdiv_rn(v.y, scale);
        v.z = __fdiv_rn(v.z, scale); v.w = __fdiv_rn(v.w, scale);
      }
      if (do_clip) {
        v.x = fminf(fmaxf(v.x, 0.f), clip_w); v.y = fminf(fmaxf(v.y, 0.f), clip_h);
        v.z = fminf(fmaxf(v.z, 0.f), clip_w); v.w = fminf(fmaxf(v.w, 0.f), clip_h);
      }
      mine = v;
      my_area = box_area(v);
      my_cls = classes ? classes[a] : 0;
      s_anchor[tid] = (int)a;
    }
    s_box[tid] = mine;
    s_area[tid] = my_area;
    s_cls[tid] = my_cls;
    // 4b. candidates beyond n are born removed; later chunks: test against everything kept.  The
    // kept boxes are staged through shared memory 512 at a time (box + area + class, in the sort
    // block, which is free until 4c), and when the chunk has fewer than 512 candidates the idle
    // threads take slices of the kept list: thread = (slice, candidate).
    bool dead = tid >= n;
    if (kept_before > 0) {
      float4* k_box = reinterpret_cast<float4*>(s_big);                 // [512]
      float* k_area = reinterpret_cast<float*>(k_box + CHUNK);          // [512]
      int* k_cls = reinterpret_cast<int*>(k_area + CHUNK);              // [512]
      const int n_pad = (n + 31) & ~31;
      const int slices = NMS_THREADS / n_pad;                           // >= 1
      const int cand = tid % n_pad, slice = tid / n_pad;
      const bool active = slice < slices && cand < n;
      __syncthreads();                                                  // s_box / s_area / s_cls of this chunk are visible
      const float4 cb = s_box[active ? cand : 0];
      const float ca = s_area[active ? cand : 0];
      const int cc = s_cls[active ? cand : 0];
      bool hit = false;
      for (int k0 = 0; k0 < kept_before; k0 += CHUNK) {
        const int kn = min(CHUNK, kept_before - k0);
        if (tid < kn) {
          const float4 kb = kept_boxes[k0 + tid];
          k_box[tid] = kb;
          k_area[tid] = box_area(kb);
          k_cls[tid] = kept_cls[k0 + tid];
        }
        __syncthreads();
        if (active && !hit) {
          const int per = (kn + slices - 1) / slices;
          const int k_end = min(kn, (slice + 1) * per);
          for (int k = slice * per; k < k_end; ++k) {
            if (aware && k_cls[k] != cc) continue;
            if (suppresses(k_box[k], k_area[k], cb, ca, thr)) { hit = true; break; }
          }
        }
        __syncthreads();
      }
      // combine the slices: one flag per candidate
      int* s_hit = s_scan;                                              // NMS_THREADS ints, free here
      s_hit[tid] = 0;
      __syncthreads();
      if (hit) s_hit[cand] = 1;
      __syncthreads();
      if (tid < n && s_hit[tid]) dead = true;
      __syncthreads();
    }
    NMS_TR(5);
    const uint32_t dead_bits = __ballot_sync(0xffffffffu, dead);
    if (lane == 0) s_removed[warp] = dead_bits;
    __syncthreads();
    // 4c. suppression bitmask, 32 x 32 tiles on or above the diagonal.  A lane owns one ROW i of the tile
    // and walks the tile's 32 columns j - every lane reads the same s_box[j] (a broadcast) - setting bit j
    // of its own word: 32 independent IoU tests per thread, no ballot and no cross-lane traffic.  (The first
    // version gave a lane one column and took a warp ballot per row: a 32-deep chain of ballots and selects
    // that ran at an IPC of 0.16 - 18.5 k cycles of the kernel's 36 k at batch 1.)
    const int nblk = (n + 31) >> 5;
    const int ntiles = nblk * (nblk + 1) / 2;
    // a task = 16 columns of a tile (two tasks per tile, written as the two halves of the row's word): with 21
    // tiles on 16 warps whole-tile tasks left 11 warps waiting for the 5 that had two
    for (int task = warp; task < 2 * ntiles; task += NMS_WARPS) {
      const int t = task >> 1, h = task & 1;
      // tile index -> (row block rb <= column block cb)
      int cb = 0;
      while ((cb + 1) * (cb + 2) / 2 <= t) ++cb;
      const int rb = t - cb * (cb + 1) / 2;
      const int i = (rb << 5) + lane;
      const float4 bi = s_box[i];
      const float ai = s_area[i];
      const int ci = s_cls[i];
      uint32_t my_word = 0, unsure_word = 0, same_cls = 0xffffu;
#pragma unroll 8
      for (int c = 0; c < 16; ++c) {
        const int j = (cb << 5) + 16 * h + c;
        float inter, denom;
        bool sure_sup, unsure;
        iou_classify(bi, ai, s_box[j], s_area[j], thr, inter, denom, sure_sup, unsure);
        my_word |= (sure_sup ? 1u : 0u) << c;
        unsure_word |= (unsure ? 1u : 0u) << c;
      }
      if (aware) {
        same_cls = 0u;
#pragma unroll 8
        for (int c = 0; c < 16; ++c) same_cls |= (s_cls[(cb << 5) + 16 * h + c] == ci ? 1u : 0u) << c;
      }
      // only later candidates (j > i) of the same class (class-aware) can be suppressed by i
      const uint32_t later32 = rb == cb ? (lane == 31 ? 0u : (0xffffffffu << (lane + 1))) : 0xffffffffu;
      const uint32_t valid = (later32 >> (16 * h)) & 0xffffu & same_cls;
      my_word &= valid & ~unsure_word;
      unsure_word &= valid;
      while (unsure_word) {                                    // the close calls: exact test, about one pair in 10^6
        const int c = __ffs(unsure_word) - 1;
        unsure_word &= unsure_word - 1;
        const int j = (cb << 5) + 16 * h + c;
        float inter, denom;
        bool sure_sup, unsure;
        iou_classify(bi, ai, s_box[j], s_area[j], thr, inter, denom, sure_sup, unsure);
        if (suppresses_exact(inter, denom, thr)) my_word |= 1u << c;
      }
      reinterpret_cast<unsigned short*>(mask)[(i * CHUNK_WORDS + cb) * 2 + h] = (unsigned short)my_word;
    }
    __syncthreads();
    NMS_TR(6);
    // 4d. resolve: warp 0, one 32-candidate block per step.  The 32 diagonal words are loaded up
    // front (broadcast reads, independent of the greedy chain), so the chain itself is 32
    // register-only steps; the rows of the kept candidates are then OR-ed into the later blocks
    // with predicated, independent loads (lane = mask word).
    if (warp == 0) {
      uint32_t removed = (lane < CHUNK_WORDS) ? s_removed[lane] : 0xffffffffu;
      int kept_chunk = 0;
      for (int wb = 0; wb < nblk; ++wb) {
        const uint32_t cur = __shfl_sync(0xffffffffu, removed, wb);
        uint32_t alive = ~cur;
        const uint32_t* rows = mask + (wb << 5) * CHUNK_WORDS;
        // every load of the block is issued before the chain starts: the diagonal words (broadcast) and
        // this lane's word of each of the 32 rows, whose use is then a register select
        uint32_t d[32], mine_w[32];
        const int my_col = (lane < CHUNK_WORDS && lane > wb) ? lane : wb;
#pragma unroll
        for (int r = 0; r < 32; ++r) {
          d[r] = rows[r * CHUNK_WORDS + wb];
          mine_w[r] = rows[r * CHUNK_WORDS + my_col];
        }
#pragma unroll
        for (int r = 0; r < 32; ++r)
          if ((alive >> r) & 1u) alive &= ~d[r];
        if (lane < CHUNK_WORDS && lane > wb) {
#pragma unroll
          for (int r = 0; r < 32; ++r)
            if ((alive >> r) & 1u) removed |= mine_w[r];
        }
        if ((alive >> lane) & 1u)
          s_kept_pos[kept_chunk + __popc(alive & ((1u << lane) - 1u))] = (wb << 5) + lane;
        kept_chunk += __popc(alive);
      }
      if (lane == 0) s_kept_chunk = kept_chunk;
    }
    __syncthreads();
    NMS_TR(7);
    // 4e. emit the survivors of this chunk (kept order == score order)
    const int kept_chunk = s_kept_chunk;
    for (int e = tid; e < kept_chunk; e += NMS_THREADS) {
      const int pos = s_kept_pos[e];
      const int k = kept_before + e;
      const float4 v = s_box[pos];
      kept_boxes[k] = v;
      kept_cls[k] = s_cls[pos];
      if (k < p.max_det) {
        const size_t o = (size_t)b * p.max_det + k;
        const int a = s_anchor[pos];
        reinterpret_cast<float4*>(p.out_boxes)[o] = v;
        if (p.out_scores) p.out_scores[o] = scores[a];
        if (p.out_classes) p.out_classes[o] = classes ? s_cls[pos] : 0;
        if (p.out_anchor) p.out_anchor[o] = a;
        if (p.out_keep)
          p.out_keep[o] = resident ? s_prefix[a >> 5] + __popc(mask_word(a >> 5) & ((1u << (a & 31)) - 1u))
                                   : prefix[a >> 5] + __popc(sel[a >> 5] & ((1u << (a & 31)) - 1u));
      }
    }
    __syncthreads();
    if (tid == 0) s_kept_total = kept_before + kept_chunk;
    __syncthreads();
  }
  NMS_TR(8);
  if (tid == 0) p.out_count[b] = min(s_kept_total, p.max_det);
}

int next_pow2_cap(int64_t anchors) {
  int p = 32;
  while (p < anchors) p <<= 1;
  return p;
}

}  // namespace
}  // namespace ovdet

extern "C" size_t ovdet_nms_workspace_bytes(int64_t batch, int64_t anchors) {
  using namespace ovdet;
  if (batch <= 0 || anchors <= 0 || anchors >= (1ll << 30)) return 0;
  const WsLayout L((int)anchors, (int)((anchors + 31) / 32), next_pow2_cap(anchors));
  return (size_t)batch * L.total;
}

int ovdet_nms_launch_internal(int pdl, int use_conf, float conf, const float* boxes, const float* scores, const int32_t* classes,
                                 const uint32_t* pass_mask, int64_t batch, int64_t anchors,
                                 const float* scale, const float* clip_wh, float iou_thr,
                                 int class_aware, int topk, int64_t max_det, float* out_boxes,
                                 float* out_scores, int32_t* out_classes, int32_t* out_anchor,
                                 int32_t* out_keep, int32_t* out_count, int32_t* out_candidates,
                                 void* workspace, size_t workspace_bytes, void* stream) {
  using namespace ovdet;
  if (batch == 0) return check_device();            // an empty batch is a no-op (its pointers may be null)
  if (!boxes || !scores || !out_boxes || !out_count || batch < 0 || anchors < 0 || max_det <= 0 || topk < 0)
    return OVDET_ERR_INVALID_ARG;
  if (class_aware && !classes) return OVDET_ERR_INVALID_ARG;
  if (anchors >= (1ll << 30) || max_det >= (1ll << 30) || batch >= (1ll << 31)) return OVDET_ERR_UNSUPPORTED_SHAPE;
  if (((uintptr_t)boxes & 15) || ((uintptr_t)out_boxes & 15)) return OVDET_ERR_INVALID_ARG;
  if (int rc = check_device()) return rc;
  if (batch == 0) return OVDET_OK;
  if (anchors == 0) {
    OVDET_CUDA_TRY(cudaMemsetAsync(out_count, 0, sizeof(int32_t) * batch, as_stream(stream)));
    if (out_candidates) OVDET_CUDA_TRY(cudaMemsetAsync(out_candidates, 0, sizeof(int32_t) * batch, as_stream(stream)));
    return OVDET_OK;
  }
  const size_t need = ovdet_nms_workspace_bytes(batch, anchors);
  if (!workspace || workspace_bytes < need || ((uintptr_t)workspace & 15)) return OVDET_ERR_WORKSPACE;
  NmsParams p{};
  p.boxes = boxes; p.scores = scores; p.classes = classes; p.pass_mask = pass_mask;
  p.use_conf = use_conf; p.conf = conf; p.pdl = pdl;
  if (use_conf && (pass_mask || (anchors + 31) / 32 > ovdet::FAST_WORDS)) return OVDET_ERR_UNSUPPORTED_SHAPE;
  p.anchors = (int)anchors; p.words = (int)((anchors + 31) / 32);
  p.scale = scale; p.clip_wh = clip_wh; p.iou_thr = iou_thr;
  p.class_aware = class_aware; p.topk = topk; p.max_det = (int)max_det;
  p.out_boxes = out_boxes; p.out_scores = out_scores; p.out_classes = out_classes;
  p.out_anchor = out_anchor; p.out_keep = out_keep; p.out_count = out_count;
  p.out_candidates = out_candidates;
  p.ws = static_cast<unsigned char*>(workspace);
  p.pow2_cap = next_pow2_cap(anchors);
  p.ws_per_image = WsLayout(p.anchors, p.words, p.pow2_cap).total;
  if (int rc = once_per_device(2, []() -> int {
        OVDET_CUDA_TRY(cudaFuncSetAttribute(nms_batched_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, NMS_SMEM_BYTES));
        return OVDET_OK;
      })) return rc;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)batch);
  cfg.blockDim = dim3(NMS_THREADS);
  cfg.dynamicSmemBytes = NMS_SMEM_BYTES;
  cfg.stream = as_stream(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  OVDET_CUDA_TRY(cudaLaunchKernelEx(&cfg, nms_batched_kernel, p));
  return OVDET_OK;
}

extern "C" int ovdet_nms_batched(const float* boxes, const float* scores, const int32_t* classes,
                                 const uint32_t* pass_mask, int64_t batch, int64_t anchors,
                                 const float* scale, const float* clip_wh, float iou_thr,
                                 int class_aware, int topk, int64_t max_det, float* out_boxes,
                                 float* out_scores, int32_t* out_classes, int32_t* out_anchor,
                                 int32_t* out_keep, int32_t* out_count, int32_t* out_candidates,
                                 void* workspace, size_t workspace_bytes, void* stream) {
  return ovdet_nms_launch_internal(0, 0, 0.f, boxes, scores, classes, pass_mask, batch, anchors, scale, clip_wh, iou_thr,
                    class_aware, topk, max_det, out_boxes, out_scores, out_classes, out_anchor, out_keep,
                    out_count, out_candidates, workspace, workspace_bytes, stream);
}

extern "C" int ovdet_nms_batched_conf(const float* boxes, const float* scores, const int32_t* classes,
                                      float conf, int64_t batch, int64_t anchors, const float* scale,
                                      const float* clip_wh, float iou_thr, int class_aware, int topk,
                                      int64_t max_det, float* out_boxes, float* out_scores,
                                      int32_t* out_classes, int32_t* out_anchor, int32_t* out_keep,
                                      int32_t* out_count, int32_t* out_candidates, void* workspace,
                                      size_t workspace_bytes, void* stream) {
  return ovdet_nms_launch_internal(0, 1, conf, boxes, scores, classes, nullptr, batch, anchors, scale, clip_wh, iou_thr,
                    class_aware, topk, max_det, out_boxes, out_scores, out_classes, out_anchor, out_keep,
                    out_count, out_candidates, workspace, workspace_bytes, stream);
}
