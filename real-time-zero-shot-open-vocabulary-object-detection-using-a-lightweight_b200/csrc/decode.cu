// K3: DFL box decode for all pyramid levels, fused with score activation and the confidence
// threshold.  Replaces BoxHead.decode_boxes (model/heads/box_head.py:150-218; the int64 grid of
// :115-148 is computed from the thread index) and `scores > conf` (inference/detector.py:184).
//
// HBM-bound streaming kernel: a thread decodes 4 consecutive cells of a level with 16-byte
// loads (coalesced along the spatial axis: a warp reads 512 contiguous bytes per bin), 272 B in,
// 16 B (+ a bit) out per anchor.  softmax expectation = sum_k k exp(x_k - max) / sum_k exp(x_k -
// max); the decode proper follows the reference's float32 operation order: centre = (cell + e)
// * stride, size = exp(e) * stride, xyxy = centre -/+ size / 2, no FMA contraction.
#include "common.cuh"

namespace ovdet {

struct DecodeParams {
  const void* pred[OVDET_MAX_LEVELS];      // fp32 or bf16 (template parameter T of the kernel)
  long long bstride[OVDET_MAX_LEVELS];
  int h[OVDET_MAX_LEVELS], w[OVDET_MAX_LEVELS], stride[OVDET_MAX_LEVELS];
  int off[OVDET_MAX_LEVELS + 1];       // anchor offset of each level; off[levels] = anchors
  int levels;
  int bins;
  float wscale, hscale;
  // programmatic dependent launch (ovdet_head_step only): the kernel was launched while the
  // similarity kernel may still be running; that kernel produces `scores` and nothing else this
  // one reads, so the grid dependency is awaited just before the scores are loaded and the box
  // decode overlaps the similarity kernel's tail
  int pdl;
};

// Expectation of the softmax over the bins of one coordinate, for V consecutive anchors at once:
//   E = sum_k k * exp(x_k - m) / sum_k exp(x_k - m)
// (one division per coordinate instead of the reference's one per bin; |dE| is a few 1e-7, which
// the box tolerance of 1e-4 relative absorbs - box_head.py:185-192).
// exp(x - m) as ex2.approx(x * log2(e) - m * log2(e)): one FFMA + one MUFU instead of expf's
// ~10 instructions.  The rounding of m * log2(e) scales every bin by the same factor (it cancels
// in the ratio); the FFMA's own rounding bounds the relative error of a term by |x - m| * 1e-7
// (< 1e-6 for every term that carries weight), far inside the box tolerance.
constexpr float kLog2e = 1.4426950408889634f;
__device__ __forceinline__ float exp_rel(float x, float neg_m_log2e) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(fmaf(x, kLog2e, neg_m_log2e)));
  return r;
}

// Typed streaming loads: the box logits are fp32 (the reference's convolutions) or bf16 (the head ran
// under autocast); the arithmetic is fp32 either way.
__device__ __forceinline__ float ld1(const float* p) { return ld_stream_f32(p); }
__device__ __forceinline__ float ld1(const __nv_bfloat16* p) {
  unsigned short u;
  asm volatile("ld.global.nc.L1::no_allocate.u16 %0, [%1];" : "=h"(u) : "l"(p));
  return __uint_as_float((uint32_t)u << 16);
}
__device__ __forceinline__ float4 ld4(const float* p) { return ld_stream_f32x4(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ld4(const __nv_bfloat16* p) {                 // 8-byte aligned
  uint32_t a, b;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(a), "=r"(b) : "l"(p));
  return make_float4(__uint_as_float(a << 16), __uint_as_float(a & 0xffff0000u),
                     __uint_as_float(b << 16), __uint_as_float(b & 0xffff0000u));
}

template <int BINS, int V, typename T>
__device__ __forceinline__ void dfl_expectation(const T* __restrict__ p, long long cstride,
                                                int bins_rt, float (&e)[V]) {
  if (BINS > 0) {
    float v[BINS > 0 ? BINS : 1][V];
#pragma unroll
    for (int k = 0; k < BINS; ++k) {
      if (V == 4) {
        const float4 q = ld4(p + k * cstride);
        v[k][0] = q.x; v[k][1 % V] = q.y; v[k][2 % V] = q.z; v[k][3 % V] = q.w;
      } else {
        v[k][0] = ld1(p + k * cstride);
      }
    }
#pragma unroll
    for (int j = 0; j < V; ++j) {
      float m = v[0][j];
#pragma unroll
      for (int k = 1; k < BINS; ++k) m = fmaxf(m, v[k][j]);
      const float nm = -m * kLog2e;
      float s = 0.f, n = 0.f;
#pragma unroll
      for (int k = 0; k < BINS; ++k) {
        const float t = exp_rel(v[k][j], nm);
        s += t;
        n = fmaf((float)k, t, n);
      }
      e[j] = __fdiv_rn(n, s);
    }
  } else {
#pragma unroll
    for (int j = 0; j < V; ++j) {
      float m = -INFINITY;
      for (int k = 0; k < bins_rt; ++k) m = fmaxf(m, ld1(p + k * cstride + j));
      const float nm = -m * kLog2e;
      float s = 0.f, n = 0.f;
      for (int k = 0; k < bins_rt; ++k) {
        const float t = exp_rel(ld1(p + k * cstride + j), nm);
        s += t;
        n = fmaf((float)k, t, n);
      }
      e[j] = __fdiv_rn(n, s);
    }
  }
}

// Latency variant for small launches (batch 1): one anchor per thread, all 4 x BINS loads issued
// before any arithmetic, so the kernel pays one memory round trip instead of four.
template <int BINS, typename T>
__device__ __forceinline__ void dfl_expectation4_hoisted(const T* __restrict__ p, long long cstride,
                                                         float (&e)[4]) {
  float v[4][BINS];
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int k = 0; k < BINS; ++k) v[c][k] = ld1(p + (long long)(c * BINS + k) * cstride);
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    float m = v[c][0];
#pragma unroll
    for (int k = 1; k < BINS; ++k) m = fmaxf(m, v[c][k]);
    const float nm = -m * kLog2e;
    float s = 0.f, n = 0.f;
#pragma unroll
    for (int k = 0; k < BINS; ++k) {
      const float t = exp_rel(v[c][k], nm);
      s += t;
      n = fmaf((float)k, t, n);
    }
    e[c] = __fdiv_rn(n, s);
  }
}

// V = 4: a thread decodes 4 consecutive cells of one level with 16-byte loads (every level's
// H*W and batch stride must be a multiple of 4 and the pointers 16-byte aligned); V = 1 is the
// general path.  blockIdx.y = image.
template <int BINS, int V, bool HOIST = false, typename T = float>
__global__ void __launch_bounds__(256)
decode_filter_kernel(const DecodeParams p, int anchors, const float* scores,
                     float conf, int activation, float* __restrict__ boxes,
                     float* scores_act, uint32_t* __restrict__ pass_mask, int words) {
  const int a = (blockIdx.x * 256 + threadIdx.x) * V;       // first anchor of this thread
  const int b = blockIdx.y;
  uint32_t pass = 0;                                         // bit j: anchor a + j passes
  if (a < anchors) {
    int l = 0;
#pragma unroll
    for (int i = 1; i < OVDET_MAX_LEVELS; ++i)
      if (i < p.levels && a >= p.off[i]) l = i;
    const int cell = a - p.off[l];
    const int wdt = p.w[l];
    const long long cstride = (long long)p.h[l] * wdt;
    const T* base = static_cast<const T*>(p.pred[l]) + b * p.bstride[l] + cell;
    const int bins = p.bins;
    float e0[V], e1[V], e2[V], e3[V];
    if constexpr (HOIST && V == 1 && BINS > 0) {
      float e[4];
      dfl_expectation4_hoisted<BINS, T>(base, cstride, e);
      e0[0] = e[0]; e1[0] = e[1]; e2[0] = e[2]; e3[0] = e[3];
    } else {
      dfl_expectation<BINS, V, T>(base, cstride, bins, e0);
      dfl_expectation<BINS, V, T>(base + 1ll * bins * cstride, cstride, bins, e1);
      dfl_expectation<BINS, V, T>(base + 2ll * bins * cstride, cstride, bins, e2);
      dfl_expectation<BINS, V, T>(base + 3ll * bins * cstride, cstride, bins, e3);
    }
    const float st = (float)p.stride[l];
    const long long ga = (long long)b * anchors + a;
    float sc[V];
    if (p.pdl) cudaGridDependencySynchronize();
    if (scores != nullptr) {
      // L2-coherent loads (ld.global.cg), never the read-only path: under PDL this kernel is already
      // running while the similarity kernel writes the scores, so they are not read-only for its
      // lifetime and an ld.global.nc may hit an L1 line left by the previous step (measured: stale
      // pass masks in 299 of 300 steps with LDG.CONSTANT here)
      if (V == 4) {
        const float4 q = __ldcg(reinterpret_cast<const float4*>(scores + ga));
        sc[0] = q.x; sc[1 % V] = q.y; sc[2 % V] = q.z; sc[3 % V] = q.w;
      } else {
        sc[0] = __ldcg(scores + ga);
      }
    }
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const int c = cell + j;
      const int gy = c / wdt, gx = c - gy * wdt;
      const float cx = __fmul_rn(__fadd_rn((float)gx, e0[j]), st);
      const float cy = __fmul_rn(__fadd_rn((float)gy, e1[j]), st);
      const float bw = __fmul_rn(__fmul_rn(expf(e2[j]), st), p.wscale);
      const float bh = __fmul_rn(__fmul_rn(expf(e3[j]), st), p.hscale);
      const float hw_ = __fmul_rn(bw, 0.5f), hh_ = __fmul_rn(bh, 0.5f);
      if (boxes != nullptr)
        reinterpret_cast<float4*>(boxes)[ga + j] =
            make_float4(__fsub_rn(cx, hw_), __fsub_rn(cy, hh_), __fadd_rn(cx, hw_), __fadd_rn(cy, hh_));
      if (scores != nullptr) {
        float s = sc[j];
        if (activation == OVDET_ACT_SIGMOID) s = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-s)));
        sc[j] = s;
        if (s > conf) pass |= 1u << j;
      }
    }
    if (scores != nullptr && scores_act != nullptr) {
      if (V == 4) *reinterpret_cast<float4*>(scores_act + ga) = make_float4(sc[0], sc[1 % V], sc[2 % V], sc[3 % V]);
      else scores_act[ga] = sc[0];
    }
  }
  if (pass_mask != nullptr) {
    const int lane = threadIdx.x & 31;
    if (V == 1) {
      const uint32_t bits = __ballot_sync(0xffffffffu, pass != 0);
      const int word = a >> 5;
      if (lane == 0 && word < words) pass_mask[(long long)b * words + word] = bits;
    } else {
      // 8 lanes x 4 anchors make one 32-bit word
      uint32_t v = pass << (4 * (lane & 7));
      v |= __shfl_xor_sync(0xffffffffu, v, 1);
      v |= __shfl_xor_sync(0xffffffffu, v, 2);
      v |= __shfl_xor_sync(0xffffffffu, v, 4);
      const int word = a >> 5;
      if ((lane & 7) == 0 && word < words) pass_mask[(long long)b * words + word] = v;
    }
  }
}

}  // namespace ovdet

int ovdet_decode_launch_internal(int pdl, int in_bf16, const void* const* box_preds, const int32_t* heights,
                                   const int32_t* widths, const int32_t* strides,
                                   const int64_t* batch_strides, int num_levels, int bins,
                                   int64_t batch, float width_scale, float height_scale,
                                   const float* scores, float conf, int activation,
                                   float* boxes, float* scores_act, uint32_t* pass_mask,
                                   void* stream) {
  using namespace ovdet;
  if (!box_preds || !heights || !widths || !strides || !batch_strides) return OVDET_ERR_INVALID_ARG;
  if (num_levels <= 0 || bins <= 0 || batch < 0) return OVDET_ERR_INVALID_ARG;
  if (num_levels > OVDET_MAX_LEVELS || batch > 65535) return OVDET_ERR_UNSUPPORTED_SHAPE;
  if (activation != OVDET_ACT_NONE && activation != OVDET_ACT_SIGMOID) return OVDET_ERR_INVALID_ARG;
  if (pass_mask && !scores) return OVDET_ERR_INVALID_ARG;
  if (boxes && ((uintptr_t)boxes & 15)) return OVDET_ERR_INVALID_ARG;
  if (int rc = check_device()) return rc;
  if (batch == 0) return OVDET_OK;                  // an empty batch is a no-op (its pointers may be null)
  DecodeParams p{};
  long long total = 0;
  for (int l = 0; l < num_levels; ++l) {
    if (!box_preds[l] || heights[l] <= 0 || widths[l] <= 0) return OVDET_ERR_INVALID_ARG;
    p.pred[l] = box_preds[l];
    p.bstride[l] = batch_strides[l];
    p.h[l] = heights[l];
    p.w[l] = widths[l];
    p.stride[l] = strides[l];
    p.off[l] = (int)total;
    total += (long long)heights[l] * widths[l];
  }
  if (total >= (1ll << 30)) return OVDET_ERR_UNSUPPORTED_SHAPE;
  p.off[num_levels] = (int)total;
  p.levels = num_levels;
  p.bins = bins;
  p.wscale = width_scale;
  p.hscale = height_scale;
  p.pdl = pdl;
  if (batch == 0) return OVDET_OK;
  const int anchors = (int)total;
  const int words = (anchors + 31) / 32;
  cudaStream_t s = as_stream(stream);
  // every variant goes through one launcher so that the PDL attribute can be attached
  auto launch = [&](auto kern, dim3 grid) -> cudaError_t {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, p, anchors, scores, conf, activation, boxes, scores_act, pass_mask, words);
  };
  bool vec4 = bins == 17 && !(((uintptr_t)scores | (uintptr_t)scores_act) & 15);
  for (int l = 0; l < num_levels && vec4; ++l)
    vec4 = ((long long)heights[l] * widths[l]) % 4 == 0 && batch_strides[l] % 4 == 0 &&
           !((uintptr_t)box_preds[l] & (in_bf16 ? 7 : 15));
  if (bins == 17 && (long long)anchors * batch <= 65536) {
    // small launch: latency matters, not bandwidth
    dim3 grid((unsigned)ceil_div(anchors, 256), (unsigned)batch);
    if (in_bf16)
      OVDET_CUDA_TRY(launch(decode_filter_kernel<17, 1, true, __nv_bfloat16>, grid));
    else
      OVDET_CUDA_TRY(launch(decode_filter_kernel<17, 1, true>, grid));
  } else if (vec4) {
    dim3 grid((unsigned)ceil_div(anchors, 1024), (unsigned)batch);
    if (in_bf16)
      OVDET_CUDA_TRY(launch(decode_filter_kernel<17, 4, false, __nv_bfloat16>, grid));
    else
      OVDET_CUDA_TRY(launch(decode_filter_kernel<17, 4>, grid));
  } else {
    dim3 grid((unsigned)ceil_div(anchors, 256), (unsigned)batch);
    if (in_bf16)
      OVDET_CUDA_TRY(launch(decode_filter_kernel<0, 1, false, __nv_bfloat16>, grid));
    else if (bins == 17)
      OVDET_CUDA_TRY(launch(decode_filter_kernel<17, 1>, grid));
    else
      OVDET_CUDA_TRY(launch(decode_filter_kernel<0, 1>, grid));
  }
  OVDET_LAUNCH_CHECK();
  return OVDET_OK;
}

extern "C" int ovdet_decode_filter(const float* const* box_preds, const int32_t* heights,
                                   const int32_t* widths, const int32_t* strides,
                                   const int64_t* batch_strides, int num_levels, int bins,
                                   int64_t batch, float width_scale, float height_scale,
                                   const float* scores, float conf, int activation,
                                   float* boxes, float* scores_act, uint32_t* pass_mask,
                                   void* stream) {
  return ovdet_decode_launch_internal(0, 0, reinterpret_cast<const void* const*>(box_preds), heights, widths, strides,
                       batch_strides, num_levels, bins, batch, width_scale, height_scale, scores, conf,
                       activation, boxes, scores_act, pass_mask, stream);
}

extern "C" int ovdet_decode_filter_bf16in(const void* const* box_preds, const int32_t* heights,
                                          const int32_t* widths, const int32_t* strides,
                                          const int64_t* batch_strides, int num_levels, int bins,
                                          int64_t batch, float width_scale, float height_scale,
                                          const float* scores, float conf, int activation,
                                          float* boxes, float* scores_act, uint32_t* pass_mask,
                                          void* stream) {
  return ovdet_decode_launch_internal(0, 1, box_preds, heights, widths, strides, batch_strides, num_levels, bins, batch,
                       width_scale, height_scale, scores, conf, activation, boxes, scores_act, pass_mask,
                       stream);
}
