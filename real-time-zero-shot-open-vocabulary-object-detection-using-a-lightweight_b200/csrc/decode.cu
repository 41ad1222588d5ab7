// K3: DFL box decode for all pyramid levels, fused with score activation and the confidence
// threshold.  Replaces BoxHead.decode_boxes (model/heads/box_head.py:150-218; the int64 grid of
// :115-148 is computed from the thread index) and `scores > conf` (inference/detector.py:184).
//
// HBM-bound streaming kernel: one thread per anchor reads its 4 x bins logits (coalesced along
// the spatial axis - consecutive threads are consecutive cells), 272 B in, 16 B (+ a bit) out.
// The float32 operation order follows the reference: softmax = exp(x - max) / sum, expectation
// sum_k p_k * k (separate multiply and add, no FMA contraction), centre = (cell + e) * stride,
// size = exp(e) * stride, xyxy = centre -/+ size / 2.
#include "common.cuh"

namespace ovdet {

struct DecodeParams {
  const float* pred[OVDET_MAX_LEVELS];
  long long bstride[OVDET_MAX_LEVELS];
  int h[OVDET_MAX_LEVELS], w[OVDET_MAX_LEVELS], stride[OVDET_MAX_LEVELS];
  int off[OVDET_MAX_LEVELS + 1];       // anchor offset of each level; off[levels] = anchors
  int levels;
  int bins;
  float wscale, hscale;
};

template <int BINS>
__device__ __forceinline__ float dfl_expectation(const float* __restrict__ p, long long cstride,
                                                 int bins_rt) {
  if (BINS > 0) {
    float v[BINS > 0 ? BINS : 1];
#pragma unroll
    for (int k = 0; k < BINS; ++k) v[k] = ld_stream_f32(p + k * cstride);
    float m = v[0];
#pragma unroll
    for (int k = 1; k < BINS; ++k) m = fmaxf(m, v[k]);
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < BINS; ++k) { v[k] = expf(v[k] - m); s = __fadd_rn(s, v[k]); }
    float e = 0.f;
#pragma unroll
    for (int k = 0; k < BINS; ++k) e = __fadd_rn(e, __fmul_rn(__fdiv_rn(v[k], s), (float)k));
    return e;
  } else {
    float m = -INFINITY;
    for (int k = 0; k < bins_rt; ++k) m = fmaxf(m, p[k * cstride]);
    float s = 0.f;
    for (int k = 0; k < bins_rt; ++k) s = __fadd_rn(s, expf(p[k * cstride] - m));
    float e = 0.f;
    for (int k = 0; k < bins_rt; ++k)
      e = __fadd_rn(e, __fmul_rn(__fdiv_rn(expf(p[k * cstride] - m), s), (float)k));
    return e;
  }
}

template <int BINS>
__global__ void __launch_bounds__(256)
decode_filter_kernel(const DecodeParams p, int anchors, const float* __restrict__ scores,
                     float conf, int activation, float* __restrict__ boxes,
                     float* __restrict__ scores_act, uint32_t* __restrict__ pass_mask, int words) {
  const int a = blockIdx.x * 256 + threadIdx.x;
  const int b = blockIdx.y;
  bool pass = false;
  if (a < anchors) {
    int l = 0;
#pragma unroll
    for (int i = 1; i < OVDET_MAX_LEVELS; ++i)
      if (i < p.levels && a >= p.off[i]) l = i;
    const int cell = a - p.off[l];
    const int wdt = p.w[l];
    const int gy = cell / wdt, gx = cell - gy * wdt;
    const long long cstride = (long long)p.h[l] * wdt;
    const float* base = p.pred[l] + b * p.bstride[l] + cell;
    const int bins = p.bins;
    const float e0 = dfl_expectation<BINS>(base, cstride, bins);
    const float e1 = dfl_expectation<BINS>(base + 1ll * bins * cstride, cstride, bins);
    const float e2 = dfl_expectation<BINS>(base + 2ll * bins * cstride, cstride, bins);
    const float e3 = dfl_expectation<BINS>(base + 3ll * bins * cstride, cstride, bins);
    const float st = (float)p.stride[l];
    const float cx = __fmul_rn(__fadd_rn((float)gx, e0), st);
    const float cy = __fmul_rn(__fadd_rn((float)gy, e1), st);
    const float bw = __fmul_rn(__fmul_rn(expf(e2), st), p.wscale);
    const float bh = __fmul_rn(__fmul_rn(expf(e3), st), p.hscale);
    const float hw_ = __fmul_rn(bw, 0.5f), hh_ = __fmul_rn(bh, 0.5f);
    const long long ga = (long long)b * anchors + a;
    if (boxes != nullptr)
      reinterpret_cast<float4*>(boxes)[ga] =
          make_float4(__fsub_rn(cx, hw_), __fsub_rn(cy, hh_), __fadd_rn(cx, hw_), __fadd_rn(cy, hh_));
    if (scores != nullptr) {
      float s = scores[ga];
      if (activation == OVDET_ACT_SIGMOID) s = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-s)));
      if (scores_act != nullptr) scores_act[ga] = s;
      pass = s > conf;
    }
  }
  if (pass_mask != nullptr) {
    const uint32_t bits = __ballot_sync(0xffffffffu, pass);
    const int word = a >> 5;
    if ((threadIdx.x & 31) == 0 && word < words) pass_mask[(long long)b * words + word] = bits;
  }
}

}  // namespace ovdet

extern "C" int ovdet_decode_filter(const float* const* box_preds, const int32_t* heights,
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
  if (batch == 0) return OVDET_OK;
  const int anchors = (int)total;
  const int words = (anchors + 31) / 32;
  dim3 grid((unsigned)ceil_div(anchors, 256), (unsigned)batch);
  cudaStream_t s = as_stream(stream);
  if (bins == 17)
    decode_filter_kernel<17><<<grid, 256, 0, s>>>(p, anchors, scores, conf, activation, boxes,
                                                  scores_act, pass_mask, words);
  else
    decode_filter_kernel<0><<<grid, 256, 0, s>>>(p, anchors, scores, conf, activation, boxes,
                                                 scores_act, pass_mask, words);
  OVDET_LAUNCH_CHECK();
  return OVDET_OK;
}
