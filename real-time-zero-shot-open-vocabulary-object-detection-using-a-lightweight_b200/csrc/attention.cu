// N1: max-sigmoid text attention of the neck's text-guided CSP layer ("next" row f-1).
// Replaces model/repvl_pan.py:77-95:
//   scores = y^T t'   ([B, HW, c] x [B, C, c]^T),  w = sigmoid(max_c scores),  out = y * w
// The score GEMM + class max reuses the fused tcgen05 kernel (fp32 NCHW activations read by TMA,
// A operand in tensor memory, row max in the epilogue; raw dot products: normalize = 0).  With
// `precise` the product is accumulated from three bf16 passes and matches the fp32 reference to
// ~1e-6; without it one bf16 pass (|dscore| ~ 4e-3 |y||t|).  A streaming kernel then scales the
// activations: 2 x 4 bytes per element, 16-byte accesses along the spatial axis.
#include "common.cuh"

namespace ovdet {
namespace {

__global__ void __launch_bounds__(256)
scale_by_sigmoid_kernel(const float* __restrict__ y, const float* __restrict__ row_max,
                        float* __restrict__ out, int channels, int hw, long long stride_b,
                        long long stride_c, long long out_stride_b, long long out_stride_c) {
  const int a = (blockIdx.x * 256 + threadIdx.x) * 4;
  if (a >= hw) return;
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * 8;
  const float4 m = *reinterpret_cast<const float4*>(row_max + (long long)b * hw + a);
  float4 w;
  w.x = __fdiv_rn(1.0f, 1.0f + expf(-m.x));
  w.y = __fdiv_rn(1.0f, 1.0f + expf(-m.y));
  w.z = __fdiv_rn(1.0f, 1.0f + expf(-m.z));
  w.w = __fdiv_rn(1.0f, 1.0f + expf(-m.w));
  float4 v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (c0 + i < channels)
      v[i] = ld_stream_f32x4(reinterpret_cast<const float4*>(y + b * stride_b + (c0 + i) * stride_c + a));
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (c0 + i < channels) {
      float4 o = make_float4(v[i].x * w.x, v[i].y * w.y, v[i].z * w.z, v[i].w * w.w);
      *reinterpret_cast<float4*>(out + b * out_stride_b + (c0 + i) * out_stride_c + a) = o;
    }
}

}  // namespace
}  // namespace ovdet

extern "C" int ovdet_max_sigmoid_attention(const float* y, int64_t batch, int64_t channels, int64_t hw,
                                           int64_t stride_b, int64_t stride_c, const void* text_op,
                                           int64_t classes, int text_batched, int precise,
                                           float* row_max, float* out, int64_t out_stride_b,
                                           int64_t out_stride_c, void* stream) {
  using namespace ovdet;
  if (!y || !text_op || !row_max || !out || batch < 0 || channels <= 0 || hw <= 0 || classes <= 0)
    return OVDET_ERR_INVALID_ARG;
  // 16-byte accesses along the spatial axis (the TMA of the fused kernel needs the same)
  if ((hw & 3) || (stride_b & 3) || (stride_c & 3) || (out_stride_b & 3) || (out_stride_c & 3) ||
      ((uintptr_t)y & 15) || ((uintptr_t)out & 15) || ((uintptr_t)row_max & 15) || batch > 65535)
    return OVDET_ERR_UNSUPPORTED_SHAPE;
  const float* levels[1] = {y};
  const int64_t hws[1] = {hw}, sb[1] = {stride_b}, sc[1] = {stride_c};
  int rc = fused_launch(levels, hws, sb, sc, 1, batch, channels, text_op, nullptr, classes, text_batched,
                        /*normalize=*/0, precise ? 1 : 0, 1.0f, 0.0f, nullptr, OVDET_F32, classes,
                        row_max, nullptr, nullptr, stream);
  if (rc != OVDET_OK || batch == 0) return rc;
  dim3 grid((unsigned)ceil_div<int64_t>(hw, 1024), (unsigned)ceil_div<int64_t>(channels, 8), (unsigned)batch);
  scale_by_sigmoid_kernel<<<grid, 256, 0, as_stream(stream)>>>(y, row_max, out, (int)channels, (int)hw,
                                                               stride_b, stride_c, out_stride_b, out_stride_c);
  OVDET_LAUNCH_CHECK();
  return OVDET_OK;
}
