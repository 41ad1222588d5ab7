// K2b: max / argmax over the class axis of materialised logits.
// Replaces similarity.max(dim=1) (model/yolo_clip.py:198-202); ties -> lowest class index.
// HBM-bound: every logit is read exactly once, one warp per anchor row, coalesced.
#include "common.cuh"

namespace ovdet {

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T>
__global__ void __launch_bounds__(256)
rowmax_kernel(const T* __restrict__ logits, int64_t rows, int classes, int64_t ldc,
              float* __restrict__ row_max, int* __restrict__ row_arg) {
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const T* src = logits + row * ldc;
  float best = -INFINITY;
  int idx = 0x7fffffff;
  int c = lane;
  for (; c + 96 < classes; c += 128) {          // four independent loads in flight
    const float v0 = to_f32(src[c]), v1 = to_f32(src[c + 32]);
    const float v2 = to_f32(src[c + 64]), v3 = to_f32(src[c + 96]);
    if (v0 > best) { best = v0; idx = c; }
    if (v1 > best) { best = v1; idx = c + 32; }
    if (v2 > best) { best = v2; idx = c + 64; }
    if (v3 > best) { best = v3; idx = c + 96; }
  }
  for (; c < classes; c += 32) {
    const float v = to_f32(src[c]);
    if (v > best) { best = v; idx = c; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (ob > best || (ob == best && oi < idx)) { best = ob; idx = oi; }
  }
  if (lane == 0) {
    row_max[row] = best;
    if (row_arg) row_arg[row] = (idx == 0x7fffffff) ? 0 : idx;
  }
}

}  // namespace ovdet

extern "C" int ovdet_rowmax(const void* logits, int logits_dtype, int64_t rows, int64_t classes,
                            int64_t ldc, float* row_max, int32_t* row_arg, void* stream) {
  using namespace ovdet;
  if (!logits || !row_max || rows < 0 || classes <= 0 || ldc < classes) return OVDET_ERR_INVALID_ARG;
  if (classes >= (1ll << 31)) return OVDET_ERR_UNSUPPORTED_SHAPE;
  if (int rc = check_device()) return rc;
  if (rows == 0) return OVDET_OK;
  const unsigned grid = (unsigned)ceil_div<int64_t>(rows, 8);
  if (logits_dtype == OVDET_F32)
    rowmax_kernel<float><<<grid, 256, 0, as_stream(stream)>>>(static_cast<const float*>(logits), rows,
                                                             (int)classes, ldc, row_max, row_arg);
  else if (logits_dtype == OVDET_BF16)
    rowmax_kernel<__nv_bfloat16><<<grid, 256, 0, as_stream(stream)>>>(
        static_cast<const __nv_bfloat16*>(logits), rows, (int)classes, ldc, row_max, row_arg);
  else
    return OVDET_ERR_INVALID_ARG;
  OVDET_LAUNCH_CHECK();
  return OVDET_OK;
}
