// E1: the `obj_embeddings` entry of the forward dict.
//
// Replaces model/yolo_clip.py:208-214: per level `embed.permute(0, 2, 3, 1).reshape(B, HW, D)`
// followed by `torch.cat(..., dim=1)` - a strided-view copy per level plus one concat copy in
// the reference; here every element is read once (NCHW, anchors contiguous) and written once
// (anchor-major [B, A, D], levels concatenated P3 | P4 | P5) by one launch per level.
// HBM-bound: 8 bytes per element.  A CTA turns a 64 (channels) x 64 (anchors) tile through shared
// memory; both the global reads (256 B per warp instruction along HW) and the global writes
// (256 B per warp instruction along D) are coalesced, the 65-float pitch keeps the column reads
// bank-conflict free.
#include "common.cuh"

namespace ovdet {

constexpr int E_TILE = 64;

__global__ void __launch_bounds__(256)
concat_embeddings_kernel(const float* __restrict__ x, int dim, int hw, int64_t stride_b, int64_t stride_d,
                         float* __restrict__ out, int64_t rows_per_batch, int64_t row_offset) {
  __shared__ float tile[E_TILE][E_TILE + 1];
  const int a0 = blockIdx.x * E_TILE, d0 = blockIdx.y * E_TILE, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* xb = x + b * stride_b;
#pragma unroll
  for (int i = 0; i < E_TILE / 8; ++i) {
    const int d = warp + 8 * i;
    if (d0 + d < dim) {
      const float* row = xb + (int64_t)(d0 + d) * stride_d + a0;
      tile[d][lane] = (a0 + lane < hw) ? ld_stream_f32(row + lane) : 0.f;
      tile[d][lane + 32] = (a0 + lane + 32 < hw) ? ld_stream_f32(row + lane + 32) : 0.f;
    }
  }
  __syncthreads();
  float* ob = out + ((int64_t)b * rows_per_batch + row_offset) * dim;
#pragma unroll
  for (int i = 0; i < E_TILE / 8; ++i) {
    const int a = warp + 8 * i;
    if (a0 + a < hw) {
      float* row = ob + (int64_t)(a0 + a) * dim + d0;
      if (d0 + lane < dim) row[lane] = tile[lane][a];
      if (d0 + lane + 32 < dim) row[lane + 32] = tile[lane + 32][a];
    }
  }
}

// E2: re-pitch the rows of a conv output so that TMA can address it.  The fused similarity kernel reads
// [64 k x 128 anchors] boxes straight out of the NCHW tensor, which needs 16-byte aligned rows: H*W a
// multiple of 4 (fp32).  Image sizes such as 416, 480 or 608 give a 13x13 / 15x15 / 19x19 P5 level; without this
// the WHOLE step fell back to the two-kernel path.  One warp per row of `row_elems` elements (4-byte
// or 2-byte), 8 rows per block, coalesced along the row.
template <typename T>
__global__ void __launch_bounds__(256)
repitch_rows_kernel(const T* __restrict__ src, int64_t rows, int row_elems, int64_t src_pitch,
                    T* __restrict__ dst, int64_t dst_pitch) {
  const int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  const T* s = src + r * src_pitch;
  T* d = dst + r * dst_pitch;
  for (int i = threadIdx.x & 31; i < dst_pitch; i += 32) d[i] = i < row_elems ? s[i] : T(0);
}

}  // namespace ovdet

extern "C" int ovdet_repitch_rows(const void* src, int64_t rows, int64_t row_elems, int64_t src_pitch,
                                  void* dst, int64_t dst_pitch, int elem_size, void* stream) {
  using namespace ovdet;
  if (rows == 0) return check_device();
  if (!src || !dst || rows < 0 || row_elems <= 0 || src_pitch < row_elems || dst_pitch < row_elems ||
      (elem_size != 2 && elem_size != 4) || row_elems >= (1ll << 31) || dst_pitch >= (1ll << 31))
    return OVDET_ERR_INVALID_ARG;
  if (ceil_div<int64_t>(rows, 8) >= (1ll << 31)) return OVDET_ERR_UNSUPPORTED_SHAPE;
  if (int rc = check_device()) return rc;
  const unsigned grid = (unsigned)ceil_div<int64_t>(rows, 8);
  if (elem_size == 4)
    repitch_rows_kernel<uint32_t><<<grid, 256, 0, as_stream(stream)>>>(
        static_cast<const uint32_t*>(src), rows, (int)row_elems, src_pitch, static_cast<uint32_t*>(dst), dst_pitch);
  else
    repitch_rows_kernel<uint16_t><<<grid, 256, 0, as_stream(stream)>>>(
        static_cast<const uint16_t*>(src), rows, (int)row_elems, src_pitch, static_cast<uint16_t*>(dst), dst_pitch);
  OVDET_LAUNCH_CHECK();
  return OVDET_OK;
}

extern "C" int ovdet_concat_embeddings(const float* x, int64_t batch, int64_t dim, int64_t hw,
                                       int64_t stride_b, int64_t stride_d, float* out,
                                       int64_t rows_per_batch, int64_t row_offset, void* stream) {
  using namespace ovdet;
  if (batch == 0 || hw == 0) return check_device();
  if (!x || !out || batch < 0 || dim <= 0 || hw < 0 || row_offset < 0 || row_offset + hw > rows_per_batch)
    return OVDET_ERR_INVALID_ARG;
  if (batch > 65535 || ceil_div<int64_t>(dim, E_TILE) > 65535) return OVDET_ERR_UNSUPPORTED_SHAPE;
  if (int rc = check_device()) return rc;
  dim3 grid((unsigned)ceil_div<int64_t>(hw, E_TILE), (unsigned)ceil_div<int64_t>(dim, E_TILE), (unsigned)batch);
  concat_embeddings_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, (int)dim, (int)hw, stride_b, stride_d, out,
                                                              rows_per_batch, row_offset);
  OVDET_LAUNCH_CHECK();
  return OVDET_OK;
}
