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

}  // namespace ovdet

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
