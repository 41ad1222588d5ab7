// K1: L2 normalisation + bf16 tensor-core operand construction.
//
// Replaces the permute + F.normalize pair of the reference
// (model/heads/text_contrastive.py:134,137-138).  HBM-bound streaming kernels:
//   regions: read  fp32 [B, D, HW] (NCHW, HW contiguous)           4 B / element
//            write bf16 [B, A, D]  (anchor-major = K-major operand) 2 B / element (4 B split)
//            write fp32 inv_norm [B, A]
//   text:    one warp per prompt row, normalised before rounding.
#include "common.cuh"
#include <cuda_fp16.h>

namespace ovdet {

// ---------------------------------------------------------------------------------------------
// Regions.  One CTA owns ANCH consecutive anchors of one image and walks the whole embedding
// dimension.  Global reads are coalesced along HW (each warp instruction covers one or two
// 128 B lines); the transpose to anchor-major happens in shared memory with a 16 B-chunk XOR
// swizzle (chunk ^ (anchor & 7)) so that both the column-wise fill and the row-wise drain are
// bank-conflict free; global writes are 16 B per lane, 512 B contiguous per warp instruction.
// ---------------------------------------------------------------------------------------------
template <int ANCH, bool SPLIT>
__global__ void __launch_bounds__(256)
l2norm_regions_kernel(const float* __restrict__ x, int dim, int hw, int64_t stride_b,
                      int64_t stride_d, __nv_bfloat16* __restrict__ operand,
                      int64_t rows_per_batch, int64_t row_offset, int kop,
                      float* __restrict__ inv_norm) {
  constexpr int APL = ANCH / 32;            // anchors per lane
  constexpr int WARPS = 8;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int chunks = dim >> 3;              // 16-byte chunks per operand row
  uint4* tile_hi = reinterpret_cast<uint4*>(smem_raw);
  uint4* tile_lo = tile_hi + (SPLIT ? ANCH * chunks : 0);
  float* partial = reinterpret_cast<float*>(tile_hi + (SPLIT ? 2 : 1) * ANCH * chunks);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int a0 = blockIdx.x * ANCH;
  const int b = blockIdx.y;
  const float* xb = x + b * stride_b;

  float ss[APL];
#pragma unroll
  for (int j = 0; j < APL; ++j) ss[j] = 0.f;

  for (int g = warp; g < chunks; g += WARPS) {
    float v[APL][8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const float* row = xb + (int64_t)(g * 8 + r) * stride_d + a0;
#pragma unroll
      for (int j = 0; j < APL; ++j) {
        const int a = lane + 32 * j;
        v[j][r] = (a0 + a < hw) ? ld_stream_f32(row + a) : 0.f;
      }
    }
#pragma unroll
    for (int j = 0; j < APL; ++j) {
      const int a = lane + 32 * j;
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int r = 0; r < 8; r += 2) {
        const float f0 = v[j][r], f1 = v[j][r + 1];
        ss[j] = fmaf(f0, f0, ss[j]);
        ss[j] = fmaf(f1, f1, ss[j]);
        const __nv_bfloat16 h0 = __float2bfloat16_rn(f0), h1 = __float2bfloat16_rn(f1);
        hi[r >> 1] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
        if (SPLIT) lo[r >> 1] = pack_bf16x2(f0 - __bfloat162float(h0), f1 - __bfloat162float(h1));
      }
      const int slot = a * chunks + (g ^ (a & 7));
      tile_hi[slot] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      if (SPLIT) tile_lo[slot] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
  }
#pragma unroll
  for (int j = 0; j < APL; ++j) partial[warp * ANCH + lane + 32 * j] = ss[j];
  __syncthreads();

  if (threadIdx.x < ANCH) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) s += partial[w * ANCH + threadIdx.x];
    const int a = a0 + threadIdx.x;
    if (a < hw)
      inv_norm[b * rows_per_batch + row_offset + a] = 1.0f / fmaxf(sqrtf(s), 1e-12f);
  }

  for (int a = warp; a < ANCH; a += WARPS) {
    if (a0 + a >= hw) break;
    uint4* dst = reinterpret_cast<uint4*>(operand + (b * rows_per_batch + row_offset + a0 + a) * kop);
    for (int c = lane; c < chunks; c += 32) {
      const int slot = a * chunks + (c ^ (a & 7));
      dst[c] = tile_hi[slot];
      if (SPLIT) dst[chunks + c] = tile_lo[slot];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Text.  One warp per prompt; rows are tiny (D floats) so the second pass re-reads L1/L2.
// ---------------------------------------------------------------------------------------------
// SEG 1: [hi]; 2: [hi | lo] (the two-kernel fp32 recipe); 3: [hi | lo | hi], the operand the fused
// kernel's three-pass mode multiplies against activation blocks laid out [hi | hi | lo].
// SEG 0: ONE fp16 segment holding 16 x the unit row (the fused kernel's fp16 tier: fp16 keeps 11
// significant bits down to 6e-5, so the unit rows - typical element 0.04 - are lifted by 2^4; the
// factor comes back out in the kernel's row scale).
template <int SEG>
__global__ void __launch_bounds__(256)
l2norm_text_kernel(const float* __restrict__ t, int64_t total_rows, int classes, int dim,
                   int64_t stride_b, int64_t stride_c, __nv_bfloat16* __restrict__ operand,
                   int kop, float* __restrict__ inv_norm) {
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= total_rows) return;
  const int64_t b = row / classes, c = row % classes;
  const float* src = t + b * stride_b + c * stride_c;
  float ss = 0.f;
  for (int i = lane; i < dim; i += 32) { const float f = src[i]; ss = fmaf(f, f, ss); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float denom = fmaxf(sqrtf(ss), 1e-12f);
  if (inv_norm != nullptr && lane == 0) inv_norm[row] = 1.0f / denom;
  __nv_bfloat16* dst = operand + row * kop;
  for (int i = lane; i < dim; i += 32) {
    const float f = __fdiv_rn(src[i], denom);
    if (SEG == 0) {
      reinterpret_cast<__half*>(dst)[i] = __float2half_rn(f * 16.0f);
      continue;
    }
    const __nv_bfloat16 h = __float2bfloat16_rn(f);
    dst[i] = h;
    if (SEG >= 2) dst[dim + i] = __float2bfloat16_rn(f - __bfloat162float(h));
    if (SEG == 3) dst[2 * dim + i] = h;
  }
}

// The same for rows of 128 * nv floats (nv <= 8) that start 16-byte aligned: the row is read ONCE with
// 16-byte loads (4 * nv per lane, kept in registers) and written with 8-byte stores.  This is the
// per-image-text case (text [B, C, 512] from the neck, model/repvl_pan.py:173-182: batch * classes =
// 308 k rows per step at the benchmark's shape), where the scalar kernel above - two passes over the
// row, 4-byte loads, 2-byte stores - ran at 0.6 of the HBM bandwidth.
template <int SEG>
__global__ void __launch_bounds__(256)
l2norm_text_vec_kernel(const float* __restrict__ t, int64_t total_rows, int classes, int nv,
                       int64_t stride_b, int64_t stride_c, __nv_bfloat16* __restrict__ operand,
                       int kop, float* __restrict__ inv_norm) {
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= total_rows) return;
  const int64_t b = row / classes, c = row % classes;
  const float4* src = reinterpret_cast<const float4*>(t + b * stride_b + c * stride_c) + lane;
  const int dim = nv * 128;
  float4 v[8];
  float ss = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (j < nv) v[j] = ld_stream_f32x4(src + 32 * j);
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (j < nv) {
      ss = fmaf(v[j].x, v[j].x, ss); ss = fmaf(v[j].y, v[j].y, ss);
      ss = fmaf(v[j].z, v[j].z, ss); ss = fmaf(v[j].w, v[j].w, ss);
    }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float denom = fmaxf(sqrtf(ss), 1e-12f);
  if (inv_norm != nullptr && lane == 0) inv_norm[row] = 1.0f / denom;
  __nv_bfloat16* dst = operand + row * kop + 4 * lane;
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (j < nv) {
      const float f0 = __fdiv_rn(v[j].x, denom), f1 = __fdiv_rn(v[j].y, denom);
      const float f2 = __fdiv_rn(v[j].z, denom), f3 = __fdiv_rn(v[j].w, denom);
      if (SEG == 0) {
        *reinterpret_cast<uint2*>(dst + 128 * j) = make_uint2(pack_f16x2_sat(f0 * 16.0f, f1 * 16.0f),
                                                              pack_f16x2_sat(f2 * 16.0f, f3 * 16.0f));
        continue;
      }
      const uint32_t h01 = pack_bf16x2(f0, f1), h23 = pack_bf16x2(f2, f3);
      *reinterpret_cast<uint2*>(dst + 128 * j) = make_uint2(h01, h23);
      if (SEG >= 2)
        *reinterpret_cast<uint2*>(dst + dim + 128 * j) =
            make_uint2(pack_bf16x2(f0 - __uint_as_float(h01 << 16), f1 - __uint_as_float(h01 & 0xffff0000u)),
                       pack_bf16x2(f2 - __uint_as_float(h23 << 16), f3 - __uint_as_float(h23 & 0xffff0000u)));
      if (SEG == 3) *reinterpret_cast<uint2*>(dst + 2 * dim + 128 * j) = make_uint2(h01, h23);
    }
}

// Raw (un-normalised) text rows -> bf16 operand, zero padded to `dpad` columns per segment.
// segments == 1: [hi]; segments == 3: [hi | lo | hi] with lo = bf16(x - hi), the layout the
// three-pass (fp32-accurate) mode of the fused kernel multiplies against [hi | hi | lo].
__global__ void __launch_bounds__(256)
cast_text_kernel(const float* __restrict__ t, int64_t total_rows, int classes, int dim, int dpad,
                 int segments, int64_t stride_b, int64_t stride_c, __nv_bfloat16* __restrict__ operand) {
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= total_rows) return;
  const int64_t b = row / classes, c = row % classes;
  const float* src = t + b * stride_b + c * stride_c;
  __nv_bfloat16* dst = operand + row * (int64_t)dpad * segments;
  for (int i = lane; i < dpad; i += 32) {
    const float f = i < dim ? src[i] : 0.f;
    const __nv_bfloat16 h = __float2bfloat16_rn(f);
    dst[i] = h;
    if (segments == 3) {
      dst[dpad + i] = __float2bfloat16_rn(f - __bfloat162float(h));
      dst[2 * dpad + i] = h;
    }
  }
}

template <int ANCH, bool SPLIT>
static int launch_regions(const float* x, int64_t batch, int dim, int hw, int64_t stride_b,
                          int64_t stride_d, __nv_bfloat16* operand, int64_t rows_per_batch,
                          int64_t row_offset, int kop, float* inv_norm, cudaStream_t stream) {
  const size_t smem = (size_t)ANCH * dim * 2 * (SPLIT ? 2 : 1) + 8 * ANCH * sizeof(float);
  if (smem > 227 * 1024) return OVDET_ERR_UNSUPPORTED_SHAPE;
  auto kern = l2norm_regions_kernel<ANCH, SPLIT>;
  OVDET_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)ceil_div<int64_t>(hw, ANCH), (unsigned)batch);
  kern<<<grid, 256, smem, stream>>>(x, dim, hw, stride_b, stride_d, operand, rows_per_batch,
                                    row_offset, kop, inv_norm);
  OVDET_LAUNCH_CHECK();
  return OVDET_OK;
}

}  // namespace ovdet

extern "C" int ovdet_l2norm_regions(const float* x, int64_t batch, int64_t dim, int64_t hw,
                                    int64_t stride_b, int64_t stride_d, void* operand,
                                    int64_t rows_per_batch, int64_t row_offset, int64_t kop,
                                    int split, float* inv_norm, void* stream) {
  using namespace ovdet;
  if (!x || !operand || !inv_norm || batch < 0 || dim <= 0 || hw < 0) return OVDET_ERR_INVALID_ARG;
  if (row_offset < 0 || row_offset + hw > rows_per_batch) return OVDET_ERR_INVALID_ARG;
  if (dim % 64 != 0 || batch > 65535) return OVDET_ERR_UNSUPPORTED_SHAPE;
  if (kop != dim * (split ? 2 : 1)) return OVDET_ERR_INVALID_ARG;
  if ((reinterpret_cast<uintptr_t>(operand) & 15) != 0) return OVDET_ERR_INVALID_ARG;
  if (int rc = check_device()) return rc;
  if (batch == 0 || hw == 0) return OVDET_OK;
  auto* op = static_cast<__nv_bfloat16*>(operand);
  cudaStream_t s = as_stream(stream);
  if (split)
    return launch_regions<32, true>(x, batch, (int)dim, (int)hw, stride_b, stride_d, op,
                                    rows_per_batch, row_offset, (int)kop, inv_norm, s);
  return launch_regions<64, false>(x, batch, (int)dim, (int)hw, stride_b, stride_d, op,
                                   rows_per_batch, row_offset, (int)kop, inv_norm, s);
}

extern "C" int ovdet_l2norm_text(const float* t, int64_t batch, int64_t classes, int64_t dim,
                                 int64_t stride_b, int64_t stride_c, void* operand, int64_t kop,
                                 int split, float* inv_norm, void* stream) {
  using namespace ovdet;
  if (!t || !operand || batch < 0 || classes < 0 || dim <= 0) return OVDET_ERR_INVALID_ARG;
  if (dim % 64 != 0) return OVDET_ERR_UNSUPPORTED_SHAPE;
  if (split < 0 || split > 3 || kop != dim * (split == 3 ? 1 : split + 1)) return OVDET_ERR_INVALID_ARG;
  if (int rc = check_device()) return rc;
  const int64_t total = batch * classes;
  if (total == 0) return OVDET_OK;
  auto* op = static_cast<__nv_bfloat16*>(operand);
  const unsigned grid = (unsigned)ceil_div<int64_t>(total, 8);
  // rows of 128 * nv floats, 16-byte aligned (and 8-byte aligned operand rows): the vectorised kernel
  if (dim % 128 == 0 && dim <= 1024 && !((uintptr_t)t & 15) && !(stride_b & 3) && !(stride_c & 3) &&
      !((uintptr_t)operand & 7) && !(kop & 3)) {
    const int nv = (int)(dim / 128);
    auto s_ = as_stream(stream);
    if (split == 3) l2norm_text_vec_kernel<0><<<grid, 256, 0, s_>>>(t, total, (int)classes, nv, stride_b, stride_c, op, (int)kop, inv_norm);
    else if (split == 2) l2norm_text_vec_kernel<3><<<grid, 256, 0, s_>>>(t, total, (int)classes, nv, stride_b, stride_c, op, (int)kop, inv_norm);
    else if (split == 1) l2norm_text_vec_kernel<2><<<grid, 256, 0, s_>>>(t, total, (int)classes, nv, stride_b, stride_c, op, (int)kop, inv_norm);
    else l2norm_text_vec_kernel<1><<<grid, 256, 0, s_>>>(t, total, (int)classes, nv, stride_b, stride_c, op, (int)kop, inv_norm);
    OVDET_LAUNCH_CHECK();
    return OVDET_OK;
  }
  if (split == 3)
    l2norm_text_kernel<0><<<grid, 256, 0, as_stream(stream)>>>(t, total, (int)classes, (int)dim,
                                                               stride_b, stride_c, op, (int)kop, inv_norm);
  else if (split == 2)
    l2norm_text_kernel<3><<<grid, 256, 0, as_stream(stream)>>>(t, total, (int)classes, (int)dim,
                                                               stride_b, stride_c, op, (int)kop, inv_norm);
  else if (split == 1)
    l2norm_text_kernel<2><<<grid, 256, 0, as_stream(stream)>>>(t, total, (int)classes, (int)dim,
                                                               stride_b, stride_c, op, (int)kop, inv_norm);
  else
    l2norm_text_kernel<1><<<grid, 256, 0, as_stream(stream)>>>(t, total, (int)classes, (int)dim,
                                                               stride_b, stride_c, op, (int)kop, inv_norm);
  OVDET_LAUNCH_CHECK();
  return OVDET_OK;
}

extern "C" int ovdet_cast_text(const float* t, int64_t batch, int64_t classes, int64_t dim,
                               int64_t stride_b, int64_t stride_c, void* operand, int64_t kop,
                               int split3, void* stream) {
  using namespace ovdet;
  if (!t || !operand || batch < 0 || classes < 0 || dim <= 0) return OVDET_ERR_INVALID_ARG;
  const int64_t dpad = ceil_div<int64_t>(dim, 64) * 64;
  if (kop != dpad * (split3 ? 3 : 1)) return OVDET_ERR_INVALID_ARG;
  if (dim > 512) return OVDET_ERR_UNSUPPORTED_SHAPE;
  if (int rc = check_device()) return rc;
  const int64_t total = batch * classes;
  if (total == 0) return OVDET_OK;
  cast_text_kernel<<<(unsigned)ceil_div<int64_t>(total, 8), 256, 0, as_stream(stream)>>>(
      t, total, (int)classes, (int)dim, (int)dpad, split3 ? 3 : 1, stride_b, stride_c,
      static_cast<__nv_bfloat16*>(operand));
  OVDET_LAUNCH_CHECK();
  return OVDET_OK;
}
