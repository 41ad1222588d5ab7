// P1: letterbox pre-processing of uint8 HWC images, bit-exact with the reference's
// cv2.resize(INTER_LINEAR) + top-left paste on a zero canvas + /255 + HWC -> CHW
// (inference/detector.py:139-156).  P2: int-truncated box records (detector.py:213-221).
//
// cv2.resize for 8-bit INTER_LINEAR is integer arithmetic (OpenCV imgproc/resize.cpp,
// HResizeLinear / VResizeLinear with INTER_RESIZE_COEF_BITS = 11):
//   fx = (float)((dx + 0.5) * scale_x - 0.5) evaluated in double, sx = floor(fx), fx -= sx,
//   columns clamp (sx < 0 -> sx = 0, fx = 0; sx >= w - 1 -> sx = w - 1, fx = 0), rows keep fy and
//   clamp the two source row indices instead;
//   alpha = rint((1 - fx) * 2048), rint(fx * 2048) as int16, likewise beta;
//   row[dx] = S[sx] * a0 + S[sx + 1] * a1                       (int32)
//   dst = (((b0 * (row0 >> 4)) >> 16) + ((b1 * (row1 >> 4)) >> 16) + 2) >> 2
// and an exact 2x decimation in both axes takes the INTER_AREA fast path (a+b+c+d+2) >> 2.
// The kernel is a streaming gather: one thread per canvas column walking 8 rows (the first
// version launched one thread per pixel and was bound by CTA launch rate: 123 k CTAs for 64
// images), three channel planes written with coalesced fp32 stores (4.9 MB per 640x640 image), source bytes served from L1/L2.
#include "common.cuh"

namespace ovdet {
namespace {

constexpr int kMaxImages = 64;        // images per launch (descriptor table passed by value, 3.6 KB)

struct ImageDesc {
  const uint8_t* data;
  long long row_stride;               // bytes
  int h, w;                           // source size
  int rh, rw;                         // resized size (pasted at the canvas origin)
  double scale_x, scale_y;            // 1 / ((double)rw / w), 1 / ((double)rh / h) as OpenCV computes them
  int area2x;                         // exact 2x decimation: INTER_AREA fast path
};

struct LetterboxParams {
  ImageDesc img[kMaxImages];
  int count;
  int out_h, out_w;
  float* out;                         // [count, 3, out_h, out_w]
};

__device__ __forceinline__ void linear_coeff(int d, double scale, int n, bool clamp_frac,
                                             int& s0, int& s1, int& a0, int& a1) {
  // no FMA contraction: OpenCV evaluates (d + 0.5) * scale - 0.5 with separate roundings
  const double t = __dsub_rn(__dmul_rn(__dadd_rn((double)d, 0.5), scale), 0.5);
  float f = __double2float_rn(t);
  int s = (int)floorf(f);
  f = __fsub_rn(f, (float)s);
  if (clamp_frac) {
    if (s < 0) { s = 0; f = 0.f; }
    if (s >= n - 1) { s = n - 1; f = 0.f; }
    s0 = s;
    s1 = min(s + 1, n - 1);
  } else {
    s0 = min(max(s, 0), n - 1);
    s1 = min(max(s + 1, 0), n - 1);
  }
  a0 = __float2int_rn(__fmul_rn(__fsub_rn(1.0f, f), 2048.0f));
  a1 = __float2int_rn(__fmul_rn(f, 2048.0f));
}

constexpr int kRowsPerBlock = 8;

// A thread owns one canvas column and walks kRowsPerBlock rows: the column coefficients (double
// precision source coordinate) are computed once, the row coefficients once per row.
__global__ void __launch_bounds__(256)
letterbox_kernel(const LetterboxParams p) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y0 = blockIdx.y * kRowsPerBlock;
  const int n = blockIdx.z;
  if (x >= p.out_w) return;
  const ImageDesc& im = p.img[n];
  const long long plane = (long long)p.out_h * p.out_w;
  float* out = p.out + (long long)n * 3 * plane + x;
  const bool in_x = x < im.rw;
  int sx0 = 0, sx1 = 0, ax0 = 0, ax1 = 0;
  if (in_x && !im.area2x) linear_coeff(x, im.scale_x, im.w, true, sx0, sx1, ax0, ax1);
  sx0 *= 3;
  sx1 *= 3;
  const int y_end = min(y0 + kRowsPerBlock, p.out_h);
#pragma unroll 4
  for (int y = y0; y < y_end; ++y) {
    float r = 0.f, g = 0.f, b = 0.f;
    if (in_x && y < im.rh) {
      int v[3];
      if (im.area2x) {
        const uint8_t* r0 = im.data + (long long)(2 * y) * im.row_stride + 6 * x;
        const uint8_t* r1 = r0 + im.row_stride;
#pragma unroll
        for (int c = 0; c < 3; ++c) v[c] = (r0[c] + r0[3 + c] + r1[c] + r1[3 + c] + 2) >> 2;
      } else {
        int sy0, sy1, ay0, ay1;
        linear_coeff(y, im.scale_y, im.h, false, sy0, sy1, ay0, ay1);
        const uint8_t* r0 = im.data + (long long)sy0 * im.row_stride;
        const uint8_t* r1 = im.data + (long long)sy1 * im.row_stride;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const int h0 = r0[sx0 + c] * ax0 + r0[sx1 + c] * ax1;
          const int h1 = r1[sx0 + c] * ax0 + r1[sx1 + c] * ax1;
          const int o = (((ay0 * (h0 >> 4)) >> 16) + ((ay1 * (h1 >> 4)) >> 16) + 2) >> 2;
          v[c] = min(max(o, 0), 255);
        }
      }
      r = __fdiv_rn((float)v[0], 255.0f);
      g = __fdiv_rn((float)v[1], 255.0f);
      b = __fdiv_rn((float)v[2], 255.0f);
    }
    float* o = out + (long long)y * p.out_w;
    o[0] = r;
    o[plane] = g;
    o[2 * plane] = b;
  }
}

__global__ void pack_boxes_kernel(const float* __restrict__ boxes, const int32_t* __restrict__ count,
                                  int max_det, int32_t* __restrict__ out, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // one box per thread
  if (i >= total) return;
  const int b = (int)(i / max_det), k = (int)(i - (long long)b * max_det);
  int4 v = make_int4(0, 0, 0, 0);
  if (k < count[b]) {
    const float4 f = reinterpret_cast<const float4*>(boxes)[i];
    v = make_int4((int)f.x, (int)f.y, (int)f.z, (int)f.w);               // toward zero, like astype(int)
  }
  reinterpret_cast<int4*>(out)[i] = v;
}

}  // namespace
}  // namespace ovdet

extern "C" int ovdet_letterbox_u8(const uint8_t* const* images, const int32_t* heights,
                                  const int32_t* widths, const int64_t* row_strides,
                                  const int32_t* resized_h, const int32_t* resized_w, int count,
                                  int out_h, int out_w, float* out, void* stream) {
  using namespace ovdet;
  if (!images || !heights || !widths || !row_strides || !resized_h || !resized_w || !out)
    return OVDET_ERR_INVALID_ARG;
  if (count < 0 || out_h <= 0 || out_w <= 0) return OVDET_ERR_INVALID_ARG;
  for (int i = 0; i < count; ++i) {
    if (!images[i] || heights[i] <= 0 || widths[i] <= 0 || row_strides[i] < 3ll * widths[i])
      return OVDET_ERR_INVALID_ARG;
    if (resized_h[i] <= 0 || resized_w[i] <= 0 || resized_h[i] > out_h || resized_w[i] > out_w)
      return OVDET_ERR_INVALID_ARG;
  }
  if (out_h > 65535 * kRowsPerBlock) return OVDET_ERR_UNSUPPORTED_SHAPE;
  if (int rc = check_device()) return rc;
  const long long plane3 = 3ll * out_h * out_w;
  for (int base = 0; base < count; base += kMaxImages) {
    LetterboxParams p{};
    p.count = count - base < kMaxImages ? count - base : kMaxImages;
    p.out_h = out_h;
    p.out_w = out_w;
    p.out = out + base * plane3;
    for (int i = 0; i < p.count; ++i) {
      ImageDesc& d = p.img[i];
      const int j = base + i;
      d.data = images[j];
      d.row_stride = row_strides[j];
      d.h = heights[j]; d.w = widths[j];
      d.rh = resized_h[j]; d.rw = resized_w[j];
      // cv::resize: inv_scale = (double)dsize / ssize; scale = 1. / inv_scale
      const double inv_x = (double)d.rw / d.w, inv_y = (double)d.rh / d.h;
      d.scale_x = 1.0 / inv_x;
      d.scale_y = 1.0 / inv_y;
      d.area2x = (d.w == 2 * d.rw && d.h == 2 * d.rh) ? 1 : 0;
    }
    dim3 grid((unsigned)ceil_div(out_w, 256), (unsigned)ceil_div(out_h, kRowsPerBlock), (unsigned)p.count);
    letterbox_kernel<<<grid, 256, 0, as_stream(stream)>>>(p);
    OVDET_LAUNCH_CHECK();
  }
  return OVDET_OK;
}

extern "C" int ovdet_pack_boxes_i32(const float* boxes, const int32_t* count, int64_t batch,
                                    int64_t max_det, int32_t* out, void* stream) {
  using namespace ovdet;
  if (!boxes || !count || !out || batch < 0 || max_det <= 0) return OVDET_ERR_INVALID_ARG;
  if (((uintptr_t)boxes & 15) || ((uintptr_t)out & 15)) return OVDET_ERR_INVALID_ARG;
  if (max_det >= (1ll << 31)) return OVDET_ERR_UNSUPPORTED_SHAPE;
  if (int rc = check_device()) return rc;
  const long long total = batch * max_det;
  if (total == 0) return OVDET_OK;
  pack_boxes_kernel<<<(unsigned)ceil_div<long long>(total, 256), 256, 0, as_stream(stream)>>>(
      boxes, count, (int)max_det, out, total);
  OVDET_LAUNCH_CHECK();
  return OVDET_OK;
}
