// Shared host/device helpers for libovdet (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>

#include "../../include/ovdet.h"

namespace ovdet {

// ---- error plumbing ------------------------------------------------------------------------
extern thread_local int g_last_cuda_error;

inline int cuda_fail(cudaError_t e) {
  g_last_cuda_error = static_cast<int>(e);
  return OVDET_ERR_CUDA;
}

#define OVDET_CUDA_TRY(expr)                                   \
  do {                                                         \
    cudaError_t _e = (expr);                                   \
    if (_e != cudaSuccess) return ::ovdet::cuda_fail(_e);      \
  } while (0)

#define OVDET_LAUNCH_CHECK() OVDET_CUDA_TRY(cudaGetLastError())

// Returns OVDET_OK when the current device is compute capability 10.x; cached per device.
int check_device();
int sm_count();
// Kernel attributes such as the dynamic shared-memory limit are per device: `init` (returning an
// ovdet_status) runs once per (current device, slot), under a mutex, and the slot is marked done only
// after it succeeded - a concurrent first caller waits instead of launching before the attribute is
// set, and a failed attempt is retried by the next call.  Slots: 0 sim_gemm, 1 sim_fused, 2 nms.
bool first_use_done(int slot);
void first_use_mark(int slot);
std::mutex& first_use_mutex();
template <class F>
inline int once_per_device(int slot, F&& init) {
  if (first_use_done(slot)) return OVDET_OK;
  std::lock_guard<std::mutex> lock(first_use_mutex());
  if (first_use_done(slot)) return OVDET_OK;
  const int rc = init();
  if (rc == OVDET_OK) first_use_mark(slot);
  return rc;
}

// sim_fused_sm100.cu: the fused normalise + GEMM + row max kernel behind ovdet_similarity_fused
// and ovdet_max_sigmoid_attention.
// `vp` (vocabulary-parallel launches only): where the packed (score, class) keys of every row go.
struct VpTarget {
  int world;
  int class_offset;                              // global index of this launch's class 0
  long long rows;                                // batch * anchors
  const unsigned long long* step;                // this rank's device-resident step counter
  unsigned long long* keys[OVDET_MAX_PEERS];     // keys[2][rows] of every rank
};
int fused_launch(const float* const* obj_embeds, const int64_t* hw, const int64_t* stride_b,
                 const int64_t* stride_d, int num_levels, int64_t batch, int64_t dim,
                 const void* text_op, const void* const* level_ops, int64_t classes, int text_batched,
                 int normalize, int split3, float alpha, float beta, void* logits, int logits_dtype,
                 int64_t ldc, float* row_max, int32_t* row_arg, float* inv_norm, void* stream,
                 int in_bf16 = 0, void* split_ws = nullptr, size_t split_ws_bytes = 0,
                 const VpTarget* vp = nullptr, int f16_operands = 0);

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

template <typename T>
__host__ __device__ constexpr T ceil_div(T a, T b) { return (a + b - 1) / b; }

// ---- small device utilities ----------------------------------------------------------------
__device__ __forceinline__ float ld_stream_f32(const float* p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float4 ld_stream_f32x4(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ uint4 ld_stream_u32x4(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);   // .x = lo (low 16 bits), .y = hi
  return *reinterpret_cast<uint32_t*>(&v);
}

// two floats -> packed fp16 pair (.x = lo in the low 16 bits), round to nearest, finite saturation
__device__ __forceinline__ uint32_t pack_f16x2_sat(float lo, float hi) {
  uint32_t v;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(v) : "f"(hi), "f"(lo));
  return v;
}

// float -> uint32 whose unsigned order equals the float order (-inf < ... < +inf).
__device__ __forceinline__ uint32_t float_to_ordered(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(uint32_t u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// (score, class) -> one 64-bit key whose unsigned order is (score ascending, class DESCENDING): the
// maximum over any set of keys is the highest score and, among equal scores, the lowest class index
// (torch.max's CPU tie rule).  -0.0 is folded into +0.0 first (equal as floats).
__device__ __forceinline__ unsigned long long vp_pack_key(float score, int cls) {
  return ((unsigned long long)float_to_ordered(score + 0.0f) << 32) | (0xffffffffu - (uint32_t)cls);
}
__device__ __forceinline__ void vp_unpack_key(unsigned long long key, float& score, int& cls) {
  score = ordered_to_float((uint32_t)(key >> 32));
  cls = (int)(0xffffffffu - (uint32_t)key);
}

}  // namespace ovdet

// Internal launchers behind ovdet_decode_filter* / ovdet_nms_batched*; `pdl` = launch with programmatic
// stream serialization (only ovdet_head_step, which knows what the preceding kernel is, sets it).
int ovdet_decode_launch_internal(int pdl, int in_bf16, const void* const* box_preds, const int32_t* heights,
                                 const int32_t* widths, const int32_t* strides, const int64_t* batch_strides,
                                 int num_levels, int bins, int64_t batch, float width_scale, float height_scale,
                                 const float* scores, float conf, int activation, float* boxes,
                                 float* scores_act, uint32_t* pass_mask, void* stream);
int ovdet_nms_launch_internal(int pdl, int use_conf, float conf, const float* boxes, const float* scores,
                              const int32_t* classes, const uint32_t* pass_mask, int64_t batch, int64_t anchors,
                              const float* scale, const float* clip_wh, float iou_thr, int class_aware, int topk,
                              int64_t max_det, float* out_boxes, float* out_scores, int32_t* out_classes,
                              int32_t* out_anchor, int32_t* out_keep, int32_t* out_count,
                              int32_t* out_candidates, void* workspace, size_t workspace_bytes, void* stream);
