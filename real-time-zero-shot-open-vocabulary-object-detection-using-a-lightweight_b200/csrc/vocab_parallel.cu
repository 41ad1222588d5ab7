// Vocabulary-parallel exchange: the prompts are sharded over the GPUs of one NVLink box and the
// per-anchor (max, argmax) of model/yolo_clip.py:198-206 is reduced over the class shards.
//
// The reduction itself happens inside the similarity kernel (sim_fused_sm100.cu, emit_row):
// every finished row is one 64-bit key, max-reduced with red.sys.max.u64 into every rank's key
// array through peer mappings.  This file holds what surrounds it:
//   * peer buffers: cudaMalloc + CUDA IPC handles (one process per GPU),
//   * the flag handshake that replaces a collective's synchronisation (signal / wait),
//   * the unpack of the merged keys into the scores / class ids K3 and K4 read,
//   * pack / unpack for the library-collective baseline (all-reduce MAX over int64).
//
// Step counters live in the rank's own buffer (ctr[0] = step its next similarity kernel
// contributes to, ctr[1] = step its next wait_unpack consumes; the signal kernel moves both), so no
// kernel argument changes from step to step and the whole sequence can be replayed from a CUDA graph.
//
// Ordering argument (why one parity bit is enough).  Rank r hands keys[p] back (writes 0) in its
// wait_unpack of step s.  A peer q touches r's keys[p] again in its similarity kernel of step
// s + 2, which its stream runs after its wait_unpack of step s + 1, which returns only after r's
// flag reached s + 1, which r stores in the signal kernel that FOLLOWS its step-s wait_unpack in
// stream order.  So the hand-back precedes every step s + 2 atomic.  The same chain with s
// instead of s + 1 shows that all step-s atomics of q are complete before r reads: q's signal
// kernel runs after q's similarity kernel has finished (stream order), fences at system scope,
// then stores the flag.
#include "common.cuh"
#include <cstring>

namespace ovdet {
namespace {

constexpr long long kSignBit = (long long)0x8000000000000000ull;

__host__ __device__ inline size_t vp_flags_offset(long long rows) {
  return ((size_t)rows * 16 + 127) / 128 * 128;            // after keys[2][rows]
}

// flags block (128 bytes): flags[8] u64, then ctr[2] u64, then the sticky "a wait expired" word
constexpr int kCtrOffset = 64;
constexpr int kStickyOffset = 80;

__global__ void vp_signal_kernel(VpTarget t, long long flags_off, int rank) {
  // the similarity kernel of this step has completed (stream order); make its atomics and this
  // store ordered for every observer
  unsigned long long* ctr = reinterpret_cast<unsigned long long*>(
      reinterpret_cast<char*>(t.keys[rank]) + flags_off + kCtrOffset);
  const unsigned long long step = __ldcg(ctr);
  __threadfence_system();
  const int g = threadIdx.x;
  if (g < t.world) {
    unsigned long long* flag = reinterpret_cast<unsigned long long*>(
        reinterpret_cast<char*>(t.keys[g]) + flags_off) + rank;
    asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(flag), "l"(step) : "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    ctr[1] = step;                  // what this rank's wait_unpack consumes next
    ctr[0] = step + 1;              // what its next similarity kernel contributes to
  }
}

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// A timeout is STICKY (a word in the rank's own flags block, so the test costs one L2 read; it is
// mirrored into `status`, which may be mapped pinned host memory that the host reads without a
// synchronisation): once a wait has expired - in this block, another block or an earlier step -
// nothing is unpacked and nothing is handed back.  A late peer's atomics would otherwise land in a
// key array that was already zeroed for step + 2 and corrupt two steps.  The rows this block owns
// get the sentinel (-inf, class 0), so that K3 / K4 produce NO detections instead of detections
// from a partly reduced maximum; the host raises on the status word (VocabParallelHead).
__global__ void __launch_bounds__(256)
vp_wait_unpack_kernel(unsigned long long* keys2, const unsigned long long* flags, int world, long long rows,
                      float* __restrict__ scores, int* __restrict__ class_ids,
                      int* status, unsigned long long timeout_ns) {
  __shared__ int s_timeout;
  unsigned long long* sticky = const_cast<unsigned long long*>(flags) + kStickyOffset / 8;
  if (threadIdx.x == 0) s_timeout = __ldcg(sticky) != 0ull;
  const unsigned long long step = __ldcg(flags + kCtrOffset / 8 + 1);
  unsigned long long* keys = keys2 + (long long)(step & 1ull) * rows;
  __syncthreads();
  if ((int)threadIdx.x < world && !s_timeout) {
    const unsigned long long t0 = global_timer_ns();
    while (ld_acquire_sys(flags + threadIdx.x) < step) {
      if (global_timer_ns() - t0 > timeout_ns) { s_timeout = 1; break; }
      __nanosleep(100);
    }
  }
  __syncthreads();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s_timeout) {
    if (threadIdx.x == 0) {
      *reinterpret_cast<volatile unsigned long long*>(sticky) = 1ull;
      if (status != nullptr) *reinterpret_cast<volatile int*>(status) = 1;
      __threadfence_system();
    }
    if (i < rows) {
      scores[i] = -INFINITY;
      class_ids[i] = 0;
    }
    return;
  }
  if (i >= rows) return;
  // the keys were written by remote (and local) atomics, performed at this GPU's L2: read there
  const unsigned long long key = __ldcg(keys + i);
  float s;
  int c;
  vp_unpack_key(key, s, c);
  scores[i] = s;
  class_ids[i] = c;
  keys[i] = 0ull;                                        // handed back for step + 2
}

__global__ void vp_init_kernel(unsigned long long* buf, long long words, long long ctr0_word) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < words) buf[i] = i == ctr0_word ? 1ull : 0ull;       // first step is 1
}

__global__ void pack_keys_kernel(const float* __restrict__ scores, const int* __restrict__ class_ids,
                                 long long n, int class_offset, long long* __restrict__ keys) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) keys[i] = (long long)vp_pack_key(scores[i], class_offset + class_ids[i]) ^ kSignBit;
}
__global__ void unpack_keys_kernel(const long long* __restrict__ keys, long long n, float* __restrict__ scores,
                                   int* __restrict__ class_ids) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s;
  int c;
  vp_unpack_key((unsigned long long)(keys[i] ^ kSignBit), s, c);
  scores[i] = s;
  class_ids[i] = c;
}

int fill_target(VpTarget& t, void* const* peer_buffers, int world, int rank, long long rows) {
  if (!peer_buffers || world < 1 || world > OVDET_MAX_PEERS || rows <= 0 || rank < 0 || rank >= world)
    return OVDET_ERR_INVALID_ARG;
  t.world = world;
  t.rows = rows;
  for (int g = 0; g < world; ++g) {
    if (!peer_buffers[g] || ((uintptr_t)peer_buffers[g] & 127)) return OVDET_ERR_INVALID_ARG;
    t.keys[g] = static_cast<unsigned long long*>(peer_buffers[g]);
  }
  t.step = reinterpret_cast<const unsigned long long*>(
      static_cast<char*>(peer_buffers[rank]) + vp_flags_offset(rows) + kCtrOffset);
  return OVDET_OK;
}

}  // namespace
}  // namespace ovdet

using namespace ovdet;

extern "C" int ovdet_peer_buffer_create(size_t bytes, void** ptr, void* handle64) {
  if (!ptr || !handle64 || bytes == 0) return OVDET_ERR_INVALID_ARG;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  if (int rc = check_device()) return rc;
  void* p = nullptr;
  OVDET_CUDA_TRY(cudaMalloc(&p, bytes));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    return cuda_fail(e);
  }
  memcpy(handle64, &h, 64);
  *ptr = p;
  return OVDET_OK;
}

extern "C" int ovdet_peer_buffer_open(const void* handle64, void** ptr) {
  if (!ptr || !handle64) return OVDET_ERR_INVALID_ARG;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  OVDET_CUDA_TRY(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return OVDET_OK;
}

extern "C" int ovdet_peer_buffer_close(void* ptr) {
  if (!ptr) return OVDET_ERR_INVALID_ARG;
  OVDET_CUDA_TRY(cudaIpcCloseMemHandle(ptr));
  return OVDET_OK;
}

extern "C" int ovdet_peer_buffer_destroy(void* ptr) {
  if (!ptr) return OVDET_ERR_INVALID_ARG;
  OVDET_CUDA_TRY(cudaFree(ptr));
  return OVDET_OK;
}

extern "C" size_t ovdet_vp_buffer_bytes(int64_t rows, int world) {
  if (rows <= 0 || world < 1 || world > OVDET_MAX_PEERS) return 0;
  return vp_flags_offset(rows) + 128;
}

extern "C" int ovdet_vp_buffer_init(void* buffer, int64_t rows, int world, void* stream) {
  if (!buffer || ((uintptr_t)buffer & 127) || rows <= 0 || world < 1 || world > OVDET_MAX_PEERS)
    return OVDET_ERR_INVALID_ARG;
  if (int rc = check_device()) return rc;
  const long long words = (long long)(ovdet_vp_buffer_bytes(rows, world) / 8);
  vp_init_kernel<<<(unsigned)ceil_div<long long>(words, 256), 256, 0, as_stream(stream)>>>(
      static_cast<unsigned long long*>(buffer), words, (long long)((vp_flags_offset(rows) + kCtrOffset) / 8));
  OVDET_LAUNCH_CHECK();
  return OVDET_OK;
}

extern "C" int ovdet_similarity_fused_vp(const void* const* obj_embeds, const int64_t* hw,
                                         const int64_t* stride_b, const int64_t* stride_d,
                                         int num_levels, int64_t batch, int64_t dim, const void* text_op,
                                         int64_t classes, int text_batched, float alpha, float beta,
                                         float* inv_norm, void* workspace, size_t workspace_bytes,
                                         int embed_dtype, int64_t class_offset,
                                         void* const* peer_buffers, int world, int rank, void* stream) {
  if (dim % 64 != 0) return OVDET_ERR_UNSUPPORTED_SHAPE;
  if (!hw || num_levels <= 0 || num_levels > OVDET_MAX_LEVELS || batch < 0) return OVDET_ERR_INVALID_ARG;
  if (class_offset < 0 || class_offset + classes >= (1ll << 31)) return OVDET_ERR_INVALID_ARG;
  long long anchors = 0;
  for (int l = 0; l < num_levels; ++l) anchors += hw[l];
  VpTarget t{};
  if (batch == 0) return check_device();
  if (int rc = fill_target(t, peer_buffers, world, rank, batch * anchors)) return rc;
  t.class_offset = (int)class_offset;
  return fused_launch(reinterpret_cast<const float* const*>(obj_embeds), hw, stride_b, stride_d, num_levels,
                      batch, dim, text_op, nullptr, classes, text_batched, 1, 0, alpha, beta, nullptr,
                      OVDET_F32, classes, nullptr, nullptr, inv_norm, stream, embed_dtype == OVDET_BF16,
                      workspace, workspace_bytes, &t);
}

extern "C" int ovdet_vp_signal(void* const* peer_buffers, int world, int rank, int64_t rows, void* stream) {
  VpTarget t{};
  if (int rc = fill_target(t, peer_buffers, world, rank, rows)) return rc;
  if (int rc = check_device()) return rc;
  vp_signal_kernel<<<1, 32, 0, as_stream(stream)>>>(t, (long long)vp_flags_offset(rows), rank);
  OVDET_LAUNCH_CHECK();
  return OVDET_OK;
}

extern "C" int ovdet_vp_wait_unpack(void* local_buffer, int world, int64_t rows,
                                    float* scores, int32_t* class_ids, int32_t* status, int timeout_ms,
                                    void* stream) {
  if (!local_buffer || ((uintptr_t)local_buffer & 127) || !scores || !class_ids) return OVDET_ERR_INVALID_ARG;
  if (world < 1 || world > OVDET_MAX_PEERS || rows <= 0 || timeout_ms < 0) return OVDET_ERR_INVALID_ARG;
  if (int rc = check_device()) return rc;
  unsigned long long* base = static_cast<unsigned long long*>(local_buffer);
  const unsigned long long* flags = reinterpret_cast<const unsigned long long*>(
      static_cast<char*>(local_buffer) + vp_flags_offset(rows));
  const unsigned long long timeout_ns = (unsigned long long)(timeout_ms ? timeout_ms : 2000) * 1000000ull;
  vp_wait_unpack_kernel<<<(unsigned)ceil_div<long long>(rows, 256), 256, 0, as_stream(stream)>>>(
      base, flags, world, rows, scores, class_ids, status, timeout_ns);
  OVDET_LAUNCH_CHECK();
  return OVDET_OK;
}

extern "C" int ovdet_pack_score_keys(const float* scores, const int32_t* class_ids, int64_t n,
                                     int64_t class_offset, int64_t* keys, void* stream) {
  if (!scores || !class_ids || !keys || n < 0 || class_offset < 0 || class_offset >= (1ll << 31))
    return OVDET_ERR_INVALID_ARG;
  if (int rc = check_device()) return rc;
  if (n == 0) return OVDET_OK;
  pack_keys_kernel<<<(unsigned)ceil_div<long long>(n, 256), 256, 0, as_stream(stream)>>>(
      scores, class_ids, n, (int)class_offset, reinterpret_cast<long long*>(keys));
  OVDET_LAUNCH_CHECK();
  return OVDET_OK;
}

extern "C" int ovdet_unpack_score_keys(const int64_t* keys, int64_t n, float* scores, int32_t* class_ids,
                                       void* stream) {
  if (!scores || !class_ids || !keys || n < 0) return OVDET_ERR_INVALID_ARG;
  if (int rc = check_device()) return rc;
  if (n == 0) return OVDET_OK;
  unpack_keys_kernel<<<(unsigned)ceil_div<long long>(n, 256), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const long long*>(keys), n, scores, class_ids);
  OVDET_LAUNCH_CHECK();
  return OVDET_OK;
}
