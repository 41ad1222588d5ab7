// Library runtime: version, error strings, device checks.
#include "common.cuh"
#include <atomic>
#include <mutex>

namespace ovdet {

thread_local int g_last_cuda_error = 0;

namespace {
struct DeviceInfo { int checked = 0; int status = 0; int sms = 0; };
DeviceInfo g_dev[64];
}  // namespace

static DeviceInfo* current_info() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  DeviceInfo& d = g_dev[dev];
  if (!d.checked) {
    int major = 0, sms = 0;
    cudaError_t e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) { g_last_cuda_error = (int)e; d.status = OVDET_ERR_CUDA; }
    else d.status = (major == 10) ? OVDET_OK : OVDET_ERR_WRONG_ARCH;
    d.sms = sms;
    d.checked = 1;
  }
  return &d;
}

int check_device() {
  DeviceInfo* d = current_info();
  if (!d) { return OVDET_ERR_CUDA; }
  return d->status;
}

static std::atomic<unsigned> g_first_use_done[64];

bool first_use_done(int slot) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return false;
  return (g_first_use_done[dev].load(std::memory_order_acquire) >> slot) & 1u;
}

void first_use_mark(int slot) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return;
  g_first_use_done[dev].fetch_or(1u << slot, std::memory_order_release);
}

std::mutex& first_use_mutex() {
  static std::mutex mu;
  return mu;
}

int sm_count() {
  DeviceInfo* d = current_info();
  return d && d->sms > 0 ? d->sms : 148;
}

}  // namespace ovdet

extern "C" int ovdet_version(void) { return OVDET_VERSION; }

extern "C" const char* ovdet_strerror(int status) {
  switch (status) {
    case OVDET_OK: return "ok";
    case OVDET_ERR_INVALID_ARG: return "invalid argument";
    case OVDET_ERR_UNSUPPORTED_SHAPE: return "unsupported shape";
    case OVDET_ERR_WRONG_ARCH: return "device is not compute capability 10.x (B200 / sm_100a required, no fallback)";
    case OVDET_ERR_CUDA: return "CUDA runtime error (see ovdet_last_cuda_error)";
    case OVDET_ERR_WORKSPACE: return "workspace too small or misaligned";
    case OVDET_ERR_DRIVER: return "cuTensorMapEncodeTiled unavailable or failed";
    default: return "unknown status";
  }
}

extern "C" int ovdet_last_cuda_error(void) { return ovdet::g_last_cuda_error; }

extern "C" int ovdet_check_device(void) { return ovdet::check_device(); }
