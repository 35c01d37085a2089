// C-ABI plumbing shared by all kernels: error text, version, device info.
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "kd_common.cuh"

namespace kd {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  set_error("%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
  return 2;
}

static std::atomic<unsigned long long> g_launches{0};

int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return check_cuda(cudaGetLastError(), what);
}

}  // namespace kd

extern "C" unsigned long long kd_launch_count(void) { return kd::g_launches.load(std::memory_order_relaxed); }

extern "C" int kd_version(void) { return KD_ABI_VERSION; }

extern "C" const char* kd_last_error(void) { return kd::g_error; }

extern "C" int kd_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  if (kd::check_cuda(cudaGetDevice(&dev), "cudaGetDevice")) return 2;
  cudaDeviceProp prop;
  if (kd::check_cuda(cudaGetDeviceProperties(&prop, dev), "cudaGetDeviceProperties")) return 2;
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  return 0;
}
