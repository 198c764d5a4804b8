// Error plumbing and device info for the C-ABI (include/neurecon_b200.h).
#include <stdarg.h>
#include <string.h>
#include "common.cuh"

#include <atomic>
static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void nr_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

extern "C" long long nr_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

void nr_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" int nr_version(void) { return 100; }

extern "C" int nr_last_error(char* buf, size_t n) {
  if (!buf || n == 0) return NR_ERR_INVALID;
  strncpy(buf, g_err, n - 1);
  buf[n - 1] = 0;
  return NR_OK;
}

extern "C" int nr_device_info(int* sm_count, int* smem_optin, int* cc_major, int* cc_minor) {
  int dev = 0;
  NR_CHECK_CUDA(cudaGetDevice(&dev));
  if (sm_count) NR_CHECK_CUDA(cudaDeviceGetAttribute(sm_count, cudaDevAttrMultiProcessorCount, dev));
  if (smem_optin) NR_CHECK_CUDA(cudaDeviceGetAttribute(smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  if (cc_major) NR_CHECK_CUDA(cudaDeviceGetAttribute(cc_major, cudaDevAttrComputeCapabilityMajor, dev));
  if (cc_minor) NR_CHECK_CUDA(cudaDeviceGetAttribute(cc_minor, cudaDevAttrComputeCapabilityMinor, dev));
  return NR_OK;
}
