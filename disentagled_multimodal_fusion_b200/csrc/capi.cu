// Library-level entry points of libdmf_b200.so (version, error text, device guard, launch counter).
#include "common.cuh"
#include <string.h>

namespace dmf {
thread_local char g_last_error[512] = {0};
std::atomic<long long> g_launches{0};
}  // namespace dmf

using namespace dmf;

extern "C" int dmf_version(void) { return 100; }  // 0.1.0

extern "C" int dmf_last_error(char* buf, size_t n) {
  if (!buf || n == 0) return (int)strlen(g_last_error);
  strncpy(buf, g_last_error, n - 1);
  buf[n - 1] = 0;
  return (int)strlen(buf);
}

extern "C" long long dmf_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" int dmf_device_check(void) {
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return fail((int)e, "dmf_device_check: cudaGetDevice: %s", cudaGetErrorString(e));
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10)
    return fail(-2, "dmf_device_check: device %d is sm_%d%d; this library contains sm_100a code only (no fallback)", dev,
                major, minor);
  return 0;
}
