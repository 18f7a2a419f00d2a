// C-ABI bookkeeping: version, thread-local error text, device probe.
#include "common.h"

namespace mmseg {
char* last_error_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}
}  // namespace mmseg

extern "C" int mmseg_version(void) { return MMSEG_ABI_VERSION; }

extern "C" const char* mmseg_last_error(void) { return mmseg::last_error_buf(); }

extern "C" int mmseg_device_ok(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n < 1) {
    cudaGetLastError();
    return 0;
  }
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10 ? 1 : 0;
}
