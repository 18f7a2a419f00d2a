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

// sizeof() of every argument struct of the ABI, so that a binding (ctypes, cgo, JNI ...) can check its own layout at load time
extern "C" int mmseg_sizeof(int which) {
  switch (which) {
    case 0: return (int)sizeof(mmseg_conv_args);
    case 1: return (int)sizeof(mmseg_wgrad_args);
    case 2: return (int)sizeof(mmseg_norm_args);
    case 3: return (int)sizeof(mmseg_norm_bwd_args);
    case 4: return (int)sizeof(mmseg_adamw_tensor);
    case 5: return (int)sizeof(mmseg_repack_desc);
    case 6: return (int)sizeof(mmseg_swin_attn_args);
    default: return -1;
  }
}
