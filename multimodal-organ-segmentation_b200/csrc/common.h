// Host-side helpers shared by the C-ABI translation units: error reporting and launch checks.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>

#include "mmseg_b200.h"

namespace mmseg {

char* last_error_buf();  // thread-local, 512 bytes (api.cu)

inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(MMSEG_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return MMSEG_OK;
}

inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace mmseg
