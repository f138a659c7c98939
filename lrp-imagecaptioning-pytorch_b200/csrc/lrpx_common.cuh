// Shared helpers for liblrpx (sm_100a).  No torch types anywhere in this library.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/lrpx.h"

namespace lrpx {

void set_error(const char* fmt, ...);

#define LRPX_CHECK_ARG(cond, msg)                                  \
  do {                                                             \
    if (!(cond)) {                                                 \
      lrpx::set_error("%s: %s", __func__, msg);                    \
      return LRPX_E_INVALID;                                       \
    }                                                              \
  } while (0)

#define LRPX_CHECK_LAUNCH()                                                        \
  do {                                                                             \
    cudaError_t e__ = cudaGetLastError();                                          \
    if (e__ != cudaSuccess) {                                                      \
      lrpx::set_error("%s: CUDA launch failed: %s", __func__, cudaGetErrorString(e__)); \
      return LRPX_E_CUDA;                                                          \
    }                                                                              \
  } while (0)

// LRPtools/utils.py:16-18
__device__ __forceinline__ float safe_div(float num, float den) {
  return num / (den + (den == 0.f ? LRPX_Z_EPSILON : 0.f));
}
// gridTDmodel.py:757-759 / lrp_modules.py:19-20 :  z + eps*sign(z), exact zero -> eps
__device__ __forceinline__ float stab(float z) {
  float s = (z > 0.f) ? LRPX_EPSILON : ((z < 0.f) ? -LRPX_EPSILON : 0.f);
  float o = s + z;
  return o == 0.f ? LRPX_EPSILON : o;
}

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }
static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

}  // namespace lrpx
