// ll_common.cuh -- shared host-side helpers of the C-ABI library (error slot, checks).
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>

#include "../../include/ll_api.h"

namespace ll {

char* err_slot();  // thread-local message buffer (ll_api.cu)
constexpr int ERR_LEN = 512;

inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_slot(), ERR_LEN, fmt, ap);
  va_end(ap);
  return code;
}

#define LL_CUDA_OK(expr)                                                                       \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess)                                                                     \
      return ll::fail(LL_ECUDA, "%s:%d %s: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
  } while (0)

#define LL_LAUNCH_OK(what)                                                                     \
  do {                                                                                         \
    cudaError_t _e = cudaGetLastError();                                                       \
    if (_e != cudaSuccess) return ll::fail(LL_ECUDA, "launch %s: %s", what, cudaGetErrorString(_e)); \
  } while (0)

int sm_count_cached();

inline cudaStream_t as_stream(ll_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

}  // namespace ll
