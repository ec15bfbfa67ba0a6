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

// GaussianConditional.forward + -log2 for one coefficient, in the reference's operation order (compressai 1.2.1; call
// sites LiftingBasedDWT_net.py:334-335,345-346,364-365).  Shared by rate.cu and the fused cgp tail of igemm_conv.cu.
__device__ __forceinline__ float gauss_bits(float xv, float sg, float mu, const float* noise, float& y) {
  if (noise) y = __fadd_rn(xv, *noise);
  else y = __fadd_rn(rintf(__fsub_rn(xv, mu)), mu);
  const float v = fabsf(__fsub_rn(y, mu));
  const float s = fmaxf(sg, 0.11f);
  const float kc = -0.70710678118654752440f;  // float(-(2 ** -0.5))
  const float up = 0.5f * erfcf(kc * __fdiv_rn(0.5f - v, s));
  const float lo = 0.5f * erfcf(kc * __fdiv_rn(-0.5f - v, s));
  const float pr = fmaxf(up - lo, 1e-9f);
  return -log2f(pr);
}

}  // namespace ll
