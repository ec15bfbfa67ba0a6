// probe.cu -- measurement helper: FP32 FMA-pipe peak of the device (register-only FFMA2 loop).
// Used by bench.py as the measured denominator of the learned-lifting kernels' roofline
// (MEASURED_PEAKS.json carries HBM and bf16 tensor peaks only).
#include "../ll_common.cuh"
#include "../../../include/ll_probe.h"

namespace ll {

__global__ void __launch_bounds__(256) fma_peak_kernel(float* out, int iters, float seed) {
  float2 acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = make_float2(seed + i, seed - i);
  const float2 a = make_float2(1.0000001f, 0.9999999f);
  const float2 b = make_float2(seed * 1e-7f, -seed * 1e-7f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int i = 0; i < 16; ++i) acc[i] = __ffma2_rn(acc[i], a, b);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i].x + acc[i].y;
  if (s == 123.456f) out[0] = s;  // keep the loop alive
}

}  // namespace ll

extern "C" {

/* Enqueues the probe; FLOPs executed = blocks * 256 * iters * 64 FFMA2 * 4 flop. */
int ll_fma_peak_probe(float* out, int blocks, int iters, ll_stream_t stream) {
  if (!out || blocks <= 0 || iters <= 0) return ll::fail(LL_EINVAL, "ll_fma_peak_probe: bad arguments");
  ll::fma_peak_kernel<<<blocks, 256, 0, ll::as_stream(stream)>>>(out, iters, 1.0f);
  LL_LAUNCH_OK("fma_peak_kernel");
  return LL_OK;
}

}  // extern "C"
