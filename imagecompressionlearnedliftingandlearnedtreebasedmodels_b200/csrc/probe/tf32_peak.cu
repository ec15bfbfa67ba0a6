// tf32_peak.cu -- measured dense TF32 tensor-pipe peak: the denominator of the 3xTF32 kernels' rooflines
// (MEASURED_PEAKS.json carries the bf16 cuBLAS figure only; "bf16 / 2" was a derived number).
// One CTA per SM; A (128 x 32 tf32 per k-block) and B (N x 32 tf32) sit in shared memory as K-major SWIZZLE_128B
// tiles (content irrelevant: zeros), the accumulator in tensor memory; a single thread issues `iters` chains of
// 4 * kblocks tcgen05.mma kind::tf32 M128 x N x K8 and one commit per chain.  No global traffic, no epilogue: what is
// timed is the tensor pipe fed from shared memory, i.e. the ceiling an SS-mode TF32 GEMM can reach on this part.
#include <stdint.h>

#include "../ll_common.cuh"
#include "../../../include/ll_probe.h"
#include "../tc_ptx.cuh"

namespace ll {

constexpr int PK_MAXKB = 2;                                  // distinct k-block buffers cycled through
constexpr int PK_A_BYTES = 128 * 128;                        // 128 rows x 128 B
constexpr int PK_B_BYTES = 256 * 128;
constexpr int PK_SMEM = 1024 + PK_MAXKB * (PK_A_BYTES + PK_B_BYTES) + 64;

__global__ void __launch_bounds__(128, 1) tf32_peak_kernel(float* out, int iters, int kblocks, int n) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - raw);
  const uint32_t bar = base + PK_MAXKB * (PK_A_BYTES + PK_B_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + PK_MAXKB * (PK_A_BYTES + PK_B_BYTES) + 16);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < PK_MAXKB * (PK_A_BYTES + PK_B_BYTES) / 4; i += 128) reinterpret_cast<float*>(gen)[i] = 0.f;
  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    tmem_alloc(smem_u32(tmem_slot), 256);
    tmem_relinquish();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (warp == 1 && elect_one()) {
    // D fp32, A/B tf32, both K-major, N = n, M = 128
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    uint32_t phase = 0;
    for (int it = 0; it < iters; ++it) {
      for (int kb = 0; kb < kblocks; ++kb) {
        const uint32_t sa = base + (kb % PK_MAXKB) * (PK_A_BYTES + PK_B_BYTES);
        const uint64_t adesc = umma_desc_sw128(sa), bdesc = umma_desc_sw128(sa + PK_A_BYTES);
#pragma unroll
        for (int k = 0; k < 4; ++k) tc_mma_tf32_ss(tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (uint32_t)((kb | k) != 0));
      }
      tc_commit(bar);
      mbar_wait(bar, phase);
      phase ^= 1;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    if (tid == 0 && out) out[0] = 1.f;
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}

}  // namespace ll

extern "C" int ll_tf32_peak_probe(float* out, int blocks, int iters, int kblocks, int n, ll_stream_t stream) {
  using namespace ll;
  if (blocks <= 0 || iters <= 0 || kblocks <= 0 || n < 16 || n > 256 || n % 16)
    return fail(LL_EINVAL, "ll_tf32_peak_probe: bad arguments (n in 16..256, multiple of 16)");
  LL_CUDA_OK(cudaFuncSetAttribute(tf32_peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PK_SMEM));
  tf32_peak_kernel<<<blocks, 128, PK_SMEM, as_stream(stream)>>>(out, iters, kblocks, n);
  LL_LAUNCH_OK("tf32_peak_kernel");
  return LL_OK;
}
