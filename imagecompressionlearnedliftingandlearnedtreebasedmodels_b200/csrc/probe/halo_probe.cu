// halo_probe.cu -- unit probe of shifted-window operand addressing for a 3x3 implicit GEMM on tcgen05 (no reference
// counterpart): an 18 x `pitch` pixel halo tile of a 32-channel fp32 NHWC tensor is brought in by ONE TMA load
// (SWIZZLE_128B, one 128-byte row per pixel) and the A operand of tap (dy, dx) -- M = 128 rows = 16 image rows x 8 pixels --
// is addressed inside it with a K-major SWIZZLE_128B descriptor whose start address is moved by whole 128-byte rows
// ((dy+1) * pitch + dx+1) and whose stride byte offset is the halo row pitch, instead of being fetched again per tap.
// Whether the tensor core applies the 128-byte swizzle to absolute shared-memory address bits (then any row offset / any
// pitch works), or relative to the descriptor start with the 3-bit `base_offset` field, is not stated precisely enough in
// the documentation available offline: this probe measures it.  D (128, 32) = A_tap (128, 32) x B (32, 32)^T.
#include <cuda.h>
#include <stdint.h>

#include "../ll_common.cuh"
#include "../../../include/ll_probe.h"
#include "../tc_ptx.cuh"

namespace ll {

constexpr int HP_ROWS = 18, HP_MAXPITCH = 16;
constexpr int HP_A_BYTES = HP_ROWS * HP_MAXPITCH * 128;
constexpr int HP_SMEM = 1024 + HP_A_BYTES + 4096 + 64;

__device__ __forceinline__ uint64_t halo_desc(uint32_t saddr, uint32_t sbo, uint32_t base_offset) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base_offset & 7) << 49;
  d |= (uint64_t)2 << 61;
  return d;
}

__global__ void __launch_bounds__(128, 1)
halo_probe_kernel(const __grid_constant__ CUtensorMap tmA, const float* __restrict__ b, float* __restrict__ d, int pitch,
                  int dy, int dx, int bo_mode) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - raw);
  const uint32_t a_s = base, b_s = base + HP_A_BYTES;
  const uint32_t bar0 = b_s + 4096, bar1 = bar0 + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + HP_A_BYTES + 4096 + 16);
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    tmem_alloc(smem_u32(tmem_slot), 32);
    tmem_relinquish();
  }
  // B (32 rows n x 32 k) -> K-major SWIZZLE_128B rows
  for (int e = tid; e < 32 * 32; e += 128) {
    const int n = e >> 5, k = e & 31;
    *reinterpret_cast<float*>(gen + HP_A_BYTES + n * 128 + ((((k >> 2) ^ (n & 7))) << 4) + ((k & 3) << 2)) = b[e];
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (tid == 0) {
    mbar_expect_tx(bar0, (uint32_t)(HP_ROWS * pitch * 128));
    tma_load_4d(a_s, &tmA, bar0, 0, 0, 0, 0);
    mbar_wait(bar0, 0);
    tc_fence_after();
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t start = a_s + (uint32_t)(((dy + 1) * pitch + dx + 1) * 128);
    const uint32_t bo = bo_mode ? (start >> 7) & 7u : 0u;
    for (int k = 0; k < 4; ++k)
      tc_mma_tf32_ss(tmem, halo_desc(start + 32 * k, (uint32_t)pitch * 128, bo), umma_desc_sw128(b_s + 32 * k), idesc, (uint32_t)(k != 0));
    tc_commit(bar1);
  }
  __syncwarp();
  mbar_wait(bar1, 0);
  tc_fence_after();
  uint32_t v[32];
  tc_ld32(tmem + ((uint32_t)(warp * 32) << 16), v);
  tc_wait_ld();
  for (int j = 0; j < 32; ++j) d[tid * 32 + j] = __uint_as_float(v[j]);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 32);
  }
}

typedef CUresult (*HaloEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace ll

using namespace ll;

// a: (18, 16, 32) fp32 channels-last halo source, b: (32, 32) [n][k], d: (128, 32); pitch in 10..16 = pixels per halo row
// fetched; (dy, dx) in -1..1; bo_mode 1 sets the descriptor's base_offset to bits 7..9 of the start address.
extern "C" int ll_halo_probe(const float* a, const float* b, float* d, int pitch, int dy, int dx, int bo_mode, ll_stream_t stream) {
  if (pitch < 10 || pitch > HP_MAXPITCH || dy < -1 || dy > 1 || dx < -1 || dx > 1) return fail(LL_EINVAL, "ll_halo_probe: bad arguments");
  void* f = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
    return fail(LL_ECUDA, "ll_halo_probe: cuTensorMapEncodeTiled not available");
  CUtensorMap tmA;
  cuuint64_t gdim[4] = {32, 16, 18, 1};
  cuuint64_t gstr[3] = {32 * 4, 16 * 32 * 4, 18 * 16 * 32 * 4};
  cuuint32_t box[4] = {32, (cuuint32_t)pitch, 18, 1};
  cuuint32_t est[4] = {1, 1, 1, 1};
  CUresult r = reinterpret_cast<HaloEncodeFn>(f)(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(a), gdim, gstr, box, est,
                                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(LL_ECUDA, "ll_halo_probe: cuTensorMapEncodeTiled failed with %d", (int)r);
  LL_CUDA_OK(cudaFuncSetAttribute(halo_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HP_SMEM));
  halo_probe_kernel<<<1, 128, HP_SMEM, as_stream(stream)>>>(tmA, b, d, pitch, dy, dx, bo_mode);
  LL_LAUNCH_OK("halo_probe_kernel");
  return LL_OK;
}
