// dwt97.cu -- CDF 9/7 fixed-filter DWT (K1): kernels + C ABI.
#include "dwt97_body.cuh"
#include "ll_common.cuh"

namespace ll {

__global__ void __launch_bounds__(DW_THREADS) dwt97_fwd_kernel(const __grid_constant__ DwtParams p) {
  extern __shared__ __align__(16) float sm[];
  const int tid = threadIdx.x;
  const DwtTile t = dwt_tile(p, blockIdx.x);
  dwtf_load(p, t, sm, tid);
  __syncthreads();
  dwtf_rows(sm, tid);
  __syncthreads();
  dwtf_cols(p, t, sm, tid);
}

__global__ void __launch_bounds__(DW_THREADS) dwt97_inv_kernel(const __grid_constant__ DwtParams p) {
  extern __shared__ __align__(16) float sm[];
  const int tid = threadIdx.x;
  const DwtTile t = dwt_tile(p, blockIdx.x);
  dwti_load(p, t, sm, tid);
  __syncthreads();
  dwti_cols(p, t, sm, tid);
  __syncthreads();
  dwti_rows(p, t, sm, tid);
}

__global__ void __launch_bounds__(DW_THREADS) dwt97_fwd_fast_kernel(const __grid_constant__ DwtParams p) {
  extern __shared__ __align__(16) float sm[];
  const int tid = threadIdx.x;
  const DwtTile t = dwt_tile(p, blockIdx.x);
  dwtff_load(p, t, sm, tid);
  __syncthreads();
  dwtff_rows(sm, tid);
  __syncthreads();
  dwtff_cols(p, t, sm, tid);
}

__global__ void __launch_bounds__(DW_THREADS) dwt97_inv_fast_kernel(const __grid_constant__ DwtParams p) {
  extern __shared__ __align__(16) float sm[];
  const int tid = threadIdx.x;
  const DwtTile t = dwt_tile(p, blockIdx.x);
  dwtif_load(p, t, sm, tid);
  __syncthreads();
  dwtif_cols(sm, tid);
  __syncthreads();
  dwtif_rows(p, t, sm, tid);
}

static int fill(DwtParams& p, int N, int h, int w, const char* who) {
  if (N < 0 || h < 0 || w < 0 || (h & 1) || (w & 1)) return fail(LL_EINVAL, "%s: h, w must be even and non-negative (got %d x %d)", who, h, w);
  p.N = N;
  p.h = h;
  p.w = w;
  p.tiles_x = (w / 2 + DW_TX - 1) / DW_TX;
  p.tiles_y = (h / 2 + DW_TY - 1) / DW_TY;
  return LL_OK;
}

}  // namespace ll

using namespace ll;

extern "C" {

int ll_dwt97_fwd_level(const float* x, int64_t x_sn, float* llp, int64_t ll_sn, float* yh, int64_t yh_sn, int N,
                       int h, int w, ll_stream_t stream) {
  DwtParams p = {};
  int rc = fill(p, N, h, w, "ll_dwt97_fwd_level");
  if (rc) return rc;
  const long long tiles = (long long)N * p.tiles_x * p.tiles_y;
  if (tiles == 0) return LL_OK;
  if (!x || !llp || !yh) return fail(LL_EINVAL, "ll_dwt97_fwd_level: null pointer");
  if (tiles > 0x7fffffffLL) return fail(LL_EINVAL, "ll_dwt97_fwd_level: too many tiles");
  p.x = x;
  p.x_sn = x_sn;
  p.llo = llp;
  p.ll_sn = ll_sn;
  p.yho = yh;
  p.yh_sn = yh_sn;
  constexpr size_t smem = DWF_SM_TOTAL * sizeof(float);
  constexpr size_t smem_fast = DFF_SM_TOTAL * sizeof(float);
  static thread_local bool attr[64] = {false};
  int dev = 0;
  LL_CUDA_OK(cudaGetDevice(&dev));
  if (dev < 64 && !attr[dev]) {
    LL_CUDA_OK(cudaFuncSetAttribute(dwt97_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LL_CUDA_OK(cudaFuncSetAttribute(dwt97_fwd_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_fast));
    attr[dev] = true;
  }
  if (dwt_fast_ok(p)) dwt97_fwd_fast_kernel<<<(unsigned)tiles, DW_THREADS, smem_fast, as_stream(stream)>>>(p);
  else dwt97_fwd_kernel<<<(unsigned)tiles, DW_THREADS, smem, as_stream(stream)>>>(p);
  LL_LAUNCH_OK("dwt97_fwd_kernel");
  return LL_OK;
}

int ll_dwt97_inv_level(const float* llp, int64_t ll_sn, const float* yh, int64_t yh_sn, float* x, int64_t x_sn, int N,
                       int h, int w, ll_stream_t stream) {
  DwtParams p = {};
  int rc = fill(p, N, h, w, "ll_dwt97_inv_level");
  if (rc) return rc;
  const long long tiles = (long long)N * p.tiles_x * p.tiles_y;
  if (tiles == 0) return LL_OK;
  if (!x || !llp || !yh) return fail(LL_EINVAL, "ll_dwt97_inv_level: null pointer");
  if (tiles > 0x7fffffffLL) return fail(LL_EINVAL, "ll_dwt97_inv_level: too many tiles");
  p.xo = x;
  p.x_sn = x_sn;
  p.ll = llp;
  p.ll_sn = ll_sn;
  p.yh = yh;
  p.yh_sn = yh_sn;
  constexpr size_t smem = DWI_SM_TOTAL * sizeof(float);
  constexpr size_t smem_fast = DIF_SM_TOTAL * sizeof(float);
  static_assert(smem <= 48 * 1024 && smem_fast <= 48 * 1024, "inverse kernels fit the default shared-memory limit");
  if (dwt_fast_ok(p)) dwt97_inv_fast_kernel<<<(unsigned)tiles, DW_THREADS, smem_fast, as_stream(stream)>>>(p);
  else dwt97_inv_kernel<<<(unsigned)tiles, DW_THREADS, smem, as_stream(stream)>>>(p);
  LL_LAUNCH_OK("dwt97_inv_kernel");
  return LL_OK;
}

}  // extern "C"
