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

// Persistent, double-buffered: smem = in[2] | lo | hi.  The phase functions address the input
// tile at sm + DFF_SM_IN and lo/hi at fixed offsets, so buffer b is presented by shifting the
// base pointer: lo/hi live at the same absolute place for both (layout below).
constexpr int DFF_IN_FLOATS = DWF_R * DFF_PI;
constexpr int DFF_PIPE_TOTAL = 2 * DFF_IN_FLOATS + DWF_R * DFF_PM;

// Tap table in global memory, statically initialised at module load and deliberately not const: the kernels load
// it into registers once, and the compiler cannot fold the values back into per-instruction immediates.
// analysis (dec_lo[k], dec_hi[k]) pairs k = 1..9, then rec_lo[0..9], rec_hi[0..9]
#define LL_DP(k) LL_DEC_LO(k), LL_DEC_HI(k)
__device__ float g_dwt_taps[18 + 20] = {
    LL_DP(1), LL_DP(2), LL_DP(3), LL_DP(4), LL_DP(5), LL_DP(6), LL_DP(7), LL_DP(8), LL_DP(9),
    LL_REC_LO(0), LL_REC_LO(1), LL_REC_LO(2), LL_REC_LO(3), LL_REC_LO(4), LL_REC_LO(5), LL_REC_LO(6), LL_REC_LO(7), LL_REC_LO(8), LL_REC_LO(9),
    LL_REC_HI(0), LL_REC_HI(1), LL_REC_HI(2), LL_REC_HI(3), LL_REC_HI(4), LL_REC_HI(5), LL_REC_HI(6), LL_REC_HI(7), LL_REC_HI(8), LL_REC_HI(9)};
#undef LL_DP

// ---- persistent fast kernels ------------------------------------------------------------------------------------
// Double-buffered: the next tile is prefetched with 16-byte LDGSTS copies while the current one is filtered (without
// it every CTA of a wave loads, filters and stores in lock-step and DRAM idles during the compute phases).  Bulk
// copies (UBLKCP) were tried and rejected: one per tile row is needed, each is a uniform-datapath instruction that the
// compiler serialises over the issuing lanes (~25 issue slots per row against ~10 for the row's 34 LDGSTS lanes).
// A single-buffer variant (44 KB, 4 resident CTAs, next copy issued behind the row pass) measured the same.
// The tile index advances incrementally (no division per tile); three phases, two block barriers per tile -- the next
// copy into a buffer and the next first-pass write into ``mid`` are both issued behind a barrier that every thread
// only reaches after it finished reading them.
// Programmatic dependent launch: the levels of a multi-level call are tiny launches in one stream (levels 2-3 of a 512x768
// batch run ~6 us each, half of it launch latency).  Every fast kernel lets its successor be scheduled right away
// (launch_dependents at entry) and itself waits for its predecessor's memory (wait) only after its own prologue, so the
// next level's CTAs are resident and set up when the previous level drains.  Launched with the programmatic-stream-
// serialization attribute; without it (or behind a kernel that never signals) both instructions are no-ops.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

struct CopyAsync16 {
  __device__ __forceinline__ void operator()(float* dst, const float* src) const {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
  }
};
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

struct TileWalk {   // tile coordinates advanced by gridDim.x tiles per iteration without dividing
  int x, y, n, sx, sy, sn;
  __device__ __forceinline__ TileWalk(const DwtParams& p, unsigned t, unsigned step) {
    const unsigned tx = p.tiles_x, ty = p.tiles_y;
    unsigned q = t / tx;
    x = t - q * tx;
    n = q / ty;
    y = q - n * ty;
    q = step / tx;
    sx = step - q * tx;
    sn = q / ty;
    sy = q - sn * ty;
  }
  __device__ __forceinline__ void advance(const DwtParams& p) {
    x += sx;
    int c = x >= p.tiles_x;
    x -= c ? p.tiles_x : 0;
    y += sy + c;
    c = y >= p.tiles_y;
    y -= c ? p.tiles_y : 0;
    n += sn + c;
  }
  __device__ __forceinline__ DwtTile tile() const { return DwtTile{n, y * DW_TY, x * DW_TX}; }
};

__global__ void __launch_bounds__(DW_THREADS) dwt97_fwd_fast_kernel(const __grid_constant__ DwtParams p) {
  extern __shared__ __align__(16) float sm[];
  // layout: [in0][in1][mid]
  const int tid = threadIdx.x;
  pdl_launch_dependents();
  DwtTaps tp;
#pragma unroll
  for (int k = 0; k < 9; ++k) tp.d[k] = f2{g_dwt_taps[2 * k], g_dwt_taps[2 * k + 1]};
  const unsigned ntiles = (unsigned)p.N * p.tiles_x * p.tiles_y;
  unsigned t = blockIdx.x;
  float* mid = sm + 2 * DFF_IN_FLOATS - DFF_SM_LO;  // so that mid + DFF_SM_LO lands after both inputs
  TileWalk cur(p, t, gridDim.x), nxt = cur;
  pdl_wait();
  dwtff_load(p, cur.tile(), sm, tid, CopyAsync16());
  cp_commit();
  for (unsigned it = 0;; ++it) {
    const unsigned b = it & 1;
    const unsigned tn = t + gridDim.x;
    nxt.advance(p);
    if (tn < ntiles) dwtff_load(p, nxt.tile(), sm + (b ^ 1) * DFF_IN_FLOATS, tid, CopyAsync16());
    cp_commit();
    cp_wait<1>();
    __syncthreads();
    dwtff_rows(sm + b * DFF_IN_FLOATS, mid, tid, tp);
    __syncthreads();
    dwtff_cols(p, cur.tile(), mid, tid, tp);
    if (tn >= ntiles) break;
    t = tn;
    cur = nxt;
  }
}

constexpr int DIF_SB_FLOATS = 4 * DWI_R * DIF_P;
constexpr int DIF_PIPE_TOTAL = 2 * DIF_SB_FLOATS + 2 * 2 * DW_TY * DIF_P;

__global__ void __launch_bounds__(DIF_THREADS) dwt97_inv_fast_kernel(const __grid_constant__ DwtParams p) {
  extern __shared__ __align__(16) float sm[];
  const int tid = threadIdx.x;
  pdl_launch_dependents();
  DwtSynTaps tp;
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    tp.l[k] = f2{g_dwt_taps[18 + 2 * k], g_dwt_taps[18 + 2 * k + 1]};
    tp.h[k] = f2{g_dwt_taps[28 + 2 * k], g_dwt_taps[28 + 2 * k + 1]};
  }
  const unsigned ntiles = (unsigned)p.N * p.tiles_x * p.tiles_y;
  unsigned t = blockIdx.x;
  float* mid = sm + 2 * DIF_SB_FLOATS - DIF_SM_LO;
  TileWalk cur(p, t, gridDim.x), nxt = cur;
  pdl_wait();
  dwtif_load(p, cur.tile(), sm, tid, CopyAsync16());
  cp_commit();
  for (unsigned it = 0;; ++it) {
    const unsigned b = it & 1;
    const unsigned tn = t + gridDim.x;
    nxt.advance(p);
    if (tn < ntiles) dwtif_load(p, nxt.tile(), sm + (b ^ 1) * DIF_SB_FLOATS, tid, CopyAsync16());
    cp_commit();
    cp_wait<1>();
    __syncthreads();
    dwtif_cols(sm + b * DIF_SB_FLOATS, mid, tid, tp);
    __syncthreads();
    dwtif_rows(p, cur.tile(), mid, tid, tp);
    if (tn >= ntiles) break;
    t = tn;
    cur = nxt;
  }
}

static int launch_pdl(void (*kernel)(const DwtParams), unsigned grid, unsigned block, size_t smem, cudaStream_t stream, const DwtParams& p) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  LL_CUDA_OK(cudaLaunchKernelEx(&cfg, kernel, p));
  return LL_OK;
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
  constexpr size_t smem_fast = DFF_PIPE_TOTAL * sizeof(float);
  static thread_local bool attr[64] = {false};
  int dev = 0;
  LL_CUDA_OK(cudaGetDevice(&dev));
  if (dev < 64 && !attr[dev]) {
    LL_CUDA_OK(cudaFuncSetAttribute(dwt97_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    LL_CUDA_OK(cudaFuncSetAttribute(dwt97_fwd_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_fast));
    attr[dev] = true;
  }
  const long long pgrid = (long long)sm_count_cached() * 3;   // 3 resident CTAs per SM (70 KB each)
  if (dwt_fast_ok(p)) {
    rc = launch_pdl(dwt97_fwd_fast_kernel, (unsigned)(tiles < pgrid ? tiles : pgrid), DW_THREADS, smem_fast, as_stream(stream), p);
    if (rc) return rc;
  } else dwt97_fwd_kernel<<<(unsigned)tiles, DW_THREADS, smem, as_stream(stream)>>>(p);
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
  constexpr size_t smem_fast = DIF_PIPE_TOTAL * sizeof(float);
  static_assert(smem <= 48 * 1024, "generic inverse kernel fits the default shared-memory limit");
  static thread_local bool attr[64] = {false};
  int dev = 0;
  LL_CUDA_OK(cudaGetDevice(&dev));
  if (dev < 64 && !attr[dev]) {
    LL_CUDA_OK(cudaFuncSetAttribute(dwt97_inv_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_fast));
    attr[dev] = true;
  }
  const long long pgrid = (long long)sm_count_cached() * 3;
  if (dwt_fast_ok(p)) {
    rc = launch_pdl(dwt97_inv_fast_kernel, (unsigned)(tiles < pgrid ? tiles : pgrid), DIF_THREADS, smem_fast, as_stream(stream), p);
    if (rc) return rc;
  } else dwt97_inv_kernel<<<(unsigned)tiles, DW_THREADS, smem, as_stream(stream)>>>(p);
  LL_LAUNCH_OK("dwt97_inv_kernel");
  return LL_OK;
}


size_t ll_dwt97_scratch_floats(int N, int h, int w, int J) {
  // intermediate LL planes of levels 0..J-2 (forward) / reconstructed LL planes of levels J-1..1 (inverse)
  size_t n = 0;
  for (int j = 1; j < J; ++j) n += (size_t)N * (h >> j) * (w >> j);
  return n;
}

int ll_dwt97_fwd(const float* x, float* yl, float* const* yh, float* scratch, int N, int h, int w, int J,
                 ll_stream_t stream) {
  if (J < 1 || J > 16) return fail(LL_EINVAL, "ll_dwt97_fwd: J out of range");
  if ((h % (1 << J)) || (w % (1 << J))) return fail(LL_EINVAL, "ll_dwt97_fwd: h, w must be divisible by 2^J (got %d x %d, J=%d)", h, w, J);
  if ((long long)N * h * w == 0) return LL_OK;
  if (!x || !yl || !yh || (J > 1 && !scratch)) return fail(LL_EINVAL, "ll_dwt97_fwd: null pointer");
  const float* cur = x;
  float* sc = scratch;
  for (int j = 0; j < J; ++j) {
    const int hh = h >> j, ww = w >> j;
    const long long sub = (long long)(hh / 2) * (ww / 2);
    float* out_ll = (j == J - 1) ? yl : sc;
    if (!yh[j]) return fail(LL_EINVAL, "ll_dwt97_fwd: null yh[%d]", j);
    int rc = ll_dwt97_fwd_level(cur, (long long)hh * ww, out_ll, sub, yh[j], 3 * sub, N, hh, ww, stream);
    if (rc) return rc;
    cur = out_ll;
    if (j < J - 1) sc += (size_t)N * sub;
  }
  return LL_OK;
}

int ll_dwt97_inv(const float* yl, const float* const* yh, float* x, float* scratch, int N, int h, int w, int J,
                 ll_stream_t stream) {
  if (J < 1 || J > 16) return fail(LL_EINVAL, "ll_dwt97_inv: J out of range");
  if ((h % (1 << J)) || (w % (1 << J))) return fail(LL_EINVAL, "ll_dwt97_inv: h, w must be divisible by 2^J (got %d x %d, J=%d)", h, w, J);
  if ((long long)N * h * w == 0) return LL_OK;
  if (!x || !yl || !yh || (J > 1 && !scratch)) return fail(LL_EINVAL, "ll_dwt97_inv: null pointer");
  const float* cur = yl;
  float* sc = scratch;
  for (int j = J - 1; j >= 0; --j) {
    const int hh = h >> j, ww = w >> j;   // output size of this level
    const long long sub = (long long)(hh / 2) * (ww / 2);
    float* out = (j == 0) ? x : sc;
    if (!yh[j]) return fail(LL_EINVAL, "ll_dwt97_inv: null yh[%d]", j);
    int rc = ll_dwt97_inv_level(cur, sub, yh[j], 3 * sub, out, (long long)hh * ww, N, hh, ww, stream);
    if (rc) return rc;
    cur = out;
    if (j > 0) sc += (size_t)N * hh * ww;
  }
  return LL_OK;
}

}  // extern "C"
