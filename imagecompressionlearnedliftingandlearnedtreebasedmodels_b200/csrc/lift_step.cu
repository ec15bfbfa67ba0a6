// lift_step.cu -- learned lifting (K2): kernel wrapper, weight packing, level composition.
#include "lift_step_body.cuh"
#include "lift_level_body.h"
#include "ll_common.cuh"

namespace ll {

int launch_lift_step_tc(const LiftParams& p, cudaStream_t stream);                                  // lift_tc.cu
int launch_pack_lift_tc(const float* w2, const float* w3, const float* w4, float* blob, cudaStream_t stream);         // lift_tc.cu
#ifdef LL_DEBUG   // timing experiments of the tensor-core kernel (scripts/gpu_lift_tc_time*.py); not in release builds
static int g_lift_dbg = 0;
static long long* g_lift_dbg_buf = nullptr;
#endif

__global__ void __launch_bounds__(LS_THREADS, 1) lift_step_kernel(const __grid_constant__ LiftParams p) {
  extern __shared__ __align__(16) float sm[];
  const int tid = threadIdx.x;
#define LL_PHASE(call) \
  do {                 \
    call;              \
    __syncthreads();   \
  } while (0)
  LL_LIFT_STEP_DRIVER(LL_PHASE, p, sm, blockIdx.x, gridDim.x);
#undef LL_PHASE
}

__global__ void pack_lift_step_kernel(const float* __restrict__ pre, const float* __restrict__ w1,
                                      const float* __restrict__ b1, const float* __restrict__ w2,
                                      const float* __restrict__ b2, const float* __restrict__ w3,
                                      const float* __restrict__ b3, const float* __restrict__ w4,
                                      const float* __restrict__ b4, float* __restrict__ blob) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < BL_TOTAL; i += gridDim.x * blockDim.x)
    blob[i] = pack_lift_elem(i, pre, w1, b1, w2, b2, w3, b3, w4, b4);
}

__global__ void scale2_kernel(float* __restrict__ a, long long a_sb, long long a_sy, long long a_sx,
                              int nb, int ny, int nx, const float* __restrict__ n, float base, int divide) {
  // wavelet_forward_v2.py:76-80 / wavelet_inverse_v2.py:70-74:  x * (base + 0.1 n)  or  x / (...)
  const float s = __fadd_rn(base, __fmul_rn(n[0], 0.1f));  // torch: python float + tensor * 0.1, two roundings
  const long long total = (long long)nb * ny * nx;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = int(i % nx);
    const long long t = i / nx;
    const int y = int(t % ny);
    const int b = int(t / ny);
    float* q = a + b * a_sb + y * a_sy + x * a_sx;
    *q = divide ? (*q / s) : (*q * s);
  }
}

static int launch_lift_step(const ll_lift_job* jobs, int njobs, const float* blob, float sign, float rw, int linear,
                            int precision, cudaStream_t stream) {
  if (precision != LL_LIFT_FP32 && precision != LL_LIFT_TC && precision != LL_LIFT_TC16)
    return fail(LL_EINVAL, "ll_lift_step: unknown precision %d (LL_LIFT_FP32 = 0, LL_LIFT_TC = 1, LL_LIFT_TC16 = 2)", precision);
  if (njobs < 1 || njobs > 2) return fail(LL_EINVAL, "ll_lift_step: njobs must be 1 or 2 (got %d)", njobs);
  if (!blob) return fail(LL_EINVAL, "ll_lift_step: null blob");
  LiftParams p;
  memset(&p, 0, sizeof(p));
  p.njobs = njobs;
  p.blob = blob;
  p.sign = sign;
  p.rw = rw;
  p.linear = linear;
  p.total_units = 0;
  for (int j = 0; j < 2; ++j) {
    if (j < njobs) {
      const ll_lift_job& J = jobs[j];
      if (J.nb < 0 || J.ny < 0 || J.nx < 0) return fail(LL_EINVAL, "ll_lift_step: negative extent");
      if (!J.src.ptr || !J.din.ptr || !J.dout.ptr) {
        if ((long long)J.nb * J.ny * J.nx != 0) return fail(LL_EINVAL, "ll_lift_step: null view pointer");
      }
      p.job[j] = J;
      p.nstrips[j] = (J.nx + LS_WT - 1) / LS_WT;
      p.nchunks[j] = (J.ny + LS_R - 1) / LS_R;
      p.units[j] = (long long)J.nb * p.nstrips[j] * p.nchunks[j];
    } else {
      p.units[j] = 0;
      p.nstrips[j] = p.nchunks[j] = 1;
    }
    p.total_units += p.units[j];
  }
  if (p.total_units == 0) return LL_OK;  // empty input: nothing to do
#ifdef LL_DEBUG
  p.dbg = g_lift_dbg;
  p.dbg_buf = g_lift_dbg_buf;
#endif
  p.f16 = precision == LL_LIFT_TC16;
  if (precision != LL_LIFT_FP32) return launch_lift_step_tc(p, stream);
  static thread_local bool attr_set[64] = {false};
  int dev = 0;
  LL_CUDA_OK(cudaGetDevice(&dev));
  if (dev < 64 && !attr_set[dev]) {
    LL_CUDA_OK(cudaFuncSetAttribute(lift_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LS_SMEM_BYTES));
    attr_set[dev] = true;
  }
  long long grid = sm_count_cached();
  if (grid > p.total_units) grid = p.total_units;
  lift_step_kernel<<<(unsigned)grid, LS_THREADS, LS_SMEM_BYTES, stream>>>(p);
  LL_LAUNCH_OK("lift_step_kernel");
  return LL_OK;
}

static int launch_scale(ll_view3 v, int nb, int ny, int nx, const float* n, float base, int divide, cudaStream_t st) {
  const long long total = (long long)nb * ny * nx;
  if (total == 0) return LL_OK;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  scale2_kernel<<<blocks, 256, 0, st>>>(v.ptr, v.sb, v.sy, v.sx, nb, ny, nx, n, base, divide);
  LL_LAUNCH_OK("scale2_kernel");
  return LL_OK;
}

struct CudaBackend {
  cudaStream_t st;
  float rw;
  int linear;
  int precision;
  int step(const ll_lift_job* jobs, int n, const float* blob, float sign) {
    return launch_lift_step(jobs, n, blob, sign, rw, linear, precision, st);
  }
  int scale(ll_view3 v, int nb, int ny, int nx, const float* n, float base, int divide) {
    return launch_scale(v, nb, ny, nx, n, base, divide, st);
  }
  int copy(float* dst, const float* src, size_t count) {
    LL_CUDA_OK(cudaMemcpyAsync(dst, src, count * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return LL_OK;
  }
};

}  // namespace ll

using namespace ll;

extern "C" {

int ll_pack_lift_step(const float* pre_w, const float* w1, const float* b1, const float* w2, const float* b2,
                      const float* w3, const float* b3, const float* w4, const float* b4, float* blob,
                      ll_stream_t stream) {
  if (!pre_w || !w1 || !b1 || !w2 || !b2 || !w3 || !b3 || !w4 || !b4 || !blob)
    return fail(LL_EINVAL, "ll_pack_lift_step: null pointer");
  pack_lift_step_kernel<<<(BL_TOTAL + 255) / 256, 256, 0, as_stream(stream)>>>(pre_w, w1, b1, w2, b2, w3, b3, w4, b4, blob);
  LL_LAUNCH_OK("pack_lift_step_kernel");
  return launch_pack_lift_tc(w2, w3, w4, blob, as_stream(stream));
}

#ifdef LL_DEBUG
int ll_dbg_lift_switches(int bits) {   // bit 0 no MMA, 1 no E-B, 2 no conv1, 3 no conv4, 4 no E-A: results are WRONG when non-zero
  g_lift_dbg = bits;
  return LL_OK;
}
int ll_dbg_lift_stamp_buffer(long long* buf) {   // per-warp phase timestamps (17 x 8 int64) of CTA 0
  g_lift_dbg_buf = buf;
  return LL_OK;
}
#endif

int ll_lift_step(const ll_lift_job* jobs, int njobs, const float* blob, float sign, float res_weight, int linear,
                 int precision, ll_stream_t stream) {
  if (!jobs) return fail(LL_EINVAL, "ll_lift_step: null jobs");
  return launch_lift_step(jobs, njobs, blob, sign, res_weight, linear, precision, as_stream(stream));
}

size_t ll_lift_level_scratch_floats(int B, int h, int w) { return lift_level_scratch_floats(B, h, w); }

int ll_lift_level_fwd(const float* x, int64_t x_sb, float* llp, int64_t ll_sb, float* yh, int64_t yh_sb,
                      float* scratch, int B, int h, int w, const float* const* blobs, float rw, int linear,
                      int scale, const float* nh, const float* nl, int precision, ll_stream_t stream) {
  if (B < 0 || h < 0 || w < 0 || (h & 1) || (w & 1)) return fail(LL_EINVAL, "ll_lift_level_fwd: h, w must be even (got %d x %d)", h, w);
  if ((long long)B * h * w == 0) return LL_OK;
  if (!x || !llp || !yh || !scratch || !blobs) return fail(LL_EINVAL, "ll_lift_level_fwd: null pointer");
  if (scale && (!nh || !nl)) return fail(LL_EINVAL, "ll_lift_level_fwd: scale=1 needs nh, nl");
  CudaBackend be{as_stream(stream), rw, linear, precision};
  return lift_level_fwd_impl(be, x, x_sb, llp, ll_sb, yh, yh_sb, scratch, B, h, w, blobs, scale, nh, nl);
}

int ll_lift_level_inv(const float* llp, int64_t ll_sb, const float* yh, int64_t yh_sb, float* x, int64_t x_sb,
                      float* scratch, int B, int h, int w, const float* const* blobs, float rw, int linear,
                      int scale, const float* nh, const float* nl, int precision, ll_stream_t stream) {
  if (B < 0 || h < 0 || w < 0 || (h & 1) || (w & 1)) return fail(LL_EINVAL, "ll_lift_level_inv: h, w must be even (got %d x %d)", h, w);
  if ((long long)B * h * w == 0) return LL_OK;
  if (!x || !llp || !yh || !scratch || !blobs) return fail(LL_EINVAL, "ll_lift_level_inv: null pointer");
  if (scale && (!nh || !nl)) return fail(LL_EINVAL, "ll_lift_level_inv: scale=1 needs nh, nl");
  CudaBackend be{as_stream(stream), rw, linear, precision};
  return lift_level_inv_impl(be, llp, ll_sb, yh, yh_sb, x, x_sb, scratch, B, h, w, blobs, scale, nh, nl);
}

}  // extern "C"
