// colour.cu -- agent-side pointwise work around the codec (SURVEY.md 8f "next #2").
//
// Reference: agents/liftingDWT_agent.py:164-186 (validate, clrch == 1 branch; the training loop :100-125 is the same
// chain) with compressai.transforms.RGB2YCbCr / YCbCr2RGB (BT.709, full range, chroma offset +0.5):
//   y    = RGB2YCbCr(x);  y[:, 0] -= 0.5                       -> model input            (:170-171)
//   xhat = YCbCr2RGB(yhat + (0.5, 0, 0)) - 0.5;  clamp(+-0.5)  -> reconstruction         (:174-181)
//   mse  = mean((x - 0.5 - xhat)^2);  PSNR = 10 log10(1 / mse)                           (:182-186, rate_dist.py:36)
// The reference runs this as ~10 elementwise launches, 3 full-image round trips and an .item() sync per scalar; here it
// is one pass in (3 planes read, 3 written) and one pass out (6 planes read, 3 optionally written, the squared error
// summed per image into a double).  Pure HBM traffic: 24 B / pixel in, 24-36 B / pixel out.
// Same operation order as the reference in IEEE fp32 (explicit _rn intrinsics: no FMA contraction).
#include "ll_common.cuh"

namespace ll {

constexpr int CL_THREADS = 256;
constexpr float CL_KR = 0.2126f, CL_KG = 0.7152f, CL_KB = 0.0722f;

__device__ __forceinline__ void rgb2ycc(float r, float g, float b, float& y, float& cb, float& cr) {
  y = __fadd_rn(__fadd_rn(__fmul_rn(CL_KR, r), __fmul_rn(CL_KG, g)), __fmul_rn(CL_KB, b));
  cb = __fadd_rn(__fdiv_rn(__fmul_rn(0.5f, __fsub_rn(b, y)), (float)(1.0 - 0.0722)), 0.5f);
  cr = __fadd_rn(__fdiv_rn(__fmul_rn(0.5f, __fsub_rn(r, y)), (float)(1.0 - 0.2126)), 0.5f);
}

__device__ __forceinline__ void ycc2rgb(float y, float cb, float cr, float& r, float& g, float& b) {
  r = __fadd_rn(y, __fmul_rn((float)(2.0 - 2.0 * 0.2126), __fsub_rn(cr, 0.5f)));
  b = __fadd_rn(y, __fmul_rn((float)(2.0 - 2.0 * 0.0722), __fsub_rn(cb, 0.5f)));
  g = __fdiv_rn(__fsub_rn(__fsub_rn(y, __fmul_rn(CL_KR, r)), __fmul_rn(CL_KB, b)), CL_KG);
}

// rgb, ycc: (B, 3, hw) planar.  VEC = 4 when hw % 4 == 0 and the bases are 16-byte aligned.
template <int VEC>
__global__ void __launch_bounds__(CL_THREADS) rgb_to_ycbcr_shift_kernel(const float* __restrict__ rgb, float* __restrict__ ycc,
                                                                        long long hw, long long total) {
  for (long long i = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * VEC; i < total; i += (long long)gridDim.x * blockDim.x * VEC) {
    const long long b = i / hw, pix = i - b * hw;
    const float* s = rgb + b * 3 * hw + pix;
    float* d = ycc + b * 3 * hw + pix;
    float r[VEC], g[VEC], bl[VEC], y[VEC], cb[VEC], cr[VEC];
    if (VEC == 4) {
      const float4 a = *reinterpret_cast<const float4*>(s), c = *reinterpret_cast<const float4*>(s + hw), e = *reinterpret_cast<const float4*>(s + 2 * hw);
      r[0] = a.x; r[1 % VEC] = a.y; r[2 % VEC] = a.z; r[3 % VEC] = a.w;
      g[0] = c.x; g[1 % VEC] = c.y; g[2 % VEC] = c.z; g[3 % VEC] = c.w;
      bl[0] = e.x; bl[1 % VEC] = e.y; bl[2 % VEC] = e.z; bl[3 % VEC] = e.w;
    } else {
      r[0] = s[0]; g[0] = s[hw]; bl[0] = s[2 * hw];
    }
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      rgb2ycc(r[k], g[k], bl[k], y[k], cb[k], cr[k]);
      y[k] = __fsub_rn(y[k], 0.5f);                       // only Y is shifted (:171)
    }
    if (VEC == 4) {
      *reinterpret_cast<float4*>(d) = make_float4(y[0], y[1 % VEC], y[2 % VEC], y[3 % VEC]);
      *reinterpret_cast<float4*>(d + hw) = make_float4(cb[0], cb[1 % VEC], cb[2 % VEC], cb[3 % VEC]);
      *reinterpret_cast<float4*>(d + 2 * hw) = make_float4(cr[0], cr[1 % VEC], cr[2 % VEC], cr[3 % VEC]);
    } else {
      d[0] = y[0]; d[hw] = cb[0]; d[2 * hw] = cr[0];
    }
  }
}

// grid.y = image: the squared error of image b is summed into sse[b] (double).
template <int VEC>
__global__ void __launch_bounds__(CL_THREADS) ycbcr_to_rgb_sse_kernel(const float* __restrict__ ycc, const float* __restrict__ rgb_ref,
                                                                      float* __restrict__ xhat, long long hw, double* __restrict__ sse) {
  __shared__ float red[CL_THREADS / 32];
  const long long b = blockIdx.y;
  const float* s = ycc + b * 3 * hw;
  const float* ref = rgb_ref ? rgb_ref + b * 3 * hw : nullptr;
  float* d = xhat ? xhat + b * 3 * hw : nullptr;
  float local = 0.f;
  for (long long pix = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * VEC; pix < hw; pix += (long long)gridDim.x * blockDim.x * VEC) {
    float y[VEC], cb[VEC], cr[VEC], o[3][VEC], x[3][VEC];
    if (VEC == 4) {
      const float4 a = *reinterpret_cast<const float4*>(s + pix), c = *reinterpret_cast<const float4*>(s + hw + pix), e = *reinterpret_cast<const float4*>(s + 2 * hw + pix);
      y[0] = a.x; y[1 % VEC] = a.y; y[2 % VEC] = a.z; y[3 % VEC] = a.w;
      cb[0] = c.x; cb[1 % VEC] = c.y; cb[2 % VEC] = c.z; cb[3 % VEC] = c.w;
      cr[0] = e.x; cr[1 % VEC] = e.y; cr[2 % VEC] = e.z; cr[3 % VEC] = e.w;
      if (ref) {
#pragma unroll
        for (int c3 = 0; c3 < 3; ++c3) {
          const float4 t = *reinterpret_cast<const float4*>(ref + c3 * hw + pix);
          x[c3][0] = t.x; x[c3][1 % VEC] = t.y; x[c3][2 % VEC] = t.z; x[c3][3 % VEC] = t.w;
        }
      }
    } else {
      y[0] = s[pix]; cb[0] = s[hw + pix]; cr[0] = s[2 * hw + pix];
      if (ref) { x[0][0] = ref[pix]; x[1][0] = ref[hw + pix]; x[2][0] = ref[2 * hw + pix]; }
    }
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      float r, g, bl;
      ycc2rgb(__fadd_rn(y[k], 0.5f), cb[k], cr[k], r, g, bl);          // add 0.5 back to Y only (:174)
      o[0][k] = fminf(fmaxf(__fsub_rn(r, 0.5f), -0.5f), 0.5f);          // (xhat - 0.5).clamp_(-0.5, 0.5) (:178,181)
      o[1][k] = fminf(fmaxf(__fsub_rn(g, 0.5f), -0.5f), 0.5f);
      o[2][k] = fminf(fmaxf(__fsub_rn(bl, 0.5f), -0.5f), 0.5f);
      if (ref) {
#pragma unroll
        for (int c3 = 0; c3 < 3; ++c3) {
          const float e = __fsub_rn(__fsub_rn(x[c3][k], 0.5f), o[c3][k]);
          local = __fadd_rn(local, __fmul_rn(e, e));
        }
      }
    }
    if (d) {
      if (VEC == 4) {
#pragma unroll
        for (int c3 = 0; c3 < 3; ++c3)
          *reinterpret_cast<float4*>(d + c3 * hw + pix) = make_float4(o[c3][0], o[c3][1 % VEC], o[c3][2 % VEC], o[c3][3 % VEC]);
      } else {
        d[pix] = o[0][0]; d[hw + pix] = o[1][0]; d[2 * hw + pix] = o[2][0];
      }
    }
  }
  if (sse) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) local += __shfl_xor_sync(0xffffffffu, local, off);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) red[w] = local;
    __syncthreads();
    if (w == 0) {
      float t = l < CL_THREADS / 32 ? red[l] : 0.f;
#pragma unroll
      for (int off = 4; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
      if (l == 0) atomicAdd(&sse[b], (double)t);
    }
  }
}

static bool vec4_ok(const void* a, const void* b, const void* c, long long hw) {
  return (hw % 4 == 0) && (((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c)) & 15) == 0);
}

}  // namespace ll

using namespace ll;

extern "C" {

int ll_rgb_to_ycbcr_shift(const float* rgb, float* ycc, int B, int64_t hw, ll_stream_t stream) {
  if (B < 0 || hw < 0) return fail(LL_EINVAL, "ll_rgb_to_ycbcr_shift: bad extents");
  const long long total = (long long)B * hw;
  if (total == 0) return LL_OK;
  if (!rgb || !ycc) return fail(LL_EINVAL, "ll_rgb_to_ycbcr_shift: null pointer");
  const bool v4 = vec4_ok(rgb, ycc, nullptr, hw);
  long long blocks = (total / (v4 ? 4 : 1) + CL_THREADS - 1) / CL_THREADS;
  const long long cap = (long long)sm_count_cached() * 16;
  if (blocks > cap) blocks = cap;
  if (v4) rgb_to_ycbcr_shift_kernel<4><<<(unsigned)blocks, CL_THREADS, 0, as_stream(stream)>>>(rgb, ycc, hw, total);
  else rgb_to_ycbcr_shift_kernel<1><<<(unsigned)blocks, CL_THREADS, 0, as_stream(stream)>>>(rgb, ycc, hw, total);
  LL_LAUNCH_OK("rgb_to_ycbcr_shift_kernel");
  return LL_OK;
}

int ll_ycbcr_to_rgb_sse(const float* ycc_hat, const float* rgb_ref, float* xhat, int B, int64_t hw, double* sse, ll_stream_t stream) {
  if (B < 0 || hw < 0) return fail(LL_EINVAL, "ll_ycbcr_to_rgb_sse: bad extents");
  if ((long long)B * hw == 0) return LL_OK;
  if (!ycc_hat || (!xhat && !sse)) return fail(LL_EINVAL, "ll_ycbcr_to_rgb_sse: null pointer (need ycc_hat and at least one output)");
  if (sse && !rgb_ref) return fail(LL_EINVAL, "ll_ycbcr_to_rgb_sse: the squared error needs the reference image");
  if (B > 65535) return fail(LL_EINVAL, "ll_ycbcr_to_rgb_sse: batch too large");
  const bool v4 = vec4_ok(ycc_hat, rgb_ref, xhat, hw);
  long long bx = (hw / (v4 ? 4 : 1) + CL_THREADS - 1) / CL_THREADS;
  const long long cap = ((long long)sm_count_cached() * 16 + B - 1) / B;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  dim3 grid((unsigned)bx, (unsigned)B);
  if (v4) ycbcr_to_rgb_sse_kernel<4><<<grid, CL_THREADS, 0, as_stream(stream)>>>(ycc_hat, rgb_ref, xhat, hw, sse);
  else ycbcr_to_rgb_sse_kernel<1><<<grid, CL_THREADS, 0, as_stream(stream)>>>(ycc_hat, rgb_ref, xhat, hw, sse);
  LL_LAUNCH_OK("ycbcr_to_rgb_sse_kernel");
  return LL_OK;
}

}  // extern "C"
