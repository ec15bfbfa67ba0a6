// rate.cu -- quantisation, likelihood and bit-rate kernels of the entropy models.
//
// ll_quantize      EntropyModel.quantize (compressai 1.2.1; call sites LiftingBasedDWT_net.py:
//                  330,341,352,719): round half-to-even (eval) or x + noise (training).
// ll_gauss_rate    GaussianConditional.forward + -log2 (:334-335,345-346,364-365,752-754,832-833):
//                  y = round(x - mu) + mu | x + noise; sigma = max(sigma, 0.11);
//                  p = max(Phi((.5-|y-mu|)/sigma) - Phi((-.5-|y-mu|)/sigma), 1e-9); bits = -log2 p.
// ll_eb_rate       EntropyBottleneck.forward + -log2 (:204-210,689-690,800-801): per-channel
//                  factorized CDF (1-3-3-3-3-1 softplus/tanh MLP).
// Each also accumulates sum(bits) into a double accumulator (TrainRDLoss.forward3's reductions,
// graphs/losses/rate_dist.py:35-42) so the bpp needs no extra pass over the self-information.
#include "ll_common.cuh"

namespace ll {

constexpr int RT_THREADS = 256;

__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) red[w] = v;
  __syncthreads();
  float t = 0.f;
  if (w == 0) {
    t = l < RT_THREADS / 32 ? red[l] : 0.f;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  }
  return t;  // valid in thread 0
}

__global__ void quantize_kernel(const float* __restrict__ x, const float* __restrict__ noise, float* __restrict__ q,
                                long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    q[i] = noise ? __fadd_rn(x[i], noise[i]) : rintf(x[i]);
}

// x: (B, C, hw) with batch stride x_sb and channel offset folded into the pointer;
// ms: (B, 2C, hw): channel 2c = sigma, 2c+1 = mu.
__global__ void __launch_bounds__(RT_THREADS) gauss_rate_kernel(
    const float* __restrict__ x, long long x_sb, const float* __restrict__ ms, long long ms_sb,
    const float* __restrict__ noise, float* __restrict__ bits, long long bits_sb, float* __restrict__ yout,
    int C, long long hw, long long total, double* __restrict__ sum_out) {
  __shared__ float red[RT_THREADS / 32];
  float local = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i % hw;
    const long long t = i / hw;
    const int c = int(t % C);
    const long long b = t / C;
    const float xv = x[b * x_sb + c * hw + pix];
    const float sg = ms[b * ms_sb + (2 * c) * hw + pix];
    const float mu = ms[b * ms_sb + (2 * c + 1) * hw + pix];
    float y;
    const float bt = gauss_bits(xv, sg, mu, noise ? &noise[(b * C + c) * hw + pix] : nullptr, y);
    bits[b * bits_sb + c * hw + pix] = bt;
    if (yout) yout[(b * C + c) * hw + pix] = y;
    local += bt;
  }
  if (sum_out) {
    const float t = block_sum(local, red);
    if (threadIdx.x == 0) atomicAdd(sum_out, (double)t);
  }
}

// cgp tail: C2 -> C3 (LeakyReLU) -> (sigma, mu) per group and pixel, then the Gaussian rate.
// grid.y = group; one thread = one pixel; weights of the group in shared memory (broadcast reads).
constexpr int TL_MAXC2 = 64, TL_MAXC3 = 32;
template <int C3P>   // C3 rounded up to the accumulator tile the instantiation keeps in registers (20 or 32)
__global__ void __launch_bounds__(RT_THREADS) cgp_tail_rate_kernel(
    const float* __restrict__ h2, long long h2_sb, const float* __restrict__ w3, const float* __restrict__ b3,
    const float* __restrict__ w4, const float* __restrict__ b4, const float* __restrict__ x, long long x_sb,
    const float* __restrict__ noise, float* __restrict__ bits, long long bits_sb, float* __restrict__ yout,
    float* __restrict__ ms_out, int B, int G, int C2, int C3, long long hw, double* __restrict__ sum_out) {
  __shared__ __align__(16) float s_w3[TL_MAXC2 * TL_MAXC3];   // transposed [c][k], zero padded to 32 k
  __shared__ float s_b3[TL_MAXC3], s_w4[2 * TL_MAXC3], s_b4[2];
  __shared__ float red[RT_THREADS / 32];
  const int g = blockIdx.y;
  for (int i = threadIdx.x; i < C2 * TL_MAXC3; i += RT_THREADS) {
    const int c = i / TL_MAXC3, k = i % TL_MAXC3;
    s_w3[i] = k < C3 ? w3[((long long)g * C3 + k) * C2 + c] : 0.f;
  }
  for (int i = threadIdx.x; i < C3; i += RT_THREADS) s_b3[i] = b3[g * C3 + i];
  for (int i = threadIdx.x; i < 2 * C3; i += RT_THREADS) s_w4[i] = w4[(long long)g * 2 * C3 + i];
  if (threadIdx.x < 2) s_b4[threadIdx.x] = b4[g * 2 + threadIdx.x];
  __syncthreads();
  float local = 0.f;
  const long long total = (long long)B * hw;
  // two samples per thread (i and i + half) in the halves of packed FFMA2: one broadcast weight load feeds both, so the
  // shared-memory loads and the FMA instructions per sample halve, and twice as many global loads are in flight per batch
  const long long half = (total + 1) / 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < half; i += (long long)gridDim.x * blockDim.x) {
    const long long i1 = i + half;
    const bool two = i1 < total;
    const long long b0 = i / hw, pix0 = i % hw;
    const long long b1 = two ? i1 / hw : b0, pix1 = two ? i1 % hw : pix0;
    const float* hp0 = h2 + b0 * h2_sb + (long long)g * C2 * hw + pix0;
    const float* hp1 = h2 + b1 * h2_sb + (long long)g * C2 * hw + pix1;
    float2 h[C3P];
#pragma unroll
    for (int k = 0; k < C3P; ++k) h[k] = k < C3 ? make_float2(s_b3[k], s_b3[k]) : make_float2(0.f, 0.f);
    // input channels in batches of 6: the twelve global loads of a batch are issued together (one load per iteration
    // made the loop a chain of C2 dependent DRAM round trips: 0.92 ms for 2.4 M samples, 10x over its instruction count)
    for (int c0 = 0; c0 < C2; c0 += 6) {
      float2 v[6];
#pragma unroll
      for (int j = 0; j < 6; ++j)
        v[j] = (c0 + j < C2) ? make_float2(hp0[(long long)(c0 + j) * hw], hp1[(long long)(c0 + j) * hw]) : make_float2(0.f, 0.f);
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        if (c0 + j < C2) {
#pragma unroll
          for (int k = 0; k < C3P; k += 4)
            if (k < C3) {
              const float4 w = *reinterpret_cast<const float4*>(&s_w3[(c0 + j) * TL_MAXC3 + k]);
              h[k] = __ffma2_rn(make_float2(w.x, w.x), v[j], h[k]);
              h[k + 1] = __ffma2_rn(make_float2(w.y, w.y), v[j], h[k + 1]);
              h[k + 2] = __ffma2_rn(make_float2(w.z, w.z), v[j], h[k + 2]);
              h[k + 3] = __ffma2_rn(make_float2(w.w, w.w), v[j], h[k + 3]);
            }
        }
      }
    }
    float2 sg = make_float2(s_b4[0], s_b4[0]), mu = make_float2(s_b4[1], s_b4[1]);
#pragma unroll
    for (int k = 0; k < C3P; ++k)
      if (k < C3) {
        const float2 a = make_float2(h[k].x < 0.f ? h[k].x * 0.01f : h[k].x, h[k].y < 0.f ? h[k].y * 0.01f : h[k].y);
        sg = __ffma2_rn(make_float2(s_w4[k], s_w4[k]), a, sg);
        mu = __ffma2_rn(make_float2(s_w4[C3 + k], s_w4[C3 + k]), a, mu);
      }
#pragma unroll
    for (int s2 = 0; s2 < 2; ++s2) {
      if (s2 == 1 && !two) break;
      const long long b = s2 ? b1 : b0, pix = s2 ? pix1 : pix0;
      const float sgv = s2 ? sg.y : sg.x, muv = s2 ? mu.y : mu.x;
      const float xv = x[b * x_sb + (long long)g * hw + pix];
      float y;
      const float bt = gauss_bits(xv, sgv, muv, noise ? &noise[(b * G + g) * hw + pix] : nullptr, y);
      bits[b * bits_sb + (long long)g * hw + pix] = bt;
      if (yout) yout[(b * G + g) * hw + pix] = y;
      if (ms_out) {
        ms_out[(b * 2 * G + 2 * g) * hw + pix] = sgv;
        ms_out[(b * 2 * G + 2 * g + 1) * hw + pix] = muv;
      }
      local += bt;
    }
  }
  if (sum_out) {
    const float t = block_sum(local, red);
    if (threadIdx.x == 0) atomicAdd(sum_out, (double)t);
  }
}

// EntropyBottleneck parameter blob per channel (64 floats):
//  [0..2] softplus(M0) [3..5] b0 [6..8] tanh(f0); layers 1..3 at 9+15(l-1): M(9, row-major [out][in]) b(3) tanh f(3);
//  [54..56] softplus(M4) [57] b4 [58] median
constexpr int EB_BLOB = 64;

__device__ __forceinline__ float softplus_t(float x) { return x > 20.f ? x : log1pf(expf(x)); }

__global__ void pack_eb_kernel(const float* __restrict__ m0, const float* __restrict__ b0, const float* __restrict__ f0,
                               const float* __restrict__ m1, const float* __restrict__ b1, const float* __restrict__ f1,
                               const float* __restrict__ m2, const float* __restrict__ b2, const float* __restrict__ f2,
                               const float* __restrict__ m3, const float* __restrict__ b3, const float* __restrict__ f3,
                               const float* __restrict__ m4, const float* __restrict__ b4,
                               const float* __restrict__ quantiles, int C, float* __restrict__ blob) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float* o = blob + (size_t)c * EB_BLOB;
  for (int i = 0; i < 3; ++i) {
    o[i] = softplus_t(m0[c * 3 + i]);
    o[3 + i] = b0[c * 3 + i];
    o[6 + i] = tanhf(f0[c * 3 + i]);
  }
  const float* ms[3] = {m1, m2, m3};
  const float* bs[3] = {b1, b2, b3};
  const float* fs[3] = {f1, f2, f3};
  for (int l = 0; l < 3; ++l) {
    float* q = o + 9 + 15 * l;
    for (int i = 0; i < 9; ++i) q[i] = softplus_t(ms[l][c * 9 + i]);
    for (int i = 0; i < 3; ++i) {
      q[9 + i] = bs[l][c * 3 + i];
      q[12 + i] = tanhf(fs[l][c * 3 + i]);
    }
  }
  for (int i = 0; i < 3; ++i) o[54 + i] = softplus_t(m4[c * 3 + i]);
  o[57] = b4[c];
  o[58] = quantiles[c * 3 + 1];
  for (int i = 59; i < EB_BLOB; ++i) o[i] = 0.f;
}

__device__ __forceinline__ float eb_logits(const float* __restrict__ w, float v) {
  float h[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float t = __fadd_rn(__fmul_rn(w[i], v), w[3 + i]);
    h[i] = __fadd_rn(t, __fmul_rn(w[6 + i], tanhf(t)));
  }
#pragma unroll
  for (int l = 0; l < 3; ++l) {
    const float* q = w + 9 + 15 * l;
    float g[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      float t = __fmul_rn(q[3 * i], h[0]);
      t = fmaf(q[3 * i + 1], h[1], t);
      t = fmaf(q[3 * i + 2], h[2], t);
      t = __fadd_rn(t, q[9 + i]);
      g[i] = __fadd_rn(t, __fmul_rn(q[12 + i], tanhf(t)));
    }
    h[0] = g[0];
    h[1] = g[1];
    h[2] = g[2];
  }
  float t = __fmul_rn(w[54], h[0]);
  t = fmaf(w[55], h[1], t);
  t = fmaf(w[56], h[2], t);
  return __fadd_rn(t, w[57]);
}

__device__ __forceinline__ float sigmoid_t(float x) { return 1.f / (1.f + expf(-x)); }

__global__ void __launch_bounds__(RT_THREADS) eb_rate_kernel(const float* __restrict__ x, const float* __restrict__ noise,
                                                            const float* __restrict__ blob, float* __restrict__ y,
                                                            float* __restrict__ bits, int C, long long hw,
                                                            long long total, double* __restrict__ sum_out) {
  __shared__ float red[RT_THREADS / 32];
  float local = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = int((i / hw) % C);
    const float* w = blob + (size_t)c * EB_BLOB;
    const float med = w[58];
    const float xv = x[i];
    const float yv = noise ? __fadd_rn(xv, noise[i]) : __fadd_rn(rintf(__fsub_rn(xv, med)), med);
    const float lo = eb_logits(w, yv - 0.5f);
    const float up = eb_logits(w, yv + 0.5f);
    const float sum = lo + up;
    const float sg = sum > 0.f ? -1.f : (sum < 0.f ? 1.f : 0.f);
    const float pr = fmaxf(fabsf(sigmoid_t(sg * up) - sigmoid_t(sg * lo)), 1e-9f);
    const float bt = -log2f(pr);
    y[i] = yv;
    bits[i] = bt;
    local += bt;
  }
  if (sum_out) {
    const float t = block_sum(local, red);
    if (threadIdx.x == 0) atomicAdd(sum_out, (double)t);
  }
}

static int grid_for(long long n) {
  long long b = (n + RT_THREADS - 1) / RT_THREADS;
  const long long cap = 148LL * 16;
  return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

}  // namespace ll

using namespace ll;

extern "C" {

int ll_quantize(const float* x, const float* noise, float* q, int64_t n, ll_stream_t stream) {
  if (n < 0) return fail(LL_EINVAL, "ll_quantize: negative size");
  if (n == 0) return LL_OK;
  if (!x || !q) return fail(LL_EINVAL, "ll_quantize: null pointer");
  quantize_kernel<<<grid_for(n), RT_THREADS, 0, as_stream(stream)>>>(x, noise, q, n);
  LL_LAUNCH_OK("quantize_kernel");
  return LL_OK;
}

int ll_gauss_rate(const float* x, int64_t x_sb, const float* ms, int64_t ms_sb, const float* noise, float* bits,
                  int64_t bits_sb, float* y, int B, int C, int64_t hw, double* sum_out, ll_stream_t stream) {
  if (B < 0 || C <= 0 || hw < 0) return fail(LL_EINVAL, "ll_gauss_rate: bad extents");
  const long long total = (long long)B * C * hw;
  if (total == 0) return LL_OK;
  if (!x || !ms || !bits) return fail(LL_EINVAL, "ll_gauss_rate: null pointer");
  gauss_rate_kernel<<<grid_for(total), RT_THREADS, 0, as_stream(stream)>>>(x, x_sb, ms, ms_sb, noise, bits, bits_sb, y, C,
                                                                           hw, total, sum_out);
  LL_LAUNCH_OK("gauss_rate_kernel");
  return LL_OK;
}

int ll_cgp_tail_rate(const float* h2, int64_t h2_sb, const float* w3, const float* b3, const float* w4, const float* b4,
                     const float* x, int64_t x_sb, const float* noise, float* bits, int64_t bits_sb, float* y,
                     float* ms_out, int B, int G, int C2, int C3, int64_t hw, double* sum_out, ll_stream_t stream) {
  if (B < 0 || G < 1 || G > 65535 || hw < 0 || C2 < 1 || C2 > TL_MAXC2 || C3 < 1 || C3 > TL_MAXC3)
    return fail(LL_EINVAL, "ll_cgp_tail_rate: bad extents (C2 <= %d, C3 <= %d)", TL_MAXC2, TL_MAXC3);
  const long long total = (long long)B * hw;
  if (total == 0) return LL_OK;
  if (!h2 || !w3 || !b3 || !w4 || !b4 || !x || !bits) return fail(LL_EINVAL, "ll_cgp_tail_rate: null pointer");
  dim3 grid((unsigned)grid_for(total), (unsigned)G);
  if (C3 <= 20)
    cgp_tail_rate_kernel<20><<<grid, RT_THREADS, 0, as_stream(stream)>>>(h2, h2_sb, w3, b3, w4, b4, x, x_sb, noise, bits, bits_sb,
                                                                         y, ms_out, B, G, C2, C3, hw, sum_out);
  else
    cgp_tail_rate_kernel<32><<<grid, RT_THREADS, 0, as_stream(stream)>>>(h2, h2_sb, w3, b3, w4, b4, x, x_sb, noise, bits, bits_sb,
                                                                         y, ms_out, B, G, C2, C3, hw, sum_out);
  LL_LAUNCH_OK("cgp_tail_rate_kernel");
  return LL_OK;
}

int ll_pack_eb(const float* const* params, int C, float* blob, ll_stream_t stream) {
  // params: _matrix0,_bias0,_factor0, ..., _matrix3,_bias3,_factor3, _matrix4,_bias4, quantiles (15 pointers)
  if (C <= 0 || !params || !blob) return fail(LL_EINVAL, "ll_pack_eb: bad arguments");
  for (int i = 0; i < 15; ++i)
    if (!params[i]) return fail(LL_EINVAL, "ll_pack_eb: null parameter %d", i);
  pack_eb_kernel<<<(C + 63) / 64, 64, 0, as_stream(stream)>>>(params[0], params[1], params[2], params[3], params[4],
                                                             params[5], params[6], params[7], params[8], params[9],
                                                             params[10], params[11], params[12], params[13], params[14],
                                                             C, blob);
  LL_LAUNCH_OK("pack_eb_kernel");
  return LL_OK;
}

int ll_eb_rate(const float* x, const float* noise, const float* blob, float* y, float* bits, int B, int C, int64_t hw,
               double* sum_out, ll_stream_t stream) {
  if (B < 0 || C <= 0 || hw < 0) return fail(LL_EINVAL, "ll_eb_rate: bad extents");
  const long long total = (long long)B * C * hw;
  if (total == 0) return LL_OK;
  if (!x || !blob || !y || !bits) return fail(LL_EINVAL, "ll_eb_rate: null pointer");
  eb_rate_kernel<<<grid_for(total), RT_THREADS, 0, as_stream(stream)>>>(x, noise, blob, y, bits, C, hw, total, sum_out);
  LL_LAUNCH_OK("eb_rate_kernel");
  return LL_OK;
}

}  // extern "C"
