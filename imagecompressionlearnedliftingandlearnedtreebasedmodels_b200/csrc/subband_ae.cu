// subband_ae.cu -- pointwise "scaling network" SubbandAutoEncoder (v1) fused with the quantiser.
//
// Reference: SubbandAutoEncoder.encode/.decode (graphs/layers/lifting_dwt_nets.py:99-125): per
// channel, grouped 1x1 convs 1 -> 32 -> 32 -> 32 -> 1 with tanh; the decoder uses
// ConvTranspose2d, i.e. the 32x32 matrices act transposed.  Symbols are torch.round of the
// encoder output (EntropyModel.quantize "dequantize", LiftingBasedDWT_net.py:330,341,352).
// ALU-bound (96 tanh + 2112 FMA per coefficient), so it is a separate kernel from the
// HBM-bound DWT: fusing it would only hide the DWT's bandwidth behind tanh latency.
#include "ll_common.cuh"

namespace ll {

constexpr int AE_H = 32;
constexpr int AE_W0 = 0, AE_B0 = 32, AE_W1 = 64, AE_B1 = AE_W1 + 1024, AE_W2 = AE_B1 + 32, AE_B2 = AE_W2 + 1024,
              AE_W3 = AE_B2 + 32, AE_B3 = AE_W3 + 32, AE_TOTAL = 2212;
static_assert(AE_B3 + 1 <= AE_TOTAL && AE_TOTAL == LL_AE1_BLOB_FLOATS, "ae blob");
constexpr int AE_THREADS = 256, AE_PER_THREAD = 2;

// blob matrices are stored [in i][out j] so that the out index is contiguous (FFMA2 pairs)
__global__ void pack_ae1_kernel(const float* __restrict__ w0, const float* __restrict__ b0,
                                const float* __restrict__ w1, const float* __restrict__ b1,
                                const float* __restrict__ w2, const float* __restrict__ b2,
                                const float* __restrict__ w3, const float* __restrict__ b3, int C, int transposed,
                                float* __restrict__ blob) {
  const int total = C * AE_TOTAL;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
    const int c = e / AE_TOTAL, i = e % AE_TOTAL;
    float v = 0.f;
    if (i < AE_B0) v = w0[c * AE_H + i];
    else if (i < AE_W1) v = b0[c * AE_H + (i - AE_B0)];
    else if (i < AE_B1) {
      const int k = i - AE_W1, in = k / AE_H, out = k % AE_H;
      // Conv2d weight (C*H, H,1,1): [c*H+out][in]; ConvTranspose2d weight (C*H, H,1,1): [c*H+in][out]
      v = transposed ? w1[(c * AE_H + in) * AE_H + out] : w1[(c * AE_H + out) * AE_H + in];
    } else if (i < AE_W2) v = b1[c * AE_H + (i - AE_B1)];
    else if (i < AE_B2) {
      const int k = i - AE_W2, in = k / AE_H, out = k % AE_H;
      v = transposed ? w2[(c * AE_H + in) * AE_H + out] : w2[(c * AE_H + out) * AE_H + in];
    } else if (i < AE_W3) v = b2[c * AE_H + (i - AE_B2)];
    else if (i < AE_B3) v = w3[c * AE_H + (i - AE_W3)];
    else if (i == AE_B3) v = b3[c];
    blob[e] = v;
  }
}

__device__ __forceinline__ void ae_layer(const float* __restrict__ W, const float* __restrict__ Bv,
                                         const float (&hin)[AE_PER_THREAD][AE_H], float (&hout)[AE_PER_THREAD][AE_H]) {
  float2 acc[AE_PER_THREAD][AE_H / 2];
#pragma unroll
  for (int j = 0; j < AE_H; j += 4) {
    const float4 b = *reinterpret_cast<const float4*>(Bv + j);
#pragma unroll
    for (int u = 0; u < AE_PER_THREAD; ++u) {
      acc[u][j / 2] = make_float2(b.x, b.y);
      acc[u][j / 2 + 1] = make_float2(b.z, b.w);
    }
  }
#pragma unroll
  for (int i = 0; i < AE_H; ++i) {
#pragma unroll
    for (int j = 0; j < AE_H; j += 4) {
      const float4 w = *reinterpret_cast<const float4*>(W + i * AE_H + j);
#pragma unroll
      for (int u = 0; u < AE_PER_THREAD; ++u) {
        acc[u][j / 2] = __ffma2_rn(make_float2(hin[u][i], hin[u][i]), make_float2(w.x, w.y), acc[u][j / 2]);
        acc[u][j / 2 + 1] = __ffma2_rn(make_float2(hin[u][i], hin[u][i]), make_float2(w.z, w.w), acc[u][j / 2 + 1]);
      }
    }
  }
#pragma unroll
  for (int u = 0; u < AE_PER_THREAD; ++u)
#pragma unroll
    for (int j = 0; j < AE_H; j += 2) {
      hout[u][j] = tanhf(acc[u][j / 2].x);
      hout[u][j + 1] = tanhf(acc[u][j / 2].y);
    }
}

__global__ void __launch_bounds__(AE_THREADS) ae1_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                         float* __restrict__ q, const float* __restrict__ blob,
                                                         int C, long long n, int chunks_per_plane) {
  __shared__ __align__(16) float sw[AE_TOTAL];
  const int plane = blockIdx.x / chunks_per_plane;  // b * C + c
  const int chunk = blockIdx.x % chunks_per_plane;
  const int c = plane % C;
  for (int i = threadIdx.x; i < AE_TOTAL; i += AE_THREADS) sw[i] = blob[(size_t)c * AE_TOTAL + i];
  __syncthreads();
  const long long base = (long long)plane * n;
  const long long i0 = ((long long)chunk * AE_THREADS + threadIdx.x) * AE_PER_THREAD;
  float xin[AE_PER_THREAD];
#pragma unroll
  for (int u = 0; u < AE_PER_THREAD; ++u) xin[u] = (i0 + u < n) ? x[base + i0 + u] : 0.f;
  float h1[AE_PER_THREAD][AE_H], h2[AE_PER_THREAD][AE_H];
#pragma unroll
  for (int u = 0; u < AE_PER_THREAD; ++u)
#pragma unroll
    for (int j = 0; j < AE_H; ++j) h1[u][j] = tanhf(fmaf(sw[AE_W0 + j], xin[u], sw[AE_B0 + j]));
  ae_layer(sw + AE_W1, sw + AE_B1, h1, h2);
  ae_layer(sw + AE_W2, sw + AE_B2, h2, h1);
#pragma unroll
  for (int u = 0; u < AE_PER_THREAD; ++u) {
    float o = sw[AE_B3];
#pragma unroll
    for (int j = 0; j < AE_H; ++j) o = fmaf(sw[AE_W3 + j], h1[u][j], o);
    if (i0 + u < n) {
      y[base + i0 + u] = o;
      if (q) q[base + i0 + u] = rintf(o);  // torch.round: half to even
    }
  }
}

}  // namespace ll

using namespace ll;

extern "C" {

int ll_pack_ae1(const float* w0, const float* b0, const float* w1, const float* b1, const float* w2, const float* b2,
                const float* w3, const float* b3, int C, int transposed, float* blob, ll_stream_t stream) {
  if (C <= 0) return fail(LL_EINVAL, "ll_pack_ae1: C must be positive");
  if (!w0 || !b0 || !w1 || !b1 || !w2 || !b2 || !w3 || !b3 || !blob) return fail(LL_EINVAL, "ll_pack_ae1: null pointer");
  pack_ae1_kernel<<<(C * AE_TOTAL + 255) / 256, 256, 0, as_stream(stream)>>>(w0, b0, w1, b1, w2, b2, w3, b3, C, transposed, blob);
  LL_LAUNCH_OK("pack_ae1_kernel");
  return LL_OK;
}

int ll_ae1_apply(const float* x, float* y, float* q, const float* blob, int B, int C, int64_t n, ll_stream_t stream) {
  if (B < 0 || C <= 0 || n < 0) return fail(LL_EINVAL, "ll_ae1_apply: bad extents");
  if ((long long)B * n == 0) return LL_OK;
  if (!x || !y || !blob) return fail(LL_EINVAL, "ll_ae1_apply: null pointer");
  const long long per_block = (long long)AE_THREADS * AE_PER_THREAD;
  const long long chunks = (n + per_block - 1) / per_block;
  const long long blocks = chunks * B * C;
  if (blocks > 0x7fffffffLL) return fail(LL_EINVAL, "ll_ae1_apply: too many blocks");
  ae1_kernel<<<(unsigned)blocks, AE_THREADS, 0, as_stream(stream)>>>(x, y, q, blob, C, n, (int)chunks);
  LL_LAUNCH_OK("ae1_kernel");
  return LL_OK;
}

}  // extern "C"
