// conv2d.cu -- fp32 SIMT direct convolution for the small context CNNs of the entropy models.
//
// Covers every conv of the tree-based entropy layers that is not a dense contraction large
// enough for the tensor cores: MaskedConv2d chains (csc_xe / csc_list, LiftingBasedDWT_net.py:
// 298-317), the masked 5x5 csc (:274-277), plc's first 3->243 conv on the nearest-2x-upsampled
// parent (:271,355), the grouped 1x1 cgp MLP (:280-290), onlyEZWT's 1x1 head (:792-794) and the
// ZTBlock CNNs (:618-680).  Cross-correlation, zero padding K/2, stride 1, optional LeakyReLU(0.01)
// on the output, optional nearest-2x upsampling of the input fused into the load (replaces
// repeat_interleave(2,2).repeat_interleave(2,3), :348,367), and an output-channel remap so that
// torch.cat((plc0,csc0,plc1,csc1,plc2,csc2)) (:357-359) is a write pattern, not a copy.
//
// Tiling: CTA = 8x32 output pixels x 32 output channels, thread = 4 pixels x 8 channels on
// packed FFMA2 (channel pairs), input channels staged 8 at a time through shared memory.
#include "ll_common.cuh"

namespace ll {

constexpr int CV_THREADS = 256;
constexpr int CV_TH = 8, CV_TW = 32;  // output tile
constexpr int CV_CO = 32;             // output channels per CTA
constexpr int CV_CI = 8;              // input channels per stage

struct ConvParams {
  const float* x;
  const float* w;
  const float* b;
  float* y;
  int B, Cin, H, W, Cout, groups;
  int upsample2, lrelu;
  long long x_sb, y_sb;       // batch strides (elements)
  int co_group, co_stride, co_off;  // output channel remap: (co / co_group) * co_stride + co_off + co % co_group
  int tiles_x, tiles_y, co_tiles_per_group;
};

template <int K>
__global__ void __launch_bounds__(CV_THREADS) conv2d_kernel(const __grid_constant__ ConvParams p) {
  constexpr int PAD = K / 2;
  constexpr int IH = CV_TH + K - 1, IW = CV_TW + K - 1;
  constexpr int IWP = IW + 1;  // pitch
  constexpr int KK = K * K;
  __shared__ __align__(16) float s_in[CV_CI * IH * IWP];
  // weights [ci][tap][co] with a 36-float pitch: the staging loop below walks (co, ci, tap) so that a warp reads runs of
  // CV_CI * KK contiguous floats of the torch layout (walking co fastest gathers one sector per lane), and the pitch
  // spreads its shared-memory stores over 8 banks instead of 1
  constexpr int WP = CV_CO + 4;
  __shared__ __align__(16) float s_w[CV_CI * KK * WP];

  const int tid = threadIdx.x;
  const int tile = blockIdx.x;
  const int tx0 = (tile % p.tiles_x) * CV_TW, ty0 = (tile / p.tiles_x) * CV_TH;
  const int g = blockIdx.y / p.co_tiles_per_group;
  const int cin_g = p.Cin / p.groups, cout_g = p.Cout / p.groups;
  const int co0 = (blockIdx.y % p.co_tiles_per_group) * CV_CO;  // within the group
  const int b = blockIdx.z;

  // thread -> (pixel group, channel group)
  const int cg = tid >> 6;         // 0..3: 8 output channels each
  const int pg = tid & 63;         // 0..63
  const int py = pg >> 3, px = (pg & 7) * 4;

  float2 acc[4][4];
#pragma unroll
  for (int c2 = 0; c2 < 4; ++c2) {
    const int co = co0 + cg * 8 + 2 * c2;
    const float b0 = (p.b && co < cout_g) ? p.b[g * cout_g + co] : 0.f;
    const float b1 = (p.b && co + 1 < cout_g) ? p.b[g * cout_g + co + 1] : 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i][c2] = make_float2(b0, b1);
  }

  const int Hs = p.upsample2 ? p.H / 2 : p.H, Ws = p.upsample2 ? p.W / 2 : p.W;
  const float* xb = p.x + (long long)b * p.x_sb + (long long)g * cin_g * Hs * Ws;

  for (int ci0 = 0; ci0 < cin_g; ci0 += CV_CI) {
    __syncthreads();
    // stage input tile (zero outside the image / beyond the channel count)
    for (int e = tid; e < CV_CI * IH * IW; e += CV_THREADS) {
      const int ci = e / (IH * IW), r = (e / IW) % IH, c = e % IW;
      const int gy = ty0 + r - PAD, gx = tx0 + c - PAD;
      float v = 0.f;
      if (ci0 + ci < cin_g && gy >= 0 && gy < p.H && gx >= 0 && gx < p.W) {
        const int sy = p.upsample2 ? (gy >> 1) : gy, sx = p.upsample2 ? (gx >> 1) : gx;
        v = xb[((long long)(ci0 + ci) * Hs + sy) * Ws + sx];
      }
      s_in[(ci * IH + r) * IWP + c] = v;
    }
    // stage weights [ci][tap][co]
    for (int e = tid; e < CV_CI * KK * CV_CO; e += CV_THREADS) {
      const int co = e / (CV_CI * KK), r = e % (CV_CI * KK), ci = r / KK, tap = r % KK;
      float v = 0.f;
      if (co0 + co < cout_g && ci0 + ci < cin_g)
        v = p.w[((long long)(g * cout_g + co0 + co) * cin_g + (ci0 + ci)) * KK + tap];
      s_w[(ci * KK + tap) * WP + co] = v;
    }
    __syncthreads();
    const int nci = min(CV_CI, cin_g - ci0);
    for (int ci = 0; ci < nci; ++ci) {
#pragma unroll
      for (int dy = 0; dy < K; ++dy) {
        const float* arow = &s_in[(ci * IH + py + dy) * IWP + px];
        float a[4 + K - 1];
#pragma unroll
        for (int k = 0; k < 4 + K - 1; ++k) a[k] = arow[k];
#pragma unroll
        for (int dx = 0; dx < K; ++dx) {
          const float* wp = &s_w[(ci * KK + dy * K + dx) * WP + cg * 8];
          const float4 w0 = *reinterpret_cast<const float4*>(wp);
          const float4 w1 = *reinterpret_cast<const float4*>(wp + 4);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 av = make_float2(a[i + dx], a[i + dx]);
            acc[i][0] = __ffma2_rn(av, make_float2(w0.x, w0.y), acc[i][0]);
            acc[i][1] = __ffma2_rn(av, make_float2(w0.z, w0.w), acc[i][1]);
            acc[i][2] = __ffma2_rn(av, make_float2(w1.x, w1.y), acc[i][2]);
            acc[i][3] = __ffma2_rn(av, make_float2(w1.z, w1.w), acc[i][3]);
          }
        }
      }
    }
  }

  const int oy = ty0 + py, ox = tx0 + px;
  if (oy >= p.H) return;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int co = co0 + cg * 8 + c;
    if (co >= cout_g) break;
    const int cog = g * cout_g + co;
    const int cm = (cog / p.co_group) * p.co_stride + p.co_off + (cog % p.co_group);
    float* yo = p.y + (long long)b * p.y_sb + ((long long)cm * p.H + oy) * p.W + ox;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (ox + i < p.W) {
        float v = (c & 1) ? acc[i][c >> 1].y : acc[i][c >> 1].x;
        if (p.lrelu) v = v > 0.f ? v : v * 0.01f;
        yo[i] = v;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Last 3x3 conv of SubbandAutoEncoderBerk (lifting_dwt_nets.py:133,143: C = 96 -> 3 or 32 -> 1, the output feeds
// the quantiser / the inverse transform) straight from the tensor-core chain's activations: z is channels-last
// fp32 [hi | lo] (B,H,W,2C); the kernel forms hi + lo on load and writes fp32 NCHW.  Replaces the
// NHWC -> NCHW conversion + generic NCHW conv pair (2.45 ms -> one HBM-bound pass over z).
// CTA = 4 warps = 8 rows x 32 columns, thread = 2 vertically adjacent pixels x CO outputs, 32 input channels per stage
// ([py][px][36] floats: the 144-byte pixel pitch makes the 16-byte channel-quad reads of 8 neighbouring pixels hit 8
// distinct bank groups; 128 contiguous bytes per pixel, half and stage).  The one-pixel-per-thread version was bound by
// shared memory (ncu: LSU data pipe 75 %, FMA 25 %: one activation and CO weight LDS.128 per 4 CO FMAs); here the four
// rows a thread's two pixels need are read once per (channel quad, dx) and every broadcast weight quad feeds both pixels:
// 13 LDS.128 per 72 FMAs at CO = 3.  With that fixed the kernel is bound by the memory system: 64-byte pieces per pixel
// and stage streamed at 2.1 TB/s, 128-byte pieces at 2.6 TB/s (an L2 prefetch of the rest of the pixel's row did not help).
constexpr int TL_TH = 8, TL_TW = 32, TL_CC = 32, TL_PITCH = 36, TL_PY = 2, TL_THREADS = 128, TL_MLP = 6;

template <int CO>
constexpr int tl_smem_bytes() { return ((TL_TH + 2) * (TL_TW + 2) * TL_PITCH + 9 * CO * TL_CC) * (int)sizeof(float); }

template <int CO>
__global__ void __launch_bounds__(TL_THREADS) nhwc_split_conv3_kernel(const float* __restrict__ z, const float* __restrict__ w,
                                                                      const float* __restrict__ bias, float* __restrict__ out,
                                                                      int B, int C, int H, int W, int tiles_x) {
  extern __shared__ __align__(16) float tl_smem[];
  float* tile = tl_smem;                                        // [TL_TH + 2][TL_TW + 2][TL_PITCH]
  float* wsm = tl_smem + (TL_TH + 2) * (TL_TW + 2) * TL_PITCH;  // [tap][co][16 ch]
  const int tid = threadIdx.x, tx = tid & 31, wy = tid >> 5;
  const int x0 = (blockIdx.x % tiles_x) * TL_TW, y0 = (blockIdx.x / tiles_x) * TL_TH, b = blockIdx.y;
  const float* zb = z + (long long)b * H * W * 2 * C;
  float acc[TL_PY][CO];
#pragma unroll
  for (int i = 0; i < TL_PY; ++i)
#pragma unroll
    for (int o = 0; o < CO; ++o) acc[i][o] = bias ? bias[o] : 0.f;
  for (int c0 = 0; c0 < C; c0 += TL_CC) {
    __syncthreads();
    // this stage's weights first: their (strided, L2-latency) loads are in flight while the tile is fetched; as a
    // load -> store loop after the tile they cost seven serial L2 round trips per stage (ncu: 7 % of the samples)
    constexpr int NW = 9 * CO * TL_CC, WPT = (NW + TL_THREADS - 1) / TL_THREADS;
    float wreg[WPT];
#pragma unroll
    for (int u = 0; u < WPT; ++u) {
      const int e = tid + u * TL_THREADS;
      const int c = e % TL_CC, o = (e / TL_CC) % CO, t = e / (TL_CC * CO);
      wreg[u] = e < NW ? __ldg(w + ((long long)o * C + c0 + c) * 9 + t) : 0.f;
    }
    // TL_MLP tile elements per thread in flight: all 2 TL_MLP 16-byte loads are issued before the first one is consumed
    // (with one element per iteration the fill ran load -> add -> store serially at 16 warps per SM)
    constexpr int NE = (TL_TH + 2) * (TL_TW + 2) * (TL_CC / 4);
    for (int e0 = tid; e0 < NE; e0 += TL_MLP * TL_THREADS) {
      float4 hi[TL_MLP], lo[TL_MLP];
#pragma unroll
      for (int u = 0; u < TL_MLP; ++u) {
        const int e = e0 + u * TL_THREADS;
        const int q = e % (TL_CC / 4), px = (e / (TL_CC / 4)) % (TL_TW + 2), py = (e / (TL_CC / 4)) / (TL_TW + 2);
        const int gy = y0 + py - 1, gx = x0 + px - 1;
        hi[u] = lo[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (e < NE && gy >= 0 && gy < H && gx >= 0 && gx < W) {
          const float* s = zb + ((long long)gy * W + gx) * (2 * C) + c0 + 4 * q;
          hi[u] = __ldg(reinterpret_cast<const float4*>(s));
          lo[u] = __ldg(reinterpret_cast<const float4*>(s + C));
        }
      }
#pragma unroll
      for (int u = 0; u < TL_MLP; ++u) {
        const int e = e0 + u * TL_THREADS;
        if (e < NE) {
          const int q = e % (TL_CC / 4), pxy = e / (TL_CC / 4);
          *reinterpret_cast<float4*>(&tile[pxy * TL_PITCH + 4 * q]) =
              make_float4(hi[u].x + lo[u].x, hi[u].y + lo[u].y, hi[u].z + lo[u].z, hi[u].w + lo[u].w);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < WPT; ++u) {
      const int e = tid + u * TL_THREADS;
      if (e < NW) wsm[e] = wreg[u];
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < TL_CC / 4; ++q) {
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        float4 a[TL_PY + 2];
#pragma unroll
        for (int r = 0; r < TL_PY + 2; ++r)
          a[r] = *reinterpret_cast<const float4*>(&tile[((wy * TL_PY + r) * (TL_TW + 2) + tx + dx) * TL_PITCH + 4 * q]);
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
          for (int o = 0; o < CO; ++o) {
            const float4 k = *reinterpret_cast<const float4*>(&wsm[((dy * 3 + dx) * CO + o) * TL_CC + 4 * q]);
#pragma unroll
            for (int i = 0; i < TL_PY; ++i) {
              acc[i][o] = fmaf(a[i + dy].x, k.x, acc[i][o]);
              acc[i][o] = fmaf(a[i + dy].y, k.y, acc[i][o]);
              acc[i][o] = fmaf(a[i + dy].z, k.z, acc[i][o]);
              acc[i][o] = fmaf(a[i + dy].w, k.w, acc[i][o]);
            }
          }
        }
      }
    }
  }
  const int gx = x0 + tx;
  if (gx < W) {
#pragma unroll
    for (int i = 0; i < TL_PY; ++i) {
      const int gy = y0 + wy * TL_PY + i;
      if (gy < H) {
#pragma unroll
        for (int o = 0; o < CO; ++o) out[(((long long)b * CO + o) * H + gy) * W + gx] = acc[i][o];
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// 1x1 head of onlyEZWT's parent context net (nn.Conv2d(243, 6, 1) behind a LeakyReLU, LiftingBasedDWT_net.py:
// 792-794) on the raw channels-last output of the 3xTF32 tensor-core conv: out[b][o][pix] = bias[o] +
// sum_c w[o][c] * lrelu(y[pix][c]).  y (B*hw, Cpad) fp32 with Cpad % 4 == 0 (channels >= C are ignored), out fp32
// NCHW (B, Cout, hw), Cout <= 8.  One thread per pixel, weights broadcast from shared memory, exact fp32 FMA.
constexpr int PW_MAXCO = 8, PW_MAXC = 256;

__global__ void __launch_bounds__(256) nhwc_lrelu_conv1_kernel(const float* __restrict__ y, const float* __restrict__ w,
                                                               const float* __restrict__ bias, float* __restrict__ out,
                                                               long long npix, long long hw, int C, int Cpad, int Cout, int lrelu) {
  __shared__ __align__(16) float wsm[PW_MAXCO * PW_MAXC];   // [c / 4][o][4]
  const int c4n = (C + 3) / 4;
  for (int e = threadIdx.x; e < c4n * PW_MAXCO * 4; e += 256) {
    const int j = e & 3, o = (e >> 2) % PW_MAXCO, c = (e / (4 * PW_MAXCO)) * 4 + j;
    wsm[e] = (o < Cout && c < C) ? w[(long long)o * C + c] : 0.f;
  }
  __syncthreads();
  for (long long pix = blockIdx.x * 256LL + threadIdx.x; pix < npix; pix += (long long)gridDim.x * 256) {
    const float* yp = y + pix * Cpad;
    float acc[PW_MAXCO];
#pragma unroll
    for (int o = 0; o < PW_MAXCO; ++o) acc[o] = (bias && o < Cout) ? bias[o] : 0.f;
    for (int c4 = 0; c4 < c4n; ++c4) {
      float4 v = *reinterpret_cast<const float4*>(yp + 4 * c4);
      if (lrelu) {
        v.x = v.x < 0.f ? v.x * 0.01f : v.x;
        v.y = v.y < 0.f ? v.y * 0.01f : v.y;
        v.z = v.z < 0.f ? v.z * 0.01f : v.z;
        v.w = v.w < 0.f ? v.w * 0.01f : v.w;
      }
#pragma unroll
      for (int o = 0; o < PW_MAXCO; ++o) {
        const float4 k = *reinterpret_cast<const float4*>(&wsm[(c4 * PW_MAXCO + o) * 4]);
        acc[o] = fmaf(v.x, k.x, acc[o]);
        acc[o] = fmaf(v.y, k.y, acc[o]);
        acc[o] = fmaf(v.z, k.z, acc[o]);
        acc[o] = fmaf(v.w, k.w, acc[o]);
      }
    }
    const long long b = pix / hw, p = pix - b * hw;
#pragma unroll
    for (int o = 0; o < PW_MAXCO; ++o)
      if (o < Cout) out[(b * Cout + o) * hw + p] = acc[o];
  }
}

}  // namespace ll

using namespace ll;

extern "C" {

int ll_conv2d(const float* x, int64_t x_sb, const float* w, const float* b, float* y, int64_t y_sb, int B, int Cin,
              int H, int W, int Cout, int K, int groups, int upsample2, int lrelu, int co_group, int co_stride,
              int co_off, ll_stream_t stream) {
  if (B < 0 || Cin <= 0 || Cout <= 0 || H < 0 || W < 0 || groups <= 0) return fail(LL_EINVAL, "ll_conv2d: bad extents");
  if (Cin % groups || Cout % groups) return fail(LL_EINVAL, "ll_conv2d: channels not divisible by groups");
  if (K != 1 && K != 3 && K != 5) return fail(LL_EINVAL, "ll_conv2d: kernel size %d not supported (1, 3, 5)", K);
  if (upsample2 && ((H & 1) || (W & 1))) return fail(LL_EINVAL, "ll_conv2d: upsample2 needs even output size");
  if ((long long)B * H * W == 0) return LL_OK;
  if (!x || !w || !y) return fail(LL_EINVAL, "ll_conv2d: null pointer");
  if (co_group <= 0) {
    co_group = Cout;
    co_stride = 0;
    co_off = 0;
  }
  ConvParams p = {};
  p.x = x; p.w = w; p.b = b; p.y = y;
  p.B = B; p.Cin = Cin; p.H = H; p.W = W; p.Cout = Cout; p.groups = groups;
  p.upsample2 = upsample2; p.lrelu = lrelu;
  p.x_sb = x_sb; p.y_sb = y_sb;
  p.co_group = co_group; p.co_stride = co_stride; p.co_off = co_off;
  p.tiles_x = (W + CV_TW - 1) / CV_TW;
  p.tiles_y = (H + CV_TH - 1) / CV_TH;
  p.co_tiles_per_group = (Cout / groups + CV_CO - 1) / CV_CO;
  if (B > 65535 || p.co_tiles_per_group * groups > 65535) return fail(LL_EINVAL, "ll_conv2d: grid too large");
  dim3 grid((unsigned)(p.tiles_x * p.tiles_y), (unsigned)(p.co_tiles_per_group * groups), (unsigned)B);
  cudaStream_t st = as_stream(stream);
  if (K == 1) conv2d_kernel<1><<<grid, CV_THREADS, 0, st>>>(p);
  else if (K == 3) conv2d_kernel<3><<<grid, CV_THREADS, 0, st>>>(p);
  else conv2d_kernel<5><<<grid, CV_THREADS, 0, st>>>(p);
  LL_LAUNCH_OK("conv2d_kernel");
  return LL_OK;
}

int ll_nhwc_split_conv3(const float* z, const float* w, const float* bias, float* out, int B, int C, int Cout, int H, int W,
                        ll_stream_t stream) {
  if (B < 0 || H < 0 || W < 0 || C <= 0 || (C % TL_CC)) return fail(LL_EINVAL, "ll_nhwc_split_conv3: C must be a positive multiple of %d (got %d)", TL_CC, C);
  if (Cout != 1 && Cout != 3) return fail(LL_EINVAL, "ll_nhwc_split_conv3: Cout must be 1 or 3 (got %d)", Cout);
  if ((long long)B * H * W == 0) return LL_OK;
  if (!z || !w || !out) return fail(LL_EINVAL, "ll_nhwc_split_conv3: null pointer");
  if (reinterpret_cast<uintptr_t>(z) & 15) return fail(LL_EINVAL, "ll_nhwc_split_conv3: z must be 16-byte aligned");
  const int tiles_x = (W + TL_TW - 1) / TL_TW, tiles_y = (H + TL_TH - 1) / TL_TH;
  if (B > 65535) return fail(LL_EINVAL, "ll_nhwc_split_conv3: batch too large");
  dim3 grid((unsigned)(tiles_x * tiles_y), (unsigned)B);
  static bool attr_set = false;      // > 48 KB of dynamic shared memory needs the opt-in, once per process (idempotent)
  if (!attr_set) {
    LL_CUDA_OK(cudaFuncSetAttribute(nhwc_split_conv3_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, tl_smem_bytes<1>()));
    LL_CUDA_OK(cudaFuncSetAttribute(nhwc_split_conv3_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, tl_smem_bytes<3>()));
    attr_set = true;
  }
  if (Cout == 1) nhwc_split_conv3_kernel<1><<<grid, TL_THREADS, tl_smem_bytes<1>(), as_stream(stream)>>>(z, w, bias, out, B, C, H, W, tiles_x);
  else nhwc_split_conv3_kernel<3><<<grid, TL_THREADS, tl_smem_bytes<3>(), as_stream(stream)>>>(z, w, bias, out, B, C, H, W, tiles_x);
  LL_LAUNCH_OK("nhwc_split_conv3_kernel");
  return LL_OK;
}

int ll_nhwc_lrelu_conv1(const float* y, const float* w, const float* bias, float* out, int B, int64_t hw, int C, int Cpad,
                        int Cout, int lrelu, ll_stream_t stream) {
  if (B < 0 || hw < 0 || C <= 0 || C > PW_MAXC || Cpad < C || (Cpad % 4) || Cout <= 0 || Cout > PW_MAXCO)
    return fail(LL_EINVAL, "ll_nhwc_lrelu_conv1: bad extents (C <= %d, Cpad %% 4 == 0 and >= C, Cout <= %d)", PW_MAXC, PW_MAXCO);
  const long long npix = (long long)B * hw;
  if (npix == 0) return LL_OK;
  if (!y || !w || !out) return fail(LL_EINVAL, "ll_nhwc_lrelu_conv1: null pointer");
  if (reinterpret_cast<uintptr_t>(y) & 15) return fail(LL_EINVAL, "ll_nhwc_lrelu_conv1: y must be 16-byte aligned");
  long long blocks = (npix + 255) / 256;
  const long long cap = (long long)sm_count_cached() * 8;
  if (blocks > cap) blocks = cap;
  nhwc_lrelu_conv1_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(y, w, bias, out, npix, hw, C, Cpad, Cout, lrelu);
  LL_LAUNCH_OK("nhwc_lrelu_conv1_kernel");
  return LL_OK;
}

}  // extern "C"
