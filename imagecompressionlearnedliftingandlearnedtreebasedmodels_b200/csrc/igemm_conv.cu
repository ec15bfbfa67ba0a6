// igemm_conv.cu -- tcgen05 / TMEM implicit-GEMM convolution for the dense context CNNs (K3).
//
// Replaces the 243->243 3x3 conv of plc_list[i] (graphs/models/LiftingBasedDWT_net.py:271-272,355;
// onlyEZWT :789-794) and the dense grouped 1x1 layers of the cgp MLP (:280-290): the only convs of
// the tree-based entropy models whose channel counts make them dense contractions (62 % of the
// entropy-model FLOPs sit in plc alone).  These nets only produce (sigma, mu)
// of the rate model -- they move bpp (tolerance 0.1 %), never a quantised symbol -- so the
// operands are BF16 with FP32 accumulation in tensor memory.
//
//   GEMM view   D[m, n] = sum_{tap, ci} A_tap[m, ci] * W[tap][n][ci]
//               m = pixel of a 8x16 tile (M = 128), n = output channel (N = Npad <= 256),
//               k = (tap, ci): 9 (or 1) taps x Kpad input channels, 64 channels (128 B) per k-block.
//   A operand   activations, NHWC bf16 (B, H, W, Kpad): one TMA box {64 ch, 16 x, 8 y, 1 b} per
//               (tap, k-block) at (x0 + dx, y0 + dy); out-of-image coordinates are zero-filled by
//               TMA = the conv's zero padding.  Lands as 128 rows x 128 B, SWIZZLE_128B, K-major.
//   B operand   weights packed [tap][Npad][Kpad] bf16 (ll_pack_igemm_weight): TMA box {64, Npad, 1}.
//   pipeline    warp 0 = TMA producer, warp 1 = MMA issuer (one thread, tcgen05.mma
//               cta_group::1 kind::f16, M128 x Npad x K16), warps 2-5 = epilogue
//               (tcgen05.ld 32x32b -> bias, LeakyReLU -> global).  4 smem stages (full/empty
//               mbarriers), 2 accumulator stages of 256 TMEM columns (tmem_full/tmem_empty), so the
//               epilogue of tile t overlaps the MMAs of tile t+1.  Persistent: grid = #SMs, static
//               round-robin over (b, tile_y, tile_x); neighbouring CTAs work on neighbouring tiles
//               at the same time so halos and weight tiles are L2 hits.
//   outputs     fp32 NCHW through the channel remap of ll_conv2d (torch.cat as a write pattern)
//               and/or bf16 NHWC at a channel offset (feeds the next igemm layer).
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdint.h>

#include <type_traits>
#include "ll_common.cuh"
#include "tc_ptx.cuh"

namespace ll {

constexpr int IG_TW = 16, IG_TH = 8;
constexpr int IG_BM = IG_TW * IG_TH;   // 128
constexpr int IG_BK = 64;              // bf16 per k-block (one 128-byte swizzle row)
constexpr int IG_STAGES = 4;
constexpr int IG_MAXN = 256;
constexpr int IG_A_BYTES = IG_BM * IG_BK * 2;     // 16 KB
constexpr int IG_B_BYTES = IG_MAXN * IG_BK * 2;   // 32 KB
constexpr int IG_STAGE_BYTES = IG_A_BYTES + IG_B_BYTES;
constexpr int IG_THREADS = 320;   // TMA warp, MMA warp, 8 epilogue warps (two per TMEM lane quarter: even / odd 32-column chunks)
constexpr int IG_TMEM_COLS = 512;
constexpr int IG_MAXG = 3;       // channel groups per launch (the cgp MLP has groups = 3)
constexpr int IG_MAXSLOTS = 24;  // k-blocks per (group, tap); the 3xTF32 chain of a 192-channel layer needs 18
constexpr int IG_SMEM_BYTES = 1024 /*align slack*/ + IG_STAGES * IG_STAGE_BYTES + 1024 /*barriers*/ + 2 * IG_MAXG * IG_MAXN * 4 +
                              IG_MAXG * (64 * 20 + 20 + 40 + 4) * 4 /*cgp tail weights (epi 4)*/;

// GDN epilogues: a = beta + gamma . y^2 is a normal positive float (beta >= the reparametrisation's pedestal), so the one-MUFU
// forms apply: rsqrt.approx.ftz returns what rsqrtf() returns for normal inputs (rsqrtf only adds the denormal rescaling:
// FSETP + two predicated FMULs per value), sqrt.approx.ftz is within 1 ulp of sqrtf() (inverse GDN = decoder side only:
// it moves the reconstruction by <= 1 ulp per layer, not the symbols) and has no slow-path call.
__device__ __forceinline__ float rsqrt_approx(float a) {
  float r;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
  return r;
}
__device__ __forceinline__ float sqrt_approx(float a) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
  return r;
}

struct IgemmParams {
  const float* bias;
  float* out_f32;          // NCHW fp32 (may be null)
  long long out_sb;        // batch stride (elements)
  int co_group, co_stride, co_off;
  __nv_bfloat16* out_bf16; // NHWC bf16 (may be null)
  int out_cstride, out_coff;
  int B, H, W, Cout, Npad, kblocks, taps, lrelu;
  int tiles_x, tiles_y;
  long long ntiles;                       // tiles x groups
  int groups, out_gstride;                // NHWC channel stride between groups
  int a_koff[IG_MAXG][IG_MAXSLOTS];       // input-channel coordinate of each k-block, per group
  int b_koff[IG_MAXSLOTS];                // K coordinate of each k-block in the packed weights
  // TF32 chain (SubbandAutoEncoderBerk): epi 1 = conv -> Y (raw) + S (hi|lo of y^2); 2 = GDN -> Z (hi|lo of
  // y * rsqrt(acc + beta), or * sqrt for the inverse GDN), reading Y; 3 = conv -> Y only
  int epi, inverse;
  // accumulator plan: nstages accumulator stages of stage_cols TMEM columns; a k-block goes to accumulator
  // acc_sel[k] (0 main, 1 small terms) at column offset acc_sel * acc_cols inside the stage
  int nstages, stage_cols, nacc, acc_cols;
  int acc_sel[IG_MAXSLOTS];
  float* y;                               // NHWC fp32 (B,H,W,Cout)
  float* sz;                              // NHWC fp32 (B,H,W,2*Cout)
  // epi 4 (BF16 kernel): cgp layer 2 + layers 3-4 + Gaussian rate in the epilogue (ll_igemm_cgp_tail)
  const float *t_w3, *t_b3, *t_w4, *t_b4, *t_x, *t_noise;
  float* t_bits;
  double* t_sum;
  long long t_xsb, t_bsb;
  int t_c3;
};
constexpr int TL_K = 20;                                   // layer-3 outputs per group, padded (float4 rows)
constexpr int TL_GROUP = IG_BK * TL_K + TL_K + 2 * TL_K + 4;   // per group: w3t[c 64][20] | b3[20] | w4 sigma[20] | w4 mu[20] | b4[2]

// ---------------------------------------------------------------------------------------- kernel
template <bool TF32>
__global__ void __launch_bounds__(IG_THREADS, 1)
igemm_conv_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const __grid_constant__ IgemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;          // SWIZZLE_128B atoms are 1024-byte aligned
  uint8_t* gen = smem_raw + (base - raw);
  const uint32_t bars = base + IG_STAGES * IG_STAGE_BYTES;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (IG_STAGES + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * IG_STAGES + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * IG_STAGES + 2 + a); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + IG_STAGES * IG_STAGE_BYTES + 8 * (2 * IG_STAGES + 4));
  float* s_bias = reinterpret_cast<float*>(gen + IG_STAGES * IG_STAGE_BYTES + 1024);
  int* s_cmap = reinterpret_cast<int*>(s_bias + IG_MAXG * IG_MAXN);
  float* s_tail = reinterpret_cast<float*>(s_cmap + IG_MAXG * IG_MAXN);
  const bool tail = !TF32 && p.epi == 4;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (tail) {
    for (int i = threadIdx.x; i < p.groups * TL_GROUP; i += IG_THREADS) {
      const int g = i / TL_GROUP, r = i % TL_GROUP;
      float v = 0.f;
      if (r < IG_BK * TL_K) {
        const int c = r / TL_K, k = r % TL_K;
        if (c < p.Cout && k < p.t_c3) v = p.t_w3[((long long)g * p.t_c3 + k) * p.Cout + c];
      } else if (r < IG_BK * TL_K + TL_K) {
        const int k = r - IG_BK * TL_K;
        if (k < p.t_c3) v = p.t_b3[g * p.t_c3 + k];
      } else if (r < IG_BK * TL_K + 3 * TL_K) {
        const int k = (r - IG_BK * TL_K - TL_K) % TL_K, which = (r - IG_BK * TL_K - TL_K) / TL_K;   // 0 sigma, 1 mu
        if (k < p.t_c3) v = p.t_w4[((long long)g * 2 + which) * p.t_c3 + k];
      } else if (r < IG_BK * TL_K + 3 * TL_K + 2) {
        v = p.t_b4[g * 2 + (r - IG_BK * TL_K - 3 * TL_K)];
      }
      s_tail[i] = v;
    }
  }

  for (int i = threadIdx.x; i < p.groups * IG_MAXN; i += IG_THREADS) {
    const int g = i / IG_MAXN, c = i % IG_MAXN, co = g * p.Cout + c;   // channel index across groups
    s_bias[i] = (p.bias && c < p.Cout) ? p.bias[co] : 0.f;
    s_cmap[i] = (co / p.co_group) * p.co_stride + p.co_off + (co % p.co_group);
  }
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    for (int s = 0; s < IG_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), tail ? 4 : 8);   // one arrival per epilogue warp (epi 4: a stage belongs to one warp parity)
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(IG_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int iters = p.taps * p.kblocks;
  const uint32_t stage_tx = IG_A_BYTES + (uint32_t)p.Npad * IG_BK * 2;
  const long long per_img = (long long)p.tiles_x * p.tiles_y;
  const int G = p.groups;

  if (warp == 0) {
    {
      int stage = 0;
      uint32_t phase = 0;
      for (long long t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
        const int g = (int)(t % G);
        const long long tt = t / G;
        const int b = (int)(tt / per_img);
        const int r = (int)(tt % per_img);
        const int y0 = (r / p.tiles_x) * IG_TH, x0 = (r % p.tiles_x) * IG_TW;
        for (int tap = 0; tap < p.taps; ++tap) {
          const int dy = p.taps == 9 ? tap / 3 - 1 : 0, dx = p.taps == 9 ? tap % 3 - 1 : 0;
          for (int kb = 0; kb < p.kblocks; ++kb) {
            mbar_wait(empty_bar(stage), phase ^ 1);
            if (elect_one()) {
              mbar_expect_tx(full_bar(stage), stage_tx);
              const uint32_t sa = base + stage * IG_STAGE_BYTES;
              tma_load_4d(sa, &tmA, full_bar(stage), p.a_koff[g][kb], x0 + dx, y0 + dy, b);
              tma_load_3d(sa + IG_A_BYTES, &tmB, full_bar(stage), p.b_koff[kb], 0, g * p.taps + tap);
            }
            __syncwarp();
            if (++stage == IG_STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // warp-uniform loops, one elected lane issues (see igemm_tf32_gdn_pair_kernel: `if (lane == 0)` wraps every MMA in an
    // ELECT / R2UR loop of ~95 cycles)
    {
      // instruction descriptor: D fp32, A/B bf16, both K-major, N = Npad, M = 128
      const uint32_t fmt = TF32 ? 2u : 1u;   // operand format: TF32 (fp32 containers) or BF16
      const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(p.Npad >> 3) << 17) | ((uint32_t)(IG_BM >> 4) << 24);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (long long t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);
        tc_fence_after();
        uint32_t started = 0;                    // bit s: accumulator s of this tile already holds a partial sum
        for (int it = 0; it < iters; ++it) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const int sel = p.acc_sel[it % p.kblocks];
          if (elect_one()) {
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.stage_cols + sel * p.acc_cols);
            const uint32_t first = ((started >> sel) & 1u) ^ 1u;
            const uint32_t sa = base + stage * IG_STAGE_BYTES;
            const uint64_t adesc = umma_desc_sw128(sa), bdesc = umma_desc_sw128(sa + IG_A_BYTES);
#pragma unroll
            for (int k = 0; k < IG_BK / 16; ++k) {  // 32 bytes (16 bf16 / 8 tf32) per MMA along K: +2 in 16-byte units
              const uint32_t accum = (uint32_t)(k != 0) | (first ^ 1u);
              if (TF32) tc_mma_tf32_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, accum);
              else tc_mma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, accum);
            }
            tc_commit(empty_bar(stage));           // smem stage reusable once these MMAs have read it
          }
          __syncwarp();
          started |= 1u << sel;
          if (++stage == IG_STAGES) { stage = 0; phase ^= 1; }
        }
        if (elect_one()) tc_commit(tfull_bar(acc));   // accumulator complete
        __syncwarp();
        if (++acc == p.nstages) { acc = 0; acc_phase ^= 1; }
      }
    }
    __syncwarp();
  } else {
    // Eight epilogue warps: the epilogue is a chain of global loads / stores per 32-column chunk (ncu: 4 warps left the
    // GDN GEMMs latency bound at 9 % warp occupancy, tensor pipe 7-12 % active), so every TMEM lane quarter gets two
    // warps that take the even and the odd chunks.
    const int q = warp & 3;                      // TMEM lane quarter this warp may read (warp id mod 4)
    const int chunk0 = (warp - 2) >> 2;          // 0: even chunks, 1: odd chunks
    const int row = q * 32 + lane;
    const int ty = row / IG_TW, tx = row % IG_TW;
    int acc = 0;
    uint32_t acc_phase = 0;
    const int nchunks = p.Npad / 32 + ((p.Npad % 32) ? 1 : 0);
    float tail_sum = 0.f;
    unsigned unit = 0;
    for (long long t = blockIdx.x; t < p.ntiles; t += gridDim.x, ++unit) {
      if (tail && (int)(unit & 1u) != chunk0) {       // epi 4: the even units (accumulator stage 0) belong to the warps of
        if (++acc == p.nstages) { acc = 0; acc_phase ^= 1; }   // parity 0, the odd ones to parity 1 -- no hand-off inside a unit
        continue;
      }
      const int g = (int)(t % G);
      const long long tt = t / G;
      const int b = (int)(tt / per_img);
      const int r = (int)(tt % per_img);
      const int y = (r / p.tiles_x) * IG_TH + ty, x = (r % p.tiles_x) * IG_TW + tx;
      const float* gb = s_bias + g * IG_MAXN;
      const int* gm = s_cmap + g * IG_MAXN;
      const bool valid = y < p.H && x < p.W;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.stage_cols);
      float* of = p.out_f32 ? p.out_f32 + (long long)b * p.out_sb + (long long)y * p.W + x : nullptr;
      __nv_bfloat16* ob = p.out_bf16 ? p.out_bf16 + (((long long)b * p.H + y) * p.W + x) * p.out_cstride + p.out_coff + g * p.out_gstride : nullptr;
      const long long plane = (long long)p.H * p.W;
      if (tail) {
        // cgp layer 2 -> (LeakyReLU) -> layer 3 (C2 -> C3, LeakyReLU) -> layer 4 (sigma, mu) -> Gaussian rate, one pixel per
        // thread, in the operation order of cgp_tail_rate_kernel (rate.cu): the 54-channel map never leaves the SM.
        uint32_t v0[32], v1[32];
        tc_ld32(taddr, v0);
        tc_ld32(taddr + 32, v1);
        tc_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(acc));        // accumulator stage free: the arithmetic below overlaps the next MMAs
        const float* tw = s_tail + g * TL_GROUP;
        float h[TL_K];
#pragma unroll
        for (int k = 0; k < TL_K; ++k) h[k] = tw[IG_BK * TL_K + k];
#pragma unroll
        for (int c = 0; c < IG_BK; ++c) {
          if (c < p.Cout) {
            float a = __uint_as_float(c < 32 ? v0[c & 31] : v1[c & 31]) + gb[c];
            a = a < 0.f ? a * 0.01f : a;
#pragma unroll
            for (int k4 = 0; k4 < TL_K / 4; ++k4) {
              const float4 w = *reinterpret_cast<const float4*>(tw + c * TL_K + 4 * k4);
              h[4 * k4 + 0] = fmaf(w.x, a, h[4 * k4 + 0]);
              h[4 * k4 + 1] = fmaf(w.y, a, h[4 * k4 + 1]);
              h[4 * k4 + 2] = fmaf(w.z, a, h[4 * k4 + 2]);
              h[4 * k4 + 3] = fmaf(w.w, a, h[4 * k4 + 3]);
            }
          }
        }
        float sg = tw[IG_BK * TL_K + 3 * TL_K], mu = tw[IG_BK * TL_K + 3 * TL_K + 1];
#pragma unroll
        for (int k = 0; k < TL_K; ++k) {
          if (k < p.t_c3) {
            const float a = h[k] < 0.f ? h[k] * 0.01f : h[k];
            sg = fmaf(tw[IG_BK * TL_K + TL_K + k], a, sg);
            mu = fmaf(tw[IG_BK * TL_K + 2 * TL_K + k], a, mu);
          }
        }
        if (valid) {
          const long long pix = (long long)y * p.W + x;
          const float xv = p.t_x[(long long)b * p.t_xsb + (long long)g * plane + pix];
          float yq;
          const float bt = gauss_bits(xv, sg, mu, p.t_noise ? &p.t_noise[((long long)b * G + g) * plane + pix] : nullptr, yq);
          p.t_bits[(long long)b * p.t_bsb + (long long)g * plane + pix] = bt;
          tail_sum += bt;
        }
        if (++acc == p.nstages) { acc = 0; acc_phase ^= 1; }
        continue;
      }
      for (int c = chunk0; c < nchunks; c += 2) {
        // GDN epilogue: the raw conv output y of this chunk is fetched first, so its global latency overlaps the
        // tensor-memory loads below (it used to be issued behind them, one dependent round trip per chunk)
        float o[32];
        const bool chunk_on = TF32 && valid && c * 32 < p.Cout;
        if (TF32 && p.epi == 2 && chunk_on) {
          const float* yq = p.y + (((long long)b * p.H + y) * p.W + x) * p.Cout + c * 32;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 t4 = *reinterpret_cast<const float4*>(yq + j);
            o[j] = t4.x; o[j + 1] = t4.y; o[j + 2] = t4.z; o[j + 3] = t4.w;
          }
        }
        uint32_t v[32];
        tc_ld32(taddr + c * 32, v);
        tc_wait_ld();
        if (TF32 && p.nacc == 2) {                 // main + small-term accumulators, summed here in round-to-nearest fp32
          uint32_t w[32];
          tc_ld32(taddr + p.acc_cols + c * 32, w);
          tc_wait_ld();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(w[j]));
        }
        if (TF32) {
          if (valid && c * 32 < p.Cout) {
            const long long px = ((long long)b * p.H + y) * p.W + x;
            float* yp = p.y + px * p.Cout + c * 32;
            float* zp = p.sz + px * (2 * p.Cout) + c * 32;
            float hi[32], lo[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const float a = __uint_as_float(v[j]) + gb[c * 32 + j];
              float r;
              if (p.epi == 2) r = o[j] * (p.inverse ? sqrt_approx(a) : rsqrt_approx(a));   // GDN / inverse GDN
              else { o[j] = a; r = a * a; }                                       // conv: raw output, square for the GDN norm
              hi[j] = tf32_rna(r);
              lo[j] = tf32_rna(r - hi[j]);
            }
            if (p.epi != 2) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(yp + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
            }
            if (p.epi != 3) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                *reinterpret_cast<float4*>(zp + j) = make_float4(hi[j], hi[j + 1], hi[j + 2], hi[j + 3]);
                *reinterpret_cast<float4*>(zp + p.Cout + j) = make_float4(lo[j], lo[j + 1], lo[j + 2], lo[j + 3]);
              }
            }
          }
        } else if (valid) {
          float f[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float a = __uint_as_float(v[j]) + gb[c * 32 + j];
            f[j] = (p.lrelu && a < 0.f) ? a * 0.01f : a;
          }
          if (of) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (c * 32 + j < p.Cout) of[(long long)gm[c * 32 + j] * plane] = f[j];
          }
          if (ob) {
            // every Npad channel is written (the padded ones are exact zeros: zero weights, zero bias), so a
            // consumer igemm layer never multiplies uninitialised memory
            if (c * 32 + 32 <= p.Npad && ((p.out_coff + g * p.out_gstride) % 8 == 0) && (p.out_cstride % 8 == 0)) {
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                uint4 pk;
                __nv_bfloat162 h0 = __floats2bfloat162_rn(f[j], f[j + 1]), h1 = __floats2bfloat162_rn(f[j + 2], f[j + 3]);
                __nv_bfloat162 h2 = __floats2bfloat162_rn(f[j + 4], f[j + 5]), h3 = __floats2bfloat162_rn(f[j + 6], f[j + 7]);
                pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
                pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
                *reinterpret_cast<uint4*>(ob + c * 32 + j) = pk;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (c * 32 + j < p.Npad) ob[c * 32 + j] = __float2bfloat16_rn(f[j]);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      if (++acc == p.nstages) { acc = 0; acc_phase ^= 1; }
    }
    if (tail && p.t_sum) {       // sum of the self-information (TrainRDLoss.forward3's reduction), one atomic per warp
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) tail_sum += __shfl_xor_sync(0xffffffffu, tail_sum, o);
      if (lane == 0) atomicAdd(p.t_sum, (double)tail_sum);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(IG_TMEM_COLS) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// 3xTF32 implicit GEMM on CTA PAIRS (cta_group::2) -- the conv / GDN kernels of SubbandAutoEncoderBerk.
//
// ncu on the single-CTA instance (profiles/r02_ncu_berk_igemm_first.csv): tensor pipe 28-34 % active with L2 and DRAM
// far from their limits -- the MMA waits for operand bytes: every 32-channel k-block was fetched three times
// (A_lo|B_hi, A_hi|B_lo, A_hi|B_hi = 3 x 40 KB per 12 MMAs at N = 192).  Here
//   * a k-block is fetched ONCE: stage = {A_hi, A_lo, B_hi, B_lo}, 12 MMAs per stage;
//   * two CTAs of a cluster run one M = 256 MMA: each holds its own 128 pixels of A and HALF of the weight rows, the
//     tensor cores of the pair read both halves -- the weight tile crosses L2 -> SM once per pair.
// 56 KB per CTA per 12 MMAs instead of 120 KB (N = 192).  Accumulators (main + small-term, see ll_igemm_tf32) and the
// epilogues are those of igemm_conv_kernel<true>; every CTA drains its own 128 TMEM lanes.
// Protocol: full[s] lives in the leader (even) CTA and counts the bytes of both CTAs' TMA loads; empty[s] / tfull[a]
// are arrived in both CTAs by multicast tcgen05.commit; tempty[a] lives in the leader, 16 arrivals (8 epilogue warps x 2).
// ------------------------------------------------------------------------------------------------
constexpr int PR_MAXST = 6;
constexpr int PR_A_BYTES = IG_BM * 128;              // 128 pixels x 32 tf32
constexpr int PR_SMEM_LIMIT = 232448;                // 227 KB
constexpr int PR_TAIL_BYTES = 256 + IG_MAXN * 4;     // barriers + tmem slot, bias

struct PairParams {
  const float* bias;
  float* y;
  float* sz;
  int B, H, W, C, Cout, Npad, taps, epi, inverse;
  int tiles_x, tiles_y;
  long long ntiles, npairs;
  int kb;                                  // 32-channel k-blocks
  int stages, stage_bytes, bhalf_bytes;    // per CTA
  int nacc_stages, stage_cols, acc_cols;   // TMEM plan (as IgemmParams)
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(IG_THREADS, 1)
igemm_tf32_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                       const __grid_constant__ PairParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - raw);
  const uint32_t ring_bytes = (uint32_t)p.stages * p.stage_bytes;
  const uint32_t bars = base + ring_bytes;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (PR_MAXST + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * PR_MAXST + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * PR_MAXST + 2 + a); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + ring_bytes + 8 * (2 * PR_MAXST + 4));
  float* s_bias = reinterpret_cast<float*>(gen + ring_bytes + 256);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  for (int i = threadIdx.x; i < IG_MAXN; i += IG_THREADS) s_bias[i] = (p.bias && i < p.Cout) ? p.bias[i] : 0.f;
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 16);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    tmem_alloc_2sm(smem_u32(tmem_slot), IG_TMEM_COLS);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                       // the peer's barriers exist before anything is signalled across the pair
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int iters = p.taps * p.kb;
  const long long per_img = (long long)p.tiles_x * p.tiles_y;
  const long long pair0 = blockIdx.x >> 1, pair_step = gridDim.x >> 1;

  if (warp == 0) {
    {
      int stage = 0;
      uint32_t phase = 0;
      const int nhalf = p.Npad / 2;
      for (long long pt = pair0; pt < p.npairs; pt += pair_step) {
        const long long t = 2 * pt + rank;             // an odd tile count leaves the last pair's second CTA on an
        const int b = (int)(t / per_img);              // out-of-range image index: TMA zero-fills, nothing is stored
        const int r = (int)(t % per_img);
        const int y0 = (r / p.tiles_x) * IG_TH, x0 = (r % p.tiles_x) * IG_TW;
        for (int tap = 0; tap < p.taps; ++tap) {
          const int dy = p.taps == 9 ? tap / 3 - 1 : 0, dx = p.taps == 9 ? tap % 3 - 1 : 0;
          for (int kb = 0; kb < p.kb; ++kb) {
            mbar_wait(empty_bar(stage), phase ^ 1);
            if (elect_one()) {
              if (leader) mbar_expect_tx(full_bar(stage), 2u * (uint32_t)p.stage_bytes);
              const uint32_t sa = base + stage * p.stage_bytes;
              tma_load_4d_2sm(sa, &tmA, full_bar(stage), 32 * kb, x0 + dx, y0 + dy, b);
              tma_load_4d_2sm(sa + PR_A_BYTES, &tmA, full_bar(stage), p.C + 32 * kb, x0 + dx, y0 + dy, b);
              tma_load_3d_2sm(sa + 2 * PR_A_BYTES, &tmB, full_bar(stage), 32 * kb, (int)rank * nhalf, tap);
              tma_load_3d_2sm(sa + 2 * PR_A_BYTES + p.bhalf_bytes, &tmB, full_bar(stage), p.C + 32 * kb, (int)rank * nhalf, tap);
            }
            __syncwarp();
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (leader) {                                  // warp-uniform loops, one elected lane issues
      // D fp32, A/B tf32 K-major, N = Npad, M = 256 across the pair
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.Npad >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (long long pt = pair0; pt < p.npairs; pt += pair_step) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_main = tmem_base + (uint32_t)(acc * p.stage_cols);
        const uint32_t d_small = d_main + (uint32_t)p.acc_cols;
        for (int it = 0; it < iters; ++it) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t sa = base + stage * p.stage_bytes;
            const uint64_t a_hi = umma_desc_sw128(sa), a_lo = umma_desc_sw128(sa + PR_A_BYTES);
            const uint64_t b_hi = umma_desc_sw128(sa + 2 * PR_A_BYTES), b_lo = umma_desc_sw128(sa + 2 * PR_A_BYTES + p.bhalf_bytes);
            const uint32_t cont = (uint32_t)(it != 0);
            // the two small terms first (own accumulator: its ulp is 2^-11 of the main one's), then A_hi * B_hi
#pragma unroll
            for (int k = 0; k < 4; ++k) tc_mma_tf32_ss_2sm(d_small, a_lo + 2 * k, b_hi + 2 * k, idesc, cont | (uint32_t)(k != 0));
#pragma unroll
            for (int k = 0; k < 4; ++k) tc_mma_tf32_ss_2sm(d_small, a_hi + 2 * k, b_lo + 2 * k, idesc, 1u);
#pragma unroll
            for (int k = 0; k < 4; ++k) tc_mma_tf32_ss_2sm(d_main, a_hi + 2 * k, b_hi + 2 * k, idesc, cont | (uint32_t)(k != 0));
            tc_commit_2sm(empty_bar(stage), 3);          // both CTAs' copies of the stage are reusable
          }
          __syncwarp();
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        if (elect_one()) tc_commit_2sm(tfull_bar(acc), 3);   // accumulators complete, in both CTAs
        __syncwarp();
        if (++acc == p.nacc_stages) { acc = 0; acc_phase ^= 1; }
      }
    }
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int chunk0 = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const int ty = row / IG_TW, tx = row % IG_TW;
    int acc = 0;
    uint32_t acc_phase = 0;
    const int nchunks = (p.Npad + 31) / 32;
    for (long long pt = pair0; pt < p.npairs; pt += pair_step) {
      const long long t = 2 * pt + rank;
      const int b = (int)(t / per_img);
      const int r = (int)(t % per_img);
      const int y = (r / p.tiles_x) * IG_TH + ty, x = (r % p.tiles_x) * IG_TW + tx;
      const bool valid = t < p.ntiles && y < p.H && x < p.W;
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.stage_cols);
      const long long px = ((long long)b * p.H + y) * p.W + x;
      for (int c = chunk0; c < nchunks; c += 2) {
        float o[32];
        const bool chunk_on = valid && c * 32 < p.Cout;
        if (p.epi == 2 && chunk_on) {
          const float* yq = p.y + px * p.Cout + c * 32;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 t4 = *reinterpret_cast<const float4*>(yq + j);
            o[j] = t4.x; o[j + 1] = t4.y; o[j + 2] = t4.z; o[j + 3] = t4.w;
          }
        }
        uint32_t v[32], w[32];
        tc_ld32(taddr + c * 32, v);
        tc_ld32(taddr + p.acc_cols + c * 32, w);
        tc_wait_ld();
        if (chunk_on) {
          float* yp = p.y + px * p.Cout + c * 32;
          float* zp = p.sz + px * (2 * p.Cout) + c * 32;
          float hi[32], lo[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float a = (__uint_as_float(v[j]) + __uint_as_float(w[j])) + s_bias[c * 32 + j];
            float rr;
            if (p.epi == 2) rr = o[j] * (p.inverse ? sqrt_approx(a) : rsqrt_approx(a));
            else { o[j] = a; rr = a * a; }
            hi[j] = tf32_rna(rr);
            lo[j] = tf32_rna(rr - hi[j]);
          }
          if (p.epi != 2) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(yp + j) = make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]);
          }
          if (p.epi != 3) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              *reinterpret_cast<float4*>(zp + j) = make_float4(hi[j], hi[j + 1], hi[j + 2], hi[j + 3]);
              *reinterpret_cast<float4*>(zp + p.Cout + j) = make_float4(lo[j], lo[j + 1], lo[j + 2], lo[j + 3]);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(tempty_bar(acc));
      if (++acc == p.nacc_stages) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                       // neither CTA frees tensor memory / exits while the pair is still working
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, IG_TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// Fused 3x3 conv + GDN / inverse GDN on CTA pairs (SubbandAutoEncoderBerk, lifting_dwt_nets.py:140-148 and
// graphs/layers/gdn.py:54-92):   z = y * rsqrt(beta + gamma . y^2)   (inverse: * sqrt),   y = conv(a) + bias.
// The un-fused chain wrote y (4 B) and the [hi|lo] split of y^2 (8 B) per value and ran the norm as a second GEMM that
// read them back (12 B) and wrote the split of z (8 B): 32 B of HBM traffic per value and two kernels that ncu showed
// bound by their epilogues' global loads / stores.  Here the conv's accumulator tile never leaves the SM:
//   E1   epilogue warps: y = main + small + bias, written back IN PLACE into tensor memory;
//   S    staging rounds: [hi | lo] split of y^2 for KS channels -> tensor memory as the A operand (tcgen05.st);
//   G    tcgen05.mma (A from tensor memory, B = gamma rows streamed through the TMA ring) accumulates the norm of NP output
//        channels, small terms and main term in separate accumulators as in the conv;
//   E2   z = y * rsqrt(norm + beta) -> [hi | lo] split -> the only global stores of the kernel (8 B per value).
// Tensor-memory plan per CTA (columns): y [0, N) | norm main [N, N + NP) | norm small [N + NP, N + 2 NP) | staging
// hi [N + 2 NP, +KS) lo [.., +KS).  N = 192: NP = 96 (two passes), KS = 64 -> 512 columns; N <= 96: NP = KS = N.
// 16 epilogue warps = 4 groups x 4 TMEM lane quarters.  Group g only touches the 16-column chunks of y whose global index
// is g (mod 4), so no y column is read by a warp other than the one that wrote it; a staging round's slots (16 channels
// each) are staged by different groups at the same time and their MMAs start slot by slot.  In E2 a warp first pulls y and
// the norm of its chunks into registers and finishes the arithmetic, THEN releases tensor memory (the next pass's norm
// MMAs / the next tile's mainloop start) and only then splits and stores: the global stores of a tile overlap the tensor
// work that follows.  (Timeline of the 8-warp version, conv 96->192: mainloop 30 k cycles, then 30 k cycles of E1 /
// staging / E2 with the tensor pipe idle.)
// ------------------------------------------------------------------------------------------------
constexpr int GD_THREADS = 576;   // TMA warp, MMA warp, 16 epilogue warps
constexpr int GD_EPI_WARPS = 16;
struct GdnPairParams {
  const float* bias;
  const float* beta;
  float* sz;
  int B, H, W, C, N, taps, inverse;
  int tiles_x, tiles_y;
  long long ntiles, npairs;
  int kb;
  int stages, stage_bytes, bhalf_bytes;
  int NP, KS, passes, rpp;                 // outputs per pass, channels per staging round, passes, rounds per pass
  int ghalf_bytes;                         // (NP / 2) rows x 128 B
  int col_nm, col_ns, col_st;
  // head mode (first layer of the scaling network: Conv2d(iC, N, 3, padding=1), iC <= 3, K = 9 iC <= 27 padded to 32): the
  // epilogue warps build the A operand -- the [hi | lo] split of every pixel's 3x3 x iC window of the fp32 NCHW input --
  // straight in tensor memory, the weights sit in shared memory for the whole launch, and the conv is twelve 3xTF32 MMAs
  // per tile; everything from E1 on is shared.  (The first version computed y on the FP32 pipe, one pixel per thread: 864
  // broadcast LDS.128 + 1 728 FFMA2 per warp and tile, 10 k of the tile's 28 k cycles.)
  int head, iC;
  const float* x;                          // (B, iC, H, W)
  const float* w0;                         // (N, iC, 3, 3) torch layout
  int wstage;                              // head mode: one extra ring-sized slot behind the ring holds this CTA's half of w0
#ifdef LL_TIMELINE
  long long* tl;                           // probe build only (csrc/probe/igemm_timeline.cu): clock64 stamps of CTA 0
#endif
};
#ifdef LL_TIMELINE
__device__ int g_nostore_dev = 0;
static long long* g_timeline = nullptr;
#define LL_TL(tile, slot) do { if (p.tl && blockIdx.x == 0 && (threadIdx.x & 31) == 0 && (tile) < 16) p.tl[(tile) * 64 + (slot)] = clock64(); } while (0)
#else
#define LL_TL(tile, slot) do { } while (0)
#endif

template <bool HEAD>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GD_THREADS, 1)
igemm_tf32_gdn_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                           const __grid_constant__ CUtensorMap tmG, const __grid_constant__ GdnPairParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - raw);
  const uint32_t ring_bytes = (uint32_t)(p.stages + p.wstage) * p.stage_bytes;
  const uint32_t wbuf = base + (uint32_t)p.stages * p.stage_bytes;     // head mode: [hi | lo] of w0, (N / 2) x 128 B each
  const uint32_t bars = base + ring_bytes;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (PR_MAXST + s); };
  const uint32_t tfull_bar = bars + 8u * (2 * PR_MAXST), tempty_bar = bars + 8u * (2 * PR_MAXST + 1);
  // GDN hand-offs, one pair of barriers per 16-channel staging slot (at most 6 slots): aready[c] = the split of y^2 of slot
  // c is in tensor memory (8 arrivals: the 4 warps that own the slot x 2 CTAs), gdone[c] = the MMAs that read it are
  // complete; e2done = every warp is done with E1 / has read the previous pass's norm (32 arrivals, once per pass)
  auto aready_bar = [&](int c) { return bars + 8u * (2 * PR_MAXST + 2 + c); };
  auto gdone_bar = [&](int c) { return bars + 8u * (2 * PR_MAXST + 8 + c); };
  const uint32_t e2done_bar = bars + 8u * (2 * PR_MAXST + 14);
  // pdone = all norm MMAs of a pass are complete.  Its own barrier, waited once per pass by every warp: a warp may only wait
  // on a barrier whose every phase it observes (a parity wait cannot tell phase k from phase k - 2), and a slot's gdone is
  // observed round by round by its owners only.
  const uint32_t pdone_bar = bars + 8u * (2 * PR_MAXST + 15);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + ring_bytes + 8 * (2 * PR_MAXST + 16));
  float* s_bias = reinterpret_cast<float*>(gen + ring_bytes + 256);
  float* s_beta = s_bias + IG_MAXN;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  for (int i = threadIdx.x; i < IG_MAXN; i += GD_THREADS) {
    s_bias[i] = (p.bias && i < p.N) ? p.bias[i] : 0.f;
    s_beta[i] = i < p.N ? p.beta[i] : 1.f;
  }
  if (HEAD) {
    // B operand of the head conv: this CTA's N / 2 output channels x K = 32 (k = ci * 9 + tap, zero beyond 9 iC), split
    // [hi | lo], in the K-major SWIZZLE_128B layout a TMA box {32, N / 2} would produce (row = 128 B, 16-byte chunk index
    // XOR row mod 8) -- written once per CTA with generic stores, made visible to the tensor core by the proxy fence
    const int K0 = 9 * p.iC, nhalf = p.N / 2;
    uint8_t* wb = gen + (size_t)p.stages * p.stage_bytes;
    for (int i = threadIdx.x; i < nhalf * 32; i += GD_THREADS) {
      const int r = i >> 5, k = i & 31;
      const float v = k < K0 ? p.w0[((int)rank * nhalf + r) * K0 + k] : 0.f;
      const float h = tf32_rna(v);
      const uint32_t off = (uint32_t)r * 128u + ((((uint32_t)k >> 2) ^ ((uint32_t)r & 7u)) << 4) + ((uint32_t)k & 3u) * 4u;
      *reinterpret_cast<float*>(wb + off) = h;
      *reinterpret_cast<float*>(wb + p.bhalf_bytes + off) = tf32_rna(v - h);
    }
    fence_proxy_async();
  }
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmG) : "memory");
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tfull_bar, 1);
    mbar_init(tempty_bar, 2 * GD_EPI_WARPS);
    for (int c = 0; c < 6; ++c) {
      mbar_init(aready_bar(c), 8);
      mbar_init(gdone_bar(c), 1);
    }
    mbar_init(e2done_bar, 2 * GD_EPI_WARPS);
    mbar_init(pdone_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    tmem_alloc_2sm(smem_u32(tmem_slot), IG_TMEM_COLS);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // (All pairs of a launch run their phases in lockstep -- every tile costs the same.  Spreading them with a one-off start
  // offset per pair was measured and is SLOWER: conv 96->192 + GDN 1.44 ms in lockstep, 1.53 / 1.65 ms at 4096 / 8192 cycles
  // x (pair mod 8): neighbouring tiles that run together share their halo rows and the weight tiles in L2.)
  const int iters = p.taps * p.kb;
  const int gkb = p.N / 32;                              // k-blocks of one GDN pass
  const long long per_img = (long long)p.tiles_x * p.tiles_y;
  const long long pair0 = blockIdx.x >> 1, pair_step = gridDim.x >> 1;

  if (warp == 0) {
    {
      int stage = 0;
      uint32_t phase = 0;
      const int nhalf = p.N / 2, ghalf = p.NP / 2;
      for (long long pt = pair0; pt < p.npairs; pt += pair_step) {
        const long long t = 2 * pt + rank;
        const int b = (int)(t / per_img);
        const int r = (int)(t % per_img);
        const int y0 = (r / p.tiles_x) * IG_TH, x0 = (r % p.tiles_x) * IG_TW;
        for (int tap = 0; tap < p.taps; ++tap) {
          const int dy = p.taps == 9 ? tap / 3 - 1 : 0, dx = p.taps == 9 ? tap % 3 - 1 : 0;
          for (int kb = 0; kb < p.kb; ++kb) {
            mbar_wait(empty_bar(stage), phase ^ 1);
            if (elect_one()) {
              if (leader) mbar_expect_tx(full_bar(stage), 2u * (uint32_t)p.stage_bytes);
              const uint32_t sa = base + stage * p.stage_bytes;
              tma_load_4d_2sm(sa, &tmA, full_bar(stage), 32 * kb, x0 + dx, y0 + dy, b);
              tma_load_4d_2sm(sa + PR_A_BYTES, &tmA, full_bar(stage), p.C + 32 * kb, x0 + dx, y0 + dy, b);
              tma_load_3d_2sm(sa + 2 * PR_A_BYTES, &tmB, full_bar(stage), 32 * kb, (int)rank * nhalf, tap);
              tma_load_3d_2sm(sa + 2 * PR_A_BYTES + p.bhalf_bytes, &tmB, full_bar(stage), p.C + 32 * kb, (int)rank * nhalf, tap);
            }
            __syncwarp();
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
        }
        // gamma rows of every pass, k-block by k-block, in the order the norm MMAs consume them
        for (int ps = 0; ps < p.passes; ++ps) {
          for (int kb = 0; kb < gkb; ++kb) {
            mbar_wait(empty_bar(stage), phase ^ 1);
            if (elect_one()) {
              if (leader) mbar_expect_tx(full_bar(stage), 4u * (uint32_t)p.ghalf_bytes);
              const uint32_t sa = base + stage * p.stage_bytes;
              tma_load_3d_2sm(sa, &tmG, full_bar(stage), 32 * kb, ps * p.NP + (int)rank * ghalf, 0);
              tma_load_3d_2sm(sa + p.ghalf_bytes, &tmG, full_bar(stage), p.N + 32 * kb, ps * p.NP + (int)rank * ghalf, 0);
            }
            __syncwarp();
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // The whole warp walks the loops (warp-uniform control flow) and one elected lane issues: descriptors and tensor-memory
    // addresses then live in uniform registers.  Issued from `if (lane == 0)` every tcgen05.mma was wrapped in an
    // ELECT / R2UR.BROADCAST / BRA.U.ANY loop (~95 cycles per MMA: the timeline probe showed the N = 96 mainloop and the
    // norm GEMMs running at half the tensor rate, issue-bound).
    if (leader) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.N >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      const uint32_t idesc_g = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.NP >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      int stage = 0;
      uint32_t phase = 0, te_phase = 0, ar_phase = 0, e2_phase = 0;
      const uint32_t d_main = tmem_base, d_small = tmem_base + (uint32_t)p.N;
      const uint32_t d_nm = tmem_base + (uint32_t)p.col_nm, d_ns = tmem_base + (uint32_t)p.col_ns;
      const uint32_t a_st = tmem_base + (uint32_t)p.col_st;
      int tli = 0;
      for (long long pt = pair0; pt < p.npairs; pt += pair_step, ++tli) {
        if (HEAD) {
          // head conv: A = the [hi | lo] split of the 3x3 x iC input windows, staged in tensor memory by the epilogue warps
          // (K = 32), B = w0 from shared memory; same three-term split and accumulator plan as the mainloop below
          LL_TL(tli, 0);
          mbar_wait_spin(tempty_bar, te_phase ^ 1);      // the previous tile's E2 has read y and the norm
          te_phase ^= 1;
          LL_TL(tli, 3);
          mbar_wait_spin(e2done_bar, e2_phase);          // every warp of the pair has staged its windows
          e2_phase ^= 1;
          tc_fence_after();
          LL_TL(tli, 1);
          if (elect_one()) {
            const uint64_t b_hi = umma_desc_sw128(wbuf), b_lo = umma_desc_sw128(wbuf + p.bhalf_bytes);
            const uint32_t ah = a_st, al = a_st + (uint32_t)p.KS;
#pragma unroll
            for (int k = 0; k < 4; ++k) tc_mma_tf32_ts_2sm(d_small, al + 8 * k, b_hi + 2 * k, idesc, (uint32_t)(k != 0));
#pragma unroll
            for (int k = 0; k < 4; ++k) tc_mma_tf32_ts_2sm(d_small, ah + 8 * k, b_lo + 2 * k, idesc, 1u);
#pragma unroll
            for (int k = 0; k < 4; ++k) tc_mma_tf32_ts_2sm(d_main, ah + 8 * k, b_hi + 2 * k, idesc, (uint32_t)(k != 0));
            tc_commit_2sm(tfull_bar, 3);
          }
          __syncwarp();
          LL_TL(tli, 2);
        } else {
          LL_TL(tli, 0);
          mbar_wait_spin(tempty_bar, te_phase ^ 1);
          te_phase ^= 1;
          tc_fence_after();
          LL_TL(tli, 1);
#ifdef LL_TIMELINE
          if (p.tl && blockIdx.x == 0 && lane == 0 && tli < 16) {   // wall clock beside the cycle stamp: the SM clock under load
            long long gt;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
            p.tl[tli * 64 + 60] = gt;
          }
#endif
          for (int it = 0; it < iters; ++it) {
            mbar_wait(full_bar(stage), phase);
            tc_fence_after();
            if (elect_one()) {
              const uint32_t sa = base + stage * p.stage_bytes;
              const uint64_t a_hi = umma_desc_sw128(sa), a_lo = umma_desc_sw128(sa + PR_A_BYTES);
              const uint64_t b_hi = umma_desc_sw128(sa + 2 * PR_A_BYTES), b_lo = umma_desc_sw128(sa + 2 * PR_A_BYTES + p.bhalf_bytes);
              const uint32_t cont = (uint32_t)(it != 0);
#pragma unroll
              for (int k = 0; k < 4; ++k) tc_mma_tf32_ss_2sm(d_small, a_lo + 2 * k, b_hi + 2 * k, idesc, cont | (uint32_t)(k != 0));
#pragma unroll
              for (int k = 0; k < 4; ++k) tc_mma_tf32_ss_2sm(d_small, a_hi + 2 * k, b_lo + 2 * k, idesc, 1u);
#pragma unroll
              for (int k = 0; k < 4; ++k) tc_mma_tf32_ss_2sm(d_main, a_hi + 2 * k, b_hi + 2 * k, idesc, cont | (uint32_t)(k != 0));
              tc_commit_2sm(empty_bar(stage), 3);
            }
            __syncwarp();
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
          }
          if (elect_one()) tc_commit_2sm(tfull_bar, 3);
          __syncwarp();
          LL_TL(tli, 2);
        }
        // norm GEMMs: A = staged split of y^2 in tensor memory, B = gamma rows from the ring.  Slot by slot: the MMAs of
        // slot c start as soon as its owners have staged it, while the other parity's warps stage the next slot.
        for (int ps = 0; ps < p.passes; ++ps) {
          // every warp of the pair has finished E1 (pass 0: the small-term accumulator shares its columns with the norm) /
          // has read the previous pass's norm: the MMAs may overwrite the norm accumulators
          mbar_wait_spin(e2done_bar, e2_phase);
          e2_phase ^= 1;
          for (int rd = 0; rd < p.rpp; ++rd) {
            for (int kbr = 0; kbr < p.KS / 32; ++kbr) {
              mbar_wait(full_bar(stage), phase);
              tc_fence_after();
              const uint32_t sa = base + stage * p.stage_bytes;
              const uint64_t g_hi = umma_desc_sw128(sa), g_lo = umma_desc_sw128(sa + p.ghalf_bytes);
#pragma unroll
              for (int h = 0; h < 2; ++h) {             // the two 16-channel slots of this 32-channel gamma k-block
                const int slot = 2 * kbr + h;
                mbar_wait_spin(aready_bar(slot), ar_phase);
                tc_fence_after();
                if (slot == 0) LL_TL(tli, 4 + 2 * (ps * p.rpp + rd));
                if (elect_one()) {
                  const uint32_t cont = (uint32_t)((rd | slot) != 0);
#pragma unroll
                  for (int kk = 0; kk < 2; ++kk) {
                    const int k = 2 * h + kk;
                    const uint32_t ah = a_st + (uint32_t)(kbr * 32 + 8 * k), al = ah + (uint32_t)p.KS;
                    tc_mma_tf32_ts_2sm(d_ns, al, g_hi + 2 * k, idesc_g, cont | (uint32_t)(kk != 0));
                    tc_mma_tf32_ts_2sm(d_ns, ah, g_lo + 2 * k, idesc_g, 1u);
                    tc_mma_tf32_ts_2sm(d_nm, ah, g_hi + 2 * k, idesc_g, cont | (uint32_t)(kk != 0));
                  }
                  if (h == 1) tc_commit_2sm(empty_bar(stage), 3);
                  tc_commit_2sm(gdone_bar(slot), 3);    // this slot's MMAs done: staging reusable / (last slot) norm complete
                }
                __syncwarp();
              }
              if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
            ar_phase ^= 1;
            LL_TL(tli, 5 + 2 * (ps * p.rpp + rd));
          }
          if (elect_one()) tc_commit_2sm(pdone_bar, 3);   // the pass's norm is complete
          __syncwarp();
        }
      }
    }
    __syncwarp();
  } else {
    const int q = warp & 3;
    const int grp = (warp - 2) >> 2;                    // this warp owns the 16-column chunks of y with index = grp (mod 4)
    const int row = q * 32 + lane;
    const int ty = row / IG_TW, tx = row % IG_TW;
    const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
    uint32_t tf_phase = 0, pd_phase = 0;
    unsigned ground = 0;                                // staging rounds so far, over all tiles: every slot's gdone completes once per round
    const int nch = p.N / 16;                           // 16-column chunks of y
    const int slots = p.KS / 16, npc = p.NP / 16;
    // head mode: the 3x3 x iC input window of this thread's pixel; the NEXT tile's window is requested as soon as the
    // current one has been consumed, so its global-load latency hides behind the GDN phases (ncu: 22 % of the kernel's
    // stall samples sat on the first use of these loads when they were issued at the top of the tile)
    // Only the 16 window samples this warp stages (k = 16 (grp & 1) + j, k = ci * 9 + tap) are loaded, with compile-time
    // (ci, dy, dx) per branch.  (With all 27 samples per thread the kernel spilled them under its 96-register cap and the
    // loads ran one after the other: the timeline probe showed 8 k of a tile's 26 k cycles between the conv's issue and
    // the epilogue seeing its result, spent inside this prefetch.)
    float win[16];
    auto load_window = [&](long long pt_) {
      const long long t_ = 2 * pt_ + rank;
      const int b_ = (int)(t_ / per_img);
      const int r_ = (int)(t_ % per_img);
      const int y_ = (r_ / p.tiles_x) * IG_TH + ty, x_ = (r_ % p.tiles_x) * IG_TW + tx;
      const bool ok_ = pt_ < p.npairs && t_ < p.ntiles && y_ < p.H && x_ < p.W;
      const float* xb_ = p.x + (long long)b_ * p.iC * p.H * p.W + (long long)y_ * p.W + x_;
      const long long hw_ = (long long)p.H * p.W;
      auto fetch = [&](auto kb0) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          constexpr int K0 = decltype(kb0)::value;
          const int k = K0 + j;
          const int ci = k / 9, dy = (k % 9) / 3 - 1, dx = k % 3 - 1;
          const int gy = y_ + dy, gx = x_ + dx;
          float v = 0.f;
          if (k < 27 && ok_ && ci < p.iC && gy >= 0 && gy < p.H && gx >= 0 && gx < p.W)
            v = __ldg(xb_ + ci * hw_ + dy * p.W + dx);
          win[j] = v;
        }
      };
      if ((grp & 1) == 0) fetch(std::integral_constant<int, 0>{});
      else fetch(std::integral_constant<int, 16>{});
    };
    int tli = 0;
#ifdef LL_TIMELINE
    const bool tlw = warp == 2 && lane == 0;
#define LL_TLE(slot) do { if (tlw) LL_TL(tli, slot); } while (0)
#else
#define LL_TLE(slot) do { } while (0)
#endif
    // head: this thread's 3x3 x iC window -> [hi | lo] split -> the A operand of the head conv in tensor memory (the
    // staging columns are free: the previous tile's norm MMAs completed before its E2).  The four groups of a lane
    // quarter share the work: groups 0 / 1 write k = 0..15 / 16..31 of the hi half, groups 2 / 3 of the lo half.
    // Called for tile t + 1 between the register half of tile t's E2 and its global stores: the stores keep the LSU busy
    // for ~6 k cycles per tile (32 scattered sectors per instruction), and the hand-off -> conv MMAs -> accumulator
    // round trip of the next tile (~3.5 k cycles) now runs underneath them instead of after them.
    auto stage_window = [&]() {
      LL_TLE(30);
      uint32_t v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float h = tf32_rna(win[j]);
        v[j] = __float_as_uint(grp < 2 ? h : tf32_rna(win[j] - h));
      }
      tmem_st16(tlane + p.col_st + (grp < 2 ? 0 : p.KS) + (grp & 1) * 16, v);
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(e2done_bar);
      LL_TLE(31);
    };
    if (HEAD && pair0 < p.npairs) {
      load_window(pair0);
      stage_window();
      load_window(pair0 + pair_step);
    }
    for (long long pt = pair0; pt < p.npairs; pt += pair_step, ++tli) {
      const long long t = 2 * pt + rank;
      const int b = (int)(t / per_img);
      const int r = (int)(t % per_img);
      const int y = (r / p.tiles_x) * IG_TH + ty, x = (r % p.tiles_x) * IG_TW + tx;
      const bool valid = t < p.ntiles && y < p.H && x < p.W;
      const long long px = ((long long)b * p.H + y) * p.W + x;
      {
        LL_TLE(32);
        mbar_wait_spin(tfull_bar, tf_phase);
        tf_phase ^= 1;
        tc_fence_after();
        LL_TLE(33);
        // E1: y = main + small + bias, in place
        for (int c = grp; c < nch; c += 4) {
          uint32_t v[16], w[16];
          tc_ld16(tlane + c * 16, v);
          tc_ld16(tlane + p.N + c * 16, w);
          tc_wait_ld();
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __float_as_uint((__uint_as_float(v[j]) + __uint_as_float(w[j])) + s_bias[c * 16 + j]);
          tmem_st16(tlane + c * 16, v);
        }
      }
      tmem_wait_st();
      LL_TLE(34);
      // this warp's E1 is done (its reads of the previous tile's norm were released at the end of that tile's E2)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(e2done_bar);
      // staging rounds of one GDN pass.  Pass ps + 1 is staged between the register half of pass ps's E2 and its global
      // stores: the stores keep the LSU busy for thousands of cycles (one 32-byte sector per lane and instruction), and the
      // next pass's norm MMAs now run underneath them instead of after them (conv 96->192: E2[0] took 7 k of the tile's
      // 59 k cycles with the tensor pipe idle)
      auto stage_pass = [&](int ps) {
        for (int rd = 0; rd < p.rpp; ++rd, ++ground) {
          LL_TLE(36 + 2 * (ps * p.rpp + rd));
          for (int cc = grp; cc < slots; cc += 4) {              // the slots this warp owns (slot index = chunk index mod 4)
            const int gc = rd * slots + cc;                      // global chunk index of these y columns
            // the MMAs that read this slot in the previous round (of this or of the previous tile) are complete
            if (ground) mbar_wait_spin(gdone_bar(cc), (ground - 1) & 1u);
            tc_fence_after();
            uint32_t v[16], lo[16];
            tc_ld16(tlane + gc * 16, v);
            tc_wait_ld();
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float yy = __uint_as_float(v[j]);
              const float sq = yy * yy;
              const float h = tf32_rna(sq);
              v[j] = __float_as_uint(h);
              lo[j] = __float_as_uint(tf32_rna(sq - h));
            }
            tmem_st16(tlane + p.col_st + cc * 16, v);
            tmem_st16(tlane + p.col_st + p.KS + cc * 16, lo);
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(aready_bar(cc));
          }
          LL_TLE(37 + 2 * (ps * p.rpp + rd));
        }
      };
      stage_pass(0);
      for (int ps = 0; ps < p.passes; ++ps) {
        // norm of channels [ps NP, ps NP + NP) complete
        mbar_wait_spin(pdone_bar, pd_phase);
        pd_phase ^= 1;
        tc_fence_after();
        LL_TLE(50 + 2 * ps);
        // E2, first half: y and the norm of this warp's (at most two) chunks of the pass -> registers -> z
        const int cbase = ps * npc;
        const int first = cbase + ((grp - cbase) & 3);
        float rr[2][16];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int gc = first + 4 * i;
          if (gc < cbase + npc) {
            const int cc = gc - cbase;
            uint32_t yv[16], nm[16], ns[16];
            tc_ld16(tlane + gc * 16, yv);
            tc_ld16(tlane + p.col_nm + cc * 16, nm);
            tc_ld16(tlane + p.col_ns + cc * 16, ns);
            tc_wait_ld();
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float a = (__uint_as_float(nm[j]) + __uint_as_float(ns[j])) + s_beta[gc * 16 + j];
              rr[i][j] = __uint_as_float(yv[j]) * (p.inverse ? sqrt_approx(a) : rsqrt_approx(a));
            }
          }
        }
        // tensor memory released: the next pass's norm MMAs / the next tile's mainloop may overwrite the accumulators
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_leader(ps + 1 < p.passes ? e2done_bar : tempty_bar);
        if (ps + 1 < p.passes) stage_pass(ps + 1);
        if (HEAD && ps + 1 == p.passes && pt + pair_step < p.npairs) {
          stage_window();
          load_window(pt + 2 * pair_step);
        }
        // E2, second half: [hi | lo] split and the kernel's only global stores, overlapping the tensor work that follows
        if (valid) {
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            const int gc = first + 4 * i;
            if (gc < cbase + npc) {
              float* zp = p.sz + px * (2 * p.N) + gc * 16;
#pragma unroll
              for (int j0 = 0; j0 < 16; j0 += 8) {
                float hi[8], lo[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  hi[j] = tf32_rna(rr[i][j0 + j]);
                  lo[j] = tf32_rna(rr[i][j0 + j] - hi[j]);
                }
#ifdef LL_TIMELINE
                if (g_nostore_dev && hi[0] != 12345.678f) continue;
#endif
                st_global_v8(zp + j0, hi);
                st_global_v8(zp + p.N + j0, lo);
              }
            }
          }
        }
        LL_TLE(51 + 2 * ps);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, IG_TMEM_COLS);
  }
}

// Small-Cin convs of the context models written channels-last in bf16 for the igemm layers:
//   plc head  Conv2d(3, 243, 3, padding=1) + LeakyReLU on the nearest-2x-upsampled parent (:271,355)
//   csc       MaskedConv2d('A', 3, 243, 5, padding=2, groups=3) on the quantised child (:274-277,353);
//             only the first `live_taps` taps in row-major order are non-zero under the mask (12 of 25).
// One thread = one pixel x 8 consecutive destination channels (one 16-byte store).  Destination
// channel d of the covered region maps to conv channel co = (d / co_gstride) * co_group + d % co_gstride
// (valid when d % co_gstride < co_group and co < Cout); every other destination channel of the
// region is written as 0, so a consumer igemm layer never reads uninitialised memory.
// Mapping: one warp = 4 consecutive pixels of a row x 32 slots of 8 destination channels (lane = slot),
// so every pixel is stored as one contiguous 512-byte run.  Weights live in shared memory k-major
// ([tap][dest channel], zero where unmapped): two LDS.128 per tap feed 8 channels x 4 pixels of packed
// FFMA2; the (K-1+4)-wide input window of the 4 pixels is loaded once into registers (warp-uniform
// addresses for a dense conv).
constexpr int HD_THREADS = 256;
constexpr int HD_PX = 4;
struct SmallConvParams {
  const float* x; const float* w; const float* bias;
  __nv_bfloat16* out;
  int B, Cin, H, W, Cout, groups, upsample2, lrelu;
  int out_cstride, out_coff, co_group, co_gstride, region;
};

template <int CIN_G, int K, int LIVE>
__global__ void __launch_bounds__(HD_THREADS) ctx_conv_nhwc_kernel(const __grid_constant__ SmallConvParams p) {
  constexpr int NIN = CIN_G * LIVE;
  constexpr int ROWS = (LIVE + K - 1) / K;           // live kernel rows (3 of 5 under mask 'A')
  constexpr int WCOLS = HD_PX + K - 1;
  constexpr int PAD = K / 2;
  extern __shared__ __align__(16) float s_w[];       // [NIN][region] + bias[region]
  const int cout_g = p.Cout / p.groups;
  float* s_b = s_w + NIN * p.region;
  // per-warp input window: [Cin][ROWS][WCOLS] samples around the warp's 4 pixels.  The 32 lanes fetch it together
  // (one bounds check + one load per sample for the whole warp) and then read it back as broadcasts; loading it per
  // lane made ~800 of the kernel's ~1 800 instructions per work item redundant address arithmetic (ncu: issue-bound
  // at 23 % occupancy, FMA pipe 45 % busy).
  constexpr int WIN_CH = ROWS * WCOLS;
  float* s_win = s_b + p.region + (threadIdx.x >> 5) * (p.Cin * WIN_CH);
  for (int i = threadIdx.x; i < (NIN + 1) * p.region; i += HD_THREADS) {
    const int k = i / p.region, d = i % p.region;
    const int gd = d / p.co_gstride, j = d % p.co_gstride, co = gd * p.co_group + j;
    float v = 0.f;
    if (j < p.co_group && co < p.Cout) {
      if (k < NIN) v = p.w[((long long)co * CIN_G + k / LIVE) * (K * K) + k % LIVE];
      else v = p.bias ? p.bias[co] : 0.f;
    }
    // tap rows are stored as [half][slot][4]: the first (second) 128-bit load of the 32 lanes of a warp then covers 512
    // contiguous bytes -- with the natural [slot][8] order the lanes sit 32 bytes apart and every load is a 2-way bank
    // conflict, which made the kernel shared-memory bound (54 conflicted LDS.128 against 432 FFMA2 per 4 pixels).
    s_w[k < NIN ? k * p.region + ((d >> 2) & 1) * (p.region / 2) + (d >> 3) * 4 + (d & 3) : i] = v;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int chunks = (p.region / 8 + 31) / 32;       // 32-slot chunks per pixel
  const int qx = (p.W + HD_PX - 1) / HD_PX;          // pixel quads per row
  const int Hs = p.upsample2 ? p.H / 2 : p.H, Ws = p.upsample2 ? p.W / 2 : p.W;
  const long long nwork = (long long)p.B * p.H * qx * chunks;
  const long long wstride = (long long)gridDim.x * (HD_THREADS / 32);
  for (long long wk = (long long)blockIdx.x * (HD_THREADS / 32) + (threadIdx.x >> 5); wk < nwork; wk += wstride) {
    const int ch = (int)(wk % chunks);
    long long r = wk / chunks;
    const int x0 = (int)(r % qx) * HD_PX;
    r /= qx;
    const int yy = (int)(r % p.H), b = (int)(r / p.H);
    const int slot = ch * 32 + lane;
    const bool inreg = slot * 8 < p.region;              // lanes past the region idle but keep the warp's barriers
    const int d0 = inreg ? slot * 8 : 0;
    const int gd = d0 / p.co_gstride, j0 = d0 % p.co_gstride, co0 = gd * p.co_group + j0;
    const bool live = inreg && j0 < p.co_group && co0 < p.Cout;
    // (dense head only: the grouped csc has one input channel per lane group and 12 taps -- too little arithmetic to
    // hide the extra shared-memory hop, measured 470 -> 541 us; the dense head went 509 -> 389 us)
    constexpr bool kSharedWindow = CIN_G > 1;
    if (kSharedWindow) __syncwarp();                     // the previous item's window reads are done
    for (int e = lane; kSharedWindow && e < p.Cin * WIN_CH; e += 32) {
      const int c = e / WIN_CH, rr = (e % WIN_CH) / WCOLS, cc = e % WCOLS;
      const int gy = yy + rr - PAD, gx = x0 + cc - PAD;
      float v = 0.f;
      if (gy >= 0 && gy < p.H && gx >= 0 && gx < p.W)
        v = __ldg(p.x + ((long long)b * p.Cin + c) * Hs * Ws + (long long)(p.upsample2 ? gy >> 1 : gy) * Ws + (p.upsample2 ? gx >> 1 : gx));
      s_win[e] = v;
    }
    if (kSharedWindow) __syncwarp();
    float2 acc[HD_PX][4];
    {
      const float4 b0 = *reinterpret_cast<const float4*>(&s_b[d0]), b1 = *reinterpret_cast<const float4*>(&s_b[d0 + 4]);
#pragma unroll
      for (int q = 0; q < HD_PX; ++q) {
        acc[q][0] = make_float2(b0.x, b0.y); acc[q][1] = make_float2(b0.z, b0.w);
        acc[q][2] = make_float2(b1.x, b1.y); acc[q][3] = make_float2(b1.z, b1.w);
      }
    }
    if (live) {
      const int g = co0 / cout_g;                    // conv group of this slot (unique: checked on the host)
#pragma unroll
      for (int ci = 0; ci < CIN_G; ++ci) {
        float win[ROWS][WCOLS];
        if (kSharedWindow) {
          const float* wc = s_win + (g * CIN_G + ci) * WIN_CH;
#pragma unroll
          for (int rr = 0; rr < ROWS; ++rr) {
#pragma unroll
            for (int cc = 0; cc < WCOLS; cc += 2) {
              const float2 v2 = *reinterpret_cast<const float2*>(wc + rr * WCOLS + cc);
              win[rr][cc] = v2.x;
              win[rr][cc + 1] = v2.y;
            }
          }
        } else {
          const float* xc = p.x + ((long long)b * p.Cin + g * CIN_G + ci) * Hs * Ws;
#pragma unroll
          for (int rr = 0; rr < ROWS; ++rr) {
            const int gy = yy + rr - PAD;
            const bool yok = gy >= 0 && gy < p.H;
            const int sy = p.upsample2 ? gy >> 1 : gy;
#pragma unroll
            for (int cc = 0; cc < WCOLS; ++cc) {
              const int gx = x0 + cc - PAD;
              float v = 0.f;
              if (yok && gx >= 0 && gx < p.W) v = __ldg(xc + (long long)sy * Ws + (p.upsample2 ? gx >> 1 : gx));
              win[rr][cc] = v;
            }
          }
        }
#pragma unroll
        for (int t = 0; t < LIVE; ++t) {
          const float* wr = s_w + (ci * LIVE + t) * p.region + slot * 4;
          const float4 w0 = *reinterpret_cast<const float4*>(wr), w1 = *reinterpret_cast<const float4*>(wr + p.region / 2);
          const float2 wa = make_float2(w0.x, w0.y), wb = make_float2(w0.z, w0.w);
          const float2 wc = make_float2(w1.x, w1.y), wd = make_float2(w1.z, w1.w);
#pragma unroll
          for (int q = 0; q < HD_PX; ++q) {
            const float v = win[t / K][q + t % K];
            const float2 vv = make_float2(v, v);
            acc[q][0] = __ffma2_rn(vv, wa, acc[q][0]);
            acc[q][1] = __ffma2_rn(vv, wb, acc[q][1]);
            acc[q][2] = __ffma2_rn(vv, wc, acc[q][2]);
            acc[q][3] = __ffma2_rn(vv, wd, acc[q][3]);
          }
        }
      }
    }
#pragma unroll
    for (int q = 0; q < HD_PX; ++q) {
      if (x0 + q >= p.W || !inreg) break;
      uint32_t pk[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float2 a = acc[q][c];
        if (p.lrelu) { a.x = a.x < 0.f ? a.x * 0.01f : a.x; a.y = a.y < 0.f ? a.y * 0.01f : a.y; }
        __nv_bfloat162 h = __floats2bfloat162_rn(a.x, a.y);
        pk[c] = *reinterpret_cast<uint32_t*>(&h);
      }
      const long long px = ((long long)b * p.H + yy) * p.W + x0 + q;
      *reinterpret_cast<uint4*>(p.out + px * p.out_cstride + p.out_coff + d0) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
  }
}

// (Co, Ci, R, S) fp32 -> [R*S][Npad][Kpad] bf16, zero padded.
// ------------------------------------------------------------------------------------------------
// Operand builder for the two small-Cin context convs as tensor-core GEMMs (LiftingBasedDWT_net.py:271,274-277):
//   plc head  Conv2d(3, 243, 3, padding=1) on the nearest-2x-upsampled parent   K = 3 x 9 = 27
//   csc       MaskedConv2d('A', 3, 243, 5, padding=2, groups=3) on the child     K = 12 live taps per group
// As SIMT convs they produced 243 channels from 3 at ~21 TFLOP/s and cost 29 % of the layer.  Here one pass writes, per
// pixel, the im2col row of both convs as a bf16 SPLIT  v = hi + lo  (hi = bf16(v), lo = bf16(v - hi): the inputs are
// quantised coefficients, integers well beyond bf16's 8 bits), and the convs become 1-tap igemm layers whose packed
// weights repeat the split on their side:   [x_hi | x_lo | x_hi] . [W_hi | W_hi | W_lo]   (16 bits per operand; the
// fp32 SIMT kernel rounded its OUTPUT to bf16, which stays the dominant error).
// Channels of the output pixel (320 bf16):  [0, 81) head: part p = c / 27 (hi, lo, hi), k = c % 27 = ci * 9 + tap;
// [81, 128) zero;  128 + 64 g + [0, 36): csc group g: part p = r / 12, tap = r % 12 (row-major 5x5);  rest zero.
// One thread = one pixel x 8 channels (a 16-byte store), 40 slots per pixel.
// ------------------------------------------------------------------------------------------------
constexpr int IM_CH = 320, IM_SLOTS = IM_CH / 8;
constexpr int IM_SPAN = 128;                       // pixels of one image row per block
constexpr int IM_THREADS = IM_SLOTS * 8;           // 40 slots x 8 pixel lanes
constexpr int IM_UPW = IM_SPAN + 2, IM_CHW = IM_SPAN + 4;
constexpr int IM_UP = 0, IM_CHD = IM_UP + 3 * 3 * IM_UPW, IM_ZERO = IM_CHD + 3 * 3 * IM_CHW, IM_TILE = IM_ZERO + IM_SPAN + 8;
// Block = one image row x IM_SPAN pixels.  The 3 x 3 rows of the upsampled parent window and the 3 x 3 child rows the live
// taps touch (dy = -2, -1, 0) are staged in shared memory with the zero padding already in place, so a channel is one
// LDS at a per-thread constant offset: thread = (slot of 8 channels, pixel lane); the channel decode (divisions) runs once.
__global__ void __launch_bounds__(IM_THREADS) ctx_im2col_kernel(const float* __restrict__ con, const float* __restrict__ q,
                                                                __nv_bfloat16* __restrict__ out, int B, int H, int W) {
  __shared__ float tile[IM_TILE];
  const int spans = (W + IM_SPAN - 1) / IM_SPAN;
  const int hs = H >> 1, ws = W >> 1;
  const int slot = threadIdx.x % IM_SLOTS, plane = threadIdx.x / IM_SLOTS;
  int off[8];
  unsigned lo_mask = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = slot * 8 + j;
    int o = IM_ZERO;                                 // a run of zeros
    if (c < 81) {
      const int pt = c / 27, k = c % 27, ci = k / 9, tap = k % 9;
      o = IM_UP + (ci * 3 + tap / 3) * IM_UPW + tap % 3;            // + local x
      if (pt == 1) lo_mask |= 1u << j;
    } else if (c >= 128) {
      const int g = (c - 128) >> 6, r = (c - 128) & 63;
      if (r < 36) {
        const int pt = r / 12, tap = r % 12;
        o = IM_CHD + (g * 3 + tap / 5) * IM_CHW + tap % 5;
        if (pt == 1) lo_mask |= 1u << j;
      }
    }
    off[j] = o;
  }
  for (long long blk = blockIdx.x; blk < (long long)B * H * spans; blk += gridDim.x) {
    const int sp = (int)(blk % spans), y = (int)((blk / spans) % H), b = (int)(blk / ((long long)spans * H));
    const int x0 = sp * IM_SPAN;
    __syncthreads();
    for (int i = threadIdx.x; i < IM_TILE; i += IM_THREADS) {
      float v = 0.f;
      if (i < IM_CHD) {
        const int ci = i / (3 * IM_UPW), r = (i / IM_UPW) % 3, cx = i % IM_UPW;
        const int yy = y + r - 1, xx = x0 + cx - 1;
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) v = __ldg(con + (((long long)b * 3 + ci) * hs + (yy >> 1)) * ws + (xx >> 1));
      } else if (i < IM_ZERO) {
        const int k = i - IM_CHD, g = k / (3 * IM_CHW), r = (k / IM_CHW) % 3, cx = k % IM_CHW;
        const int yy = y + r - 2, xx = x0 + cx - 2;
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) v = __ldg(q + (((long long)b * 3 + g) * H + yy) * W + xx);
      }
      tile[i] = v;
    }
    __syncthreads();
    __nv_bfloat16* orow = out + (((long long)b * H + y) * W + x0) * IM_CH + slot * 8;
    for (int xl = plane; xl < IM_SPAN && x0 + xl < W; xl += 8) {
      __nv_bfloat16 o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float v = tile[off[j] + xl];
        const __nv_bfloat16 hi = __float2bfloat16_rn(v);
        o[j] = ((lo_mask >> j) & 1u) ? __float2bfloat16_rn(v - __bfloat162float(hi)) : hi;
      }
      *reinterpret_cast<uint4*>(orow + (long long)xl * IM_CH) = *reinterpret_cast<const uint4*>(o);
    }
  }
}

__global__ void pack_igemm_weight_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wp, int Co, int Ci,
                                         int taps, int Npad, int Kpad) {
  const long long total = (long long)taps * Npad * Kpad;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)(e % Kpad), co = (int)((e / Kpad) % Npad), tap = (int)(e / ((long long)Kpad * Npad));
    float v = 0.f;
    if (co < Co && ci < Ci) v = w[((long long)co * Ci + ci) * taps + tap];
    wp[e] = __float2bfloat16_rn(v);
  }
}

// fp32 NCHW channel slice -> bf16 NHWC channel slice (used to drop csc's output into the cgp input layout)
__global__ void nchw_to_nhwc_bf16_kernel(const float* __restrict__ x, long long x_sb, __nv_bfloat16* __restrict__ out,
                                         int B, int C, int H, int W, int out_cstride, int out_coff) {
  const long long hw = (long long)H * W;
  const long long total = (long long)B * hw * C;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(e % C);
    const long long px = e / C;
    const int b = (int)(px / hw);
    const long long s = px % hw;
    out[px * out_cstride + out_coff + c] = __float2bfloat16_rn(x[(long long)b * x_sb + (long long)c * hw + s]);
  }
}

// (Co, Ci, R, S) fp32 -> [R*S][Npad][2*Kpad] fp32: [hi | lo] TF32 halves along K, zero padded.  ``transposed``:
// the source is a ConvTranspose2d weight (Ci_t = Co of the equivalent conv ... stored (in, out, R, S)); stride 1,
// padding 1: equivalent conv weight w'[o][i][tap] = w[i][o][8 - tap].
__global__ void pack_tf32_weight_kernel(const float* __restrict__ w, float* __restrict__ wp, int Co, int Ci, int taps,
                                        int Npad, int Kpad, int transposed) {
  const long long total = (long long)taps * Npad * Kpad;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int ci = (int)(e % Kpad), co = (int)((e / Kpad) % Npad), tap = (int)(e / ((long long)Kpad * Npad));
    float v = 0.f;
    if (co < Co && ci < Ci) v = transposed ? w[((long long)ci * Co + co) * taps + (taps - 1 - tap)] : w[((long long)co * Ci + ci) * taps + tap];
    const float hi = tf32_rna(v);
    float* row = wp + ((long long)tap * Npad + co) * (2 * Kpad);
    row[ci] = hi;
    row[Kpad + ci] = tf32_rna(v - hi);
  }
}

// 32 pixels x 32 channels per CTA through a shared-memory tile: reads run along the pixels (contiguous in NCHW), writes
// along the channels (contiguous in NHWC).  (The first version read with a stride of one plane per thread.)
__global__ void __launch_bounds__(256) nchw_to_nhwc_split_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                                 float* __restrict__ sz, int B, int C, int H, int W, int mode) {
  __shared__ float tile[32][33];
  const long long hw = (long long)H * W;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;       // 32 x 8
  const long long ptiles = (hw + 31) / 32;
  const int ctiles = (C + 31) / 32;
  const long long ntiles = (long long)B * ptiles * ctiles;
  for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int ct = (int)(t % ctiles);
    const long long r = t / ctiles;
    const long long pt = r % ptiles, b = r / ptiles;
    const long long p0 = pt * 32;
    const int c0 = ct * 32;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = c0 + ty + 8 * k;
      const long long pix = p0 + tx;
      tile[ty + 8 * k][tx] = (c < C && pix < hw) ? x[(b * C + c) * hw + pix] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long long pix = p0 + ty + 8 * k;
      const int c = c0 + tx;
      if (c < C && pix < hw) {
        const float v = tile[tx][ty + 8 * k];
        const long long px = b * hw + pix;
        if (y) y[px * C + c] = v;
        const float rr = mode ? v : v * v;
        const float hi = tf32_rna(rr);
        sz[px * (2 * C) + c] = hi;
        sz[px * (2 * C) + C + c] = tf32_rna(rr - hi);
      }
    }
    __syncthreads();
  }
}

__global__ void nhwc_split_to_nchw_kernel(const float* __restrict__ z, float* __restrict__ out, int B, int C, int H, int W) {
  const long long hw = (long long)H * W, total = (long long)B * hw * C;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long s = e % hw;
    const int c = (int)((e / hw) % C);
    const long long b = e / (hw * C);
    const float* q = z + (b * hw + s) * (2 * C);
    out[e] = q[c] + q[C + c];
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess) return nullptr;
    if (q != cudaDriverEntryPointSuccess) return nullptr;
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}

}  // namespace ll

using namespace ll;

extern "C" {

// con (B,3,H/2,W/2) fp32 quantised parent, q (B,3,H,W) fp32 quantised child -> out (B,H,W,320) bf16: the split im2col rows of
// the plc head (on the nearest-2x-upsampled parent) and of the masked csc (see ctx_im2col_kernel)
int ll_ctx_im2col(const float* con, const float* q, void* out, int B, int H, int W, ll_stream_t stream) {
  if (B < 0 || H < 0 || W < 0 || (H & 1) || (W & 1)) return fail(LL_EINVAL, "ll_ctx_im2col: bad extents (even H, W)");
  if ((long long)B * H * W == 0) return LL_OK;
  if (!con || !q || !out || ((uintptr_t)out & 15)) return fail(LL_EINVAL, "ll_ctx_im2col: null / misaligned pointer");
  long long blocks = (long long)B * H * ((W + IM_SPAN - 1) / IM_SPAN);
  const long long cap = (long long)sm_count_cached() * 12;
  if (blocks > cap) blocks = cap;
  ctx_im2col_kernel<<<(unsigned)blocks, IM_THREADS, 0, as_stream(stream)>>>(con, q, reinterpret_cast<__nv_bfloat16*>(out), B, H, W);
  LL_LAUNCH_OK("ctx_im2col_kernel");
  return LL_OK;
}

int ll_ctx_conv_nhwc(const float* x, const float* w, const float* bias, void* out, int B, int Cin, int H, int W, int Cout,
                     int K, int groups, int live_taps, int upsample2, int lrelu, int out_cstride, int out_coff,
                     int co_group, int co_gstride, int region, ll_stream_t stream) {
  if (B < 0 || H < 0 || W < 0 || Cin < 1 || Cout < 1 || groups < 1 || Cin % groups || Cout % groups)
    return fail(LL_EINVAL, "ll_ctx_conv_nhwc: bad extents");
  const int cin_g = Cin / groups;
  const bool head = (cin_g == 3 && K == 3 && live_taps == 9), csc = (cin_g == 1 && K == 5 && live_taps == 12);
  if (!head && !csc)
    return fail(LL_EINVAL, "ll_ctx_conv_nhwc: built for (Cin/groups=3, K=3, 9 taps) and (Cin/groups=1, K=5, 12 live taps), got (%d, %d, %d)",
                cin_g, K, live_taps);
  if (co_group <= 0) { co_group = Cout; co_gstride = region > 0 ? region : Cout; }
  if (region <= 0 || region % 8 || out_coff % 8 || out_cstride % 8 || out_coff < 0 || out_cstride < out_coff + region || co_gstride < co_group)
    return fail(LL_EINVAL, "ll_ctx_conv_nhwc: region / offsets must be multiples of 8 channels inside the output pixel");
  if (groups > 1 && (co_group != Cout / groups || co_gstride % 8))
    return fail(LL_EINVAL, "ll_ctx_conv_nhwc: grouped convs need co_group == Cout/groups and an 8-aligned group stride");
  if (upsample2 && ((H & 1) || (W & 1))) return fail(LL_EINVAL, "ll_ctx_conv_nhwc: upsample2 needs even output size");
  if ((long long)B * H * W == 0) return LL_OK;
  if (!x || !w || !out) return fail(LL_EINVAL, "ll_ctx_conv_nhwc: null pointer");
  SmallConvParams p = {};
  p.x = x; p.w = w; p.bias = bias; p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.B = B; p.Cin = Cin; p.H = H; p.W = W; p.Cout = Cout; p.groups = groups;
  p.upsample2 = upsample2; p.lrelu = lrelu;
  p.out_cstride = out_cstride; p.out_coff = out_coff; p.co_group = co_group; p.co_gstride = co_gstride; p.region = region;
  const int win_rows = (live_taps + K - 1) / K, win_cols = HD_PX + K - 1;   // per-warp input window, see the kernel
  const size_t smem = ((size_t)(cin_g * live_taps + 1) * region + (size_t)(HD_THREADS / 32) * Cin * win_rows * win_cols) * sizeof(float);
  if (smem > 200 * 1024) return fail(LL_EINVAL, "ll_ctx_conv_nhwc: weights do not fit shared memory");
  const int chunks = (region / 8 + 31) / 32;
  const long long nwork = (long long)B * H * ((W + HD_PX - 1) / HD_PX) * chunks;
  long long blocks = (nwork + HD_THREADS / 32 - 1) / (HD_THREADS / 32);
  const long long cap = (long long)sm_count_cached() * 4;
  if (blocks > cap) blocks = cap;
  if (head) {
    if (smem > 48 * 1024) LL_CUDA_OK(cudaFuncSetAttribute(ctx_conv_nhwc_kernel<3, 3, 9>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ctx_conv_nhwc_kernel<3, 3, 9><<<(unsigned)blocks, HD_THREADS, smem, as_stream(stream)>>>(p);
  } else {
    if (smem > 48 * 1024) LL_CUDA_OK(cudaFuncSetAttribute(ctx_conv_nhwc_kernel<1, 5, 12>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ctx_conv_nhwc_kernel<1, 5, 12><<<(unsigned)blocks, HD_THREADS, smem, as_stream(stream)>>>(p);
  }
  LL_LAUNCH_OK("ctx_conv_nhwc_kernel");
  return LL_OK;
}

int ll_pack_igemm_weight(const float* w, void* wp, int Co, int Ci, int taps, int Npad, int Kpad, ll_stream_t stream) {
  if (Co < 1 || Ci < 1 || (taps != 1 && taps != 9) || Npad < Co || Kpad < Ci || Npad % 16 || Npad > IG_MAXN || Kpad % IG_BK)
    return fail(LL_EINVAL, "ll_pack_igemm_weight: bad extents (taps 1|9, Npad %%16 <= 256, Kpad %%64)");
  if (!w || !wp) return fail(LL_EINVAL, "ll_pack_igemm_weight: null pointer");
  const long long total = (long long)taps * Npad * Kpad;
  pack_igemm_weight_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(
      w, reinterpret_cast<__nv_bfloat16*>(wp), Co, Ci, taps, Npad, Kpad);
  LL_LAUNCH_OK("pack_igemm_weight_kernel");
  return LL_OK;
}

int ll_nchw_to_nhwc_bf16(const float* x, int64_t x_sb, void* out, int B, int C, int H, int W, int out_cstride, int out_coff,
                         ll_stream_t stream) {
  if (B < 0 || C < 1 || H < 0 || W < 0 || out_cstride < C + out_coff || out_coff < 0)
    return fail(LL_EINVAL, "ll_nchw_to_nhwc_bf16: bad extents");
  const long long total = (long long)B * H * W * C;
  if (total == 0) return LL_OK;
  if (!x || !out) return fail(LL_EINVAL, "ll_nchw_to_nhwc_bf16: null pointer");
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)sm_count_cached() * 16;
  if (blocks > cap) blocks = cap;
  nchw_to_nhwc_bf16_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(x, x_sb, reinterpret_cast<__nv_bfloat16*>(out), B, C,
                                                                          H, W, out_cstride, out_coff);
  LL_LAUNCH_OK("nchw_to_nhwc_bf16_kernel");
  return LL_OK;
}

struct TailArgs {
  const float *w3, *b3, *w4, *b4, *x, *noise;
  float* bits;
  double* sum;
  long long x_sb, bits_sb;
  int c3;
};

static int launch_igemm_bf16(const void* x_nhwc, const void* wp, const float* bias, int B, int H, int W, int Cin_total, int Kpad,
                  int Npad, int Cout, int taps, int groups, const int* koff, int lrelu, float* out_f32, int64_t out_sb,
                  int co_group, int co_stride, int co_off, void* out_bf16, int out_cstride, int out_coff, int out_gstride,
                  const TailArgs* tail, ll_stream_t stream) {
  if (B < 0 || H < 0 || W < 0 || Kpad < IG_BK || Kpad % IG_BK || Npad < 16 || Npad % 16 || Npad > IG_MAXN || Cout < 1 ||
      Cout > Npad || (taps != 1 && taps != 9))
    return fail(LL_EINVAL, "ll_igemm_conv: bad extents (Kpad %%64, Npad %%16 in 16..256, Cout <= Npad, taps 1|9)");
  if (groups < 1 || groups > IG_MAXG || Kpad / IG_BK > IG_MAXSLOTS)
    return fail(LL_EINVAL, "ll_igemm_conv: groups in 1..%d, at most %d k-blocks per group", IG_MAXG, IG_MAXSLOTS);
  if (Cin_total < IG_BK || Cin_total % 8) return fail(LL_EINVAL, "ll_igemm_conv: input channel count must be >= 64 and a multiple of 8");
  if (!koff && (groups != 1 || Cin_total != Kpad)) return fail(LL_EINVAL, "ll_igemm_conv: koff is required when groups > 1 or Cin_total != Kpad");
  if (B > 65535 * 4 || H > (1 << 20) || W > (1 << 20)) return fail(LL_EINVAL, "ll_igemm_conv: extents too large");
  if ((long long)B * H * W == 0) return LL_OK;
  if (!x_nhwc || !wp || (!out_f32 && !out_bf16 && !tail)) return fail(LL_EINVAL, "ll_igemm_conv: null pointer");
  if (((uintptr_t)x_nhwc & 15) || ((uintptr_t)wp & 15)) return fail(LL_EINVAL, "ll_igemm_conv: operands must be 16-byte aligned");
  if (out_bf16 && (out_coff < 0 || out_gstride < 0 || out_cstride < out_coff + (groups - 1) * out_gstride + Npad))
    return fail(LL_EINVAL, "ll_igemm_conv: bad NHWC output slice (all Npad channels of every group are written)");
  if (groups > 1 && out_bf16 && out_gstride < Npad) return fail(LL_EINVAL, "ll_igemm_conv: NHWC group stride < Npad");
  if (co_group <= 0) { co_group = 1 << 30; co_stride = 0; co_off = 0; }
  EncodeTiledFn enc = encode_fn();
  if (!enc) return fail(LL_ECUDA, "ll_igemm_conv: cuTensorMapEncodeTiled not available from the driver");

  CUtensorMap tmA, tmB;
  {
    cuuint64_t gdim[4] = {(cuuint64_t)Cin_total, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t gstr[3] = {(cuuint64_t)Cin_total * 2, (cuuint64_t)W * Cin_total * 2, (cuuint64_t)H * W * Cin_total * 2};
    cuuint32_t box[4] = {IG_BK, IG_TW, IG_TH, 1};
    cuuint32_t est[4] = {1, 1, 1, 1};
    CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x_nhwc), gdim, gstr, box, est,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(LL_ECUDA, "ll_igemm_conv: cuTensorMapEncodeTiled(A) failed with %d", (int)r);
  }
  {
    cuuint64_t gdim[3] = {(cuuint64_t)Kpad, (cuuint64_t)Npad, (cuuint64_t)taps * groups};
    cuuint64_t gstr[2] = {(cuuint64_t)Kpad * 2, (cuuint64_t)Npad * Kpad * 2};
    cuuint32_t box[3] = {IG_BK, (cuuint32_t)Npad, 1};
    cuuint32_t est[3] = {1, 1, 1};
    CUresult r = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(wp), gdim, gstr, box, est,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(LL_ECUDA, "ll_igemm_conv: cuTensorMapEncodeTiled(B) failed with %d", (int)r);
  }
  IgemmParams p = {};
  p.bias = bias;
  p.out_f32 = out_f32; p.out_sb = out_sb;
  p.co_group = co_group; p.co_stride = co_stride; p.co_off = co_off;
  p.out_bf16 = reinterpret_cast<__nv_bfloat16*>(out_bf16); p.out_cstride = out_cstride; p.out_coff = out_coff;
  p.B = B; p.H = H; p.W = W; p.Cout = Cout; p.Npad = Npad; p.kblocks = Kpad / IG_BK; p.taps = taps; p.lrelu = lrelu;
  p.tiles_x = (W + IG_TW - 1) / IG_TW;
  p.tiles_y = (H + IG_TH - 1) / IG_TH;
  p.ntiles = (long long)B * p.tiles_x * p.tiles_y * groups;
  p.groups = groups; p.out_gstride = out_gstride;
  for (int g = 0; g < groups; ++g)
    for (int k = 0; k < p.kblocks; ++k) {
      const int off = koff ? koff[g * p.kblocks + k] : k * IG_BK;
      if (off < 0 || off + IG_BK > Cin_total || off % 8)
        return fail(LL_EINVAL, "ll_igemm_conv: k-block offset %d outside the %d input channels or not a multiple of 8", off, Cin_total);
      p.a_koff[g][k] = off;
    }
  static thread_local bool attr[64] = {false};
  int dev = 0;
  LL_CUDA_OK(cudaGetDevice(&dev));
  if (dev < 64 && !attr[dev]) {
    LL_CUDA_OK(cudaFuncSetAttribute(igemm_conv_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, IG_SMEM_BYTES));
    attr[dev] = true;
  }
  for (int k = 0; k < p.kblocks; ++k) p.b_koff[k] = k * IG_BK;
  p.nstages = 2; p.stage_cols = IG_MAXN; p.nacc = 1; p.acc_cols = 0;
  if (tail) {
    p.epi = 4;
    p.t_w3 = tail->w3; p.t_b3 = tail->b3; p.t_w4 = tail->w4; p.t_b4 = tail->b4; p.t_x = tail->x; p.t_noise = tail->noise;
    p.t_bits = tail->bits; p.t_sum = tail->sum; p.t_xsb = tail->x_sb; p.t_bsb = tail->bits_sb; p.t_c3 = tail->c3;
  }
  const long long sms = sm_count_cached();
  const unsigned grid = (unsigned)(p.ntiles < sms ? p.ntiles : sms);
  igemm_conv_kernel<false><<<grid, IG_THREADS, IG_SMEM_BYTES, as_stream(stream)>>>(tmA, tmB, p);
  LL_LAUNCH_OK("igemm_conv_kernel");
  return LL_OK;
}

int ll_igemm_conv(const void* x_nhwc, const void* wp, const float* bias, int B, int H, int W, int Cin_total, int Kpad,
                  int Npad, int Cout, int taps, int groups, const int* koff, int lrelu, float* out_f32, int64_t out_sb,
                  int co_group, int co_stride, int co_off, void* out_bf16, int out_cstride, int out_coff, int out_gstride,
                  ll_stream_t stream) {
  return launch_igemm_bf16(x_nhwc, wp, bias, B, H, W, Cin_total, Kpad, Npad, Cout, taps, groups, koff, lrelu, out_f32, out_sb,
                           co_group, co_stride, co_off, out_bf16, out_cstride, out_coff, out_gstride, nullptr, stream);
}

// cgp layers 2-4 + Gaussian rate in one launch (LiftingBasedDWT_net.py:362-365): h1 (B,H,W,Cin_total) bf16 -> layer 2 as a
// grouped 1x1 tcgen05 GEMM (weights wp [groups][1][Npad = 64][Kpad], bias2 (groups*C2), LeakyReLU) whose accumulator is
// consumed in place by layers 3 (w3 (groups*C3, C2), b3, LeakyReLU) and 4 (w4 (2*groups, C3): row 2g = sigma, 2g+1 = mu; b4)
// and the rate of x (B,groups,H,W; batch stride x_sb) -> bits (batch stride bits_sb), sum of bits added to *sum_out.
int ll_igemm_cgp_tail(const void* h1, const void* wp, const float* bias2, int B, int H, int W, int Cin_total, int Kpad, int C2,
                      int groups, const int* koff, const float* w3, const float* b3, const float* w4, const float* b4, int C3,
                      const float* x, int64_t x_sb, const float* noise, float* bits, int64_t bits_sb, double* sum_out,
                      ll_stream_t stream) {
  if (C2 < 1 || C2 > IG_BK || C3 < 1 || C3 > TL_K) return fail(LL_EINVAL, "ll_igemm_cgp_tail: C2 <= %d and C3 <= %d", IG_BK, TL_K);
  if ((long long)B * H * W == 0) return LL_OK;
  if (!w3 || !b3 || !w4 || !b4 || !x || !bits || !bias2) return fail(LL_EINVAL, "ll_igemm_cgp_tail: null pointer");
  TailArgs t = {w3, b3, w4, b4, x, noise, bits, sum_out, (long long)x_sb, (long long)bits_sb, C3};
  return launch_igemm_bf16(h1, wp, bias2, B, H, W, Cin_total, Kpad, IG_BK, C2, 1, groups, koff, 1, nullptr, 0, 0, 0, 0, nullptr, 0, 0, 0,
                           &t, stream);
}

// ------------------------------------------------------------------------------------------------
// 3xTF32 chain for SubbandAutoEncoderBerk (conv3x3 / GDN 1x1 with fp32-level accuracy)
// ------------------------------------------------------------------------------------------------
int ll_pack_tf32_weight(const float* w, float* wp, int Co, int Ci, int taps, int Npad, int Kpad, int transposed,
                        ll_stream_t stream) {
  if (Co < 1 || Ci < 1 || (taps != 1 && taps != 9) || Npad < Co || Kpad < Ci || Npad % 16 || Npad > IG_MAXN || Kpad % 32)
    return fail(LL_EINVAL, "ll_pack_tf32_weight: bad extents (taps 1|9, Npad %%16 <= 256, Kpad %%32)");
  if (!w || !wp) return fail(LL_EINVAL, "ll_pack_tf32_weight: null pointer");
  const long long total = (long long)taps * Npad * Kpad;
  pack_tf32_weight_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(w, wp, Co, Ci, taps, Npad, Kpad, transposed);
  LL_LAUNCH_OK("pack_tf32_weight_kernel");
  return LL_OK;
}

static int launch_igemm_tf32_pair(const float* a_nhwc, const float* wp, const float* bias, int B, int H, int W, int C, int Npad,
                                  int Cout, int taps, int epi, int inverse, float* y, float* sz, cudaStream_t stream) {
  EncodeTiledFn enc = encode_fn();
  if (!enc) return fail(LL_ECUDA, "ll_igemm_tf32: cuTensorMapEncodeTiled not available from the driver");
  CUtensorMap tmA, tmB;
  const int Ca = 2 * C;   // [hi | lo]
  {
    cuuint64_t gdim[4] = {(cuuint64_t)Ca, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t gstr[3] = {(cuuint64_t)Ca * 4, (cuuint64_t)W * Ca * 4, (cuuint64_t)H * W * Ca * 4};
    cuuint32_t box[4] = {32, IG_TW, IG_TH, 1};
    cuuint32_t est[4] = {1, 1, 1, 1};
    CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(a_nhwc), gdim, gstr, box, est,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(LL_ECUDA, "ll_igemm_tf32: cuTensorMapEncodeTiled(A) failed with %d", (int)r);
  }
  {
    cuuint64_t gdim[3] = {(cuuint64_t)Ca, (cuuint64_t)Npad, (cuuint64_t)taps};
    cuuint64_t gstr[2] = {(cuuint64_t)Ca * 4, (cuuint64_t)Npad * Ca * 4};
    cuuint32_t box[3] = {32, (cuuint32_t)(Npad / 2), 1};     // each CTA of a pair fetches half of the weight rows
    cuuint32_t est[3] = {1, 1, 1};
    CUresult r = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(wp), gdim, gstr, box, est,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(LL_ECUDA, "ll_igemm_tf32: cuTensorMapEncodeTiled(B) failed with %d", (int)r);
  }
  PairParams p = {};
  p.bias = bias; p.y = y; p.sz = sz;
  p.B = B; p.H = H; p.W = W; p.C = C; p.Cout = Cout; p.Npad = Npad; p.taps = taps; p.epi = epi; p.inverse = inverse;
  p.tiles_x = (W + IG_TW - 1) / IG_TW;
  p.tiles_y = (H + IG_TH - 1) / IG_TH;
  p.ntiles = (long long)B * p.tiles_x * p.tiles_y;
  p.npairs = (p.ntiles + 1) / 2;
  p.kb = C / 32;
  p.bhalf_bytes = (Npad / 2) * 128;
  p.stage_bytes = 2 * PR_A_BYTES + 2 * p.bhalf_bytes;
  p.stages = (PR_SMEM_LIMIT - 1024 - PR_TAIL_BYTES) / p.stage_bytes;
  if (p.stages > PR_MAXST) p.stages = PR_MAXST;
  if (p.stages < 2) return fail(LL_EINVAL, "ll_igemm_tf32: stage of %d bytes leaves fewer than 2 pipeline stages", p.stage_bytes);
  p.acc_cols = (Npad + 31) / 32 * 32;
  if (2 * p.acc_cols <= IG_MAXN) { p.nacc_stages = 2; p.stage_cols = IG_MAXN; }
  else { p.nacc_stages = 1; p.stage_cols = IG_TMEM_COLS; }
  const int smem = 1024 + p.stages * p.stage_bytes + PR_TAIL_BYTES;
  static thread_local bool attr[64] = {false};
  int dev = 0;
  LL_CUDA_OK(cudaGetDevice(&dev));
  if (dev < 64 && !attr[dev]) {
    LL_CUDA_OK(cudaFuncSetAttribute(igemm_tf32_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PR_SMEM_LIMIT));
    attr[dev] = true;
  }
  long long pairs = sm_count_cached() / 2;
  if (pairs > p.npairs) pairs = p.npairs;
  if (pairs < 1) pairs = 1;
  igemm_tf32_pair_kernel<<<(unsigned)(2 * pairs), IG_THREADS, smem, stream>>>(tmA, tmB, p);
  LL_LAUNCH_OK("igemm_tf32_pair_kernel");
  return LL_OK;
}

int ll_igemm_tf32(const float* a_nhwc, const float* wp, const float* bias, int B, int H, int W, int C, int Npad, int Cout,
                  int taps, int epi, int inverse, float* y, float* sz, int pair, ll_stream_t stream) {
  if (B < 0 || H < 0 || W < 0 || C < 32 || C % 32 || Npad < 16 || Npad % 16 || Npad > IG_MAXN || Cout < 32 || Cout % 32 ||
      Cout > Npad || (taps != 1 && taps != 9) || epi < 1 || epi > 3)
    return fail(LL_EINVAL, "ll_igemm_tf32: bad extents (C, Cout multiples of 32, Npad %%16 <= 256, taps 1|9, epi 1..3)");
  if (3 * (C / 32) > IG_MAXSLOTS) return fail(LL_EINVAL, "ll_igemm_tf32: at most %d input channels", IG_MAXSLOTS / 3 * 32);
  if ((long long)B * H * W == 0) return LL_OK;
  if (!a_nhwc || !wp || !y || (epi != 3 && !sz)) return fail(LL_EINVAL, "ll_igemm_tf32: null pointer");
  if (((uintptr_t)a_nhwc & 15) || ((uintptr_t)wp & 15) || ((uintptr_t)y & 15) || ((uintptr_t)sz & 15))
    return fail(LL_EINVAL, "ll_igemm_tf32: buffers must be 16-byte aligned");
  if (pair) return launch_igemm_tf32_pair(a_nhwc, wp, bias, B, H, W, C, Npad, Cout, taps, epi, inverse, y, sz, as_stream(stream));
  EncodeTiledFn enc = encode_fn();
  if (!enc) return fail(LL_ECUDA, "ll_igemm_tf32: cuTensorMapEncodeTiled not available from the driver");
  CUtensorMap tmA, tmB;
  {
    const int Ca = 2 * C;   // [hi | lo]
    cuuint64_t gdim[4] = {(cuuint64_t)Ca, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t gstr[3] = {(cuuint64_t)Ca * 4, (cuuint64_t)W * Ca * 4, (cuuint64_t)H * W * Ca * 4};
    cuuint32_t box[4] = {32, IG_TW, IG_TH, 1};
    cuuint32_t est[4] = {1, 1, 1, 1};
    CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(a_nhwc), gdim, gstr, box, est,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(LL_ECUDA, "ll_igemm_tf32: cuTensorMapEncodeTiled(A) failed with %d", (int)r);
  }
  {
    const int Kw = 2 * C;   // packed weights: [hi | lo] along K
    cuuint64_t gdim[3] = {(cuuint64_t)Kw, (cuuint64_t)Npad, (cuuint64_t)taps};
    cuuint64_t gstr[2] = {(cuuint64_t)Kw * 4, (cuuint64_t)Npad * Kw * 4};
    cuuint32_t box[3] = {32, (cuuint32_t)Npad, 1};
    cuuint32_t est[3] = {1, 1, 1};
    CUresult r = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(wp), gdim, gstr, box, est,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(LL_ECUDA, "ll_igemm_tf32: cuTensorMapEncodeTiled(B) failed with %d", (int)r);
  }
  IgemmParams p = {};
  p.bias = bias;
  p.co_group = 1 << 30;
  p.B = B; p.H = H; p.W = W; p.Cout = Cout; p.Npad = Npad; p.taps = taps;
  p.tiles_x = (W + IG_TW - 1) / IG_TW;
  p.tiles_y = (H + IG_TH - 1) / IG_TH;
  p.ntiles = (long long)B * p.tiles_x * p.tiles_y;
  p.groups = 1;
  // per 32-channel block: A_lo*B_hi, A_hi*B_lo (the small terms first), then A_hi*B_hi
  const int kb = C / 32;
  p.kblocks = 3 * kb;
  for (int k = 0; k < kb; ++k) {
    p.a_koff[0][3 * k + 0] = C + 32 * k; p.b_koff[3 * k + 0] = 32 * k;
    p.a_koff[0][3 * k + 1] = 32 * k;     p.b_koff[3 * k + 1] = C + 32 * k;
    p.a_koff[0][3 * k + 2] = 32 * k;     p.b_koff[3 * k + 2] = 32 * k;
  }
  // The tensor core's FP32 accumulator rounds toward zero: every accumulation step costs up to one ulp of the
  // running sum, in one direction.  The two small-term products of each block go to a second accumulator (whose
  // magnitude, hence ulp, is 2^-11 of the main one) so that the main accumulator takes one step per block, not three.
  p.nacc = 2; p.acc_cols = (Npad + 31) / 32 * 32;
  if (2 * p.acc_cols <= IG_MAXN) { p.nstages = 2; p.stage_cols = IG_MAXN; }
  else { p.nstages = 1; p.stage_cols = IG_TMEM_COLS; }
  for (int k = 0; k < kb; ++k) { p.acc_sel[3 * k + 0] = 1; p.acc_sel[3 * k + 1] = 1; p.acc_sel[3 * k + 2] = 0; }
  p.epi = epi; p.inverse = inverse; p.y = y; p.sz = sz;
  static thread_local bool attr[64] = {false};
  int dev = 0;
  LL_CUDA_OK(cudaGetDevice(&dev));
  if (dev < 64 && !attr[dev]) {
    LL_CUDA_OK(cudaFuncSetAttribute(igemm_conv_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, IG_SMEM_BYTES));
    attr[dev] = true;
  }
  const long long sms = sm_count_cached();
  const unsigned grid = (unsigned)(p.ntiles < sms ? p.ntiles : sms);
  igemm_conv_kernel<true><<<grid, IG_THREADS, IG_SMEM_BYTES, as_stream(stream)>>>(tmA, tmB, p);
  LL_LAUNCH_OK("igemm_conv_kernel<tf32>");
  return LL_OK;
}

// Fused conv (taps 1 | 9) + GDN / inverse GDN: a_nhwc (B,H,W,2C) [hi|lo] -> sz (B,H,W,2N) [hi|lo] of
// y * rsqrt(beta + gamma . y^2) (inverse: * sqrt), y = conv(a) + bias.  wp: conv weights from ll_pack_tf32_weight
// (Npad == N, Kpad == C); gp: gamma (N,N,1,1) packed the same way (Npad == Kpad == N); beta: N reparametrised values.
static int launch_gdn_pair(const float* a_nhwc, const float* wp, const float* bias, const float* gp, const float* beta, int B, int H,
                           int W, int C, int N, int taps, int inverse, float* sz, const float* x_head, const float* w0_head, int iC,
                           cudaStream_t stream) {
  const bool head = x_head != nullptr;
  if (B < 0 || H < 0 || W < 0) return fail(LL_EINVAL, "ll_igemm_tf32_gdn: negative extent");
  if (!head && (C < 32 || C % 32 || C > 256 || (taps != 1 && taps != 9)))
    return fail(LL_EINVAL, "ll_igemm_tf32_gdn: bad extents (C a multiple of 32 up to 256, taps 1|9)");
  if (head && (iC < 1 || iC > 3)) return fail(LL_EINVAL, "ll_conv3_gdn_head: 1..3 input channels, got %d", iC);
  if (N != 32 && N != 64 && N != 96 && N != 192)
    return fail(LL_EINVAL, "ll_igemm_tf32_gdn: N must be 32, 64, 96 or 192 (tensor-memory plan), got %d", N);
  if ((long long)B * H * W == 0) return LL_OK;
  if ((!head && (!a_nhwc || !wp)) || (head && !w0_head) || !gp || !beta || !sz) return fail(LL_EINVAL, "ll_igemm_tf32_gdn: null pointer");
  if (((uintptr_t)a_nhwc & 15) || ((uintptr_t)wp & 15) || ((uintptr_t)gp & 15) || ((uintptr_t)sz & 31))
    return fail(LL_EINVAL, "ll_igemm_tf32_gdn: operands must be 16-byte aligned, the output 32-byte aligned");
  if (head) { C = 32; taps = 0; }
  EncodeTiledFn enc = encode_fn();
  if (!enc) return fail(LL_ECUDA, "ll_igemm_tf32_gdn: cuTensorMapEncodeTiled not available from the driver");
  GdnPairParams p = {};
  p.NP = N == 192 ? 96 : N;
  p.KS = N == 192 ? 64 : N;
  p.passes = N / p.NP;
  p.rpp = N / p.KS;
  p.col_nm = N; p.col_ns = N + p.NP; p.col_st = N + 2 * p.NP;
  if (p.col_st + 2 * p.KS > IG_TMEM_COLS) return fail(LL_EINVAL, "ll_igemm_tf32_gdn: tensor-memory plan does not fit");
  // chunk ownership (16 columns, owner = index mod 4) must not move between staging rounds; <= 6 slots, <= 2 chunks per group and pass
  if ((p.rpp > 1 && (p.KS / 16) % 4) || p.KS / 16 > 6 || p.NP / 16 > 8)
    return fail(LL_EINVAL, "ll_igemm_tf32_gdn: staging plan (KS %d, NP %d) does not fit the epilogue's chunk ownership", p.KS, p.NP);
  CUtensorMap tmA, tmB, tmG;
  const int Ca = 2 * C;
  if (!head) {
    cuuint64_t gdim[4] = {(cuuint64_t)Ca, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t gstr[3] = {(cuuint64_t)Ca * 4, (cuuint64_t)W * Ca * 4, (cuuint64_t)H * W * Ca * 4};
    cuuint32_t box[4] = {32, IG_TW, IG_TH, 1};
    cuuint32_t est[4] = {1, 1, 1, 1};
    CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(a_nhwc), gdim, gstr, box, est,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(LL_ECUDA, "ll_igemm_tf32_gdn: cuTensorMapEncodeTiled(A) failed with %d", (int)r);
  }
  if (!head) {
    cuuint64_t gdim[3] = {(cuuint64_t)Ca, (cuuint64_t)N, (cuuint64_t)taps};
    cuuint64_t gstr[2] = {(cuuint64_t)Ca * 4, (cuuint64_t)N * Ca * 4};
    cuuint32_t box[3] = {32, (cuuint32_t)(N / 2), 1};
    cuuint32_t est[3] = {1, 1, 1};
    CUresult r = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(wp), gdim, gstr, box, est,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(LL_ECUDA, "ll_igemm_tf32_gdn: cuTensorMapEncodeTiled(B) failed with %d", (int)r);
  }
  {
    cuuint64_t gdim[3] = {(cuuint64_t)(2 * N), (cuuint64_t)N, 1};
    cuuint64_t gstr[2] = {(cuuint64_t)(2 * N) * 4, (cuuint64_t)N * (2 * N) * 4};
    cuuint32_t box[3] = {32, (cuuint32_t)(p.NP / 2), 1};
    cuuint32_t est[3] = {1, 1, 1};
    CUresult r = enc(&tmG, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(gp), gdim, gstr, box, est,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(LL_ECUDA, "ll_igemm_tf32_gdn: cuTensorMapEncodeTiled(G) failed with %d", (int)r);
  }
  if (head) { tmA = tmG; tmB = tmG; }      // never dereferenced in head mode
  p.head = head ? 1 : 0; p.iC = iC; p.x = x_head; p.w0 = w0_head;
#ifdef LL_TIMELINE
  p.tl = g_timeline;
#endif
  p.bias = bias; p.beta = beta; p.sz = sz;
  p.B = B; p.H = H; p.W = W; p.C = C; p.N = N; p.taps = taps; p.inverse = inverse;
  p.tiles_x = (W + IG_TW - 1) / IG_TW;
  p.tiles_y = (H + IG_TH - 1) / IG_TH;
  p.ntiles = (long long)B * p.tiles_x * p.tiles_y;
  p.npairs = (p.ntiles + 1) / 2;
  p.kb = head ? 0 : C / 32;
  p.bhalf_bytes = (N / 2) * 128;
  p.ghalf_bytes = (p.NP / 2) * 128;
  p.stage_bytes = head ? 2 * p.ghalf_bytes : 2 * PR_A_BYTES + 2 * p.bhalf_bytes;
  p.wstage = head ? 1 : 0;                   // head: [hi | lo] of this CTA's half of w0 = 2 bhalf_bytes = one gamma stage
  const int tail = PR_TAIL_BYTES + IG_MAXN * 4;
  p.stages = (PR_SMEM_LIMIT - 1024 - tail) / p.stage_bytes - p.wstage;
  if (p.stages > PR_MAXST) p.stages = PR_MAXST;
  if (p.stages < 2) return fail(LL_EINVAL, "ll_igemm_tf32_gdn: stage of %d bytes leaves fewer than 2 pipeline stages", p.stage_bytes);
  const int smem = 1024 + (p.stages + p.wstage) * p.stage_bytes + tail;
  static thread_local bool attr[64] = {false};
  int dev = 0;
  LL_CUDA_OK(cudaGetDevice(&dev));
  if (dev < 64 && !attr[dev]) {
    LL_CUDA_OK(cudaFuncSetAttribute(igemm_tf32_gdn_pair_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, PR_SMEM_LIMIT));
    LL_CUDA_OK(cudaFuncSetAttribute(igemm_tf32_gdn_pair_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, PR_SMEM_LIMIT));
    attr[dev] = true;
  }
  long long pairs = sm_count_cached() / 2;
  if (pairs > p.npairs) pairs = p.npairs;
  if (pairs < 1) pairs = 1;
  if (head) igemm_tf32_gdn_pair_kernel<true><<<(unsigned)(2 * pairs), GD_THREADS, smem, stream>>>(tmA, tmB, tmG, p);
  else igemm_tf32_gdn_pair_kernel<false><<<(unsigned)(2 * pairs), GD_THREADS, smem, stream>>>(tmA, tmB, tmG, p);
  LL_LAUNCH_OK("igemm_tf32_gdn_pair_kernel");
  return LL_OK;
}

int ll_igemm_tf32_gdn(const float* a_nhwc, const float* wp, const float* bias, const float* gp, const float* beta, int B, int H,
                      int W, int C, int N, int taps, int inverse, float* sz, ll_stream_t stream) {
  return launch_gdn_pair(a_nhwc, wp, bias, gp, beta, B, H, W, C, N, taps, inverse, sz, nullptr, nullptr, 0, as_stream(stream));
}

// First layer of SubbandAutoEncoderBerk fused with its GDN: x (B,iC,H,W) fp32 NCHW, w0 (N,iC,3,3) (for the decoder: the
// equivalent conv of the ConvTranspose2d), 3xTF32 conv (K = 9 iC padded to 32, operands split in the kernel) -> GDN /
// inverse GDN on the tensor cores -> sz (B,H,W,2N).
int ll_conv3_gdn_head(const float* x, const float* w0, const float* bias, const float* gp, const float* beta, int B, int iC, int H,
                      int W, int N, int inverse, float* sz, ll_stream_t stream) {
  if (!x) return fail(LL_EINVAL, "ll_conv3_gdn_head: null input");
  return launch_gdn_pair(nullptr, nullptr, bias, gp, beta, B, H, W, 32, N, 0, inverse, sz, x, w0, iC, as_stream(stream));
}

// fp32 NCHW (B,C,H,W) -> Y NHWC (B,H,W,C) raw and S NHWC (B,H,W,2C) = [hi | lo] of x^2 (mode 0) or of x (mode 1)
int ll_nchw_to_nhwc_split(const float* x, float* y, float* sz, int B, int C, int H, int W, int mode, ll_stream_t stream) {
  if (B < 0 || C < 1 || H < 0 || W < 0) return fail(LL_EINVAL, "ll_nchw_to_nhwc_split: bad extents");
  const long long total = (long long)B * H * W * C;
  if (total == 0) return LL_OK;
  if (!x || !sz) return fail(LL_EINVAL, "ll_nchw_to_nhwc_split: null pointer");
  long long blocks = (long long)B * (((long long)H * W + 31) / 32) * ((C + 31) / 32);
  const long long cap = (long long)sm_count_cached() * 16;
  if (blocks > cap) blocks = cap;
  nchw_to_nhwc_split_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(x, y, sz, B, C, H, W, mode);
  LL_LAUNCH_OK("nchw_to_nhwc_split_kernel");
  return LL_OK;
}

// Z NHWC (B,H,W,2C) [hi | lo] -> fp32 NCHW (B,C,H,W) = hi + lo
int ll_nhwc_split_to_nchw(const float* z, float* out, int B, int C, int H, int W, ll_stream_t stream) {
  if (B < 0 || C < 1 || H < 0 || W < 0) return fail(LL_EINVAL, "ll_nhwc_split_to_nchw: bad extents");
  const long long total = (long long)B * H * W * C;
  if (total == 0) return LL_OK;
  if (!z || !out) return fail(LL_EINVAL, "ll_nhwc_split_to_nchw: null pointer");
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)sm_count_cached() * 16;
  if (blocks > cap) blocks = cap;
  nhwc_split_to_nchw_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(z, out, B, C, H, W);
  LL_LAUNCH_OK("nhwc_split_to_nchw_kernel");
  return LL_OK;
}

}  // extern "C"
