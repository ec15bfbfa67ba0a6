// lift_step_body.cuh -- one learned lifting step as a fused, marching stencil.
//
//   dout = din + sign * (skip + rw * CNN(skip)),   skip = 3-tap pre-filter of src along y
//   CNN  = conv5x5(1->16) -> tanh -> conv5x5(16->16) -> tanh -> conv5x5(16->16) + o1 -> conv5x5(16->1)
//
// Reference: lifting_forward_row_2_stage_lifting (graphs/layers/wavelet_forward_v2.py:58-74),
// lifting_inverse_row_2_stage_lifting (graphs/layers/wavelet_inverse_v2.py:76-90),
// P_block_v2.forward (graphs/layers/P_block_v2.py:40-55).
//
// Design (DESIGN.md "K2"): a persistent CTA owns a contiguous range of (job, image, strip,
// row-chunk) units.  A strip is LS_WT output columns wide; the CTA marches down it LS_R rows at
// a time keeping the last rows of every intermediate layer in shared-memory ring buffers, so no
// activation is recomputed vertically and none ever touches HBM.  Each conv layer zero-pads ITS
// OWN input, so every intermediate is forced to 0 outside the plane.
// The 16->16 layers run as an 8-pixel x 4-channel register tile per thread on packed FFMA2
// (fma.rn.f32x2, sm_100): activation scalar broadcast x weight pair, fp32 exact.
//
// The body is written as "phases" separated by block barriers, with all control flow outside
// the phases uniform across the CTA.  The same source compiles for the host (tests/emul) where a
// phase is a loop over thread ids -- used only to check indexing against the oracle on CPU.
#pragma once
#include <math.h>
#include <stdint.h>

#include "../../include/ll_api.h"

#if defined(__CUDACC__)
#define LL_HD __host__ __device__ __forceinline__
#else
#define LL_HD inline
struct alignas(16) float4 {  // host emulation build only (tests/emul)
  float x, y, z, w;
};
#endif

namespace ll {

constexpr int LS_THREADS = 256;
constexpr int LS_R = 8;     // rows per marching chunk
constexpr int LS_WT = 56;   // output columns per strip
constexpr int LS_NR = 12;   // ring rows of a1/a2/a3
constexpr int LS_NRS = 16;  // ring rows of skip
constexpr int LS_P1 = 68;   // pitches, all == 4 (mod 8) so 16-byte row/group accesses spread over banks
constexpr int LS_P2 = 68;
constexpr int LS_P3 = 60;
constexpr int LS_PS = 76;
constexpr int LS_W1COLS = 68, LS_W2COLS = 64, LS_W3COLS = 60, LS_WSCOLS = 72;

// blob layout (floats) == LL_LIFT_BLOB_FLOATS
constexpr int BL_PRE = 0;            // 3 taps (+1 pad)
constexpr int BL_W1 = 4;             // [tap 25][co 16]
constexpr int BL_B1 = BL_W1 + 400;   // 16
constexpr int BL_W2 = BL_B1 + 16;    // [ci 16][tap 25][co 16]
constexpr int BL_B2 = BL_W2 + 6400;  // 16
constexpr int BL_W3 = BL_B2 + 16;    // [ci][tap][co]
constexpr int BL_B3 = BL_W3 + 6400;  // 16
constexpr int BL_W4 = BL_B3 + 16;    // [ci 16][tap 25]
constexpr int BL_B4 = BL_W4 + 400;   // 1 (+3 pad)
constexpr int BL_TOTAL = BL_B4 + 4;   // end of the part the SIMT kernel stages in shared memory
// tensor-core weight block (lift_tc.cu): [layer 2][term hi|lo][row dx*16+co (80)][col dy*16+ci (80)]
constexpr int BL_TC = BL_TOTAL;
constexpr int BL_TC16 = BL_TC + 4 * 80 * 80;   // fp16 weight block of the 3xFP16 kernel: 4 x 80 x 40 32-bit words (lift_tc.cu)
constexpr int BL_TC16_4 = BL_TC16 + 4 * 80 * 40;   // conv4 block of the 3xFP16 kernel: 2 x 16 x 40 words
constexpr int BL_ALL = BL_TC16_4 + 2 * 16 * 40;
static_assert(BL_ALL == LL_LIFT_BLOB_FLOATS, "blob layout");

// shared memory layout (floats)
constexpr int SM_A1 = 0;
constexpr int SM_A2 = SM_A1 + 16 * LS_NR * LS_P1;
constexpr int SM_A3 = SM_A2 + 16 * LS_NR * LS_P2;
constexpr int SM_SK = SM_A3 + 16 * LS_NR * LS_P3;
constexpr int SM_BLOB = SM_SK + LS_NRS * LS_PS;      // whole blob, same offsets as BL_*
constexpr int SM_RED = SM_BLOB + BL_TOTAL;           // conv4 partial sums [4][LS_R][LS_WT]
constexpr int SM_TOTAL = SM_RED + 4 * LS_R * LS_WT;
constexpr size_t LS_SMEM_BYTES = size_t(SM_TOTAL) * sizeof(float);
static_assert(LS_SMEM_BYTES <= 227 * 1024, "shared memory budget");
static_assert(SM_A2 % 4 == 0 && SM_A3 % 4 == 0 && SM_SK % 4 == 0 && SM_BLOB % 4 == 0 && SM_RED % 4 == 0, "align");

struct LiftParams {
  ll_lift_job job[2];
  int njobs;
  const float* blob;
  float sign, rw;
  int linear;
  // schedule (filled by the launcher)
  int nstrips[2], nchunks[2];
  long long units[2];  // nb * nstrips * nchunks per job
  long long total_units;
  int f16;             // tensor-core kernel: 3xFP16 operands (LL_LIFT_TC16) instead of 3xTF32
  int dbg;             // timing experiments only (lift_tc.cu): bit 0 no MMA, 1 no E-B, 2 no conv1, 3 no conv4, 4 no E-A
  long long* dbg_buf;  // optional [17 warps][8] clock64 stamps of CTA 0 at global step 40 (lift_tc.cu)
};

struct f2 {
  float x, y;
};

#if defined(__CUDA_ARCH__)
#define LL_FMA2(acc, a, w)                                                       \
  do {                                                                           \
    float2 _r = __ffma2_rn(make_float2((a), (a)), make_float2((w).x, (w).y),     \
                           make_float2((acc).x, (acc).y));                       \
    (acc).x = _r.x;                                                              \
    (acc).y = _r.y;                                                              \
  } while (0)
#define LL_TANH(x) tanhf(x)
#define LL_MUL(a, b) __fmul_rn((a), (b))
#define LL_ADD(a, b) __fadd_rn((a), (b))
#else
#define LL_FMA2(acc, a, w)                  \
  do {                                      \
    (acc).x = fmaf((a), (w).x, (acc).x);    \
    (acc).y = fmaf((a), (w).y, (acc).y);    \
  } while (0)
#define LL_TANH(x) tanhf(x)
#define LL_MUL(a, b) ((a) * (b))
#define LL_ADD(a, b) ((a) + (b))
#endif

LL_HD int ring(int r, int n) {
  int m = r % n;
  return m < 0 ? m + n : m;
}

// unit of work: which strip / chunk this CTA is on (uniform across the CTA)
struct Unit {
  int j;       // job
  int b;       // image in batch
  int x0;      // first output column of the strip
  int y0;      // first output row of the chunk
  int ny, nx;
};

LL_HD Unit decode_unit(const LiftParams& p, long long u) {
  Unit t;
  int j = 0;
  if (u >= p.units[0]) {
    u -= p.units[0];
    j = 1;
  }
  const int nch = p.nchunks[j], nst = p.nstrips[j];
  const int c = int(u % nch);
  const long long v = u / nch;
  const int s = int(v % nst);
  t.j = j;
  t.b = int(v / nst);
  t.x0 = s * LS_WT;
  t.y0 = c * LS_R;
  t.ny = p.job[j].ny;
  t.nx = p.job[j].nx;
  return t;
}

// ---- phase: skip rows [ra, rb) -----------------------------------------------------------
LL_HD void phase_skip(const LiftParams& p, const Unit& t, float* sm, int ra, int rb, int tid) {
  const ll_view3& S = p.job[t.j].src;
  const float* base = S.ptr + (long long)t.b * S.sb;
  const float w0 = sm[SM_BLOB + BL_PRE + 0], w1 = sm[SM_BLOB + BL_PRE + 1], w2 = sm[SM_BLOB + BL_PRE + 2];
  const int n = (rb - ra) * LS_WSCOLS;
  for (int e = tid; e < n; e += LS_THREADS) {
    const int rr = ra + e / LS_WSCOLS;
    const int j = e % LS_WSCOLS;
    const int c = t.x0 - 8 + j;
    float v = 0.f;
    if (rr >= 0 && rr < t.ny && c >= 0 && c < t.nx) {
      const float* q = base + (long long)rr * S.sy + (long long)c * S.sx;
      const float s0 = rr > 0 ? q[-S.sy] : 0.f;
      const float s1 = q[0];
      const float s2 = rr + 1 < t.ny ? q[S.sy] : 0.f;
      v = fmaf(w2, s2, fmaf(w1, s1, LL_MUL(w0, s0)));
    }
    sm[SM_SK + ring(rr, LS_NRS) * LS_PS + j] = v;
  }
}

// ---- phase: conv1 (1 -> 16) + tanh, rows [ra, rb) of a1 ----------------------------------
LL_HD void phase_conv1(const LiftParams& p, const Unit& t, float* sm, int ra, int rb, int tid) {
  constexpr int NG = 9;  // 8-pixel groups covering 68 (72) columns
  const int n = (rb - ra) * NG * 8;
  for (int it = tid; it < n; it += LS_THREADS) {
    const int cp = it & 7;  // channel pair
    const int g = (it >> 3) % NG;
    const int r = ra + (it >> 3) / NG;
    const int cbase = t.x0 - 6 + 8 * g;  // plane column of pixel 0
    f2 acc[8];
    const bool live = (r >= 0 && r < t.ny && cbase + 8 > 0 && cbase < t.nx);
    if (live) {
      const f2 b = *reinterpret_cast<const f2*>(&sm[SM_BLOB + BL_B1 + 2 * cp]);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = b;
#pragma unroll
      for (int dy = 0; dy < 5; ++dy) {
        const float* arow = &sm[SM_SK + ring(r + dy - 2, LS_NRS) * LS_PS + 8 * g];
        float a[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) a[k] = arow[k];
#pragma unroll
        for (int dx = 0; dx < 5; ++dx) {
          const f2 w = *reinterpret_cast<const f2*>(&sm[SM_BLOB + BL_W1 + (dy * 5 + dx) * 16 + 2 * cp]);
#pragma unroll
          for (int k = 0; k < 8; ++k) LL_FMA2(acc[k], a[k + dx], w);
        }
      }
    }
    float* o0 = &sm[SM_A1 + (2 * cp) * (LS_NR * LS_P1) + ring(r, LS_NR) * LS_P1 + 8 * g];
    float* o1 = o0 + LS_NR * LS_P1;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (8 * g + k < LS_W1COLS) {
        const int c = cbase + k;
        float vx = 0.f, vy = 0.f;
        if (live && c >= 0 && c < t.nx) {
          vx = p.linear ? acc[k].x : LL_TANH(acc[k].x);
          vy = p.linear ? acc[k].y : LL_TANH(acc[k].y);
        }
        o0[k] = vx;
        o1[k] = vy;
      }
    }
  }
}

// ---- 8 px x 4 ch register tile over NCI input channels ------------------------------------
// in: ring buffer [ci][NRING][PITCH]; rowoff[dy] = ring slot offset of input row r+dy-2;
// xo: first input column (buffer index) of pixel 0; W: [ci][25][16]; q: channel quarter.
template <int PITCH, int NRING>
LL_HD void conv16_tile(const float* in, const int* rowoff, int xo, const float* W, int q, f2 (&acc)[8][2]) {
#pragma unroll 1
  for (int ci = 0; ci < 16; ++ci) {
    const float* ic = in + ci * (NRING * PITCH) + xo;
    const float* wc = W + ci * 400 + 4 * q;
#pragma unroll
    for (int dy = 0; dy < 5; ++dy) {
      const float* arow = ic + rowoff[dy];
      float a[12];
#pragma unroll
      for (int k = 0; k < 12; k += 4) {
        const float4 v = *reinterpret_cast<const float4*>(arow + k);
        a[k] = v.x;
        a[k + 1] = v.y;
        a[k + 2] = v.z;
        a[k + 3] = v.w;
      }
#pragma unroll
      for (int dx = 0; dx < 5; ++dx) {
        const float4 w = *reinterpret_cast<const float4*>(wc + (dy * 5 + dx) * 16);
        const f2 w01 = {w.x, w.y}, w23 = {w.z, w.w};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          LL_FMA2(acc[k][0], a[k + dx], w01);
          LL_FMA2(acc[k][1], a[k + dx], w23);
        }
      }
    }
  }
}

// lane -> (pixel group, row within a block of 4): a quarter-warp covers 4 groups x 2 rows so
// that its 16-byte accesses land on 8 distinct bank groups (pitch == 4 mod 8).
LL_HD void lane_map(int lane, int& g, int& rl) {
  g = (lane & 3) | (((lane >> 3) & 1) << 2);
  rl = ((lane >> 2) & 1) | (((lane >> 4) & 1) << 1);
}

// ---- phase: conv2 (16 -> 16) + tanh, rows [ra, rb) of a2 ----------------------------------
LL_HD void phase_conv2(const LiftParams& p, const Unit& t, float* sm, int ra, int rb, int tid) {
  const int warp = tid >> 5, lane = tid & 31;
  int g, rl;
  lane_map(lane, g, rl);
  const int npairs = ((rb - ra + 3) >> 2) * 4;
  for (int pr = warp; pr < npairs; pr += LS_THREADS / 32) {
    const int q = pr & 3;
    const int r = ra + (pr >> 2) * 4 + rl;
    if (r >= rb) continue;
    const int cbase = t.x0 - 4 + 8 * g;
    const bool live = (r >= 0 && r < t.ny && cbase + 8 > 0 && cbase < t.nx);
    f2 acc[8][2];
    if (live) {
      const float4 b = *reinterpret_cast<const float4*>(&sm[SM_BLOB + BL_B2 + 4 * q]);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        acc[k][0] = f2{b.x, b.y};
        acc[k][1] = f2{b.z, b.w};
      }
      int rowoff[5];
#pragma unroll
      for (int dy = 0; dy < 5; ++dy) rowoff[dy] = ring(r + dy - 2, LS_NR) * LS_P1;
      conv16_tile<LS_P1, LS_NR>(&sm[SM_A1], rowoff, 8 * g, &sm[SM_BLOB + BL_W2], q, acc);
    }
    float* o = &sm[SM_A2 + (4 * q) * (LS_NR * LS_P2) + ring(r, LS_NR) * LS_P2 + 8 * g];
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      float v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int c = cbase + k;
        float x = 0.f;
        if (live && c >= 0 && c < t.nx) {
          x = (ch & 1) ? acc[k][ch >> 1].y : acc[k][ch >> 1].x;
          if (!p.linear) x = LL_TANH(x);
        }
        v[k] = x;
      }
      float* oc = o + ch * (LS_NR * LS_P2);
      *reinterpret_cast<float4*>(oc) = float4{v[0], v[1], v[2], v[3]};
      *reinterpret_cast<float4*>(oc + 4) = float4{v[4], v[5], v[6], v[7]};
    }
  }
}

// ---- phase: conv3 (16 -> 16) + o1 residual, rows [ra, rb) of a3 ---------------------------
LL_HD void phase_conv3(const LiftParams& p, const Unit& t, float* sm, int ra, int rb, int tid) {
  const int warp = tid >> 5, lane = tid & 31;
  int g, rl;
  lane_map(lane, g, rl);
  const int npairs = ((rb - ra + 3) >> 2) * 4;
  for (int pr = warp; pr < npairs; pr += LS_THREADS / 32) {
    const int q = pr & 3;
    const int r = ra + (pr >> 2) * 4 + rl;
    if (r >= rb) continue;
    const int cbase = t.x0 - 2 + 8 * g;
    const bool live = (r >= 0 && r < t.ny && cbase + 8 > 0 && cbase < t.nx);
    f2 acc[8][2];
    if (live) {
      const float4 b = *reinterpret_cast<const float4*>(&sm[SM_BLOB + BL_B3 + 4 * q]);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        acc[k][0] = f2{b.x, b.y};
        acc[k][1] = f2{b.z, b.w};
      }
      int rowoff[5];
#pragma unroll
      for (int dy = 0; dy < 5; ++dy) rowoff[dy] = ring(r + dy - 2, LS_NR) * LS_P2;
      conv16_tile<LS_P2, LS_NR>(&sm[SM_A2], rowoff, 8 * g, &sm[SM_BLOB + BL_W3], q, acc);
      // residual o1 = conv1(skip) + b1 (the PRE-tanh conv1 output, P_block_v2.py:41,53), recomputed
      // from the skip ring and accumulated on top of conv3: a3 = ((b3 + conv3) + conv1) + b1
#pragma unroll
      for (int dy = 0; dy < 5; ++dy) {
        const float* arow = &sm[SM_SK + ring(r + dy - 2, LS_NRS) * LS_PS + 8 * g + 4];
        float a[12];
#pragma unroll
        for (int k = 0; k < 12; k += 4) {
          const float4 v = *reinterpret_cast<const float4*>(arow + k);
          a[k] = v.x;
          a[k + 1] = v.y;
          a[k + 2] = v.z;
          a[k + 3] = v.w;
        }
#pragma unroll
        for (int dx = 0; dx < 5; ++dx) {
          const float4 w = *reinterpret_cast<const float4*>(&sm[SM_BLOB + BL_W1 + (dy * 5 + dx) * 16 + 4 * q]);
          const f2 w01 = {w.x, w.y}, w23 = {w.z, w.w};
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            LL_FMA2(acc[k][0], a[k + dx], w01);
            LL_FMA2(acc[k][1], a[k + dx], w23);
          }
        }
      }
      const float4 b1 = *reinterpret_cast<const float4*>(&sm[SM_BLOB + BL_B1 + 4 * q]);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        acc[k][0].x = LL_ADD(acc[k][0].x, b1.x);
        acc[k][0].y = LL_ADD(acc[k][0].y, b1.y);
        acc[k][1].x = LL_ADD(acc[k][1].x, b1.z);
        acc[k][1].y = LL_ADD(acc[k][1].y, b1.w);
      }
    }
    float* o = &sm[SM_A3 + (4 * q) * (LS_NR * LS_P3) + ring(r, LS_NR) * LS_P3 + 8 * g];
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      float v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int c = cbase + k;
        float x = 0.f;
        if (live && c >= 0 && c < t.nx) x = (ch & 1) ? acc[k][ch >> 1].y : acc[k][ch >> 1].x;
        v[k] = x;
      }
      float* oc = o + ch * (LS_NR * LS_P3);
      *reinterpret_cast<float4*>(oc) = float4{v[0], v[1], v[2], v[3]};
      if (8 * g + 4 < LS_W3COLS) *reinterpret_cast<float4*>(oc + 4) = float4{v[4], v[5], v[6], v[7]};
    }
  }
}

// ---- phase: conv4 (16 -> 1) partial sums over 4 input-channel quarters --------------------
LL_HD void phase_conv4(const LiftParams& p, const Unit& t, float* sm, int tid) {
  (void)p;
  constexpr int NG = LS_WT / 8;  // 7
  const int n = LS_R * NG * 4;
  for (int it = tid; it < n; it += LS_THREADS) {
    const int cq = it & 3;
    const int g = (it >> 2) % NG;
    const int rl = (it >> 2) / NG;
    const int r = t.y0 + rl;
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    if (r < t.ny && t.x0 + 8 * g < t.nx) {
#pragma unroll 1
      for (int ci = 4 * cq; ci < 4 * cq + 4; ++ci) {
#pragma unroll
        for (int dy = 0; dy < 5; ++dy) {
          const float* arow = &sm[SM_A3 + ci * (LS_NR * LS_P3) + ring(r + dy - 2, LS_NR) * LS_P3 + 8 * g];
          float a[12];
#pragma unroll
          for (int k = 0; k < 12; k += 4) {
            const float4 v = *reinterpret_cast<const float4*>(arow + k);
            a[k] = v.x;
            a[k + 1] = v.y;
            a[k + 2] = v.z;
            a[k + 3] = v.w;
          }
#pragma unroll
          for (int dx = 0; dx < 5; ++dx) {
            const float w = sm[SM_BLOB + BL_W4 + ci * 25 + dy * 5 + dx];
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[k] = fmaf(a[k + dx], w, acc[k]);
          }
        }
      }
    }
    float* o = &sm[SM_RED + (cq * LS_R + rl) * LS_WT + 8 * g];
    *reinterpret_cast<float4*>(o) = float4{acc[0], acc[1], acc[2], acc[3]};
    *reinterpret_cast<float4*>(o + 4) = float4{acc[4], acc[5], acc[6], acc[7]};
  }
}

// ---- phase: combine and write dout ---------------------------------------------------------
LL_HD void phase_out(const LiftParams& p, const Unit& t, float* sm, int tid) {
  const ll_lift_job& J = p.job[t.j];
  const float b4 = sm[SM_BLOB + BL_B4];
  const int n = LS_R * LS_WT;
  for (int e = tid; e < n; e += LS_THREADS) {
    const int rl = e / LS_WT, j = e % LS_WT;
    const int r = t.y0 + rl, c = t.x0 + j;
    if (r < t.ny && c < t.nx) {
      const float* red = &sm[SM_RED + rl * LS_WT + j];
      const float s01 = LL_ADD(red[0], red[1 * LS_R * LS_WT]);
      const float s23 = LL_ADD(red[2 * LS_R * LS_WT], red[3 * LS_R * LS_WT]);
      const float net = LL_ADD(LL_ADD(s01, s23), b4);
      const float sk = sm[SM_SK + ring(r, LS_NRS) * LS_PS + j + 8];
      const float d = J.din.ptr[(long long)t.b * J.din.sb + (long long)r * J.din.sy + (long long)c * J.din.sx];
      const float tn = LL_MUL(net, p.rw);
      // forward: (dst + skip) + net*w ; inverse: (dst - skip) - net*w  -- the reference's order
      // sign == 0: raw CNN output (stand-alone P_block_v2.forward)
      const float o = p.sign > 0.f ? LL_ADD(LL_ADD(d, sk), tn) : p.sign < 0.f ? LL_ADD(LL_ADD(d, -sk), -tn) : net;
      J.dout.ptr[(long long)t.b * J.dout.sb + (long long)r * J.dout.sy + (long long)c * J.dout.sx] = o;
    }
  }
}

// torch-layout P_block_v2 parameters -> element i of the blob (see BL_*)
LL_HD float pack_lift_elem(int i, const float* pre, const float* w1, const float* b1, const float* w2,
                           const float* b2, const float* w3, const float* b3, const float* w4, const float* b4) {
  if (i < BL_W1) return i < 3 ? pre[i] : 0.f;
  if (i < BL_B1) {
    const int k = i - BL_W1, tap = k / 16, co = k % 16;
    return w1[co * 25 + tap];
  }
  if (i < BL_W2) return b1[i - BL_B1];
  if (i < BL_B2) {
    const int k = i - BL_W2, ci = k / 400, tap = (k % 400) / 16, co = k % 16;
    return w2[(co * 16 + ci) * 25 + tap];
  }
  if (i < BL_W3) return b2[i - BL_B2];
  if (i < BL_B3) {
    const int k = i - BL_W3, ci = k / 400, tap = (k % 400) / 16, co = k % 16;
    return w3[(co * 16 + ci) * 25 + tap];
  }
  if (i < BL_W4) return b3[i - BL_B3];
  if (i < BL_B4) return w4[i - BL_W4];  // (1,16,5,5) -> [ci][tap]
  if (i == BL_B4) return b4[0];
  return 0.f;
}

LL_HD void phase_load_blob(const LiftParams& p, float* sm, int tid) {
  for (int i = tid; i < BL_TOTAL; i += LS_THREADS) sm[SM_BLOB + i] = p.blob[i];
}

// Drives the phases for one CTA.  PHASE(call) runs `call` for every thread and then acts as a
// block barrier (device: call; __syncthreads();  host: for tid ... call).
#define LL_LIFT_STEP_DRIVER(PHASE, p, sm, cta, ncta)                                        \
  do {                                                                                      \
    PHASE(ll::phase_load_blob(p, sm, tid));                                                 \
    const long long _lo = (p).total_units * (long long)(cta) / (ncta);                      \
    const long long _hi = (p).total_units * (long long)((cta) + 1) / (ncta);                \
    for (long long _u = _lo; _u < _hi; ++_u) {                                              \
      const ll::Unit _t = ll::decode_unit(p, _u);                                           \
      const int _y = _t.y0;                                                                 \
      /* stage 0 = prologue after a fresh start (fills the rings), stage 1 = steady chunk */   \
      for (int _st = (_u == _lo || _y == 0) ? 0 : 1; _st < 2; ++_st) {                          \
        const int _o = _st ? 8 : -8, _n = _st ? 8 : 16;                                         \
        PHASE(ll::phase_skip(p, _t, sm, _y + _o, _y + _o + _n, tid));                           \
        PHASE(ll::phase_conv1(p, _t, sm, _y + _o + 2 - 4 * _st, _y + 6 + 8 * _st, tid));        \
        PHASE(ll::phase_conv2(p, _t, sm, _y + _o + 4 - 8 * _st, _y + 4 + 8 * _st, tid));        \
        PHASE(ll::phase_conv3(p, _t, sm, _y + _o + 6 - 12 * _st, _y + 2 + 8 * _st, tid));       \
      }                                                                                         \
      PHASE(ll::phase_conv4(p, _t, sm, tid));                                               \
      PHASE(ll::phase_out(p, _t, sm, tid));                                                 \
    }                                                                                       \
  } while (0)

}  // namespace ll
