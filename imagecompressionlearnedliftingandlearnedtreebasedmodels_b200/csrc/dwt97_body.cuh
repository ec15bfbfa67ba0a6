// dwt97_body.cuh -- CDF 9/7 ("bior4.4") periodised filter bank, one 2-D level per launch.
//
// Reference: DWTPytorchWaveletsLayer (graphs/layers/lifting_dwt_nets.py:228-231,250,274) ->
// pytorch_wavelets.DWTForward/DWTInverse(mode='periodization', wave='bior4.4'); taps pinned
// in-repo by get_cdf97_filters (lifting_dwt_nets.py:415-418).  SURVEY.md appendix A.2:
//   analysis : lo[n] = sum_k dec_lo[k] x[(2n+5-k) mod N], hi likewise; width axis, then height
//   synthesis: x[m]  = sum_{n,k:(2n+k-4) mod N = m} rec_lo[k] lo[n] + rec_hi[k] hi[n]; height, then width
//
// HBM-bound (DESIGN.md "K1"): a CTA stages an input tile + 4-sample halo in shared memory with
// 16-byte loads, runs the row pass into shared memory, the column pass from shared memory, and
// writes the four subbands once.  Algorithmic traffic 8 B per input pixel per level.
// Written as barrier-separated phases so tests/emul can run the same code on the host.
#pragma once
#include <stdint.h>

#include "lift_step_body.cuh"  // LL_HD, host float4

namespace ll {

constexpr int DW_THREADS = 256;
constexpr int DW_TY = 16, DW_TX = 64;        // subband-domain tile
constexpr int DWF_R = 2 * DW_TY + 8;         // input rows staged by the forward kernel
constexpr int DWF_C = 2 * DW_TX + 8;         // input columns staged (multiple of 4)
constexpr int DWF_SM_IN = 0;
constexpr int DWF_SM_LO = DWF_SM_IN + DWF_R * DWF_C;
constexpr int DWF_SM_HI = DWF_SM_LO + DWF_R * DW_TX;
constexpr int DWF_SM_TOTAL = DWF_SM_HI + DWF_R * DW_TX;
constexpr int DWI_P = DW_TX + 4;             // subband tile pitch of the inverse kernel
constexpr int DWI_R = DW_TY + 4;
constexpr int DWI_SM_SB = 0;                 // [4][DWI_R][DWI_P]: ll, lh, hl, hh
constexpr int DWI_SM_LO = DWI_SM_SB + 4 * DWI_R * DWI_P;   // [2 TY][DWI_P]
constexpr int DWI_SM_HI = DWI_SM_LO + 2 * DW_TY * DWI_P;
constexpr int DWI_SM_TOTAL = DWI_SM_HI + 2 * DW_TY * DWI_P;

// fp32 taps (the reference holds the filters in fp32)
#define LL_DEC_LO(k) ((k) == 1 || (k) == 9 ? 0.037828455507264f : (k) == 2 || (k) == 8 ? -0.023849465019557f : \
                      (k) == 3 || (k) == 7 ? -0.110624404418437f : (k) == 4 || (k) == 6 ? 0.377402855612831f : \
                      (k) == 5 ? 0.852698679008894f : 0.f)
#define LL_DEC_HI(k) ((k) == 1 || (k) == 7 ? -0.064538882628697f : (k) == 2 || (k) == 6 ? 0.040689417609164f : \
                      (k) == 3 || (k) == 5 ? 0.418092273221617f : (k) == 4 ? -0.788485616405583f : 0.f)
#define LL_REC_LO(k) ((k) == 1 || (k) == 7 ? -0.064538882628697f : (k) == 2 || (k) == 6 ? -0.040689417609164f : \
                      (k) == 3 || (k) == 5 ? 0.418092273221617f : (k) == 4 ? 0.788485616405583f : 0.f)
#define LL_REC_HI(k) ((k) == 1 || (k) == 9 ? -0.037828455507264f : (k) == 2 || (k) == 8 ? -0.023849465019557f : \
                      (k) == 3 || (k) == 7 ? 0.110624404418437f : (k) == 4 || (k) == 6 ? 0.377402855612831f : \
                      (k) == 5 ? -0.852698679008894f : 0.f)

struct DwtParams {
  const float* x;   // forward input / inverse output (N, h, w)
  float* xo;
  long long x_sn;
  const float* ll;  // (N, h/2, w/2)
  float* llo;
  long long ll_sn;
  const float* yh;  // (N, 3, h/2, w/2)
  float* yho;
  long long yh_sn;
  int N, h, w;
  int tiles_x, tiles_y;
};

struct DwtTile {
  int n, y0, x0;  // plane, first subband row / column of the tile
};

LL_HD DwtTile dwt_tile(const DwtParams& p, long long t) {
  DwtTile d;
  d.x0 = int(t % p.tiles_x) * DW_TX;
  t /= p.tiles_x;
  d.y0 = int(t % p.tiles_y) * DW_TY;
  d.n = int(t / p.tiles_y);
  return d;
}

LL_HD int wrapi(int a, int n) {
  int m = a % n;
  return m < 0 ? m + n : m;
}

// ---------------- forward ----------------
LL_HD void dwtf_load(const DwtParams& p, const DwtTile& t, float* sm, int tid) {
  const float* base = p.x + (long long)t.n * p.x_sn;
  constexpr int C4 = DWF_C / 4;
  const bool vec = (p.w % 4 == 0) && ((reinterpret_cast<uintptr_t>(base) & 15) == 0);
  for (int e = tid; e < DWF_R * C4; e += DW_THREADS) {
    const int rr = e / C4, c4 = e % C4;
    const int ur = 2 * t.y0 - 4 + rr;  // unwrapped row / column
    const int gr = wrapi(ur, p.h);
    const int gc = 2 * t.x0 - 4 + 4 * c4;
    const float* row = base + (long long)gr * p.w;
    float4 v;
    if (vec) {
      v = *reinterpret_cast<const float4*>(row + wrapi(gc, p.w));
    } else {
      v.x = row[wrapi(gc, p.w)];
      v.y = row[wrapi(gc + 1, p.w)];
      v.z = row[wrapi(gc + 2, p.w)];
      v.w = row[wrapi(gc + 3, p.w)];
    }
    // pytorch_wavelets folds the overhang back ONCE (lowlevel.afb1d, mode 'per'): for extents
    // shorter than the filter (N < 10) taps reaching below 5 - N are dropped, not wrapped twice.
    if (ur < 5 - p.h) v = float4{0.f, 0.f, 0.f, 0.f};
    if (gc < 5 - p.w) v.x = 0.f;
    if (gc + 1 < 5 - p.w) v.y = 0.f;
    if (gc + 2 < 5 - p.w) v.z = 0.f;
    if (gc + 3 < 5 - p.w) v.w = 0.f;
    *reinterpret_cast<float4*>(&sm[DWF_SM_IN + rr * DWF_C + 4 * c4]) = v;
  }
}

LL_HD void dwtf_rows(float* sm, int tid) {
  for (int e = tid; e < DWF_R * DW_TX; e += DW_THREADS) {
    const int rr = e / DW_TX, nl = e % DW_TX;
    const float* q = &sm[DWF_SM_IN + rr * DWF_C + 2 * nl];
    float v[10];
#pragma unroll
    for (int i = 0; i < 10; i += 2) {
      v[i] = q[i];
      v[i + 1] = q[i + 1];
    }
    // lo[n] = sum_k dec_lo[k] * in[2 nl + 9 - k]
    float lo = LL_DEC_LO(1) * v[8];
    float hi = LL_DEC_HI(1) * v[8];
#pragma unroll
    for (int k = 2; k <= 9; ++k) lo = fmaf(LL_DEC_LO(k), v[9 - k], lo);
#pragma unroll
    for (int k = 2; k <= 7; ++k) hi = fmaf(LL_DEC_HI(k), v[9 - k], hi);
    sm[DWF_SM_LO + rr * DW_TX + nl] = lo;
    sm[DWF_SM_HI + rr * DW_TX + nl] = hi;
  }
}

LL_HD void dwtf_cols(const DwtParams& p, const DwtTile& t, const float* sm, int tid) {
  const int h2 = p.h / 2, w2 = p.w / 2;
  const long long sub = (long long)h2 * w2;
  for (int e = tid; e < DW_TY * DW_TX; e += DW_THREADS) {
    const int ml = e / DW_TX, nl = e % DW_TX;
    const int gy = t.y0 + ml, gx = t.x0 + nl;
    if (gy >= h2 || gx >= w2) continue;
    const float* lo = &sm[DWF_SM_LO + (2 * ml) * DW_TX + nl];
    const float* hi = &sm[DWF_SM_HI + (2 * ml) * DW_TX + nl];
    float a[9], b[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      a[i] = lo[i * DW_TX];
      b[i] = hi[i * DW_TX];
    }
    float LLv = LL_DEC_LO(1) * a[8], LHv = LL_DEC_HI(1) * a[8];
    float HLv = LL_DEC_LO(1) * b[8], HHv = LL_DEC_HI(1) * b[8];
#pragma unroll
    for (int k = 2; k <= 9; ++k) {
      LLv = fmaf(LL_DEC_LO(k), a[9 - k], LLv);
      HLv = fmaf(LL_DEC_LO(k), b[9 - k], HLv);
    }
#pragma unroll
    for (int k = 2; k <= 7; ++k) {
      LHv = fmaf(LL_DEC_HI(k), a[9 - k], LHv);
      HHv = fmaf(LL_DEC_HI(k), b[9 - k], HHv);
    }
    const long long o = (long long)gy * w2 + gx;
    p.llo[(long long)t.n * p.ll_sn + o] = LLv;
    float* yh = p.yho + (long long)t.n * p.yh_sn + o;
    yh[0] = LHv;        // high along H, low along W
    yh[sub] = HLv;      // low along H, high along W
    yh[2 * sub] = HHv;
  }
}

// ---------------- inverse ----------------
LL_HD void dwti_load(const DwtParams& p, const DwtTile& t, float* sm, int tid) {
  const int h2 = p.h / 2, w2 = p.w / 2;
  const long long sub = (long long)h2 * w2;
  for (int e = tid; e < 4 * DWI_R * DWI_P; e += DW_THREADS) {
    const int s = e / (DWI_R * DWI_P);
    const int i = (e / DWI_P) % DWI_R, c = e % DWI_P;
    const int gy = wrapi(t.y0 - 2 + i, h2), gx = wrapi(t.x0 - 2 + c, w2);
    const long long o = (long long)gy * w2 + gx;
    const float v = (s == 0) ? p.ll[(long long)t.n * p.ll_sn + o]
                             : p.yh[(long long)t.n * p.yh_sn + (s == 1 ? 0 : s == 2 ? sub : 2 * sub) + o];
    sm[DWI_SM_SB + e] = v;
  }
}

// 5-tap polyphase synthesis of one (even, odd) output pair from lo[i..i+4], hi[i..i+4] (stride st)
// nu0 = unwrapped subband index of lo[0]; n2 = subband extent.  pytorch_wavelets (lowlevel.sfb1d,
// mode 'per') folds the transposed-conv overhang back once, so for n2 < 4 a (sample, tap) pair
// whose position 2n+k reaches 2N is dropped instead of wrapped a second time.
LL_HD void syn_pair(const float* lo, const float* hi, int st, int nu0, int n2, float& even, float& odd) {
  // x[2j]   = sum_t rec_lo[2t]   lo[j+2-t] + rec_hi[2t]   hi[j+2-t]
  // x[2j+1] = sum_t rec_lo[2t+1] lo[j+2-t] + rec_hi[2t+1] hi[j+2-t],  t = 0..4, index i = 4 - t
  float el = 0.f, eh = 0.f, ol = 0.f, oh = 0.f;
#pragma unroll
  for (int t = 0; t < 5; ++t) {
    float l = lo[(4 - t) * st], h = hi[(4 - t) * st];
    if (n2 < 4) {
      const int nu = nu0 + 4 - t;
      const int q = (nu >= 0) ? nu / n2 : -((-nu + n2 - 1) / n2);
      const int pos = 2 * nu + 2 * t - 2 * q * n2;  // position of the even tap; odd tap is pos + 1
      float le = l, he = h;
      if (pos >= 4 * n2) le = he = 0.f;
      if (pos + 1 >= 4 * n2) l = h = 0.f;
      el = fmaf(LL_REC_LO(2 * t), le, el);
      eh = fmaf(LL_REC_HI(2 * t), he, eh);
      ol = fmaf(LL_REC_LO(2 * t + 1), l, ol);
      oh = fmaf(LL_REC_HI(2 * t + 1), h, oh);
      continue;
    }
    el = fmaf(LL_REC_LO(2 * t), l, el);
    eh = fmaf(LL_REC_HI(2 * t), h, eh);
    ol = fmaf(LL_REC_LO(2 * t + 1), l, ol);
    oh = fmaf(LL_REC_HI(2 * t + 1), h, oh);
  }
  even = el + eh;
  odd = ol + oh;
}

LL_HD void dwti_cols(const DwtParams& p, const DwtTile& t, float* sm, int tid) {
  // height axis first: mid_lo = syn(ll, lh), mid_hi = syn(hl, hh)
  for (int e = tid; e < 2 * DW_TY * DWI_P; e += DW_THREADS) {
    const int which = e / (DW_TY * DWI_P);
    const int jl = (e / DWI_P) % DW_TY, c = e % DWI_P;
    const float* lo = &sm[DWI_SM_SB + (which ? 2 : 0) * DWI_R * DWI_P + jl * DWI_P + c];
    const float* hi = &sm[DWI_SM_SB + (which ? 3 : 1) * DWI_R * DWI_P + jl * DWI_P + c];
    float ev, od;
    syn_pair(lo, hi, DWI_P, t.y0 - 2 + jl, p.h / 2, ev, od);
    float* o = &sm[(which ? DWI_SM_HI : DWI_SM_LO) + (2 * jl) * DWI_P + c];
    o[0] = ev;
    o[DWI_P] = od;
  }
}

LL_HD void dwti_rows(const DwtParams& p, const DwtTile& t, const float* sm, int tid) {
  const int h2 = p.h / 2, w2 = p.w / 2;
  for (int e = tid; e < 2 * DW_TY * DW_TX; e += DW_THREADS) {
    const int r = e / DW_TX, jl = e % DW_TX;
    const int gy = 2 * t.y0 + r, gx = t.x0 + jl;
    if (gy >= 2 * h2 || gx >= w2) continue;
    float ev, od;
    syn_pair(&sm[DWI_SM_LO + r * DWI_P + jl], &sm[DWI_SM_HI + r * DWI_P + jl], 1, t.x0 - 2 + jl, w2, ev, od);
    float* o = p.xo + (long long)t.n * p.x_sn + (long long)gy * p.w + 2 * gx;
    o[0] = ev;
    o[1] = od;
  }
}

}  // namespace ll

// =============================================================================================
// Fast path (register-tiled): used when w % 8 == 0, h/2 >= 4, w/2 >= 4 and all plane bases are
// 16-byte aligned; otherwise the generic phases above run.  Same arithmetic per output sample
// (same tap order), 2-2.5x fewer shared-memory transactions and instructions per pixel:
//   forward : row pass 4 outputs per thread from 4 float4 loads; column pass 4 vertical outputs
//             per thread from a 15-row register window;
//   inverse : column synthesis 8 output rows per thread from an 8-row window; row synthesis
//             4 output pairs per thread from 3+3 float4 loads, 2 float4 stores.
// Pitches are chosen so that the float4 accesses of a quarter-warp hit 8 distinct bank groups.
// =============================================================================================
namespace ll {

constexpr int DFF_PI = 2 * DW_TX + 12;   // 140: input tile pitch (== 12 mod 32)
constexpr int DFF_PM = 2 * DW_TX + 8;    // 136: mid pitch; a mid row holds (lo, hi) pairs interleaved per column
constexpr int DFF_SM_IN = 0;
constexpr int DFF_SM_LO = DFF_SM_IN + DWF_R * DFF_PI;   // mid base
constexpr int DFF_SM_TOTAL = DFF_SM_LO + DWF_R * DFF_PM;

// analysis / synthesis tap pairs for packed FFMA2: one broadcast sample feeds the low-pass and the
// high-pass (or the even and the odd polyphase) accumulator in one instruction.  A zero tap leaves its
// accumulator untouched (fma(0, v, acc) == acc), so the results equal the scalar tap loops bit for bit.
#define LL_DEC_PAIR(k) (f2{LL_DEC_LO(k), LL_DEC_HI(k)})

// One fold.  The fast path runs on extents >= 10, where a tile's halo reaches at most 4 samples past either
// end; indices further out only occur in the unused part of a partial tile and may map to any valid sample.
LL_HD int wrap1(int a, int n) {
  a += (a < 0) ? n : 0;
  a -= (a >= n) ? n : 0;
  return (unsigned)a < (unsigned)n ? a : 0;
}

// 32-bit tile decode (the launchers reject grids of 2^31 tiles or more)
LL_HD DwtTile dwt_tile32(const DwtParams& p, unsigned t) {
  DwtTile d;
  const unsigned tx = (unsigned)p.tiles_x, ty = (unsigned)p.tiles_y;
  const unsigned q = t / tx;
  d.x0 = int(t - q * tx) * DW_TX;
  const unsigned n = q / ty;
  d.y0 = int(q - n * ty) * DW_TY;
  d.n = int(n);
  return d;
}

struct CopySync16 {  // host emulation / generic: plain 16-byte copy
  LL_HD void operator()(float* dst, const float* src) const {
    *reinterpret_cast<float4*>(dst) = *reinterpret_cast<const float4*>(src);
  }
};

// The nine analysis tap pairs live in registers for the whole (persistent) kernel: passing them in from memory the
// compiler cannot see through keeps it from re-materialising two 32-bit immediates in front of every FFMA2.
struct DwtTaps {
  f2 d[9];   // d[k-1] = (dec_lo[k], dec_hi[k]), k = 1..9
};
LL_HD void dwt_taps_init(float* t18) {
  const float lo[9] = {LL_DEC_LO(1), LL_DEC_LO(2), LL_DEC_LO(3), LL_DEC_LO(4), LL_DEC_LO(5), LL_DEC_LO(6), LL_DEC_LO(7), LL_DEC_LO(8), LL_DEC_LO(9)};
  const float hi[9] = {LL_DEC_HI(1), LL_DEC_HI(2), LL_DEC_HI(3), LL_DEC_HI(4), LL_DEC_HI(5), LL_DEC_HI(6), LL_DEC_HI(7), LL_DEC_HI(8), LL_DEC_HI(9)};
  for (int k = 0; k < 9; ++k) {
    t18[2 * k] = lo[k];
    t18[2 * k + 1] = hi[k];
  }
}

template <class CP>
LL_HD void dwtff_load(const DwtParams& p, const DwtTile& t, float* sm, int tid, CP cp) {
  constexpr int C4 = DWF_C / 4;                   // 34 column groups, 7 rows in flight per pass
  constexpr int RPP = DW_THREADS / C4;
  if (tid >= C4 * RPP) return;
  const int c4 = tid % C4, r0 = tid / C4;
  int gr = wrap1(2 * t.y0 - 4 + r0, p.h);       // rows advance by RPP < h: one conditional fold per step
  const float* src = p.x + (long long)t.n * p.x_sn + wrap1(2 * t.x0 - 4 + 4 * c4, p.w) + (long long)gr * p.w;
  const long long step = (long long)RPP * p.w, fold = step - (long long)p.h * p.w;
  float* dst = sm + DFF_SM_IN + r0 * DFF_PI + 4 * c4;
#pragma unroll
  for (int k = 0; k < (DWF_R + RPP - 1) / RPP; ++k) {
    if (r0 + RPP * k < DWF_R) cp(dst, src);
    dst += RPP * DFF_PI;
    gr += RPP;
    const bool f = gr >= p.h;
    gr -= f ? p.h : 0;
    src += f ? fold : step;
  }
}

// ``in`` / ``mid`` are view bases: the input tile is at in + DFF_SM_IN, the mid rows at mid + DFF_SM_LO
// (same base in the single-buffer layout, different bases in the double-buffered kernel).
LL_HD void dwtff_rows(const float* in, float* mid, int tid, const DwtTaps& tp) {
  // item = (row pair, 16 groups of 4 outputs); a warp covers 2 rows x 16 groups
  const int warp = tid >> 5, lane = tid & 31;
  const int g = (lane & 3) | ((lane >> 3) << 2);
  const int r = (lane >> 2) & 1;
  const float* q = &in[DFF_SM_IN + (2 * warp + r) * DFF_PI + 8 * g];
  float* o = &mid[DFF_SM_LO + (2 * warp + r) * DFF_PM + 8 * g];
#pragma unroll
  for (int rp = 0; rp < (DWF_R / 2 + DW_THREADS / 32 - 1) / (DW_THREADS / 32); ++rp) {
    if (warp + rp * (DW_THREADS / 32) < DWF_R / 2) {
      float v[16];
#pragma unroll
      for (int i = 0; i < 16; i += 4) {
        const float4 f = *reinterpret_cast<const float4*>(q + i);
        v[i] = f.x;
        v[i + 1] = f.y;
        v[i + 2] = f.z;
        v[i + 3] = f.w;
      }
      f2 acc[4];   // (lo, hi) of output 4g + i
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc[i] = f2{0.f, 0.f};
#pragma unroll
        for (int k = 1; k <= 9; ++k) LL_FMA2(acc[i], v[2 * i + 9 - k], tp.d[k - 1]);
      }
      *reinterpret_cast<float4*>(o) = float4{acc[0].x, acc[0].y, acc[1].x, acc[1].y};
      *reinterpret_cast<float4*>(o + 4) = float4{acc[2].x, acc[2].y, acc[3].x, acc[3].y};
    }
    q += 2 * (DW_THREADS / 32) * DFF_PI;
    o += 2 * (DW_THREADS / 32) * DFF_PM;
  }
}

LL_HD void dwtff_cols(const DwtParams& p, const DwtTile& t, const float* sm, int tid, const DwtTaps& tp) {
  const int h2 = p.h / 2, w2 = p.w / 2;
  const long long sub = (long long)h2 * w2;
  const int nl = tid % DW_TX, mg = tid / DW_TX;   // 64 columns x 4 groups of 4 output rows
  const int gx = t.x0 + nl, gy0 = t.y0 + 4 * mg;
  if (gx >= w2 || gy0 >= h2) return;
  f2 ab[15];   // (row-lo, row-hi) of this column over the 15-row window
  const float* q = &sm[DFF_SM_LO + (8 * mg) * DFF_PM + 2 * nl];
#pragma unroll
  for (int i = 0; i < 15; ++i) ab[i] = *reinterpret_cast<const f2*>(q + i * DFF_PM);
  const long long o0 = (long long)gy0 * w2 + gx;
  float* pLL = p.llo + (long long)t.n * p.ll_sn + o0;
  float* pLH = p.yho + (long long)t.n * p.yh_sn + o0;
  float* pHL = pLH + sub;
  float* pHH = pHL + sub;
  const int nrow = h2 - gy0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f2 A = f2{0.f, 0.f}, B = f2{0.f, 0.f};   // A = (LL, LH) from the row-low samples, B = (HL, HH) from the row-high ones
#pragma unroll
    for (int k = 1; k <= 9; ++k) {
      LL_FMA2(A, ab[2 * i + 9 - k].x, tp.d[k - 1]);
      LL_FMA2(B, ab[2 * i + 9 - k].y, tp.d[k - 1]);
    }
    if (i < nrow) {                          // rows below the plane (partial tile) are computed on unused samples and dropped
      const unsigned off = (unsigned)(i * w2);
      pLL[off] = A.x;
      pLH[off] = A.y;
      pHL[off] = B.x;
      pHH[off] = B.y;
    }
  }
}

// ---- inverse fast path ----
constexpr int DIF_THREADS = 288;          // 9 warps: 288 column-synthesis items, 80 subband lines = 5 x 16 load passes
constexpr int DIF_P = DW_TX + 8;          // 72: subband tile / mid pitch; column c <-> subband col x0 - 4 + c
constexpr int DIF_SM_SB = 0;              // [4][DWI_R][DIF_P]
constexpr int DIF_SM_LO = DIF_SM_SB + 4 * DWI_R * DIF_P;   // [2 TY][DIF_P]
constexpr int DIF_SM_HI = DIF_SM_LO + 2 * DW_TY * DIF_P;
constexpr int DIF_SM_TOTAL = DIF_SM_HI + 2 * DW_TY * DIF_P;

// synthesis tap pairs, register-resident like DwtTaps: l[t] = (rec_lo[2t], rec_lo[2t+1]), h[t] = (rec_hi[2t], rec_hi[2t+1])
struct DwtSynTaps {
  f2 l[5], h[5];
};
LL_HD void dwt_syn_taps_init(float* t20) {
  const float lo[10] = {LL_REC_LO(0), LL_REC_LO(1), LL_REC_LO(2), LL_REC_LO(3), LL_REC_LO(4), LL_REC_LO(5), LL_REC_LO(6), LL_REC_LO(7), LL_REC_LO(8), LL_REC_LO(9)};
  const float hi[10] = {LL_REC_HI(0), LL_REC_HI(1), LL_REC_HI(2), LL_REC_HI(3), LL_REC_HI(4), LL_REC_HI(5), LL_REC_HI(6), LL_REC_HI(7), LL_REC_HI(8), LL_REC_HI(9)};
  for (int k = 0; k < 10; ++k) {
    t20[k] = lo[k];
    t20[10 + k] = hi[k];
  }
}

template <class CP>
LL_HD void dwtif_load(const DwtParams& p, const DwtTile& t, float* sm, int tid, CP cp) {
  const int h2 = p.h / 2, w2 = p.w / 2;
  const long long sub = (long long)h2 * w2;
  constexpr int C4 = DIF_P / 4;                   // 18 column groups, 16 (subband,row) lines in flight per pass
  constexpr int RPP = DIF_THREADS / C4;
  const int c4 = tid % C4, r0 = tid / C4;
  const int gx = wrap1(t.x0 - 4 + 4 * c4, w2);
  const float* b0 = p.ll + (long long)t.n * p.ll_sn + gx;
  const float* b1 = p.yh + (long long)t.n * p.yh_sn + gx;
  float* dst = sm + DIF_SM_SB + r0 * DIF_P + 4 * c4;
#pragma unroll
  for (int k = 0; k < 4 * DWI_R / RPP; ++k) {
    const int rs = r0 + RPP * k, s = rs / DWI_R, i = rs - s * DWI_R;
    const int gy = wrap1(t.y0 - 2 + i, h2);
    const float* src = (s == 0 ? b0 : b1 + (s - 1) * sub) + (long long)gy * w2;
    cp(dst, src);
    dst += RPP * DIF_P;
  }
}
static_assert((4 * DWI_R) % (DIF_THREADS / (DIF_P / 4)) == 0, "inverse load passes cover the subband lines exactly");

LL_HD void dwtif_cols(const float* sb, float* mid, int tid, const DwtSynTaps& tp) {
  // item = (lo|hi, group of 4 j's, column pair): 2 x 4 x 36 = one per thread
  const int cp = tid % (DIF_P / 2), jg = (tid / (DIF_P / 2)) & 3, which = tid / (2 * DIF_P);
  const float* lo = &sb[DIF_SM_SB + ((which ? 2 : 0) * DWI_R + 4 * jg) * DIF_P + 2 * cp];
  const float* hi = lo + DWI_R * DIF_P;          // lh follows ll, hh follows hl
  f2 l[8], h[8];                                 // (column 2cp, column 2cp+1) of the 8-row window
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    l[i] = *reinterpret_cast<const f2*>(lo + i * DIF_P);
    h[i] = *reinterpret_cast<const f2*>(hi + i * DIF_P);
  }
  float* o = &mid[(which ? DIF_SM_HI : DIF_SM_LO) + (8 * jg) * DIF_P + 2 * cp];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    f2 al0 = f2{0.f, 0.f}, ah0 = f2{0.f, 0.f}, al1 = f2{0.f, 0.f}, ah1 = f2{0.f, 0.f};   // (even, odd) partial sums
#pragma unroll
    for (int tt = 0; tt < 5; ++tt) {
      LL_FMA2(al0, l[j + 4 - tt].x, tp.l[tt]);
      LL_FMA2(ah0, h[j + 4 - tt].x, tp.h[tt]);
      LL_FMA2(al1, l[j + 4 - tt].y, tp.l[tt]);
      LL_FMA2(ah1, h[j + 4 - tt].y, tp.h[tt]);
    }
    *reinterpret_cast<f2*>(o + (2 * j) * DIF_P) = f2{al0.x + ah0.x, al1.x + ah1.x};
    *reinterpret_cast<f2*>(o + (2 * j + 1) * DIF_P) = f2{al0.y + ah0.y, al1.y + ah1.y};
  }
}
static_assert(2 * (DW_TY / 4) * (DIF_P / 2) == DIF_THREADS, "one column-synthesis item per thread");

LL_HD void dwtif_rows(const DwtParams& p, const DwtTile& t, const float* sm, int tid, const DwtSynTaps& tp) {
  const int w2 = p.w / 2;
  float* plane = p.xo + (long long)t.n * p.x_sn;
  // an item's 8 outputs are 32 contiguous bytes; they are 32-byte aligned when the plane base and pitches are
  // (w % 8 == 0 is a fast-path condition).  Uniform per launch.
  const bool wide = ((reinterpret_cast<uintptr_t>(p.xo) & 31) == 0) && (p.x_sn % 8 == 0);
  (void)wide;
  // item = (output row, group of 4 output pairs): 32 x 16 = 512
#pragma unroll
  for (int k = 0; k < (2 * DW_TY * (DW_TX / 4) + DIF_THREADS - 1) / DIF_THREADS; ++k) {
    const int e = tid + k * DIF_THREADS;
    const int jg = e % (DW_TX / 4), r = e / (DW_TX / 4);
    const int gy = 2 * t.y0 + r, gx = t.x0 + 4 * jg;
    if (e >= 2 * DW_TY * (DW_TX / 4) || gy >= p.h || gx >= w2) continue;
    float l[12], h[12];
#pragma unroll
    for (int i = 0; i < 12; i += 4) {
      const float4 a = *reinterpret_cast<const float4*>(&sm[DIF_SM_LO + r * DIF_P + 4 * jg + i]);
      const float4 b = *reinterpret_cast<const float4*>(&sm[DIF_SM_HI + r * DIF_P + 4 * jg + i]);
      l[i] = a.x; l[i + 1] = a.y; l[i + 2] = a.z; l[i + 3] = a.w;
      h[i] = b.x; h[i + 1] = b.y; h[i + 2] = b.z; h[i + 3] = b.w;
    }
    float out[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      // output pair jl = 4 jg + j needs subband cols jl-2..jl+2 -> tile index c = jl + 2 .. jl + 6 -> local j + 2 .. j + 6
      f2 al = f2{0.f, 0.f}, ah = f2{0.f, 0.f};
#pragma unroll
      for (int tt = 0; tt < 5; ++tt) {
        LL_FMA2(al, l[j + 6 - tt], tp.l[tt]);
        LL_FMA2(ah, h[j + 6 - tt], tp.h[tt]);
      }
      out[2 * j] = al.x + ah.x;
      out[2 * j + 1] = al.y + ah.y;
    }
    float* o = plane + (long long)gy * p.w + 2 * gx;
#if defined(__CUDA_ARCH__)
    if (wide) {   // one 256-bit store per item: every store instruction writes whole 32-byte sectors
      asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(o), "f"(out[0]), "f"(out[1]), "f"(out[2]),
                   "f"(out[3]), "f"(out[4]), "f"(out[5]), "f"(out[6]), "f"(out[7])
                   : "memory");
      continue;
    }
#endif
    *reinterpret_cast<float4*>(o) = float4{out[0], out[1], out[2], out[3]};
    *reinterpret_cast<float4*>(o + 4) = float4{out[4], out[5], out[6], out[7]};
  }
}

// fast-path eligibility (uniform per launch)
LL_HD bool dwt_fast_ok(const DwtParams& p) {
  const bool al = ((reinterpret_cast<uintptr_t>(p.x) | reinterpret_cast<uintptr_t>(p.xo) |
                    reinterpret_cast<uintptr_t>(p.ll) | reinterpret_cast<uintptr_t>(p.llo) |
                    reinterpret_cast<uintptr_t>(p.yh) | reinterpret_cast<uintptr_t>(p.yho)) & 15) == 0;
  return al && (p.w % 8 == 0) && (p.h / 2 >= 4) && (p.w / 2 >= 4) && (p.x_sn % 4 == 0) && (p.ll_sn % 4 == 0) &&
         (p.yh_sn % 4 == 0) && (((long long)(p.h / 2) * (p.w / 2)) % 4 == 0) && (p.h >= 10) && (p.w >= 10);
}

}  // namespace ll
