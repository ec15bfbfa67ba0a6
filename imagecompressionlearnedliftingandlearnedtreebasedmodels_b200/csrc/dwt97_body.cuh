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
