// lift_level_body.h -- composition of one 2-D lifting level out of lifting steps.
//
// Forward: wavelet_forward_v2.one_level_lifting (graphs/layers/wavelet_forward_v2.py:26-54);
// inverse: wavelet_inverse_v2.one_level_lifting + reconstruct_fun
// (graphs/layers/wavelet_inverse_v2.py:20-56).  The reference materialises even/odd slices,
// transposes and interleaves; here every one of those is a strided VIEW handed to the step
// kernel (row pass: y = plane row pair, x = plane column; column pass: the transposed view,
// y = plane column pair, x = plane row), so a level is 8 launches and no copies.
// Templated on a backend (CUDA launches in lift_step.cu, host emulation in tests/emul).
#pragma once
#include <stddef.h>
#include <stdint.h>

#include "../../include/ll_api.h"

namespace ll {

constexpr float K_NH = 0.869864451624781f, K_NL = 1.149604398860241f;  // lifting_coeff[4], [5]

inline ll_view3 view(const float* p, long long sb, long long sy, long long sx) {
  ll_view3 v;
  v.ptr = const_cast<float*>(p);
  v.sb = sb;
  v.sy = sy;
  v.sx = sx;
  return v;
}
inline ll_lift_job job(ll_view3 s, ll_view3 di, ll_view3 dn, int nb, int ny, int nx) {
  ll_lift_job j;
  j.src = s;
  j.din = di;
  j.dout = dn;
  j.nb = nb;
  j.ny = ny;
  j.nx = nx;
  return j;
}

inline size_t lift_level_scratch_floats(int B, int h, int w) {
  // bufL, bufH: (B, h/2, w) each; 4 x (B, h/2, w/2) intermediates for the inverse
  return (size_t)B * (h / 2) * w * 2 + (size_t)B * (h / 2) * (w / 2) * 4;
}

#define LL_TRY(expr)        \
  do {                      \
    int _rc = (expr);       \
    if (_rc) return _rc;    \
  } while (0)

template <class BE>
int lift_level_fwd_impl(BE& be, const float* x, int64_t x_sb, float* llp, int64_t ll_sb, float* yh, int64_t yh_sb,
                        float* scratch, int B, int h, int w, const float* const* blobs, int scale,
                        const float* nh, const float* nl) {
  const int h2 = h / 2, w2 = w / 2;
  float* bufL = scratch;
  float* bufH = scratch + (size_t)B * h2 * w;
  const long long sbuf = (long long)h2 * w;
  // rows: L = x[0::2], H = x[1::2]   (wavelet_forward_v2.py:27-29)
  ll_view3 xe = view(x, x_sb, 2LL * w, 1), xo = view(x + w, x_sb, 2LL * w, 1);
  ll_view3 vL = view(bufL, sbuf, w, 1), vH = view(bufH, sbuf, w, 1);
  ll_lift_job j1 = job(xe, xo, vH, B, h2, w);  // H = H + f1(L)
  LL_TRY(be.step(&j1, 1, blobs[0], 1.f));
  ll_lift_job j2 = job(vH, xe, vL, B, h2, w);  // L = L + f2(H)
  LL_TRY(be.step(&j2, 1, blobs[1], 1.f));
  ll_lift_job j3 = job(vL, vH, vH, B, h2, w);  // H = H + f3(L)
  LL_TRY(be.step(&j3, 1, blobs[2], 1.f));
  ll_lift_job j4 = job(vH, vL, vL, B, h2, w);  // L = L + f4(H)
  LL_TRY(be.step(&j4, 1, blobs[3], 1.f));
  if (scale) {
    LL_TRY(be.scale(vH, B, h2, w, nh, K_NH, 0));
    LL_TRY(be.scale(vL, B, h2, w, nl, K_NL, 0));
  }
  // columns on L -> (LL, HL) and on H -> (LH, HH), both in one launch per step (:32-51)
  const long long sub = (long long)h2 * w2;
  ll_view3 Le = view(bufL, sbuf, 2, w), Lo = view(bufL + 1, sbuf, 2, w);
  ll_view3 He = view(bufH, sbuf, 2, w), Ho = view(bufH + 1, sbuf, 2, w);
  ll_view3 LL = view(llp, ll_sb, 1, w2);
  ll_view3 LH = view(yh, yh_sb, 1, w2), HL = view(yh + sub, yh_sb, 1, w2), HH = view(yh + 2 * sub, yh_sb, 1, w2);
  ll_lift_job c1[2] = {job(Le, Lo, HL, B, w2, h2), job(He, Ho, HH, B, w2, h2)};
  LL_TRY(be.step(c1, 2, blobs[0], 1.f));
  ll_lift_job c2[2] = {job(HL, Le, LL, B, w2, h2), job(HH, He, LH, B, w2, h2)};
  LL_TRY(be.step(c2, 2, blobs[1], 1.f));
  ll_lift_job c3[2] = {job(LL, HL, HL, B, w2, h2), job(LH, HH, HH, B, w2, h2)};
  LL_TRY(be.step(c3, 2, blobs[2], 1.f));
  ll_lift_job c4[2] = {job(HL, LL, LL, B, w2, h2), job(HH, LH, LH, B, w2, h2)};
  LL_TRY(be.step(c4, 2, blobs[3], 1.f));
  if (scale) {
    LL_TRY(be.scale(HL, B, w2, h2, nh, K_NH, 0));
    LL_TRY(be.scale(HH, B, w2, h2, nh, K_NH, 0));
    LL_TRY(be.scale(LL, B, w2, h2, nl, K_NL, 0));
    LL_TRY(be.scale(LH, B, w2, h2, nl, K_NL, 0));
  }
  return LL_OK;
}

template <class BE>
int lift_level_inv_impl(BE& be, const float* llp, int64_t ll_sb, const float* yh, int64_t yh_sb, float* x,
                        int64_t x_sb, float* scratch, int B, int h, int w, const float* const* blobs, int scale,
                        const float* nh, const float* nl) {
  const int h2 = h / 2, w2 = w / 2;
  float* bufL = scratch;
  float* bufH = scratch + (size_t)B * h2 * w;
  float* tmp = scratch + (size_t)B * h2 * w * 2;
  const long long sbuf = (long long)h2 * w;
  const long long sub = (long long)h2 * w2;
  // inputs as transposed views (wavelet_inverse_v2.py:21-22, 29-30)
  ll_view3 LL = view(llp, ll_sb, 1, w2);
  ll_view3 LH = view(yh, yh_sb, 1, w2), HL = view(yh + sub, yh_sb, 1, w2), HH = view(yh + 2 * sub, yh_sb, 1, w2);
  // dense intermediates (B, h2, w2), viewed transposed
  ll_view3 tLL = view(tmp, sub, 1, w2), tHL = view(tmp + (size_t)B * sub, sub, 1, w2);
  ll_view3 tLH = view(tmp + (size_t)B * sub * 2, sub, 1, w2), tHH = view(tmp + (size_t)B * sub * 3, sub, 1, w2);
  ll_view3 Le = view(bufL, sbuf, 2, w), Lo = view(bufL + 1, sbuf, 2, w);
  ll_view3 He = view(bufH, sbuf, 2, w), Ho = view(bufH + 1, sbuf, 2, w);
  ll_view3 aL = LL, aH = HL, bL = LH, bH = HH;
  if (scale) {
    // un-scale copies first (:70-74); the caller's inputs are never modified
    for (int b = 0; b < B; ++b) {
      LL_TRY(be.copy(tmp + (size_t)b * sub, llp + (size_t)b * ll_sb, (size_t)sub));
      LL_TRY(be.copy(tmp + (size_t)(B + b) * sub, yh + (size_t)b * yh_sb + sub, (size_t)sub));
      LL_TRY(be.copy(tmp + (size_t)(2 * B + b) * sub, yh + (size_t)b * yh_sb, (size_t)sub));
      LL_TRY(be.copy(tmp + (size_t)(3 * B + b) * sub, yh + (size_t)b * yh_sb + 2 * sub, (size_t)sub));
    }
    LL_TRY(be.scale(tHL, B, w2, h2, nh, K_NH, 1));
    LL_TRY(be.scale(tHH, B, w2, h2, nh, K_NH, 1));
    LL_TRY(be.scale(tLL, B, w2, h2, nl, K_NL, 1));
    LL_TRY(be.scale(tLH, B, w2, h2, nl, K_NL, 1));
    aL = tLL;
    aH = tHL;
    bL = tLH;
    bH = tHH;
  }
  // column inverse on (LL as L, HL as H) and (LH as L, HH as H): steps 4,3,2,1 (:76-90); the last
  // update of each half is written straight into the interleaved buffers (reconstruct_fun)
  ll_lift_job c4[2] = {job(aH, aL, tLL, B, w2, h2), job(bH, bL, tLH, B, w2, h2)};  // L -= f4(H)
  LL_TRY(be.step(c4, 2, blobs[3], -1.f));
  ll_lift_job c3[2] = {job(tLL, aH, tHL, B, w2, h2), job(tLH, bH, tHH, B, w2, h2)};  // H -= f3(L)
  LL_TRY(be.step(c3, 2, blobs[2], -1.f));
  ll_lift_job c2[2] = {job(tHL, tLL, Le, B, w2, h2), job(tHH, tLH, He, B, w2, h2)};  // L -= f2(H)
  LL_TRY(be.step(c2, 2, blobs[1], -1.f));
  ll_lift_job c1[2] = {job(Le, tHL, Lo, B, w2, h2), job(He, tHH, Ho, B, w2, h2)};  // H -= f1(L)
  LL_TRY(be.step(c1, 2, blobs[0], -1.f));
  // row inverse on (bufL, bufH) -> even / odd rows of x
  ll_view3 vL = view(bufL, sbuf, w, 1), vH = view(bufH, sbuf, w, 1);
  ll_view3 xe = view(x, x_sb, 2LL * w, 1), xo = view(x + w, x_sb, 2LL * w, 1);
  if (scale) {
    LL_TRY(be.scale(vH, B, h2, w, nh, K_NH, 1));
    LL_TRY(be.scale(vL, B, h2, w, nl, K_NL, 1));
  }
  ll_lift_job r4 = job(vH, vL, vL, B, h2, w);
  LL_TRY(be.step(&r4, 1, blobs[3], -1.f));
  ll_lift_job r3 = job(vL, vH, vH, B, h2, w);
  LL_TRY(be.step(&r3, 1, blobs[2], -1.f));
  ll_lift_job r2 = job(vH, vL, xe, B, h2, w);
  LL_TRY(be.step(&r2, 1, blobs[1], -1.f));
  ll_lift_job r1 = job(xe, vH, xo, B, h2, w);
  LL_TRY(be.step(&r1, 1, blobs[0], -1.f));
  return LL_OK;
}

}  // namespace ll
