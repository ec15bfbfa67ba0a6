// lift_emul.cpp -- HOST emulation of the lifting-step kernel body.  TEST TOOL ONLY.
// Compiles csrc/lift_step_body.cuh for the host: a "phase" is a loop over the 256 thread ids
// of one CTA, CTAs run one after another.  It exists so that tile/ring/halo indexing can be
// checked against the oracle in the GPU-less dev container; it is never loaded by the product
// package and is far too slow to be a fallback.
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../imagecompressionlearnedliftingandlearnedtreebasedmodels_b200/csrc/lift_step_body.cuh"
#include "../../imagecompressionlearnedliftingandlearnedtreebasedmodels_b200/csrc/lift_level_body.h"

using namespace ll;

extern "C" {

void ll_emul_pack_lift_step(const float* pre, const float* w1, const float* b1, const float* w2, const float* b2,
                            const float* w3, const float* b3, const float* w4, const float* b4, float* blob) {
  for (int i = 0; i < BL_TOTAL; ++i) blob[i] = pack_lift_elem(i, pre, w1, b1, w2, b2, w3, b3, w4, b4);
}

int ll_emul_lift_step(const ll_lift_job* jobs, int njobs, const float* blob, float sign, float rw, int linear,
                      int ncta) {
  LiftParams p;
  memset(&p, 0, sizeof(p));
  p.njobs = njobs;
  p.blob = blob;
  p.sign = sign;
  p.rw = rw;
  p.linear = linear;
  for (int j = 0; j < 2; ++j) {
    if (j < njobs) {
      p.job[j] = jobs[j];
      p.nstrips[j] = (jobs[j].nx + LS_WT - 1) / LS_WT;
      p.nchunks[j] = (jobs[j].ny + LS_R - 1) / LS_R;
      p.units[j] = (long long)jobs[j].nb * p.nstrips[j] * p.nchunks[j];
    } else {
      p.nstrips[j] = p.nchunks[j] = 1;
    }
    p.total_units += p.units[j];
  }
  if (p.total_units == 0) return 0;
  if (ncta > p.total_units) ncta = (int)p.total_units;
  // poison shared memory so that reads of never-written slots show up as NaNs in the output
  std::vector<float> smv(SM_TOTAL);
  float* sm = smv.data();
  for (int cta = 0; cta < ncta; ++cta) {
    for (int i = 0; i < SM_TOTAL; ++i) sm[i] = __builtin_nanf("");
#define LL_PHASE(call) \
  for (int tid = 0; tid < LS_THREADS; ++tid) { call; }
    LL_LIFT_STEP_DRIVER(LL_PHASE, p, sm, cta, ncta);
#undef LL_PHASE
  }
  return 0;
}


struct HostBackend {
  float rw;
  int linear;
  int ncta;
  int step(const ll_lift_job* jobs, int n, const float* blob, float sign) {
    return ll_emul_lift_step(jobs, n, blob, sign, rw, linear, ncta);
  }
  int scale(ll_view3 v, int nb, int ny, int nx, const float* n, float base, int divide) {
    const float s = base + n[0] * 0.1f;
    for (int b = 0; b < nb; ++b)
      for (int y = 0; y < ny; ++y)
        for (int x = 0; x < nx; ++x) {
          float* q = v.ptr + b * v.sb + y * v.sy + x * v.sx;
          *q = divide ? (*q / s) : (*q * s);
        }
    return 0;
  }
  int copy(float* dst, const float* src, size_t count) {
    memcpy(dst, src, count * sizeof(float));
    return 0;
  }
};

size_t ll_emul_lift_level_scratch_floats(int B, int h, int w) { return lift_level_scratch_floats(B, h, w); }

int ll_emul_lift_level_fwd(const float* x, int64_t x_sb, float* llp, int64_t ll_sb, float* yh, int64_t yh_sb,
                           float* scratch, int B, int h, int w, const float* const* blobs, float rw, int linear,
                           int scale, const float* nh, const float* nl, int ncta) {
  HostBackend be{rw, linear, ncta};
  return lift_level_fwd_impl(be, x, x_sb, llp, ll_sb, yh, yh_sb, scratch, B, h, w, blobs, scale, nh, nl);
}

int ll_emul_lift_level_inv(const float* llp, int64_t ll_sb, const float* yh, int64_t yh_sb, float* x, int64_t x_sb,
                           float* scratch, int B, int h, int w, const float* const* blobs, float rw, int linear,
                           int scale, const float* nh, const float* nl, int ncta) {
  HostBackend be{rw, linear, ncta};
  return lift_level_inv_impl(be, llp, ll_sb, yh, yh_sb, x, x_sb, scratch, B, h, w, blobs, scale, nh, nl);
}

}  // extern "C"

// ---- CDF 9/7 emulation ---------------------------------------------------------------------
#include "../../imagecompressionlearnedliftingandlearnedtreebasedmodels_b200/csrc/dwt97_body.cuh"

extern "C" {

int ll_emul_dwt97_fwd_level(const float* x, int64_t x_sn, float* llp, int64_t ll_sn, float* yh, int64_t yh_sn, int N,
                            int h, int w) {
  DwtParams p = {};
  p.N = N; p.h = h; p.w = w;
  p.tiles_x = (w / 2 + DW_TX - 1) / DW_TX;
  p.tiles_y = (h / 2 + DW_TY - 1) / DW_TY;
  p.x = x; p.x_sn = x_sn; p.llo = llp; p.ll_sn = ll_sn; p.yho = yh; p.yh_sn = yh_sn;
  std::vector<float> smv((DWF_SM_TOTAL > DFF_SM_TOTAL ? DWF_SM_TOTAL : DFF_SM_TOTAL) + 4);
  float* sm = (float*)(((uintptr_t)smv.data() + 15) & ~(uintptr_t)15);
  const long long tiles = (long long)N * p.tiles_x * p.tiles_y;
  for (long long b = 0; b < tiles; ++b) {
    for (int i = 0; i < (DWF_SM_TOTAL > DFF_SM_TOTAL ? DWF_SM_TOTAL : DFF_SM_TOTAL); ++i) sm[i] = __builtin_nanf("");
    const DwtTile t = dwt_fast_ok(p) ? dwt_tile32(p, (unsigned)b) : dwt_tile(p, b);
    if (dwt_fast_ok(p)) {
      for (int tid = 0; tid < DW_THREADS; ++tid) dwtff_load(p, t, sm, tid, CopySync16());
      DwtTaps tp;
      {
        float t18[18];
        dwt_taps_init(t18);
        for (int k = 0; k < 9; ++k) tp.d[k] = f2{t18[2 * k], t18[2 * k + 1]};
      }
      for (int tid = 0; tid < DW_THREADS; ++tid) dwtff_rows(sm, sm, tid, tp);
      for (int tid = 0; tid < DW_THREADS; ++tid) dwtff_cols(p, t, sm, tid, tp);
      continue;
    }
    for (int tid = 0; tid < DW_THREADS; ++tid) dwtf_load(p, t, sm, tid);
    for (int tid = 0; tid < DW_THREADS; ++tid) dwtf_rows(sm, tid);
    for (int tid = 0; tid < DW_THREADS; ++tid) dwtf_cols(p, t, sm, tid);
  }
  return 0;
}

int ll_emul_dwt97_inv_level(const float* llp, int64_t ll_sn, const float* yh, int64_t yh_sn, float* x, int64_t x_sn,
                            int N, int h, int w) {
  DwtParams p = {};
  p.N = N; p.h = h; p.w = w;
  p.tiles_x = (w / 2 + DW_TX - 1) / DW_TX;
  p.tiles_y = (h / 2 + DW_TY - 1) / DW_TY;
  p.xo = x; p.x_sn = x_sn; p.ll = llp; p.ll_sn = ll_sn; p.yh = yh; p.yh_sn = yh_sn;
  std::vector<float> smv((DWI_SM_TOTAL > DIF_SM_TOTAL ? DWI_SM_TOTAL : DIF_SM_TOTAL) + 4);
  float* sm = (float*)(((uintptr_t)smv.data() + 15) & ~(uintptr_t)15);
  const long long tiles = (long long)N * p.tiles_x * p.tiles_y;
  for (long long b = 0; b < tiles; ++b) {
    for (int i = 0; i < (DWI_SM_TOTAL > DIF_SM_TOTAL ? DWI_SM_TOTAL : DIF_SM_TOTAL); ++i) sm[i] = __builtin_nanf("");
    const DwtTile t = dwt_fast_ok(p) ? dwt_tile32(p, (unsigned)b) : dwt_tile(p, b);
    if (dwt_fast_ok(p)) {
      DwtSynTaps tp;
      {
        float t20[20];
        dwt_syn_taps_init(t20);
        for (int k = 0; k < 5; ++k) {
          tp.l[k] = f2{t20[2 * k], t20[2 * k + 1]};
          tp.h[k] = f2{t20[10 + 2 * k], t20[10 + 2 * k + 1]};
        }
      }
      for (int tid = 0; tid < DIF_THREADS; ++tid) dwtif_load(p, t, sm, tid, CopySync16());
      for (int tid = 0; tid < DIF_THREADS; ++tid) dwtif_cols(sm, sm, tid, tp);
      for (int tid = 0; tid < DIF_THREADS; ++tid) dwtif_rows(p, t, sm, tid, tp);
      continue;
    }
    for (int tid = 0; tid < DW_THREADS; ++tid) dwti_load(p, t, sm, tid);
    for (int tid = 0; tid < DW_THREADS; ++tid) dwti_cols(p, t, sm, tid);
    for (int tid = 0; tid < DW_THREADS; ++tid) dwti_rows(p, t, sm, tid);
  }
  return 0;
}

}  // extern "C"
