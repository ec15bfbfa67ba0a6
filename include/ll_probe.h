/*
 * ll_probe.h -- measurement / unit-probe entry points of libll_probe.so (TEST AND BENCH TOOLING, not part of the
 * product ABI in ll_api.h; none of them has a reference counterpart).  Built from csrc/probe/ by build.py.
 */
#ifndef LL_PROBE_H
#define LL_PROBE_H
#include "ll_api.h"
#ifdef __cplusplus
extern "C" {
#endif

/* Unit probe of the tensor-core building block of the learned-lifting kernel (no reference
 * counterpart): D (128,64) = A (128, 8*kblocks) x B (8*kblocks, 64), fp32 in/out, computed with
 * tcgen05.mma kind::tf32 (A resident in tensor memory, B in MN-major SWIZZLE_128B shared-memory atoms);
 * split != 0 uses the 3xTF32 hi/lo split (fp32-level accuracy).  cycles (device, optional) receives
 * the SM cycles of one chain.  kblocks <= 20. */
int ll_tc_tf32_probe(const float* A, const float* B, float* D, int kblocks, int split, int reps, long long* cycles,
                     ll_stream_t stream);

/* Measurement helper (no reference counterpart): register-only FFMA2 loop used by bench.py to
 * measure the device's FP32 FMA-pipe peak.  FLOPs = blocks * 256 * iters * 256. */
int ll_fma_peak_probe(float* out, int blocks, int iters, ll_stream_t stream);

/* Dense tcgen05 kind::tf32 GEMM loop (A, B in shared memory, accumulator in tensor memory, no global traffic):
 * the TF32 tensor-pipe peak of the device, measured.  Every CTA issues `iters` chains of `kblocks` MMAs of shape
 * M128 x N x K8 (N in 16..256, multiple of 16).  FLOPs = blocks * iters * kblocks * 2 * 128 * N * 8. */
int ll_tf32_peak_probe(float* out, int blocks, int iters, int kblocks, int n, ll_stream_t stream);

/* Shifted-window operand addressing for a 3x3 implicit GEMM (csrc/probe/halo_probe.cu): one TMA load of an 18 x pitch
 * pixel halo tile of a (18,16,32) fp32 channels-last tensor `a`, A operand of tap (dy,dx) addressed inside it by a
 * K-major SWIZZLE_128B descriptor with start address moved by whole rows and stride byte offset = pitch * 128.
 * d (128,32) = A_tap (128 = 16 rows x 8 pixels, 32) x b (32,32)^T.  bo_mode 1: descriptor base_offset = start bits 7..9. */
int ll_halo_probe(const float* a, const float* b, float* d, int pitch, int dy, int dx, int bo_mode, ll_stream_t stream);

/* libll_probe.so carries a copy of csrc/igemm_conv.cu compiled with LL_TIMELINE (csrc/probe/igemm_timeline.cu): when a
 * device buffer of 16 x 64 int64 is registered here, CTA 0 of igemm_tf32_gdn_pair_kernel launched THROUGH THE PROBE
 * LIBRARY's ll_igemm_tf32_gdn / ll_conv3_gdn_head stamps clock64() at every hand-off of its first 16 tiles. */
int ll_probe_set_timeline(long long* buf);
/* Timeline experiments only: E2 of the probe library's fused kernel skips its global stores (results are then garbage). */
int ll_probe_set_nostore(int on);

/* libll_probe.so also carries csrc/lift_step.cu + csrc/lift_tc.cu compiled with LL_DEBUG (csrc/probe/lift_dbg.cu), reached
 * through the probe library's own ll_lift_step: role ablation of lift_step_tc_kernel (bit 0 no MMA, 1 no E-B, 2 no conv1,
 * 3 no conv4, 4 no E-A: results are WRONG when non-zero) and the clock64 stamps of CTA 0's 17 warps at global step 40
 * (device buffer of 17 x 8 int64, NULL = off).  scripts/gpu_lift_tc_timeline.py. */
int ll_dbg_lift_switches(int bits);
int ll_dbg_lift_stamp_buffer(long long* buf);

#ifdef __cplusplus
}
#endif
#endif /* LL_PROBE_H */
