// lift_dbg.cu -- probe build of the learned-lifting kernels (csrc/lift_step.cu + csrc/lift_tc.cu) with LL_DEBUG: exposes
// ll_dbg_lift_switches (role ablation) and ll_dbg_lift_stamp_buffer (clock64 stamps of CTA 0's 17 warps at global step 40)
// for scripts/gpu_lift_tc_timeline.py.  Measurement tooling only; the product library is compiled without LL_DEBUG.
#define LL_DEBUG 1
#include "../lift_step.cu"
#include "../lift_tc.cu"
