// igemm_timeline.cu -- probe build of the scaling-network kernels (csrc/igemm_conv.cu) with the LL_TIMELINE stamps compiled
// in: CTA 0 of igemm_tf32_gdn_pair_kernel records clock64() at every hand-off of its first 16 tiles (MMA thread: slots
// 0..15, epilogue warp 2: slots 32..53 of a 64-slot row per tile).  Test / measurement tooling only (include/ll_probe.h);
// the product library is compiled without the macro and carries none of this.
#define LL_TIMELINE 1
#include "../igemm_conv.cu"

extern "C" int ll_probe_set_timeline(long long* buf) {
  ll::g_timeline = buf;       // device buffer of 16 x 64 int64, or NULL to switch the stamps off
  return LL_OK;
}

extern "C" int ll_probe_set_nostore(int on) {   // timeline experiments: E2 without its global stores (results are then garbage)
  return cudaMemcpyToSymbol(ll::g_nostore_dev, &on, sizeof(int)) == cudaSuccess ? LL_OK : LL_ECUDA;
}
