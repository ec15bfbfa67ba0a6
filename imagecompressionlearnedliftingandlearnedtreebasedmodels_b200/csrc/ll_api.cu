// ll_api.cu -- library-wide entry points: version, device check, last error.
#include "ll_common.cuh"

namespace ll {
char* err_slot() {
  static thread_local char buf[ERR_LEN] = {0};
  return buf;
}
int sm_count_cached() {
  static thread_local int dev_cached = -1, n = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != dev_cached) {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return 148;
    n = prop.multiProcessorCount;
    dev_cached = dev;
  }
  return n;
}
}  // namespace ll

extern "C" {

const char* ll_last_error(void) { return ll::err_slot(); }

int ll_version(void) { return 100; }

int ll_check_device(void) {
  int dev = 0;
  LL_CUDA_OK(cudaGetDevice(&dev));
  int major = 0;
  LL_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10) return ll::fail(LL_EARCH, "device %d has compute capability %d.x; this library is sm_100a only", dev, major);
  return LL_OK;
}

int ll_sm_count(void) { return ll::sm_count_cached(); }

}  // extern "C"
