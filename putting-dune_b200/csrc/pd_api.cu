// Library-level entry points and error plumbing of the C ABI.
#include <stdarg.h>
#include <stdio.h>

#include "pd_common.cuh"

namespace pd {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

int check_cuda(cudaError_t err, const char* what) {
  if (err == cudaSuccess) return PD_OK;
  set_error("CUDA error %d (%s) in %s", static_cast<int>(err),
            cudaGetErrorString(err), what);
  return (err == cudaErrorNoDevice || err == cudaErrorInsufficientDriver)
             ? PD_ERR_NO_DEVICE
             : PD_ERR_CUDA;
}

int sm_count() {
  static thread_local int cached_dev = -1;
  static thread_local int cached = 148;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return cached;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) ==
            cudaSuccess &&
        n > 0)
      cached = n;
    cached_dev = dev;
  }
  return cached;
}

}  // namespace pd

extern "C" int pd_abi_version(void) { return PDUNE_B200_ABI_VERSION; }

extern "C" const char* pd_last_error(void) { return pd::g_error; }

extern "C" int pd_device_sm_count(int* out_sm_count) {
  PD_REQUIRE(out_sm_count != nullptr, "null output");
  int dev = 0;
  PD_CUDA_OK(cudaGetDevice(&dev));
  PD_CUDA_OK(cudaDeviceGetAttribute(out_sm_count,
                                    cudaDevAttrMultiProcessorCount, dev));
  return PD_OK;
}
