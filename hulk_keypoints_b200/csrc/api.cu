// Library-level entry points: version, error text, device check.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "hk_common.cuh"

namespace hk {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(HK_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return HK_OK;
}

bool pdl_enabled() {
  const char* env = getenv("HK_PDL");  // read per launch: A/B runs toggle it
  return !(env && env[0] == '0');
}

int sm_count() {
  static thread_local int cached_dev = -1;
  static thread_local int cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached = n;
    cached_dev = dev;
  }
  return cached;
}

}  // namespace hk

extern "C" {

int hk_version(void) { return HK_ABI_VERSION; }

const char* hk_last_error(void) { return hk::g_err; }

int hk_check_device(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return hk::fail(HK_ERR_CUDA, "cudaGetDevice: %s", cudaGetErrorString(e));
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10 || minor != 0)
    return hk::fail(HK_ERR_UNSUPPORTED_ARCH, "device %d is sm_%d%d; libhulk_sm100 is built for sm_100a only", dev, major, minor);
  return HK_OK;
}

}  // extern "C"
