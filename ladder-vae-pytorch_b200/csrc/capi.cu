// C-ABI plumbing shared by all kernels: error string, launch counter, device probe.
#include "common.cuh"
#include <stdarg.h>
#include <string.h>

static thread_local char g_err[512] = "";
unsigned long long g_lvae_launches = 0;
int g_lvae_pdl = 1;

void lvae_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

LVAE_API const char* lvae_last_error(void) { return g_err; }
LVAE_API int lvae_abi_version(void) { return 1; }
// 1 (default): launch every kernel with programmatic dependent launch allowed; 0: plain stream order
LVAE_API void lvae_set_pdl(int enabled) { g_lvae_pdl = enabled ? 1 : 0; }
LVAE_API int lvae_get_pdl(void) { return g_lvae_pdl; }
LVAE_API unsigned long long lvae_launch_count(void) { return g_lvae_launches; }
LVAE_API void lvae_reset_launch_count(void) { g_lvae_launches = 0; }

// 0 when a CUDA device with compute capability 10.x is current; the product refuses to run otherwise
LVAE_API int lvae_device_check(void) {
  int dev = 0, major = 0, minor = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    lvae_set_error("no CUDA device: %s", cudaGetErrorString(e));
    return LVAE_ERR_CUDA;
  }
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10) {
    lvae_set_error("lvae_b200 kernels are built for sm_100a only; device is sm_%d%d", major, minor);
    return LVAE_ERR_UNSUPPORTED;
  }
  return LVAE_OK;
}
