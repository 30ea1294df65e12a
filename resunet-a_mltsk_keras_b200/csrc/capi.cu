// capi.cu — version / error plumbing of the C-ABI.
#include <stdarg.h>
#include "common.cuh"

static thread_local char g_err[512] = "";

void rsa_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* rsa_version(void) { return "resuneta-b200 0.1 (sm_100a)"; }
extern "C" const char* rsa_last_error(void) { return g_err; }

extern "C" int rsa_device_check(void) {
  int dev = 0;
  cudaDeviceProp prop;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
    rsa_set_error("no CUDA device");
    return RSA_ERR_CUDA;
  }
  if (prop.major != 10) {
    rsa_set_error("libresuneta is built for sm_100a (B200) only; found sm_%d%d", prop.major, prop.minor);
    return RSA_ERR_ARCH;
  }
  return RSA_OK;
}
