// errors.cu -- thread-local error message + misc ABI entry points.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace efm {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("EFM_PDL"); v = (e && e[0] == '0') ? 0 : 1; }
  return v == 1;
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return EFM_E_CUDA;
}
}  // namespace efm

extern "C" const char* efm_last_error(void) { return efm::g_err; }
extern "C" int efm_abi_version(void) { return EFM_ABI_VERSION; }
extern "C" int efm_device_arch(void) {
  int dev = 0, major = 0, minor = 0;
  EFM_CUDA(cudaGetDevice(&dev));
  EFM_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  EFM_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  return major * 10 + minor;
}
