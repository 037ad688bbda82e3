// api.cu — error plumbing and library identification.
#include <stdarg.h>
#include "common.cuh"

namespace gcd {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int32_t cuda_fail(cudaError_t e, const char* what) {
  set_error("%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
  return GCD_ERR_CUDA;
}
}  // namespace gcd

extern "C" const char* gcd_last_error_string(void) { return gcd::g_err; }
extern "C" int32_t gcd_abi_version(void) { return 1; }
extern "C" int32_t gcd_has_tcgen05(void) {
#ifdef GCD_WITH_TCGEN05
  return 1;
#else
  return 0;
#endif
}
