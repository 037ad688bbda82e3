// api.cu — error plumbing and library identification.
#include <stdarg.h>
#include <stdlib.h>
#include <atomic>
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

// ---- options: one atomic per key; the environment is read exactly once (thread-safe function-local static)
namespace {
struct Options {
  std::atomic<int32_t> v[GCD_OPT_COUNT_];
  Options() {
    static const struct { const char* env; int32_t dflt; } kInit[GCD_OPT_COUNT_] = {
        {"GCD_PAIRS_FUSED", 2}, {"GCD_GATHER_FLAT", 1}, {"GCD_TC_STAGES", 0}, {"GCD_TC_GROUP", 0}, {"GCD_WG_CHUNK_MIN", 0}, {"GCD_TC_WARPS", 8}, {"GCD_BN_FUSED", 1}, {"GCD_KMAP_COOP", 0}, {"GCD_PDL", 1}, {"GCD_WGRAD_SIDE", 1}, {"GCD_DYN_TILES", 0}, {"GCD_DYN_AHEAD", 0}, {"GCD_BN_MASK_FROM_X", 1}, {"GCD_SCAN_LOOKBACK", 1}};
    for (int i = 0; i < GCD_OPT_COUNT_; ++i) {
      const char* e = getenv(kInit[i].env);
      v[i].store(e ? atoi(e) : kInit[i].dflt, std::memory_order_relaxed);
    }
  }
};
Options& options() { static Options o; return o; }
}  // namespace
int32_t option(int32_t key) { return options().v[key].load(std::memory_order_relaxed); }
}  // namespace gcd

extern "C" int32_t gcd_set_option(int32_t option, int32_t value) {
  GCD_REQUIRE(option >= 0 && option < GCD_OPT_COUNT_, "gcd_set_option: unknown option %d", option);
  gcd::options().v[option].store(value, std::memory_order_relaxed);
  return GCD_OK;
}
extern "C" int32_t gcd_get_option(int32_t option) {
  if (option < 0 || option >= GCD_OPT_COUNT_) return 0;
  return gcd::option(option);
}

extern "C" const char* gcd_last_error_string(void) { return gcd::g_err; }
extern "C" int32_t gcd_abi_version(void) { return 1; }
extern "C" int32_t gcd_has_tcgen05(void) {
#ifdef GCD_WITH_TCGEN05
  return 1;
#else
  return 0;
#endif
}
