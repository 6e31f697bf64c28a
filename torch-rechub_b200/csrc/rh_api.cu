// rh_api.cu — ABI bookkeeping: version + thread-local error message.
#include <stdarg.h>

#include "rh_common.cuh"

namespace rh {
static thread_local char g_err[512] = "";
unsigned long long g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace rh

extern "C" int rh_abi_version(void) { return RH_ABI_VERSION; }
extern "C" const char* rh_last_error(void) { return rh::g_err; }
extern "C" unsigned long long rh_launch_count(void) { return rh::g_launches; }
