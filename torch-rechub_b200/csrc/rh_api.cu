// rh_api.cu — ABI bookkeeping: version + thread-local error message.
#include <stdarg.h>

#include "rh_common.cuh"

namespace rh {
static thread_local char g_err[512] = "";
unsigned long long g_launches = 0;
int g_pdl = 0;

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
extern "C" int rh_set_pdl(int on) {
  const int old = rh::g_pdl;
  if (on >= 0) rh::g_pdl = on != 0;
  return old;
}

// cudaLimitMaxL2FetchGranularity of the current device: how many bytes L2 pulls from DRAM on a sector miss (32 / 64 / 128).
// Random 64-byte table rows are the engine's dominant access: at 128 bytes every row costs a second, never-used sector pair
// (measured with tools/microbench_gather.cu + ncu: 14.4 MB read for 7.7 MB of rows + ids at 128, 7.7 MB at 64 and at 32).
// `set_bytes` > 0 sets the limit first (a driver hint; the value in force is returned either way, negative on a CUDA error).
extern "C" int rh_l2_fetch_granularity(int set_bytes) {
  if (set_bytes > 0) {
    cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)set_bytes);
    if (e != cudaSuccess) {
      rh::set_error("cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, %d): %s", set_bytes, cudaGetErrorString(e));
      cudaGetLastError();
      return -1;
    }
  }
  size_t v = 0;
  cudaError_t e = cudaDeviceGetLimit(&v, cudaLimitMaxL2FetchGranularity);
  if (e != cudaSuccess) {
    rh::set_error("cudaDeviceGetLimit(cudaLimitMaxL2FetchGranularity): %s", cudaGetErrorString(e));
    cudaGetLastError();
    return -1;
  }
  return (int)v;
}
