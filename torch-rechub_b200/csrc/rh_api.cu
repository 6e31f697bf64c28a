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

// ---- one launch copying up to four contiguous device buffers (a batch's id block, numeric block, sequence block and labels into the
// captured step's static buffers: trainers/ctr_trainer.py:84 moves one tensor per column; packed batches move <= 4, and at ~1 MB each
// of those copies is launch latency) ------------------------------------------------------------------------------------------------
namespace rh {
struct CopySegs {
  void* dst[4];
  const void* src[4];
  int64_t bytes[4];
  int n;
};
__global__ void __launch_bounds__(256) copy_segs_kernel(CopySegs s) {
  pdl_wait();
  const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nt = (int64_t)gridDim.x * blockDim.x;
  for (int k = 0; k < s.n; ++k) {
    const int64_t b = s.bytes[k];
    const bool v16 = ((reinterpret_cast<uintptr_t>(s.dst[k]) | reinterpret_cast<uintptr_t>(s.src[k])) & 15) == 0;
    if (v16) {
      const int4* src = static_cast<const int4*>(s.src[k]);
      int4* dst = static_cast<int4*>(s.dst[k]);
      const int64_t n16 = b >> 4;
      for (int64_t i = t0; i < n16; i += nt) dst[i] = __ldg(src + i);
      const char* sc = static_cast<const char*>(s.src[k]);
      char* dc = static_cast<char*>(s.dst[k]);
      for (int64_t i = (n16 << 4) + t0; i < b; i += nt) dc[i] = sc[i];
    } else {
      const char* sc = static_cast<const char*>(s.src[k]);
      char* dc = static_cast<char*>(s.dst[k]);
      for (int64_t i = t0; i < b; i += nt) dc[i] = sc[i];
    }
  }
}
}  // namespace rh

extern "C" int rh_copy_segments(int n, void* const* dst, const void* const* src, const int64_t* bytes, void* stream) {
  RH_REQUIRE(n >= 1 && n <= 4 && dst && src && bytes, RH_ERR_INVALID_ARG, "rh_copy_segments: 1..4 segments");
  rh::CopySegs s{};
  s.n = n;
  int64_t total = 0;
  for (int k = 0; k < n; ++k) {
    RH_REQUIRE(dst[k] && src[k] && bytes[k] >= 0, RH_ERR_INVALID_ARG, "rh_copy_segments: null segment");
    s.dst[k] = dst[k], s.src[k] = src[k], s.bytes[k] = bytes[k];
    total += bytes[k];
  }
  if (total == 0) return RH_OK;
  int64_t g = (total / 16 + 255) / 256;
  g = g < 1 ? 1 : (g > 148 * 4 ? 148 * 4 : g);
  launch_k(rh::copy_segs_kernel, dim3((unsigned)g), dim3(256), 0, (cudaStream_t)stream, s);
  RH_LAUNCH_CHECK();
  return RH_OK;
}
