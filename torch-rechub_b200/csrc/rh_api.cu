// rh_api.cu — ABI bookkeeping: version + thread-local error message.
#include <stdarg.h>

#include <mutex>
#include <unordered_set>

#ifndef RH_PDL_FAMILY  // (the trace tools include several of these files into one unit: the first one names the family)
#define RH_PDL_FAMILY 32  /* rh_set_pdl mask bit of this file's kernels */
#endif
#include "rh_common.cuh"

namespace rh {
static thread_local char g_err[512] = "";
unsigned long long g_launches = 0;
int g_pdl = 0;
int g_carveout = -1;

// rh_set_smem_carveout: apply the preferred shared-memory carveout to every kernel of the library, once per kernel, at its first
// launch after the setting (an experiment switch: does a uniform carveout remove an L1 / shared-memory reconfiguration between the
// 198 KB GEMM CTAs and their small-shared-memory neighbours?)
void note_kernel(const void* fn) {
  static std::mutex mu;
  static std::unordered_set<const void*> seen;
  std::lock_guard<std::mutex> lock(mu);
  if (seen.insert(fn).second) {
    cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, g_carveout);
    cudaGetLastError();
  }
}

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
extern "C" int rh_set_smem_carveout(int percent) {
  const int old = rh::g_carveout;
  if (percent >= -1 && percent <= 100) rh::g_carveout = percent;
  return old;
}
extern "C" int rh_set_pdl(int on) {
  const int old = rh::g_pdl;
  if (on >= 0) rh::g_pdl = on;  // a mask of kernel families (include/rechub_b200.h); 1 is NOT "all": use 0xff
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
  // ONE index space over all segments (16-byte units): a thread's loads of different segments are independent of its stores, so
  // the whole copy is a single load -> store round trip.  (Segment after segment, the stores of one — which may alias the next
  // one's source as far as the compiler knows — serialised three round trips: 4.9 us for 1 MB.)
  const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nt = (int64_t)gridDim.x * blockDim.x;
  int64_t first[5];
  bool v16 = true;
  first[0] = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int64_t b = k < s.n ? s.bytes[k] : 0;
    first[k + 1] = first[k] + ((b + 15) >> 4);
    if (k < s.n) v16 &= ((reinterpret_cast<uintptr_t>(s.dst[k]) | reinterpret_cast<uintptr_t>(s.src[k]) | (uintptr_t)b) & 15) == 0;
  }
  if (v16) {
    constexpr int U = 4;
    for (int64_t i0 = t0; i0 < first[4]; i0 += nt * U) {
      int4 v[U];
      int4* d[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t i = i0 + u * nt;
        d[u] = nullptr;
        if (i < first[4]) {
          const int k = (i >= first[1]) + (i >= first[2]) + (i >= first[3]);
          v[u] = __ldg(static_cast<const int4*>(s.src[k]) + (i - first[k]));
          d[u] = static_cast<int4*>(s.dst[k]) + (i - first[k]);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (d[u] != nullptr) *d[u] = v[u];
    }
    return;
  }
  for (int k = 0; k < s.n; ++k) {  // a ragged last batch: 4-byte words where possible, bytes otherwise
    if (((reinterpret_cast<uintptr_t>(s.dst[k]) | reinterpret_cast<uintptr_t>(s.src[k]) | (uintptr_t)s.bytes[k]) & 3) == 0) {
      const int32_t* sw = static_cast<const int32_t*>(s.src[k]);
      int32_t* dw = static_cast<int32_t*>(s.dst[k]);
      for (int64_t i = t0; i < (s.bytes[k] >> 2); i += nt) dw[i] = __ldg(sw + i);
    } else {
      const char* sc = static_cast<const char*>(s.src[k]);
      char* dc = static_cast<char*>(s.dst[k]);
      for (int64_t i = t0; i < s.bytes[k]; i += nt) dc[i] = sc[i];
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
