// rh_common.cuh — shared device/host helpers of the sm_100a engine.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "rechub_b200.h"

namespace rh {

// ---- host-side error plumbing ------------------------------------------------------------------
void set_error(const char* fmt, ...);   // rh_api.cu (thread-local message)

#define RH_REQUIRE(cond, status, ...)            \
  do {                                           \
    if (!(cond)) {                               \
      ::rh::set_error(__VA_ARGS__);              \
      return (status);                           \
    }                                            \
  } while (0)

extern unsigned long long g_launches;  // rh_api.cu: kernels launched by this library (rh_launch_count)
extern int g_pdl;                      // rh_api.cu: launch the hot-path kernels with programmatic dependent launch (rh_set_pdl)
extern int g_carveout;                 // rh_api.cu: preferred shared-memory carveout for every kernel, -1 = the driver's choice (rh_set_smem_carveout)
void note_kernel(const void* fn);      // rh_api.cu

// Launch with the programmatic-stream-serialization attribute when rh_set_pdl(1): the kernel may be scheduled while its predecessor
// in the stream drains; it calls pdl_wait() (griddepcontrol.wait) before its first global access, which returns once the predecessor
// has completed and flushed — so only launch latency and the memory-free prologue overlap.  Captured by CUDA graphs as a
// programmatic edge.  Without the attribute pdl_wait() is a no-op.
#ifndef RH_PDL_FAMILY
#define RH_PDL_FAMILY 128
#endif
template <typename... KArgs, typename... Args>
static inline void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (g_pdl & RH_PDL_FAMILY) ? 1 : 0;
  if (g_carveout >= 0) note_kernel(reinterpret_cast<const void*>(kernel));
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

#define RH_LAUNCH_CHECK()                                                        \
  do {                                                                           \
    ++::rh::g_launches;                                                          \
    cudaError_t e__ = cudaGetLastError();                                        \
    if (e__ != cudaSuccess) {                                                    \
      ::rh::set_error("%s:%d launch failed: %s", __FILE__, __LINE__,             \
                      cudaGetErrorString(e__));                                  \
      return RH_ERR_CUDA;                                                        \
    }                                                                            \
  } while (0)

static inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

static inline int pow2_ceil(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

// ---- device helpers ----------------------------------------------------------------------------
// Programmatic dependent launch, both halves (no-ops for a launch without the attribute):
//   pdl_trigger  "my dependents may be scheduled": once EVERY CTA of this grid has said so (i.e. has started) the next kernel's
//                CTAs take whatever SM resources are free and run their memory-free prologue up to their own pdl_wait.  Issued at
//                the top of every hot-path kernel: without it the dependent is only scheduled when this grid has drained, and the
//                attribute buys nothing (first PDL A/B of round 2: 24.03 vs 24.46 M samples/s).
//   pdl_wait     returns once the prerequisite grid has COMPLETED and its memory is visible; every kernel calls it before its
//                first global access, so triggering early can never expose unfinished data (and completion is transitive: a grid
//                only completes after its own pdl_wait returned).
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() {
  pdl_trigger();
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

// 128-bit read-only load that does not allocate in L1: embedding rows are touched once per launch.
__device__ __forceinline__ float4 ldg_row16(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

// 128-bit streaming store (tile rows are consumed by the next kernel out of L2, never by this SM).
__device__ __forceinline__ void stg_row16(float* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}

// 128-bit vector reduction into global memory (REDG.E.ADD.F32x4 on sm_90+): one L2 atomic
// transaction per 16 B instead of four.
__device__ __forceinline__ void red_add_row16(float* p, const float4& v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

// system-scope release / acquire on a flag word in peer-mapped memory (the cross-GPU hand-overs of rh_dist.cu and rh_fields.cu)
__device__ __forceinline__ void st_release_sys(int32_t* p, int32_t v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int32_t ld_acquire_sys(const int32_t* p) {
  int32_t v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ int64_t load_id(const void* ids, int64_t idx, bool is_i32) {
  return is_i32 ? (int64_t)__ldg(reinterpret_cast<const int32_t*>(ids) + idx)
                : (int64_t)__ldg(reinterpret_cast<const long long*>(ids) + idx);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float4 f4_zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float4 f4_add(const float4& a, const float4& b) {
  return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}
__device__ __forceinline__ float4 f4_fma(const float4& a, const float4& b, const float4& c) {
  return make_float4(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y), fmaf(a.z, b.z, c.z), fmaf(a.w, b.w, c.w));
}
__device__ __forceinline__ float4 f4_scale(const float4& a, float s) {
  return make_float4(a.x * s, a.y * s, a.z * s, a.w * s);
}
__device__ __forceinline__ float f4_dot(const float4& a, const float4& b) {
  return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
}

}  // namespace rh
