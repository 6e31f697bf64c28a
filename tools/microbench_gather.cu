// Microbenchmark: random 64-byte row gathers (4 lanes x 16 B per row) vs table footprint and request count.
// Answers, for the roofline of rh_fields_fwd: what does the memory system deliver for random 64 B rows, and where
// does it fall off (TLB reach / DRAM row activations)?   nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ float4 ldg16(const float* p, int hint) {
  float4 r;
  if (hint == 64) {
    asm volatile("ld.global.nc.L1::no_allocate.L2::64B.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  } else if (hint == 128) {
    asm volatile("ld.global.nc.L1::no_allocate.L2::128B.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  } else {
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  }
  return r;
}

template <int U>
__global__ void gather(const float* __restrict__ table, const int* __restrict__ idx, int64_t n_rows_req, float* __restrict__ out, int hint) {
  // thread = (row request, quarter); U requests per thread in flight
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int q = t & 3;
  const int64_t r0 = (t >> 2) * U;
  float4 acc = make_float4(0, 0, 0, 0);
  float4 v[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int64_t r = r0 + u;
    v[u] = make_float4(0, 0, 0, 0);
    if (r < n_rows_req) v[u] = ldg16(table + (int64_t)idx[r] * 16 + q * 4, hint);
  }
#pragma unroll
  for (int u = 0; u < U; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
  if (acc.x == 123456.f) out[t] = acc.x + acc.y + acc.z + acc.w;
}

int main() {
  const int64_t max_rows = (int64_t)128 << 20;  // 128 Mi rows x 64 B = 8 GiB
  float* table;
  cudaMalloc(&table, max_rows * 64);
  cudaMemset(table, 0, max_rows * 64);
  const int64_t pool = 64 << 20;
  int* idx_h = (int*)malloc(pool * sizeof(int));
  int* idx_d;
  cudaMalloc(&idx_d, pool * sizeof(int));
  float* out;
  cudaMalloc(&out, 1 << 20);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int64_t reqs[] = {106496, 1703936, 13631488};
  const int64_t foot[] = {(int64_t)1 << 20, (int64_t)4 << 20, (int64_t)26 << 20, (int64_t)104 << 20};  // rows: 64 MB, 256 MB, 1.66 GB, 6.6 GB
  printf("requests,footprint_MB,unroll,hint,us_per_launch,GBps,Mreq_per_s\n");
  for (int fi = 0; fi < 4; ++fi) {
    uint64_t s = 88172645463325252ull;
    for (int64_t i = 0; i < pool; ++i) {
      s ^= s << 13; s ^= s >> 7; s ^= s << 17;
      idx_h[i] = (int)(s % (uint64_t)foot[fi]);
    }
    cudaMemcpy(idx_d, idx_h, pool * sizeof(int), cudaMemcpyHostToDevice);
    for (int ri = 0; ri < 3; ++ri) {
      for (int U = 1; U <= 4; U *= 4) {
        for (int hint = 0; hint <= 128; hint += 64) {
          const int64_t n = reqs[ri];
          const int64_t threads = ((n + U - 1) / U) * 4;
          const int block = 128;
          const int grid = (int)((threads + block - 1) / block);
          const int iters = 20;
          float best = 1e30f, total = 0;
          for (int it = 0; it < iters; ++it) {
            const int* ip = idx_d + ((int64_t)it * n) % (pool - n);
            cudaEventRecord(e0);
            if (U == 1) gather<1><<<grid, block>>>(table, ip, n, out, hint);
            else gather<4><<<grid, block>>>(table, ip, n, out, hint);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (it >= 2) { total += ms; if (ms < best) best = ms; }
          }
          const float us = total / (iters - 2) * 1000.f;
          printf("%lld,%lld,%d,%d,%.2f,%.1f,%.0f\n", (long long)n, (long long)(foot[fi] * 64 >> 20), U, hint, us, n * 64.0 / us / 1e3, n / us);
        }
      }
    }
  }
  return 0;
}
