// Microbenchmark of the memory system under the access pattern of rh_fields_fwd: random 64-byte table rows (4 lanes x 16 B per
// row), ids read first (a dependent load), rows optionally written back as a contiguous tile.
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/microbench_gather tools/microbench_gather.cu
//   tools/microbench_gather            CSV on stdout: every variant timed as a CUDA-graph replay of kLaunches launches with
//                                      CUDA events around the replay, next to a NULL kernel of the same grid (the launch /
//                                      ramp cost that is not memory time) -> us_per_launch, us_null, us_net = difference
//   tools/microbench_gather ncu        one plain launch per variant, for `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum`
//                                      (read and write DRAM bytes per launch, separately)
//
// Questions it answers: (1) what a 106 496-row launch (batch 4096 x 26 fields) costs beyond an empty launch;
// (2) whether the DRAM read over-fetch of 64-B rows (ncu: 128 B fetched per row) moves with cudaLimitMaxL2FetchGranularity or
// with the ld.global .L2::64B/.L2::128B qualifiers; (3) what the same rows cost when they are L2-resident; (4) the sustained
// random-row rate at large request counts.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static const int kLaunches = 20;

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

// thread = (row request, 16-byte quarter); U requests per thread in flight; ids are a dependent load (int32 or int64);
// row_floats = 16 (64-B rows) or 32 (128-B records of which the first 64 B are read)
template <int U, bool I64>
__global__ void gather(const float* __restrict__ table, const void* __restrict__ idx, int64_t n, float* __restrict__ tile, int hint, int row_floats,
                       int write_tile) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int q = t & 3;
  const int64_t r0 = (t >> 2) * U;
  int64_t id[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int64_t r = r0 + u;
    id[u] = -1;
    if (r < n) id[u] = I64 ? (int64_t)__ldg(reinterpret_cast<const long long*>(idx) + r) : (int64_t)__ldg(reinterpret_cast<const int*>(idx) + r);
  }
  float4 v[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    v[u] = make_float4(0, 0, 0, 0);
    if (id[u] >= 0) v[u] = ldg16(table + id[u] * row_floats + q * 4, hint);
  }
  float4 acc = make_float4(0, 0, 0, 0);
#pragma unroll
  for (int u = 0; u < U; ++u) {
    if (write_tile && id[u] >= 0) {
      asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(tile + (r0 + u) * 16 + q * 4), "f"(v[u].x), "f"(v[u].y), "f"(v[u].z),
                   "f"(v[u].w)
                   : "memory");
    }
    acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w;
  }
  if (acc.x == 123456.f) tile[t] = acc.x + acc.y + acc.z + acc.w;
}

__global__ void null_kernel(float* out) {
  if (out == nullptr && threadIdx.x == 9999) out[0] = 0.f;
}

struct Variant {
  int64_t n;       // row requests per launch
  int64_t rows;    // table footprint in rows
  int U, i64, hint, row_floats, write_tile;
};

static void launch(const Variant& v, const float* table, const void* idx, float* tile, cudaStream_t st, bool null) {
  const int64_t threads = ((v.n + v.U - 1) / v.U) * 4;
  const int block = 128;
  const int grid = (int)((threads + block - 1) / block);
  if (null) {
    null_kernel<<<grid, block, 0, st>>>(tile);
    return;
  }
#define GO(UU, II) gather<UU, II><<<grid, block, 0, st>>>(table, idx, v.n, tile, v.hint, v.row_floats, v.write_tile)
  if (v.U == 1 && v.i64) GO(1, true);
  else if (v.U == 1) GO(1, false);
  else if (v.i64) GO(4, true);
  else GO(4, false);
#undef GO
}

int main(int argc, char** argv) {
  const bool ncu_mode = argc > 1 && strcmp(argv[1], "ncu") == 0;
  const int64_t max_rows = (int64_t)104 << 20;  // 104 Mi rows x 64 B = 6.6 GB
  float* table;
  if (cudaMalloc(&table, max_rows * 64) != cudaSuccess) { fprintf(stderr, "cudaMalloc failed\n"); return 1; }
  cudaMemset(table, 0, max_rows * 64);
  const int64_t pool = 32 << 20;
  long long* idx_h = (long long*)malloc(pool * sizeof(long long));
  int* idx32_h = (int*)malloc(pool * sizeof(int));
  long long* idx64_d;
  int* idx32_d;
  cudaMalloc(&idx64_d, pool * sizeof(long long));
  cudaMalloc(&idx32_d, pool * sizeof(int));
  float* tile;
  cudaMalloc(&tile, (int64_t)13631488 * 64 + 4096);
  cudaStream_t st;
  cudaStreamCreate(&st);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  size_t gran0 = 0;
  cudaDeviceGetLimit(&gran0, cudaLimitMaxL2FetchGranularity);

  const int64_t foots[] = {(int64_t)1 << 20, (int64_t)26 << 20, (int64_t)104 << 20};  // rows: 64 MB (L2-resident), 1.66 GB, 6.6 GB
  const int64_t reqs[] = {106496, 1703936, 13631488};                                 // batch 4096 / 65536 / 524288 x 26 fields
  const int grans[] = {0, 32, 64, 128};                                               // 0 = leave the driver default
  printf("requests,footprint_MB,row_bytes,ids,unroll,ld_hint,l2_fetch_granularity,write_tile,us_per_launch,us_null,us_net,GBps_net_algorithmic,Mrows_per_s_net\n");
  for (int fi = 0; fi < 3; ++fi) {
    uint64_t s = 88172645463325252ull;
    for (int64_t i = 0; i < pool; ++i) {
      s ^= s << 13; s ^= s >> 7; s ^= s << 17;
      idx_h[i] = (long long)(s % (uint64_t)foots[fi]);
      idx32_h[i] = (int)idx_h[i];
    }
    cudaMemcpy(idx64_d, idx_h, pool * sizeof(long long), cudaMemcpyHostToDevice);
    cudaMemcpy(idx32_d, idx32_h, pool * sizeof(int), cudaMemcpyHostToDevice);
    for (int ri = 0; ri < 3; ++ri) {
      for (int gi = 0; gi < 4; ++gi) {
        if (grans[gi] != 0) cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)grans[gi]);
        else cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran0);
        size_t gran = 0;
        cudaDeviceGetLimit(&gran, cudaLimitMaxL2FetchGranularity);
        // (U, i64, hint, row_floats, write_tile)
        const int cfg[][5] = {{1, 1, 0, 16, 1}, {1, 0, 0, 16, 1}, {4, 1, 0, 16, 1}, {1, 1, 64, 16, 1}, {1, 1, 128, 16, 1}, {1, 1, 0, 16, 0}, {1, 1, 0, 32, 1}};
        for (int ci = 0; ci < 7; ++ci) {
          if (gi != 0 && ci != 0 && ci != 5) continue;  // the granularity sweep on the two base variants only
          Variant v = {reqs[ri], foots[fi], cfg[ci][0], cfg[ci][1], cfg[ci][2], cfg[ci][3], cfg[ci][4]};
          if (v.row_floats == 32 && foots[fi] * 2 > max_rows) continue;  // 128-B records need twice the footprint
          const void* ib = v.i64 ? (const void*)idx64_d : (const void*)idx32_d;
          const size_t isz = v.i64 ? 8 : 4;
          if (ncu_mode) {
            if (ri != 0 || fi != 1) continue;
            launch(v, table, ib, tile, st, false);
            cudaStreamSynchronize(st);
            continue;
          }
          float us[2];
          for (int null = 0; null < 2; ++null) {
            cudaGraph_t g;
            cudaGraphExec_t ge;
            cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal);
            for (int it = 0; it < kLaunches; ++it) {
              const int64_t off = ((int64_t)it * v.n) % (pool - v.n);  // every launch of the replay gathers different rows
              launch(v, table, (const char*)ib + off * isz, tile, st, null != 0);
            }
            cudaStreamEndCapture(st, &g);
            cudaGraphInstantiate(&ge, g, 0);
            for (int w = 0; w < 3; ++w) cudaGraphLaunch(ge, st);
            cudaStreamSynchronize(st);
            float best = 1e30f;
            for (int rep = 0; rep < 5; ++rep) {
              cudaEventRecord(e0, st);
              cudaGraphLaunch(ge, st);
              cudaEventRecord(e1, st);
              cudaEventSynchronize(e1);
              float ms;
              cudaEventElapsedTime(&ms, e0, e1);
              if (ms < best) best = ms;
            }
            us[null] = best * 1000.f / kLaunches;
            cudaGraphExecDestroy(ge);
            cudaGraphDestroy(g);
          }
          const float net = us[0] - us[1];
          const double bytes = (double)v.n * (64.0 + isz + (v.write_tile ? 64.0 : 0.0));
          printf("%lld,%lld,%d,%s,%d,%d,%zu,%d,%.2f,%.2f,%.2f,%.1f,%.0f\n", (long long)v.n, (long long)(foots[fi] * 64 >> 20), v.row_floats * 4, v.i64 ? "i64" : "i32", v.U,
                 v.hint, gran, v.write_tile, us[0], us[1], net, bytes / net / 1e3, v.n / net);
          fflush(stdout);
        }
      }
    }
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { fprintf(stderr, "CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
