// tools/bnfuse_trace.cu — where the time of rh_bn_act_fused_fwd goes (clock64 stamps per CTA, RH_BN_TRACE build of the production kernel).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -DRH_BN_TRACE -Iinclude -Itorch-rechub_b200/csrc -o tools/bnfuse_trace tools/bnfuse_trace.cu
#include <algorithm>
#include <vector>

#include "../torch-rechub_b200/csrc/rh_api.cu"
#include "../torch-rechub_b200/csrc/rh_bnfuse.cu"

int main() {
  const char* names[6] = {"entry", "rows loaded + block reduce in smem", "column atomics issued", "grid barrier passed", "phase 2 stored", "exit (last CTA: finalise)"};
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  for (int cols : {256, 128}) {
    for (float pdrop : {0.f, 0.2f}) {
      const int64_t rows = 4096;
      float *h, *y, *gamma, *beta, *stats, *scratch, *rm, *rv;
      long long* nbt;
      cudaMalloc(&h, rows * cols * 4); cudaMalloc(&y, rows * cols * 4); cudaMalloc(&gamma, cols * 4); cudaMalloc(&beta, cols * 4);
      cudaMalloc(&stats, (2 * cols + 1) * 4); cudaMalloc(&scratch, rh_bn_fused_scratch_floats(cols) * 4); cudaMalloc(&rm, cols * 4); cudaMalloc(&rv, cols * 4); cudaMalloc(&nbt, 8);
      cudaMemset(h, 0, rows * cols * 4); cudaMemset(gamma, 0, cols * 4); cudaMemset(beta, 0, cols * 4); cudaMemset(scratch, 0, rh_bn_fused_scratch_floats(cols) * 4);
      cudaMemset(rm, 0, cols * 4); cudaMemset(rv, 0, cols * 4); cudaMemset(nbt, 0, 8);
      unsigned long long* trace;
      cudaMalloc(&trace, 148 * 8 * 8);
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      g_bn_trace = nullptr;
      for (int w = 0; w < 3; ++w)
        rh_bn_act_fused_fwd(h, cols, rows, cols, 1e-5f, gamma, beta, 1, nullptr, 1e-3f, pdrop, 7u, rm, rv, (int64_t*)nbt, 0.1f, stats, scratch, y, cols, nullptr, nullptr, nullptr, nullptr, 0, nullptr, nullptr);
      cudaDeviceSynchronize();
      cudaMemset(trace, 0, 148 * 8 * 8);
      g_bn_trace = trace;
      cudaEventRecord(e0);
      int rc = rh_bn_act_fused_fwd(h, cols, rows, cols, 1e-5f, gamma, beta, 1, nullptr, 1e-3f, pdrop, 7u, rm, rv, (int64_t*)nbt, 0.1f, stats, scratch, y, cols, nullptr, nullptr, nullptr, nullptr, 0, nullptr, nullptr);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      const int ctas = (int)((rows + 31) / 32);  // 16 warps x 2 rows per CTA
      std::vector<unsigned long long> t((size_t)148 * 8);
      cudaMemcpy(t.data(), trace, t.size() * 8, cudaMemcpyDeviceToHost);
      printf("bn_fused_fwd rows=%lld cols=%d p_drop=%.1f rc=%d ctas=%d event time %.2f us\n", (long long)rows, cols, pdrop, rc, ctas, ms * 1e3);
      for (int ev = 1; ev < 6; ++ev) {
        std::vector<double> d;
        for (int c = 0; c < ctas; ++c)
          if (t[(size_t)c * 8] && t[(size_t)c * 8 + ev]) d.push_back((double)(t[(size_t)c * 8 + ev] - t[(size_t)c * 8]));
        if (d.empty()) continue;
        std::sort(d.begin(), d.end());
        printf("   %-40s median +%8.0f cyc (%6.2f us)  max +%8.0f\n", names[ev], d[d.size() / 2], d[d.size() / 2] / (khz * 1e-3), d.back());
      }
    }
  }
  return 0;
}
