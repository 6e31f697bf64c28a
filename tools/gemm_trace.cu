// tools/gemm_trace.cu — where the time of one rh_gemm_tf32x3 launch goes: the production kernel compiled with RH_GEMM_TRACE
// stamps clock64() at its pipeline milestones (per CTA); this driver runs the six tower shapes and prints, per shape, the
// median over CTAs of each interval in cycles and microseconds (SM clock from cudaDevAttrClockRate), next to the launch's
// CUDA-event time.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -DRH_GEMM_TRACE -Iinclude -Itorch-rechub_b200/csrc \
//        -o tools/gemm_trace tools/gemm_trace.cu -lcuda
#include <algorithm>
#include <vector>

#include "../torch-rechub_b200/csrc/rh_api.cu"
#include "../torch-rechub_b200/csrc/rh_gemm.cu"

static const char* kNames[12] = {"entry", "setup done", "TMA first issued", "TMA last issued", "split: first tile landed", "split: first tile split", "MMA: first issue",
                                 "MMA: last commit", "split: last tile split", "epilogue: accumulator ready", "epilogue: stores issued", "exit"};

int main() {
  struct Shape { const char* name; int M, N, K, a_mn, b_mn, split; bool bias; };
  const Shape shapes[] = {{"L1_fwd", 4096, 256, 429, 0, 0, 1, true}, {"L2_fwd", 4096, 128, 256, 0, 0, 1, true}, {"L1_dX", 4096, 429, 256, 0, 1, 1, false},
                          {"L2_dX", 4096, 256, 128, 0, 1, 1, false}, {"L1_dW", 256, 429, 4096, 1, 1, 16, false}, {"L2_dW", 128, 256, 4096, 1, 1, 32, false}};
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  float *A, *B, *C, *bias;
  const size_t big = (size_t)4096 * 448 * 4;
  cudaMalloc(&A, big);
  cudaMalloc(&B, big);
  cudaMalloc(&C, big);
  cudaMalloc(&bias, 4096);
  cudaMemset(A, 0, big);
  cudaMemset(B, 0, big);
  cudaMemset(bias, 0, 4096);
  unsigned long long* trace;
  const int max_ctas = 4096;
  cudaMalloc(&trace, (size_t)max_ctas * 16 * 8);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int tn : {128, 64})
  for (const Shape& s : shapes) {
    rh_gemm_tile_n(tn);
    const int lda = s.a_mn ? (s.M + 3) / 4 * 4 : (s.K + 3) / 4 * 4;
    const int ldb = s.b_mn ? (s.N + 3) / 4 * 4 : (s.K + 3) / 4 * 4;
    const int ldc = (s.N + 3) / 4 * 4;
    g_gemm_trace = nullptr;
    for (int w = 0; w < 3; ++w) rh_gemm_tf32x3(A, lda, s.a_mn, B, ldb, s.b_mn, C, ldc, s.M, s.N, s.K, s.bias ? bias : nullptr, s.split, nullptr);
    cudaDeviceSynchronize();
    cudaMemset(trace, 0, (size_t)max_ctas * 16 * 8);
    g_gemm_trace = trace;
    cudaEventRecord(e0);
    int rc = rh_gemm_tf32x3(A, lda, s.a_mn, B, ldb, s.b_mn, C, ldc, s.M, s.N, s.K, s.bias ? bias : nullptr, s.split, nullptr);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rc != 0) { printf("%s: rc %d %s\n", s.name, rc, rh_last_error()); continue; }
    const int kb = (s.K + 31) / 32;
    int split = s.split < kb ? s.split : kb;
    const int per = (kb + split - 1) / split;
    split = (kb + per - 1) / per;
    const int ctas = ((s.M + 127) / 128) * ((s.N + tn - 1) / tn) * split;
    std::vector<unsigned long long> h((size_t)ctas * 16);
    cudaMemcpy(h.data(), trace, h.size() * 8, cudaMemcpyDeviceToHost);
    printf("%s tile 128x%d  M=%d N=%d K=%d split=%d ctas=%d k-blocks/cta=%d  event time (single launch, idle GPU) %.2f us\n", s.name, tn, s.M, s.N, s.K, split, ctas, per, ms * 1e3);
    for (int ev = 1; ev < 12; ++ev) {
      std::vector<double> d;
      for (int c = 0; c < ctas; ++c) {
        const unsigned long long t0 = h[(size_t)c * 16], t = h[(size_t)c * 16 + ev];
        if (t0 != 0 && t != 0) d.push_back((double)(t - t0));
      }
      if (d.empty()) continue;
      std::sort(d.begin(), d.end());
      printf("   %-30s median +%8.0f cyc (%6.2f us)   max +%8.0f cyc\n", kNames[ev], d[d.size() / 2], d[d.size() / 2] / (khz * 1e-3), d.back());
    }
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
