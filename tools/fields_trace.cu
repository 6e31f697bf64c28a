// tools/fields_trace.cu — where the time of one rh_fields_fwd launch goes at the headline shape (B = 4096, 26 fields x 1M rows x 16,
// 13 numeric columns): (1) the launch's time inside a CUDA graph of 20 launches over rotating id sets (tables 1.66 GB >> L2), for
// the full kernel and with parts switched off (no FM/LR, no numeric columns, no tile), next to an empty kernel of the same grid and
// parameter block; (2) clock64 stamps of thread 0 of every block (RH_FIELDS_TRACE build of the production kernel).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -DRH_FIELDS_TRACE -Iinclude -Itorch-rechub_b200/csrc -o tools/fields_trace tools/fields_trace.cu
#include <algorithm>
#include <random>
#include <vector>

#include "../torch-rechub_b200/csrc/rh_api.cu"
#include "../torch-rechub_b200/csrc/rh_fields.cu"

__global__ void __launch_bounds__(256) empty_kernel(const __grid_constant__ rh::FwdParams p) {
  if (p.batch < 0) p.tile[threadIdx.x] = 0.f;
}

static const int B = 4096, F = 26, D = 16, ND = 13, VOCAB = 1000000, SETS = 8, TILE_LD = 432;

int main() {
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  std::vector<float*> tables(F);
  for (int f = 0; f < F; ++f) {
    cudaMalloc(&tables[f], (size_t)VOCAB * D * 4);
    cudaMemset(tables[f], 0, (size_t)VOCAB * D * 4);
  }
  std::mt19937_64 rng(1);
  std::vector<int64_t*> ids(SETS);
  std::vector<float*> nums(SETS);
  for (int s = 0; s < SETS; ++s) {
    std::vector<int64_t> h((size_t)B * F);
    for (auto& v : h) v = (int64_t)(rng() % VOCAB);
    cudaMalloc(&ids[s], h.size() * 8);
    cudaMemcpy(ids[s], h.data(), h.size() * 8, cudaMemcpyHostToDevice);
    cudaMalloc(&nums[s], (size_t)B * ND * 4);
    cudaMemset(nums[s], 0, (size_t)B * ND * 4);
  }
  float *tile, *lrw, *lrb, *yfm, *ylr, *fsum;
  int32_t* err;
  cudaMalloc(&tile, (size_t)B * TILE_LD * 4); cudaMalloc(&lrw, F * D * 4); cudaMalloc(&lrb, 4); cudaMalloc(&yfm, B * 4); cudaMalloc(&ylr, B * 4);
  cudaMalloc(&fsum, (size_t)B * D * 4); cudaMalloc(&err, 4);
  cudaMemset(lrw, 0, F * D * 4); cudaMemset(lrb, 0, 4); cudaMemset(err, 0, 4);
  unsigned long long* trace;
  cudaMalloc(&trace, (size_t)B * 16 * 8);
  cudaStream_t st;
  cudaStreamCreate(&st);

  auto launch = [&](int set, int variant) -> int {
    rh_field fl[F];
    rh_dense dn[ND];
    for (int f = 0; f < F; ++f) {
      fl[f] = rh_field{tables[f], nullptr, ids[set] + f, F, 0, VOCAB, -1, variant == 3 ? -1 : f * D, variant == 1 ? -1 : f};
    }
    for (int j = 0; j < ND; ++j) dn[j] = rh_dense{nums[set] + j, ND, 0, 1, F * D + j};
    const bool fm = variant != 1;
    return rh_fields_fwd(fl, F, D, dn, (variant == 2 || variant == 3) ? 0 : ND, B, variant == 3 ? nullptr : tile, TILE_LD, fm ? lrw : nullptr, fm ? lrb : nullptr, fm ? yfm : nullptr,
                         fm ? ylr : nullptr, fm ? fsum : nullptr, err, st);
  };
  const char* vname[5] = {"full (tile + numeric columns + FM + LR + field_sum)", "no FM / LR / field_sum", "no numeric columns", "FM / LR only (no tile, no numeric columns)", "empty kernel, same grid + parameter block"};
  for (int variant = 0; variant < 5; ++variant) {
    g_fields_trace = nullptr;
    static rh::FwdParams ep;
    for (int w = 0; w < 3; ++w) {
      if (variant < 4) launch(w, variant);
      else empty_kernel<<<B / 8, 256, 0, st>>>(ep);
    }
    cudaStreamSynchronize(st);
    cudaGraph_t g;
    cudaGraphExec_t ge;
    cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal);
    for (int i = 0; i < 20; ++i) {
      if (variant < 4) launch(i % SETS, variant);
      else empty_kernel<<<B / 8, 256, 0, st>>>(ep);
    }
    cudaStreamEndCapture(st, &g);
    cudaGraphInstantiate(&ge, g, 0);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaGraphLaunch(ge, st);
    cudaStreamSynchronize(st);
    float best = 1e9f, sum = 0;
    for (int r = 0; r < 10; ++r) {
      cudaEventRecord(e0, st);
      cudaGraphLaunch(ge, st);
      cudaEventRecord(e1, st);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      best = std::min(best, ms);
      sum += ms;
    }
    printf("%-60s  %6.2f us / launch (mean of 10 graph replays x 20 launches; best %.2f)   err=%s\n", vname[variant], sum / 10 / 20 * 1e3, best / 20 * 1e3, cudaGetErrorString(cudaGetLastError()));
  }
  // stamps
  const char* names[7] = {"entry", "ids of the warp's 4 fields loaded", "4 rows loaded", "tile stored, FM/LR partials accumulated", "numeric columns done", "block barrier passed", "exit"};
  cudaMemset(trace, 0, (size_t)B * 16 * 8);
  g_fields_trace = trace;
  launch(5, 0);
  cudaStreamSynchronize(st);
  g_fields_trace = nullptr;
  const int blocks = B / 8;
  std::vector<unsigned long long> t((size_t)blocks * 16);
  cudaMemcpy(t.data(), trace, t.size() * 8, cudaMemcpyDeviceToHost);
  printf("stamps of thread 0 per block (%d blocks), cycles after the block's own entry:\n", blocks);
  for (int ev = 1; ev < 7; ++ev) {
    std::vector<double> d;
    for (int c = 0; c < blocks; ++c) d.push_back((double)(t[(size_t)c * 16 + ev] - t[(size_t)c * 16]));
    std::sort(d.begin(), d.end());
    printf("   %-44s median +%7.0f cyc (%5.2f us)   p90 +%7.0f   max +%7.0f\n", names[ev], d[d.size() / 2], d[d.size() / 2] / (khz * 1e-3), d[d.size() * 9 / 10], d.back());
  }
  {
    unsigned long long g0 = ~0ull, g1 = 0;
    for (int c = 0; c < blocks; ++c) {
      g0 = std::min(g0, t[(size_t)c * 16 + 8]);
      g1 = std::max(g1, t[(size_t)c * 16 + 9]);
    }
    std::vector<double> d, e;
    for (int c = 0; c < blocks; ++c) {
      d.push_back((double)(t[(size_t)c * 16 + 8] - g0));
      e.push_back((double)(t[(size_t)c * 16 + 9] - g0));
    }
    std::sort(d.begin(), d.end());
    std::sort(e.begin(), e.end());
    printf("   globaltimer, ns after the first block's entry:  block entry median %.0f  p90 %.0f  max %.0f;   block exit median %.0f  p90 %.0f  max %.0f\n", d[d.size() / 2], d[d.size() * 9 / 10],
           d.back(), e[e.size() / 2], e[e.size() * 9 / 10], e.back());
  }
  return 0;
}
