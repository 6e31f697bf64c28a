// rh_head.cu — the output head of a ranking tower: Linear(K -> 1) + per-sample side terms + sigmoid, one pass each way.
//
// Reference arithmetic replaced: the MLP's output layer nn.Linear(input_dim, 1) (basic/layers.py:279-280) followed by
// the model's tail, e.g. DeepFM  y = y_linear + y_fm + y_deep; sigmoid(y.squeeze(1))  (models/ranking/deepfm.py:41-43).
// Stock PyTorch runs this as gemv + add + add + sigmoid forward and sigmoid_backward + an outer product + a
// transposed gemv (+ split-K reduce) + a bias reduction backward — ten launches moving a (batch, 128) activation that one
// warp per row reads once.  Both directions are HBM/latency-bound row maps; backward accumulates the weight gradient in
// registers across a block's rows and finishes with one RED per column per block.
#include "rh_common.cuh"

namespace rh {

// CPL = columns per lane (K <= 32 * CPL); lane l owns columns l, l + 32, ... : every warp load is one coalesced run.
template <int CPL>
__global__ void __launch_bounds__(256) head_fwd_kernel(const float* __restrict__ x, int64_t x_ld, int64_t rows, int k,
                                                       const float* __restrict__ w, const float* __restrict__ bias,
                                                       const float* __restrict__ e0, const float* __restrict__ e1, int apply_sigmoid,
                                                       float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  float wr[CPL];
#pragma unroll
  for (int i = 0; i < CPL; ++i) {
    const int c = i * 32 + lane;
    wr[i] = c < k ? __ldg(w + c) : 0.f;
  }
  const float b = bias != nullptr ? __ldg(bias) : 0.f;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < rows; r += warps) {
    const float* xr = x + r * x_ld;
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < CPL; ++i) {
      const int c = i * 32 + lane;
      if (c < k) acc = fmaf(__ldg(xr + c), wr[i], acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) {
      float y = acc + b;
      if (e0 != nullptr) y += __ldg(e0 + r);
      if (e1 != nullptr) y += __ldg(e1 + r);
      out[r] = apply_sigmoid ? 1.f / (1.f + expf(-y)) : y;
    }
  }
}

// Backward.  d_logit = d_out * p (1 - p)  (torch's sigmoid_backward form);  d_x[r, :] = d_logit[r] * w;
// d_w = sum_r d_logit[r] * x[r, :];  d_b = sum_r d_logit[r];  d_extra[r] = d_logit[r].
template <int CPL>
__global__ void __launch_bounds__(256) head_bwd_kernel(const float* __restrict__ x, int64_t x_ld, int64_t rows, int k,
                                                       const float* __restrict__ w, const float* __restrict__ out,
                                                       const float* __restrict__ d_out, int apply_sigmoid, float* __restrict__ d_x,
                                                       int64_t d_x_ld, float* __restrict__ d_w, float* __restrict__ d_b,
                                                       float* __restrict__ d_extra, int rows_per_block) {
  __shared__ float sm_dw[8][32 * CPL + 1];
  __shared__ float sm_db[8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  float wr[CPL], dw[CPL];
#pragma unroll
  for (int i = 0; i < CPL; ++i) {
    const int c = i * 32 + lane;
    wr[i] = c < k ? __ldg(w + c) : 0.f;
    dw[i] = 0.f;
  }
  float db = 0.f;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  for (int64_t r = r0 + wid; r < r1; r += 8) {
    float g = __ldg(d_out + r);
    if (apply_sigmoid) {
      const float p = __ldg(out + r);
      g = g * (1.f - p) * p;
    }
    if (lane == 0) {
      db += g;
      if (d_extra != nullptr) d_extra[r] = g;
    }
    const float* xr = x + r * x_ld;
#pragma unroll
    for (int i = 0; i < CPL; ++i) {
      const int c = i * 32 + lane;
      if (c < k) {
        if (d_w != nullptr) dw[i] = fmaf(g, __ldg(xr + c), dw[i]);
        if (d_x != nullptr) d_x[r * d_x_ld + c] = g * wr[i];
      }
    }
  }
  if (d_w == nullptr && d_b == nullptr) return;
#pragma unroll
  for (int i = 0; i < CPL; ++i) sm_dw[wid][i * 32 + lane] = dw[i];
  if (lane == 0) sm_db[wid] = db;
  __syncthreads();
  for (int c = threadIdx.x; c < k; c += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += sm_dw[j][c];
    if (d_w != nullptr) atomicAdd(d_w + c, s);
  }
  if (threadIdx.x == 0 && d_b != nullptr) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += sm_db[j];
    atomicAdd(d_b, s);
  }
}

}  // namespace rh

using namespace rh;

extern "C" int rh_head_fwd(const float* x, int64_t x_ld, int64_t rows, int k, const float* w, const float* bias, const float* extra0,
                           const float* extra1, int apply_sigmoid, float* out, void* stream) {
  RH_REQUIRE(rows >= 0 && k > 0 && k <= 1024, RH_ERR_UNSUPPORTED, "rh_head_fwd: k %d not in [1,1024]", k);
  RH_REQUIRE(x != nullptr && w != nullptr && out != nullptr && x_ld >= k, RH_ERR_INVALID_ARG, "rh_head_fwd: NULL pointer or x_ld < k");
  if (rows == 0) return RH_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t want = (rows + 7) / 8;
  const int grid = (int)(want < 148 * 8 ? want : 148 * 8);
#define RH_HEAD_FWD(CPL) head_fwd_kernel<CPL><<<grid, 256, 0, st>>>(x, x_ld, rows, k, w, bias, extra0, extra1, apply_sigmoid, out)
  if (k <= 128) {
    RH_HEAD_FWD(4);
  } else if (k <= 256) {
    RH_HEAD_FWD(8);
  } else if (k <= 512) {
    RH_HEAD_FWD(16);
  } else {
    RH_HEAD_FWD(32);
  }
#undef RH_HEAD_FWD
  RH_LAUNCH_CHECK();
  return RH_OK;
}

extern "C" int rh_head_bwd(const float* x, int64_t x_ld, int64_t rows, int k, const float* w, const float* out, const float* d_out,
                           int apply_sigmoid, float* d_x, int64_t d_x_ld, float* d_w, float* d_b, float* d_extra, void* stream) {
  RH_REQUIRE(rows >= 0 && k > 0 && k <= 1024, RH_ERR_UNSUPPORTED, "rh_head_bwd: k %d not in [1,1024]", k);
  RH_REQUIRE(x != nullptr && w != nullptr && d_out != nullptr && x_ld >= k, RH_ERR_INVALID_ARG, "rh_head_bwd: NULL pointer or x_ld < k");
  RH_REQUIRE(!apply_sigmoid || out != nullptr, RH_ERR_INVALID_ARG, "rh_head_bwd: the sigmoid derivative needs the forward output");
  RH_REQUIRE(d_x == nullptr || d_x_ld >= k, RH_ERR_INVALID_ARG, "rh_head_bwd: d_x_ld < k");
  if (rows == 0) return RH_OK;
  cudaStream_t st = (cudaStream_t)stream;
  // ~2 blocks per SM: few enough REDs per column, enough warps to hide the row loads
  int64_t rpb = (rows + 295) / 296;
  rpb = (rpb + 7) / 8 * 8;
  const int grid = (int)((rows + rpb - 1) / rpb);
#define RH_HEAD_BWD(CPL) \
  head_bwd_kernel<CPL><<<grid, 256, 0, st>>>(x, x_ld, rows, k, w, out, d_out, apply_sigmoid, d_x, d_x_ld, d_w, d_b, d_extra, (int)rpb)
  if (k <= 128) {
    RH_HEAD_BWD(4);
  } else if (k <= 256) {
    RH_HEAD_BWD(8);
  } else if (k <= 512) {
    RH_HEAD_BWD(16);
  } else {
    RH_HEAD_BWD(32);
  }
#undef RH_HEAD_BWD
  RH_LAUNCH_CHECK();
  return RH_OK;
}
