// rh_head.cu — the output head of a ranking tower: Linear(K -> 1) + per-sample side terms + sigmoid, one pass each way.
//
// Reference arithmetic replaced: the MLP's output layer nn.Linear(input_dim, 1) (basic/layers.py:279-280) followed by
// the model's tail, e.g. DeepFM  y = y_linear + y_fm + y_deep; sigmoid(y.squeeze(1))  (models/ranking/deepfm.py:41-43).
// Stock PyTorch runs this as gemv + add + add + sigmoid forward and sigmoid_backward + an outer product + a
// transposed gemv (+ split-K reduce) + a bias reduction backward — ten launches moving a (batch, 128) activation that one
// warp per row reads once.  Both directions are HBM/latency-bound row maps; backward accumulates the weight gradient in
// registers across a block's rows and finishes with one RED per column per block.
#ifndef RH_PDL_FAMILY  // (the trace tools include several of these files into one unit: the first one names the family)
#define RH_PDL_FAMILY 16  /* rh_set_pdl mask bit of this file's kernels */
#endif
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

// ---- BCELoss(mean) on probabilities (trainers/ctr_trainer.py:68,88: torch.nn.BCELoss) as one launch each way -----------------------
// torch's arithmetic: loss_i = -(y log p + (1 - y) log(1 - p)) with both logs clamped at -100; d loss / d p = (p - y) / max(p (1 - p), 1e-12) / N.
// Stock PyTorch runs four launches per step for it (element-wise loss, mean reduction, its backward, a fill); at batch 4096 each is pure
// launch latency.  Forward: per-block partial sums + a last-block finalisation in a fixed order (deterministic); `partial` holds
// gridDim floats + a ticket counter (zero on entry, left zero).
namespace rh {
__global__ void __launch_bounds__(1024) bce_fwd_kernel(const float* __restrict__ p, const float* __restrict__ y, int64_t n, float* __restrict__ partial,
                                                      unsigned* __restrict__ ticket, float* __restrict__ loss) {
  __shared__ float sm[32];
  __shared__ bool is_last;
  pdl_wait();
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float pi = __ldg(p + i), yi = __ldg(y + i);
    const float lp = fmaxf(logf(pi), -100.f), lq = fmaxf(log1pf(-pi), -100.f);
    acc -= yi * lp + (1.f - yi) * lq;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sm[w];
    if (gridDim.x == 1) {
      *loss = t / (float)n;
      is_last = false;
    } else {
      partial[blockIdx.x] = t;
      __threadfence();
      is_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
  }
  __syncthreads();
  if (!is_last || threadIdx.x != 0) return;
  __threadfence();
  float t = 0.f;
  for (unsigned b = 0; b < gridDim.x; ++b) t += __ldcg(partial + b);
  *loss = t / (float)n;
  *ticket = 0u;
}

__global__ void __launch_bounds__(256) bce_bwd_kernel(const float* __restrict__ p, const float* __restrict__ y, const float* __restrict__ d_loss, int64_t n,
                                                      float* __restrict__ d_p) {
  pdl_wait();
  const float s = __ldg(d_loss) / (float)n;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float pi = __ldg(p + i);
    d_p[i] = s * (pi - __ldg(y + i)) / fmaxf(pi * (1.f - pi), 1e-12f);
  }
}
}  // namespace rh

extern "C" int rh_bce_fwd(const float* prob, const float* target, int64_t n, float* scratch, float* loss, void* stream) {
  RH_REQUIRE(prob && target && scratch && loss && n > 0, RH_ERR_INVALID_ARG, "rh_bce_fwd: bad arguments");
  // up to 16 k probabilities: ONE 1024-thread block (no ticket, no fence: a training batch's loss is launch latency, not bandwidth)
  const bool one = n <= 16384;
  int64_t g = one ? 1 : (n + 2047) / 2048;
  if (g > 64) g = 64;
  launch_k(rh::bce_fwd_kernel, dim3((unsigned)g), dim3(one ? 1024 : 256), 0, (cudaStream_t)stream, prob, target, n, scratch, reinterpret_cast<unsigned*>(scratch + 64), loss);
  RH_LAUNCH_CHECK();
  return RH_OK;
}

extern "C" int rh_bce_bwd(const float* prob, const float* target, const float* d_loss, int64_t n, float* d_prob, void* stream) {
  RH_REQUIRE(prob && target && d_loss && d_prob && n > 0, RH_ERR_INVALID_ARG, "rh_bce_bwd: bad arguments");
  int64_t g = (n + 255) / 256;
  if (g > 148 * 8) g = 148 * 8;
  launch_k(rh::bce_bwd_kernel, dim3((unsigned)g), dim3(256), 0, (cudaStream_t)stream, prob, target, d_loss, n, d_prob);
  RH_LAUNCH_CHECK();
  return RH_OK;
}
