// rh_match.cu — the in-batch-negative branch of the two-tower trainer as engine kernels.
//
// Reference arithmetic replaced (trainers/match_trainer.py:118-140, utils/match.py:104-161): scores = U V^T (B x B);
// inbatch_negative_sampling — a Python loop over the B rows with one randperm(B - 1) per row (or a top-k per row for hard negatives);
// gather_inbatch_logits — cat(diag, gather) to (B, 1 + K); CrossEntropyLoss with the positive in column 0; and, backward, a dense
// (B, B) zero-filled score gradient with 1 + K non-zeros per row feeding two (B, B) x (B, D) products.
// Here:
//   rh_inbatch_sample_random   thread = row: K distinct off-diagonal columns by sequential rejection from a counter-based hash
//                              stream keyed by a DEVICE seed (no host sync) — uniform over K-subsets, order included
//   rh_inbatch_sample_hard     warp = row: the K best-scoring off-diagonal columns, best first (ties: lower column first), by K
//                              ordered max-scans of the L1-resident score row
//   rh_inbatch_ce_fwd          warp = row: the 1 + K logits are DOT PRODUCTS <u_i, v_c> taken straight from the tower outputs (random
//                              negatives never need the (B, B) score matrix), softmax + loss in registers
//   rh_inbatch_ce_bwd          warp = row: d_u_i = sum_j g_ij v_c(ij) written, d_v_c += g_ij u_i by vector REDs — the (B, B) score
//                              gradient never exists.
// All HBM/L2-bound maps over (B, D <= 256) tower outputs and (B, K) picks.
#include "rh_common.cuh"

namespace rh {

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x7feb352dU;
  x ^= x >> 15;
  x *= 0x846ca68bU;
  x ^= x >> 16;
  return x;
}

constexpr int kMaxRejectK = 128;

__global__ void __launch_bounds__(128) inbatch_sample_random_kernel(int B, int K, const long long* __restrict__ seed_dev, long long* __restrict__ picks) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  const unsigned long long seed = (unsigned long long)*seed_dev;
  const uint32_t s0 = hash32((uint32_t)seed ^ 0x9E3779B9U), s1 = hash32((uint32_t)(seed >> 32) + 0x85EBCA6BU);
  int got[kMaxRejectK];
  uint32_t ctr = 0;
  for (int j = 0; j < K; ++j) {
    int c;
    bool dup;
    do {
      // a uniform draw from the B - 1 other columns: 64-bit multiply-shift of a 32-bit hash
      const uint32_t h = hash32(hash32((uint32_t)i * 0x9E3779B1U + s0) ^ hash32(ctr * 0x85EBCA77U + s1));
      ++ctr;
      c = (int)(((unsigned long long)h * (unsigned long long)(B - 1)) >> 32);
      c += (c >= i);
      dup = false;
      for (int t = 0; t < j; ++t) dup |= (got[t] == c);
    } while (dup);
    got[j] = c;
    picks[(int64_t)i * K + j] = c;
  }
}

__global__ void __launch_bounds__(256) inbatch_all_others_kernel(int B, long long* __restrict__ picks) {
  const int64_t total = (int64_t)B * (B - 1);
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(t / (B - 1)), j = (int)(t - (int64_t)i * (B - 1));
    picks[t] = j + (j >= i);
  }
}

// order: larger score first, then lower column.  (v, c) "comes after" (pv, pc) when v < pv or (v == pv and c > pc).
__global__ void __launch_bounds__(256) inbatch_sample_hard_kernel(const float* __restrict__ scores, int64_t ld, int B, int K, long long* __restrict__ picks) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < B; i += warps) {
    const float* row = scores + i * ld;
    float pv = INFINITY;
    int pc = -1;
    for (int j = 0; j < K; ++j) {
      float bv = -INFINITY;
      int bc = 0x7fffffff;
      for (int c = lane; c < B; c += 32) {
        if (c == (int)i) continue;
        const float v = __ldg(row + c);
        const bool after_prev = (v < pv) || (v == pv && c > pc);
        const bool better = (v > bv) || (v == bv && c < bc);
        if (after_prev && better) {
          bv = v;
          bc = c;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oc = __shfl_xor_sync(0xffffffffu, bc, o);
        if (ov > bv || (ov == bv && oc < bc)) {
          bv = ov;
          bc = oc;
        }
      }
      if (lane == 0) picks[i * K + j] = bc;
      pv = bv;
      pc = bc;
    }
  }
}

// warp = row i.  logit_0 = <u_i, v_i>, logit_j = <u_i, v_picks[i, j-1]>;  prob = softmax(logits);  loss_i = -log prob_0
__global__ void __launch_bounds__(256) inbatch_ce_fwd_kernel(const float* __restrict__ u, int64_t ldu, const float* __restrict__ v, int64_t ldv, int D,
                                                             const long long* __restrict__ picks, int B, int K, float* __restrict__ prob,
                                                             float* __restrict__ loss_rows) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < B; i += warps) {
    const float* ui = u + i * ldu;
    float mx = -INFINITY;
    // pass 1: logits (lane = candidate j, j + 32, ...), kept in prob[] as scratch
    for (int j = lane; j <= K; j += 32) {
      const int64_t c = j == 0 ? i : picks[i * K + j - 1];
      const float* vc = v + c * ldv;
      float acc = 0.f;
      for (int d = 0; d < D; ++d) acc = fmaf(__ldg(ui + d), __ldg(vc + d), acc);
      prob[i * (K + 1) + j] = acc;
      mx = fmaxf(mx, acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    __syncwarp();
    float sum = 0.f;
    for (int j = lane; j <= K; j += 32) sum += expf(prob[i * (K + 1) + j] - mx);
    sum = warp_sum(sum);
    const float l0 = prob[i * (K + 1)];
    __syncwarp();
    for (int j = lane; j <= K; j += 32) prob[i * (K + 1) + j] = expf(prob[i * (K + 1) + j] - mx) / sum;
    if (lane == 0) loss_rows[i] = (mx + logf(sum)) - l0;
  }
}

// warp = row i.  g_j = (prob_j - [j == 0]) * scale, scale = *d_loss / B (mean reduction);  d_u_i = sum_j g_j v_c(j);  d_v_c(j) += g_j u_i
__global__ void __launch_bounds__(256) inbatch_ce_bwd_kernel(const float* __restrict__ u, int64_t ldu, const float* __restrict__ v, int64_t ldv, int D,
                                                             const long long* __restrict__ picks, const float* __restrict__ prob, const float* __restrict__ d_loss,
                                                             int B, int K, float* __restrict__ d_u, int64_t lddu, float* __restrict__ d_v, int64_t lddv) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const float scale = __ldg(d_loss) / (float)B;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < B; i += warps) {
    // lane owns dimensions d = lane, lane + 32, ... of the rows (D <= 256 -> 8 per lane)
    float ui[8], acc[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const int d = lane + 32 * t;
      ui[t] = d < D ? __ldg(u + i * ldu + d) : 0.f;
      acc[t] = 0.f;
    }
    for (int j = 0; j <= K; ++j) {
      const int64_t c = j == 0 ? i : picks[i * K + j - 1];
      const float g = (__ldg(prob + i * (K + 1) + j) - (j == 0 ? 1.f : 0.f)) * scale;
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const int d = lane + 32 * t;
        if (d < D) {
          acc[t] = fmaf(g, __ldg(v + c * ldv + d), acc[t]);
          atomicAdd(d_v + c * lddv + d, g * ui[t]);
        }
      }
    }
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const int d = lane + 32 * t;
      if (d < D) d_u[i * lddu + d] = acc[t];
    }
  }
}

static int rows_warp_grid(int64_t rows) {
  int64_t g = (rows + 7) / 8;
  const int64_t cap = (int64_t)num_sms() * 8;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace rh

using namespace rh;

extern "C" int rh_inbatch_sample_random(int batch, int k, const int64_t* seed_dev, int64_t* picks, void* stream) {
  RH_REQUIRE(seed_dev && picks, RH_ERR_INVALID_ARG, "rh_inbatch_sample_random: NULL pointer");
  RH_REQUIRE(batch > 1 && k > 0 && k <= batch - 1, RH_ERR_INVALID_ARG, "rh_inbatch_sample_random: need 0 < k <= batch - 1 (batch %d, k %d)", batch, k);
  cudaStream_t st = (cudaStream_t)stream;
  if (k == batch - 1) {  // every other column: nothing to draw
    int64_t g = ((int64_t)batch * k + 255) / 256;
    if (g > (int64_t)num_sms() * 8) g = (int64_t)num_sms() * 8;
    inbatch_all_others_kernel<<<(int)g, 256, 0, st>>>(batch, reinterpret_cast<long long*>(picks));
  } else {
    RH_REQUIRE(k <= kMaxRejectK, RH_ERR_UNSUPPORTED, "rh_inbatch_sample_random: k %d > %d (and < batch - 1) is outside the rejection sampler", k, kMaxRejectK);
    inbatch_sample_random_kernel<<<(batch + 127) / 128, 128, 0, st>>>(batch, k, reinterpret_cast<const long long*>(seed_dev), reinterpret_cast<long long*>(picks));
  }
  RH_LAUNCH_CHECK();
  return RH_OK;
}

extern "C" int rh_inbatch_sample_hard(const float* scores, int64_t ld, int batch, int k, int64_t* picks, void* stream) {
  RH_REQUIRE(scores && picks && ld >= batch, RH_ERR_INVALID_ARG, "rh_inbatch_sample_hard: NULL pointer or ld < batch");
  RH_REQUIRE(batch > 1 && k > 0 && k <= batch - 1, RH_ERR_INVALID_ARG, "rh_inbatch_sample_hard: need 0 < k <= batch - 1");
  inbatch_sample_hard_kernel<<<rows_warp_grid(batch), 256, 0, (cudaStream_t)stream>>>(scores, ld, batch, k, reinterpret_cast<long long*>(picks));
  RH_LAUNCH_CHECK();
  return RH_OK;
}

extern "C" int rh_inbatch_ce_fwd(const float* user, int64_t ldu, const float* item, int64_t ldv, int dim, const int64_t* picks, int batch, int k, float* prob,
                                 float* loss_rows, void* stream) {
  RH_REQUIRE(user && item && picks && prob && loss_rows && dim > 0 && ldu >= dim && ldv >= dim && batch > 0 && k > 0, RH_ERR_INVALID_ARG,
             "rh_inbatch_ce_fwd: bad arguments");
  inbatch_ce_fwd_kernel<<<rows_warp_grid(batch), 256, 0, (cudaStream_t)stream>>>(user, ldu, item, ldv, dim, reinterpret_cast<const long long*>(picks), batch, k, prob,
                                                                                  loss_rows);
  RH_LAUNCH_CHECK();
  return RH_OK;
}

extern "C" int rh_inbatch_ce_bwd(const float* user, int64_t ldu, const float* item, int64_t ldv, int dim, const int64_t* picks, const float* prob,
                                 const float* d_loss, int batch, int k, float* d_user, int64_t lddu, float* d_item, int64_t lddv, void* stream) {
  RH_REQUIRE(user && item && picks && prob && d_loss && d_user && d_item && dim > 0 && batch > 0 && k > 0, RH_ERR_INVALID_ARG, "rh_inbatch_ce_bwd: bad arguments");
  RH_REQUIRE(dim <= 256, RH_ERR_UNSUPPORTED, "rh_inbatch_ce_bwd: dim %d > 256", dim);
  RH_REQUIRE(ldu >= dim && ldv >= dim && lddu >= dim && lddv >= dim, RH_ERR_INVALID_ARG, "rh_inbatch_ce_bwd: leading dimension < dim");
  inbatch_ce_bwd_kernel<<<rows_warp_grid(batch), 256, 0, (cudaStream_t)stream>>>(user, ldu, item, ldv, dim, reinterpret_cast<const long long*>(picks), prob, d_loss, batch,
                                                                                  k, d_user, lddu, d_item, lddv);
  RH_LAUNCH_CHECK();
  return RH_OK;
}
