// rh_mlp.cu — the glue between the tower GEMMs: BatchNorm1d statistics and the fused
// BatchNorm-apply + activation (ReLU / Dice / PReLU / sigmoid / LeakyReLU) + dropout pass, forward
// and backward.
//
// Reference arithmetic replaced: the [BatchNorm1d -> activation -> Dropout] part of
// MLP.forward (basic/layers.py:276-292) and Dice.forward (basic/activation.py:15-25).  These are
// HBM-bound: the reference runs ~12 elementwise ATen kernels per Dice layer over a (B*L, 256) tensor
// (DIN: 204 800 x 256 = 210 MB per pass); here a layer is one read of h and one write of y.
//
// Mapping: one warp owns one row (lane <-> column c = k*32 + lane, k < KMAX) so Dice's per-row
// statistics are two warp reductions; per-column constants are hoisted into registers once per warp.
#include "rh_common.cuh"

namespace rh {

enum { ACT_NONE = 0, ACT_RELU = 1, ACT_DICE = 2, ACT_PRELU = 3, ACT_SIGMOID = 4, ACT_LEAKY = 5 };

__device__ __forceinline__ float sigmoidf_precise(float u) { return 1.f / (1.f + expf(-u)); }

// ---- column statistics (BatchNorm1d training forward) -------------------------------------------
// grid = (column tiles of 32, row chunks).  Shifted sums (shift = row 0) keep fp32 accurate when
// |mean| >> std.  The last block to finish (ticket in scratch[2*cols]) finalises mean/var, updates
// the running statistics and re-zeroes the scratch, so the scratch needs zeroing only once, ever.
__global__ void __launch_bounds__(256) colstats_kernel(const float* __restrict__ h, int64_t h_ld, int64_t rows, int cols,
                                                       int64_t rows_per_block, float* __restrict__ mean, float* __restrict__ var,
                                                       float* __restrict__ scratch, float* __restrict__ running_mean,
                                                       float* __restrict__ running_var, long long* __restrict__ num_batches_tracked,
                                                       float momentum) {
  __shared__ float sm1[8][32], sm2[8][32];
  __shared__ bool is_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  int64_t r1 = r0 + rows_per_block;
  if (r1 > rows) r1 = rows;
  float s1 = 0.f, s2 = 0.f;
  if (c < cols) {
    const float shift = __ldg(h + c);
    for (int64_t r = r0 + warp; r < r1; r += nwarp) {
      const float d = __ldg(h + r * h_ld + c) - shift;
      s1 += d;
      s2 = fmaf(d, d, s2);
    }
  }
  sm1[warp][lane] = s1;
  sm2[warp][lane] = s2;
  __syncthreads();
  if (warp == 0 && c < cols) {
    for (int k = 1; k < nwarp; ++k) {
      s1 += sm1[k][lane];
      s2 += sm2[k][lane];
    }
    atomicAdd(scratch + c, s1);
    atomicAdd(scratch + cols + c, s2);
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned total = gridDim.x * gridDim.y;
    const unsigned ticket = atomicAdd(reinterpret_cast<unsigned*>(scratch + 2 * cols), 1u);
    is_last = (ticket == total - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  const float n = (float)rows;
  for (int cc = threadIdx.x; cc < cols; cc += blockDim.x) {
    const float t1 = __ldcg(scratch + cc), t2 = __ldcg(scratch + cols + cc);
    const float shift = __ldg(h + cc);
    const float m = shift + t1 / n;
    float v = (t2 - t1 * t1 / n) / n;
    if (v < 0.f) v = 0.f;
    mean[cc] = m;
    var[cc] = v;
    if (running_mean != nullptr) running_mean[cc] = (1.f - momentum) * running_mean[cc] + momentum * m;
    if (running_var != nullptr) {
      const float unbiased = rows > 1 ? v * (n / (n - 1.f)) : v;
      running_var[cc] = (1.f - momentum) * running_var[cc] + momentum * unbiased;
    }
    scratch[cc] = 0.f;
    scratch[cols + cc] = 0.f;
  }
  if (threadIdx.x == 0) {
    *reinterpret_cast<unsigned*>(scratch + 2 * cols) = 0u;
    if (num_batches_tracked != nullptr) *num_batches_tracked += 1;
  }
}

// ---- fused BN-apply + activation + dropout ------------------------------------------------------
struct BnActP {
  const float* h;
  int64_t h_ld;
  int64_t rows;
  int cols;
  const float* mean;
  const float* var;
  float bn_eps;
  const float* gamma;
  const float* beta;
  int act;
  const float* alpha;
  float dice_eps;
  const uint8_t* mask;
  float p_drop;
  // forward
  float* y;
  int64_t y_ld;
  // backward
  const float* d_y;
  int64_t d_y_ld;
  int training;
  float* d_h;
  int64_t d_h_ld;
  float* d_gamma;
  float* d_beta;
  float* d_alpha;
};

template <int KMAX>
__global__ void __launch_bounds__(256) bn_act_fwd_kernel(const BnActP p) {
  const int lane = threadIdx.x & 31;
  const int64_t warps_total = ((int64_t)gridDim.x * blockDim.x) >> 5;
  float mu[KMAX], sc[KMAX], sh[KMAX];
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    const int c = k * 32 + lane;
    mu[k] = 0.f;
    sc[k] = 1.f;
    sh[k] = 0.f;
    if (c < p.cols) {
      if (p.mean != nullptr) {
        mu[k] = __ldg(p.mean + c);
        sc[k] = 1.f / sqrtf(__ldg(p.var + c) + p.bn_eps);
      }
      if (p.gamma != nullptr) sc[k] *= __ldg(p.gamma + c);
      if (p.beta != nullptr) sh[k] = __ldg(p.beta + c);
    }
  }
  const float alpha = (p.alpha != nullptr) ? __ldg(p.alpha) : 0.f;
  const float keep_scale = p.mask != nullptr ? 1.f / (1.f - p.p_drop) : 1.f;
  const float inv_n = 1.f / (float)p.cols;

  for (int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < p.rows; row += warps_total) {
    float z[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      const int c = k * 32 + lane;
      z[k] = c < p.cols ? fmaf(__ldg(p.h + row * p.h_ld + c) - mu[k], sc[k], sh[k]) : 0.f;
    }
    float m = 0.f, inv_s = 0.f;
    if (p.act == ACT_DICE) {
      float t = 0.f;
#pragma unroll
      for (int k = 0; k < KMAX; ++k) t += z[k];  // padding lanes hold 0
      m = warp_sum(t) * inv_n;
      float qv = 0.f;
#pragma unroll
      for (int k = 0; k < KMAX; ++k) {
        const int c = k * 32 + lane;
        if (c < p.cols) {
          const float d = z[k] - m;
          qv += fmaf(d, d, p.dice_eps);
        }
      }
      inv_s = 1.f / sqrtf(warp_sum(qv));
    }
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      const int c = k * 32 + lane;
      if (c >= p.cols) continue;
      float o;
      switch (p.act) {
        case ACT_RELU: o = fmaxf(z[k], 0.f); break;
        case ACT_DICE: {
          const float ps = sigmoidf_precise((z[k] - m) * inv_s);
          o = ps * z[k] + (1.f - ps) * alpha * z[k];
        } break;
        case ACT_PRELU: o = z[k] > 0.f ? z[k] : alpha * z[k]; break;
        case ACT_SIGMOID: o = sigmoidf_precise(z[k]); break;
        case ACT_LEAKY: o = z[k] > 0.f ? z[k] : 0.01f * z[k]; break;
        default: o = z[k]; break;
      }
      if (p.mask != nullptr) o = p.mask[row * p.cols + c] ? o * keep_scale : 0.f;
      p.y[row * p.y_ld + c] = o;
    }
  }
}

// backward pass 1: d_z (stored in d_h) + column sums d_beta, d_gamma (+ d_alpha).
// In eval mode (training == 0) BN is an affine map, so d_h = d_z * gamma * rstd is final here.
template <int KMAX>
__global__ void __launch_bounds__(256) bn_act_bwd_kernel(const BnActP p) {
  extern __shared__ float smem[];  // [warps][2*cols]
  __shared__ float sm_alpha[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int64_t warps_total = (int64_t)gridDim.x * wpb;
  float mu[KMAX], rstd[KMAX], gam[KMAX], bet[KMAX], acc_b[KMAX], acc_g[KMAX];
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    const int c = k * 32 + lane;
    mu[k] = 0.f;
    rstd[k] = 1.f;
    gam[k] = 1.f;
    bet[k] = 0.f;
    acc_b[k] = acc_g[k] = 0.f;
    if (c < p.cols) {
      if (p.mean != nullptr) {
        mu[k] = __ldg(p.mean + c);
        rstd[k] = 1.f / sqrtf(__ldg(p.var + c) + p.bn_eps);
      }
      if (p.gamma != nullptr) gam[k] = __ldg(p.gamma + c);
      if (p.beta != nullptr) bet[k] = __ldg(p.beta + c);
    }
  }
  const float alpha = (p.alpha != nullptr) ? __ldg(p.alpha) : 0.f;
  const float keep_scale = p.mask != nullptr ? 1.f / (1.f - p.p_drop) : 1.f;
  const float inv_n = 1.f / (float)p.cols;
  float acc_alpha = 0.f;

  for (int64_t row = (int64_t)blockIdx.x * wpb + warp; row < p.rows; row += warps_total) {
    float xh[KMAX], z[KMAX], da[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      const int c = k * 32 + lane;
      xh[k] = z[k] = da[k] = 0.f;
      if (c < p.cols) {
        xh[k] = (__ldg(p.h + row * p.h_ld + c) - mu[k]) * rstd[k];
        z[k] = fmaf(xh[k], gam[k], bet[k]);
        da[k] = __ldg(p.d_y + row * p.d_y_ld + c);
        if (p.mask != nullptr) da[k] = p.mask[row * p.cols + c] ? da[k] * keep_scale : 0.f;
      }
    }
    float dz[KMAX];
    if (p.act == ACT_DICE) {
      float t = 0.f;
#pragma unroll
      for (int k = 0; k < KMAX; ++k) t += z[k];
      const float m = warp_sum(t) * inv_n;
      float qv = 0.f;
#pragma unroll
      for (int k = 0; k < KMAX; ++k) {
        const int c = k * 32 + lane;
        if (c < p.cols) {
          const float d = z[k] - m;
          qv += fmaf(d, d, p.dice_eps);
        }
      }
      const float s = sqrtf(warp_sum(qv));
      const float inv_s = 1.f / s;
      // out_j = z_j (alpha + (1-alpha) p_j), p_j = sigmoid((z_j - m)/s)
      float a1 = 0.f, a2 = 0.f;
      float aj[KMAX], pj[KMAX];
#pragma unroll
      for (int k = 0; k < KMAX; ++k) {
        const int c = k * 32 + lane;
        aj[k] = pj[k] = 0.f;
        if (c < p.cols) {
          pj[k] = sigmoidf_precise((z[k] - m) * inv_s);
          aj[k] = da[k] * z[k] * (1.f - alpha) * pj[k] * (1.f - pj[k]);
          a1 += aj[k];
          a2 = fmaf(aj[k], z[k] - m, a2);
          acc_alpha = fmaf(da[k] * z[k], 1.f - pj[k], acc_alpha);
        }
      }
      a1 = warp_sum(a1);
      a2 = warp_sum(a2);
      const float k1 = a1 * inv_n * inv_s, k2 = a2 * inv_s * inv_s * inv_s;
#pragma unroll
      for (int k = 0; k < KMAX; ++k) dz[k] = da[k] * (alpha + (1.f - alpha) * pj[k]) + aj[k] * inv_s - k1 - (z[k] - m) * k2;
    } else {
#pragma unroll
      for (int k = 0; k < KMAX; ++k) {
        float d;
        switch (p.act) {
          case ACT_RELU: d = z[k] > 0.f ? da[k] : 0.f; break;
          case ACT_PRELU:
            d = z[k] > 0.f ? da[k] : alpha * da[k];
            if (z[k] <= 0.f) acc_alpha = fmaf(da[k], z[k], acc_alpha);
            break;
          case ACT_SIGMOID: {
            const float sg = sigmoidf_precise(z[k]);
            d = da[k] * sg * (1.f - sg);
          } break;
          case ACT_LEAKY: d = z[k] > 0.f ? da[k] : 0.01f * da[k]; break;
          default: d = da[k]; break;
        }
        dz[k] = d;
      }
    }
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      const int c = k * 32 + lane;
      if (c < p.cols) {
        acc_b[k] += dz[k];
        acc_g[k] = fmaf(dz[k], xh[k], acc_g[k]);
        // training: keep d_z for pass 2;  eval: BN is affine, finish now
        p.d_h[row * p.d_h_ld + c] = p.training ? dz[k] : dz[k] * gam[k] * rstd[k];
      }
    }
  }
  // block reduction of the column sums, one RED per column per block
  float* my = smem + (int64_t)warp * 2 * p.cols;
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    const int c = k * 32 + lane;
    if (c < p.cols) {
      my[c] = acc_b[k];
      my[p.cols + c] = acc_g[k];
    }
  }
  acc_alpha = warp_sum(acc_alpha);
  if (lane == 0) sm_alpha[warp] = acc_alpha;
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * p.cols; i += blockDim.x) {
    float t = 0.f;
    for (int wv = 0; wv < wpb; ++wv) t += smem[(int64_t)wv * 2 * p.cols + i];
    if (i < p.cols) {
      if (p.d_beta != nullptr) atomicAdd(p.d_beta + i, t);
    } else if (p.d_gamma != nullptr) {
      atomicAdd(p.d_gamma + (i - p.cols), t);
    }
  }
  if (threadIdx.x == 0 && p.d_alpha != nullptr && (p.act == ACT_DICE || p.act == ACT_PRELU)) {
    float t = 0.f;
    for (int wv = 0; wv < wpb; ++wv) t += sm_alpha[wv];
    atomicAdd(p.d_alpha, t);
  }
}

// backward pass 2 (training only): d_h = gamma*rstd * (d_z - sum(d_z)/N - xhat * sum(d_z*xhat)/N), in place.
template <int KMAX>
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const BnActP p) {
  const int lane = threadIdx.x & 31;
  const int64_t warps_total = ((int64_t)gridDim.x * blockDim.x) >> 5;
  float mu[KMAX], rstd[KMAX], gs[KMAX], mb[KMAX], mg[KMAX];
  const float inv_rows = 1.f / (float)p.rows;
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    const int c = k * 32 + lane;
    mu[k] = mb[k] = mg[k] = 0.f;
    rstd[k] = gs[k] = 1.f;
    if (c < p.cols) {
      mu[k] = __ldg(p.mean + c);
      rstd[k] = 1.f / sqrtf(__ldg(p.var + c) + p.bn_eps);
      gs[k] = (p.gamma != nullptr ? __ldg(p.gamma + c) : 1.f) * rstd[k];
      mb[k] = __ldcg(p.d_beta + c) * inv_rows;
      mg[k] = __ldcg(p.d_gamma + c) * inv_rows;
    }
  }
  for (int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < p.rows; row += warps_total) {
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      const int c = k * 32 + lane;
      if (c < p.cols) {
        const float xh = (__ldg(p.h + row * p.h_ld + c) - mu[k]) * rstd[k];
        float* dst = p.d_h + row * p.d_h_ld + c;
        *dst = gs[k] * (*dst - mb[k] - xh * mg[k]);
      }
    }
  }
}

static int rows_grid(int64_t rows, int wpb, int blocks_per_sm) {
  int64_t g = (rows + wpb - 1) / wpb;
  const int64_t cap = (int64_t)num_sms() * blocks_per_sm;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace rh

using namespace rh;

extern "C" int rh_colstats(const float* h, int64_t h_ld, int64_t rows, int cols, float* mean, float* var, float* scratch,
                           float* running_mean, float* running_var, int64_t* num_batches_tracked, float momentum, void* stream) {
  RH_REQUIRE(h && mean && var && scratch, RH_ERR_INVALID_ARG, "rh_colstats: NULL pointer");
  RH_REQUIRE(rows > 0 && cols > 0 && h_ld >= cols, RH_ERR_INVALID_ARG, "rh_colstats: bad sizes");
  const int col_tiles = (cols + 31) / 32;
  // aim at ~4 blocks per SM; at least 8 rows per warp so the partial sums amortise the reduction
  int64_t want_chunks = ((int64_t)num_sms() * 4 + col_tiles - 1) / col_tiles;
  int64_t rpb = (rows + want_chunks - 1) / want_chunks;
  if (rpb < 64) rpb = 64;
  const int64_t chunks = (rows + rpb - 1) / rpb;
  RH_REQUIRE(chunks <= 65535, RH_ERR_UNSUPPORTED, "rh_colstats: too many row chunks");
  dim3 grid(col_tiles, (unsigned)chunks);
  colstats_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(h, h_ld, rows, cols, rpb, mean, var, scratch, running_mean, running_var,
                                                          reinterpret_cast<long long*>(num_batches_tracked), momentum);
  RH_LAUNCH_CHECK();
  return RH_OK;
}

#define RH_DISPATCH_KMAX(cols, CALL)                                      \
  do {                                                                    \
    const int kk__ = ((cols) + 31) / 32;                                  \
    if (kk__ <= 1) { CALL(1); }                                           \
    else if (kk__ <= 2) { CALL(2); }                                      \
    else if (kk__ <= 4) { CALL(4); }                                      \
    else if (kk__ <= 8) { CALL(8); }                                      \
    else if (kk__ <= 16) { CALL(16); }                                    \
    else if (kk__ <= 32) { CALL(32); }                                    \
    else {                                                                \
      set_error("bn_act: %d columns > 1024 not supported", (cols));       \
      return RH_ERR_UNSUPPORTED;                                          \
    }                                                                     \
  } while (0)

extern "C" int rh_bn_act_fwd(const float* h, int64_t h_ld, int64_t rows, int cols, const float* mean, const float* var, float bn_eps,
                             const float* gamma, const float* beta, int act, const float* act_param, float dice_eps,
                             const uint8_t* keep_mask, float p_drop, float* y, int64_t y_ld, void* stream) {
  RH_REQUIRE(h && y, RH_ERR_INVALID_ARG, "rh_bn_act_fwd: NULL pointer");
  RH_REQUIRE((mean == nullptr) == (var == nullptr), RH_ERR_INVALID_ARG, "rh_bn_act_fwd: mean and var go together");
  RH_REQUIRE(rows >= 0 && cols > 0 && h_ld >= cols && y_ld >= cols, RH_ERR_INVALID_ARG, "rh_bn_act_fwd: bad sizes");
  RH_REQUIRE(act >= 0 && act <= 5, RH_ERR_INVALID_ARG, "rh_bn_act_fwd: act %d unknown", act);
  RH_REQUIRE(!((act == ACT_DICE || act == ACT_PRELU) && act_param == nullptr), RH_ERR_INVALID_ARG, "rh_bn_act_fwd: Dice/PReLU need act_param");
  RH_REQUIRE(keep_mask == nullptr || (p_drop >= 0.f && p_drop < 1.f), RH_ERR_INVALID_ARG, "rh_bn_act_fwd: p_drop must be in [0,1)");
  if (rows == 0) return RH_OK;
  BnActP p;
  memset(&p, 0, sizeof(p));
  p.h = h; p.h_ld = h_ld; p.rows = rows; p.cols = cols; p.mean = mean; p.var = var; p.bn_eps = bn_eps;
  p.gamma = gamma; p.beta = beta; p.act = act; p.alpha = act_param; p.dice_eps = dice_eps;
  p.mask = keep_mask; p.p_drop = p_drop; p.y = y; p.y_ld = y_ld;
  const int grid = rows_grid(rows, 8, 8);
#define RH_CALL(K) bn_act_fwd_kernel<K><<<grid, 256, 0, (cudaStream_t)stream>>>(p)
  RH_DISPATCH_KMAX(cols, RH_CALL);
#undef RH_CALL
  RH_LAUNCH_CHECK();
  return RH_OK;
}

extern "C" int rh_bn_act_bwd(const float* h, int64_t h_ld, int64_t rows, int cols, const float* mean, const float* var, float bn_eps,
                             const float* gamma, const float* beta, int act, const float* act_param, float dice_eps,
                             const uint8_t* keep_mask, float p_drop, const float* d_y, int64_t d_y_ld, int training, float* d_h,
                             int64_t d_h_ld, float* d_gamma, float* d_beta, float* d_act_param, void* stream) {
  RH_REQUIRE(h && d_y && d_h, RH_ERR_INVALID_ARG, "rh_bn_act_bwd: NULL pointer");
  RH_REQUIRE((mean == nullptr) == (var == nullptr), RH_ERR_INVALID_ARG, "rh_bn_act_bwd: mean and var go together");
  RH_REQUIRE(rows >= 0 && cols > 0 && h_ld >= cols && d_y_ld >= cols && d_h_ld >= cols, RH_ERR_INVALID_ARG, "rh_bn_act_bwd: bad sizes");
  RH_REQUIRE(act >= 0 && act <= 5, RH_ERR_INVALID_ARG, "rh_bn_act_bwd: act %d unknown", act);
  RH_REQUIRE(!((act == ACT_DICE || act == ACT_PRELU) && act_param == nullptr), RH_ERR_INVALID_ARG, "rh_bn_act_bwd: Dice/PReLU need act_param");
  const bool bn_train = training != 0 && mean != nullptr;
  RH_REQUIRE(!bn_train || (d_gamma != nullptr && d_beta != nullptr), RH_ERR_INVALID_ARG,
             "rh_bn_act_bwd: training-mode BN needs d_gamma/d_beta (zeroed) as the column-sum buffers");
  if (rows == 0) return RH_OK;
  BnActP p;
  memset(&p, 0, sizeof(p));
  p.h = h; p.h_ld = h_ld; p.rows = rows; p.cols = cols; p.mean = mean; p.var = var; p.bn_eps = bn_eps;
  p.gamma = gamma; p.beta = beta; p.act = act; p.alpha = act_param; p.dice_eps = dice_eps;
  p.mask = keep_mask; p.p_drop = p_drop; p.d_y = d_y; p.d_y_ld = d_y_ld; p.training = bn_train ? 1 : 0;
  p.d_h = d_h; p.d_h_ld = d_h_ld; p.d_gamma = d_gamma; p.d_beta = d_beta; p.d_alpha = d_act_param;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid1 = rows_grid(rows, 8, 2);
  const size_t smem = (size_t)8 * 2 * cols * sizeof(float);
#define RH_CALL(K)                                                                                                   \
  do {                                                                                                               \
    if (smem > 48 * 1024) cudaFuncSetAttribute(bn_act_bwd_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    bn_act_bwd_kernel<K><<<grid1, 256, smem, st>>>(p);                                                                \
  } while (0)
  RH_DISPATCH_KMAX(cols, RH_CALL);
#undef RH_CALL
  RH_LAUNCH_CHECK();
  if (bn_train) {
    const int grid2 = rows_grid(rows, 8, 8);
#define RH_CALL(K) bn_bwd_apply_kernel<K><<<grid2, 256, 0, st>>>(p)
    RH_DISPATCH_KMAX(cols, RH_CALL);
#undef RH_CALL
    RH_LAUNCH_CHECK();
  }
  return RH_OK;
}
