// rh_mlp.cu — the glue between the tower GEMMs: BatchNorm1d statistics and the fused
// BatchNorm-apply + activation (ReLU / Dice / PReLU / sigmoid / LeakyReLU) + dropout pass, forward
// and backward, plus the multi-tensor Adam/SGD/Adagrad step for the (small) dense parameters.
//
// Reference arithmetic replaced: the [BatchNorm1d -> activation -> Dropout] part of
// MLP.forward (basic/layers.py:276-292), Dice.forward (basic/activation.py:15-25) and the dense half of
// optimizer.step() (trainers/ctr_trainer.py:99).  All HBM/L2-bound streaming maps: the reference runs
// ~12 elementwise ATen kernels per Dice layer over a (B*L, 256) tensor (DIN: 204 800 x 256 = 210 MB per
// pass); here a layer is one read of h and one write of y.
//
// Mapping: one warp owns one row; lane owns VEC consecutive columns per step (VEC = 4: 16-byte
// accesses), KMAX steps cover the row, so Dice's per-row statistics are two warp reductions and the
// per-column constants are hoisted into registers once per warp.
#ifndef RH_PDL_FAMILY  // (the trace tools include several of these files into one unit: the first one names the family)
#define RH_PDL_FAMILY 8  /* rh_set_pdl mask bit of this file's kernels */
#endif
#include "rh_bn_common.cuh"

namespace rh {

// ---- column statistics (BatchNorm1d training forward) -------------------------------------------
// grid = (column tiles of 128, row chunks); lane owns 4 consecutive columns (16-byte loads), warps
// stride over rows.  Shifted sums (shift = row 0) keep fp32 accurate when |mean| >> std.  The last
// block to finish (ticket in scratch[2*cols]) finalises mean/var, updates the running statistics,
// stores the step counter (dropout stream id) behind the statistics and re-zeroes the scratch.
template <int VEC>
__global__ void __launch_bounds__(256) colstats_kernel(const float* __restrict__ h, int64_t h_ld, int64_t rows, int cols,
                                                       int64_t rows_per_block, float* __restrict__ stats, float* __restrict__ scratch,
                                                       float* __restrict__ running_mean, float* __restrict__ running_var,
                                                       long long* __restrict__ num_batches_tracked, float momentum) {
  __shared__ float sm1[8][32 * VEC], sm2[8][32 * VEC];
  __shared__ bool is_last;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int c0 = (blockIdx.x * 32 + lane) * VEC;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  int64_t r1 = r0 + rows_per_block;
  if (r1 > rows) r1 = rows;
  float s1[VEC], s2[VEC], shift[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) s1[j] = s2[j] = shift[j] = 0.f;
  if (c0 < cols) {
    if (VEC == 4) {
      const float4 sh = __ldg(reinterpret_cast<const float4*>(h + c0));
      shift[0] = sh.x; shift[1 % VEC] = sh.y; shift[2 % VEC] = sh.z; shift[3 % VEC] = sh.w;
      int64_t r = r0 + warp;
      for (; r + nwarp < r1; r += 2 * nwarp) {  // two rows in flight per warp
        const float4 a = __ldg(reinterpret_cast<const float4*>(h + r * h_ld + c0));
        const float4 b = __ldg(reinterpret_cast<const float4*>(h + (r + nwarp) * h_ld + c0));
        const float da[4] = {a.x - sh.x, a.y - sh.y, a.z - sh.z, a.w - sh.w};
        const float db[4] = {b.x - sh.x, b.y - sh.y, b.z - sh.z, b.w - sh.w};
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          s1[j] += da[j] + db[j];
          s2[j] = fmaf(da[j], da[j], fmaf(db[j], db[j], s2[j]));
        }
      }
      for (; r < r1; r += nwarp) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(h + r * h_ld + c0));
        const float da[4] = {a.x - sh.x, a.y - sh.y, a.z - sh.z, a.w - sh.w};
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          s1[j] += da[j];
          s2[j] = fmaf(da[j], da[j], s2[j]);
        }
      }
    } else {
      shift[0] = __ldg(h + c0);
      for (int64_t r = r0 + warp; r < r1; r += nwarp) {
        const float d = __ldg(h + r * h_ld + c0) - shift[0];
        s1[0] += d;
        s2[0] = fmaf(d, d, s2[0]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    sm1[warp][lane * VEC + j] = s1[j];
    sm2[warp][lane * VEC + j] = s2[j];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 32 * VEC; i += blockDim.x) {
    const int c = blockIdx.x * 32 * VEC + i;
    if (c < cols) {
      float t1 = 0.f, t2 = 0.f;
      for (int k = 0; k < nwarp; ++k) {
        t1 += sm1[k][i];
        t2 += sm2[k][i];
      }
      atomicAdd(scratch + c, t1);
      atomicAdd(scratch + cols + c, t2);
    }
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
  float* mean = stats;
  float* var = stats + cols;
  for (int cc = threadIdx.x; cc < cols; cc += blockDim.x) {
    const float t1 = __ldcg(scratch + cc), t2 = __ldcg(scratch + cols + cc);
    const float sh = __ldg(h + cc);
    const float m = sh + t1 / n;
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
    long long count = 0;
    if (num_batches_tracked != nullptr) {
      count = *num_batches_tracked + 1;
      *num_batches_tracked = count;
    }
    stats[2 * cols] = __int_as_float((int)(count & 0x7fffffff));  // dropout stream id of this forward
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
  // dropout: keep mask is a pure function of (seed, *counter, element index): nothing is stored
  float p_drop;
  uint32_t seed;
  const float* counter;  // device: bit pattern of the stream id (written by rh_colstats), or NULL -> 0
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

template <int VEC>
__device__ __forceinline__ void load_cols(const float* p, float (&out)[VEC]) {
  if (VEC == 4) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(p));
    out[0] = v.x; out[1 % VEC] = v.y; out[2 % VEC] = v.z; out[3 % VEC] = v.w;
  } else {
    out[0] = __ldg(p);
  }
}
template <int VEC>
__device__ __forceinline__ void store_cols(float* p, const float (&v)[VEC]) {
  if (VEC == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1 % VEC], v[2 % VEC], v[3 % VEC]);
  } else {
    p[0] = v[0];
  }
}

template <int KMAX, int VEC>
__global__ void __launch_bounds__(256) bn_act_fwd_kernel(const BnActP p) {
  const int lane = threadIdx.x & 31;
  const int64_t warps_total = ((int64_t)gridDim.x * blockDim.x) >> 5;
  float mu[KMAX][VEC], sc[KMAX][VEC], sh[KMAX][VEC];
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    const int c0 = (k * 32 + lane) * VEC;
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      mu[k][j] = 0.f;
      sc[k][j] = 1.f;
      sh[k][j] = 0.f;
    }
    if (c0 < p.cols) {
      if (p.mean != nullptr) {
        float vr[VEC];
        load_cols<VEC>(p.mean + c0, mu[k]);
        load_cols<VEC>(p.var + c0, vr);
#pragma unroll
        for (int j = 0; j < VEC; ++j) sc[k][j] = 1.f / sqrtf(vr[j] + p.bn_eps);
      }
      if (p.gamma != nullptr) {
        float g[VEC];
        load_cols<VEC>(p.gamma + c0, g);
#pragma unroll
        for (int j = 0; j < VEC; ++j) sc[k][j] *= g[j];
      }
      if (p.beta != nullptr) load_cols<VEC>(p.beta + c0, sh[k]);
    }
  }
  const float alpha = (p.alpha != nullptr) ? __ldg(p.alpha) : 0.f;
  const bool drop = p.p_drop > 0.f;
  const float keep_scale = drop ? 1.f / (1.f - p.p_drop) : 1.f;
  const uint32_t counter = p.counter != nullptr ? (uint32_t)__float_as_int(__ldg(p.counter)) : 0u;
  const float inv_n = 1.f / (float)p.cols;

  for (int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < p.rows; row += warps_total) {
    float z[KMAX][VEC];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      const int c0 = (k * 32 + lane) * VEC;
#pragma unroll
      for (int j = 0; j < VEC; ++j) z[k][j] = 0.f;
      if (c0 < p.cols) {
        float hv[VEC];
        load_cols<VEC>(p.h + row * p.h_ld + c0, hv);
#pragma unroll
        for (int j = 0; j < VEC; ++j) z[k][j] = fmaf(hv[j] - mu[k][j], sc[k][j], sh[k][j]);
      }
    }
    float m = 0.f, inv_s = 0.f;
    if (p.act == ACT_DICE) {
      float t = 0.f;
#pragma unroll
      for (int k = 0; k < KMAX; ++k)
#pragma unroll
        for (int j = 0; j < VEC; ++j) t += z[k][j];  // padding lanes hold 0
      m = warp_sum(t) * inv_n;
      float qv = 0.f;
#pragma unroll
      for (int k = 0; k < KMAX; ++k) {
        if ((k * 32 + lane) * VEC < p.cols) {
#pragma unroll
          for (int j = 0; j < VEC; ++j) {
            const float d = z[k][j] - m;
            qv += fmaf(d, d, p.dice_eps);
          }
        }
      }
      inv_s = 1.f / sqrtf(warp_sum(qv));
    }
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      const int c0 = (k * 32 + lane) * VEC;
      if (c0 >= p.cols) continue;
      float o[VEC];
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        const float zz = z[k][j];
        float v;
        switch (p.act) {
          case ACT_RELU: v = fmaxf(zz, 0.f); break;
          case ACT_DICE: {
            const float ps = sigmoidf_precise((zz - m) * inv_s);
            v = ps * zz + (1.f - ps) * alpha * zz;
          } break;
          case ACT_PRELU: v = zz > 0.f ? zz : alpha * zz; break;
          case ACT_SIGMOID: v = sigmoidf_precise(zz); break;
          case ACT_LEAKY: v = zz > 0.f ? zz : 0.01f * zz; break;
          default: v = zz; break;
        }
        if (drop) v = dropout_keep(p.seed, counter, (uint64_t)row * p.cols + c0 + j, p.p_drop) ? v * keep_scale : 0.f;
        o[j] = v;
      }
      store_cols<VEC>(p.y + row * p.y_ld + c0, o);
    }
  }
}

// backward pass 1: d_z (stored in d_h) + column sums d_beta, d_gamma (+ d_alpha).
// In eval mode (training == 0) BN is an affine map, so d_h = d_z * gamma * rstd is final here.
template <int KMAX, int VEC>
__global__ void __launch_bounds__(256) bn_act_bwd_kernel(const BnActP p) {
  extern __shared__ float smem[];  // [warps][2*cols]
  __shared__ float sm_alpha[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int64_t warps_total = (int64_t)gridDim.x * wpb;
  float mu[KMAX][VEC], rstd[KMAX][VEC], gam[KMAX][VEC], bet[KMAX][VEC], acc_b[KMAX][VEC], acc_g[KMAX][VEC];
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    const int c0 = (k * 32 + lane) * VEC;
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      mu[k][j] = 0.f;
      rstd[k][j] = 1.f;
      gam[k][j] = 1.f;
      bet[k][j] = 0.f;
      acc_b[k][j] = acc_g[k][j] = 0.f;
    }
    if (c0 < p.cols) {
      if (p.mean != nullptr) {
        float vr[VEC];
        load_cols<VEC>(p.mean + c0, mu[k]);
        load_cols<VEC>(p.var + c0, vr);
#pragma unroll
        for (int j = 0; j < VEC; ++j) rstd[k][j] = 1.f / sqrtf(vr[j] + p.bn_eps);
      }
      if (p.gamma != nullptr) load_cols<VEC>(p.gamma + c0, gam[k]);
      if (p.beta != nullptr) load_cols<VEC>(p.beta + c0, bet[k]);
    }
  }
  const float alpha = (p.alpha != nullptr) ? __ldg(p.alpha) : 0.f;
  const bool drop = p.p_drop > 0.f;
  const float keep_scale = drop ? 1.f / (1.f - p.p_drop) : 1.f;
  const uint32_t counter = p.counter != nullptr ? (uint32_t)__float_as_int(__ldg(p.counter)) : 0u;
  const float inv_n = 1.f / (float)p.cols;
  float acc_alpha = 0.f;

  for (int64_t row = (int64_t)blockIdx.x * wpb + warp; row < p.rows; row += warps_total) {
    float xh[KMAX][VEC], z[KMAX][VEC], da[KMAX][VEC], dz[KMAX][VEC];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      const int c0 = (k * 32 + lane) * VEC;
#pragma unroll
      for (int j = 0; j < VEC; ++j) xh[k][j] = z[k][j] = da[k][j] = 0.f;
      if (c0 < p.cols) {
        float hv[VEC];
        load_cols<VEC>(p.h + row * p.h_ld + c0, hv);
        load_cols<VEC>(p.d_y + row * p.d_y_ld + c0, da[k]);
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          xh[k][j] = (hv[j] - mu[k][j]) * rstd[k][j];
          z[k][j] = fmaf(xh[k][j], gam[k][j], bet[k][j]);
          if (drop) da[k][j] = dropout_keep(p.seed, counter, (uint64_t)row * p.cols + c0 + j, p.p_drop) ? da[k][j] * keep_scale : 0.f;
        }
      }
    }
    if (p.act == ACT_DICE) {
      float t = 0.f;
#pragma unroll
      for (int k = 0; k < KMAX; ++k)
#pragma unroll
        for (int j = 0; j < VEC; ++j) t += z[k][j];
      const float m = warp_sum(t) * inv_n;
      float qv = 0.f;
#pragma unroll
      for (int k = 0; k < KMAX; ++k) {
        if ((k * 32 + lane) * VEC < p.cols) {
#pragma unroll
          for (int j = 0; j < VEC; ++j) {
            const float d = z[k][j] - m;
            qv += fmaf(d, d, p.dice_eps);
          }
        }
      }
      const float s = sqrtf(warp_sum(qv));
      const float inv_s = 1.f / s;
      // out_j = z_j (alpha + (1-alpha) p_j), p_j = sigmoid((z_j - m)/s)
      float a1 = 0.f, a2 = 0.f;
      float aj[KMAX][VEC], pj[KMAX][VEC];
#pragma unroll
      for (int k = 0; k < KMAX; ++k) {
        const bool on = (k * 32 + lane) * VEC < p.cols;
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          aj[k][j] = pj[k][j] = 0.f;
          if (on) {
            pj[k][j] = sigmoidf_precise((z[k][j] - m) * inv_s);
            aj[k][j] = da[k][j] * z[k][j] * (1.f - alpha) * pj[k][j] * (1.f - pj[k][j]);
            a1 += aj[k][j];
            a2 = fmaf(aj[k][j], z[k][j] - m, a2);
            acc_alpha = fmaf(da[k][j] * z[k][j], 1.f - pj[k][j], acc_alpha);
          }
        }
      }
      a1 = warp_sum(a1);
      a2 = warp_sum(a2);
      const float k1 = a1 * inv_n * inv_s, k2 = a2 * inv_s * inv_s * inv_s;
#pragma unroll
      for (int k = 0; k < KMAX; ++k)
#pragma unroll
        for (int j = 0; j < VEC; ++j) dz[k][j] = da[k][j] * (alpha + (1.f - alpha) * pj[k][j]) + aj[k][j] * inv_s - k1 - (z[k][j] - m) * k2;
    } else {
#pragma unroll
      for (int k = 0; k < KMAX; ++k) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          const float zz = z[k][j], g = da[k][j];
          float d;
          switch (p.act) {
            case ACT_RELU: d = zz > 0.f ? g : 0.f; break;
            case ACT_PRELU:
              d = zz > 0.f ? g : alpha * g;
              if (zz <= 0.f) acc_alpha = fmaf(g, zz, acc_alpha);
              break;
            case ACT_SIGMOID: {
              const float sg = sigmoidf_precise(zz);
              d = g * sg * (1.f - sg);
            } break;
            case ACT_LEAKY: d = zz > 0.f ? g : 0.01f * g; break;
            default: d = g; break;
          }
          dz[k][j] = d;
        }
      }
    }
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      const int c0 = (k * 32 + lane) * VEC;
      if (c0 < p.cols) {
        float o[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          acc_b[k][j] += dz[k][j];
          acc_g[k][j] = fmaf(dz[k][j], xh[k][j], acc_g[k][j]);
          // training: keep d_z for pass 2;  eval: BN is affine, finish now
          o[j] = p.training ? dz[k][j] : dz[k][j] * gam[k][j] * rstd[k][j];
        }
        store_cols<VEC>(p.d_h + row * p.d_h_ld + c0, o);
      }
    }
  }
  // block reduction of the column sums, one RED per column per block
  float* my = smem + (int64_t)warp * 2 * p.cols;
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    const int c0 = (k * 32 + lane) * VEC;
    if (c0 < p.cols) {
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        my[c0 + j] = acc_b[k][j];
        my[p.cols + c0 + j] = acc_g[k][j];
      }
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
template <int KMAX, int VEC>
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const BnActP p) {
  const int lane = threadIdx.x & 31;
  const int64_t warps_total = ((int64_t)gridDim.x * blockDim.x) >> 5;
  float mu[KMAX][VEC], rstd[KMAX][VEC], gs[KMAX][VEC], mb[KMAX][VEC], mg[KMAX][VEC];
  const float inv_rows = 1.f / (float)p.rows;
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    const int c0 = (k * 32 + lane) * VEC;
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      mu[k][j] = mb[k][j] = mg[k][j] = 0.f;
      rstd[k][j] = gs[k][j] = 1.f;
    }
    if (c0 < p.cols) {
      float vr[VEC], g[VEC];
      load_cols<VEC>(p.mean + c0, mu[k]);
      load_cols<VEC>(p.var + c0, vr);
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        rstd[k][j] = 1.f / sqrtf(vr[j] + p.bn_eps);
        g[j] = 1.f;
      }
      if (p.gamma != nullptr) load_cols<VEC>(p.gamma + c0, g);
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        gs[k][j] = g[j] * rstd[k][j];
        mb[k][j] = __ldcg(p.d_beta + c0 + j) * inv_rows;
        mg[k][j] = __ldcg(p.d_gamma + c0 + j) * inv_rows;
      }
    }
  }
  for (int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < p.rows; row += warps_total) {
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      const int c0 = (k * 32 + lane) * VEC;
      if (c0 < p.cols) {
        float hv[VEC], dz[VEC];
        load_cols<VEC>(p.h + row * p.h_ld + c0, hv);
        float* dst = p.d_h + row * p.d_h_ld + c0;
        if (VEC == 4) {
          const float4 t = *reinterpret_cast<const float4*>(dst);
          dz[0] = t.x; dz[1 % VEC] = t.y; dz[2 % VEC] = t.z; dz[3 % VEC] = t.w;
        } else {
          dz[0] = dst[0];
        }
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          const float xh = (hv[j] - mu[k][j]) * rstd[k][j];
          dz[j] = gs[k][j] * (dz[j] - mb[k][j] - xh * mg[k][j]);
        }
        store_cols<VEC>(dst, dz);
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

// ---- multi-tensor optimiser step for the dense parameters ---------------------------------------
constexpr int kMaxDenseTensors = 96;
struct DenseOptP {
  float* p[kMaxDenseTensors];
  const float* g[kMaxDenseTensors];
  float* s1[kMaxDenseTensors];
  float* s2[kMaxDenseTensors];
  int32_t n[kMaxDenseTensors];
};

__global__ void __launch_bounds__(256) dense_update_kernel(const __grid_constant__ DenseOptP t, int kind, float beta1, float beta2, float eps,
                                                           float wd, const float* __restrict__ lr_dev, const float* __restrict__ bc_dev) {
  pdl_wait();
  const int ti = blockIdx.y;
  const int n = t.n[ti];
  const float lr = *lr_dev;
  const float bc1 = kind == 1 ? bc_dev[0] : 1.f, bc2s = kind == 1 ? bc_dev[1] : 1.f;
  float* p = t.p[ti];
  const float* g = t.g[ti];
  float* s1 = t.s1[ti];
  float* s2 = t.s2[ti];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float w = p[i];
    float gr = fmaf(wd, w, g[i]);
    if (kind == 0) {
      w -= lr * gr;
    } else if (kind == 1) {
      const float m = beta1 * s1[i] + (1.f - beta1) * gr;
      const float v = beta2 * s2[i] + (1.f - beta2) * gr * gr;
      s1[i] = m;
      s2[i] = v;
      w -= (lr / bc1) * (m / (sqrtf(v) / bc2s + eps));
    } else {
      const float acc = s1[i] + gr * gr;
      s1[i] = acc;
      w -= lr * gr / (sqrtf(acc) + eps);
    }
    p[i] = w;
  }
}

}  // namespace rh

using namespace rh;

static bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

extern "C" int rh_colstats(const float* h, int64_t h_ld, int64_t rows, int cols, float* stats, float* scratch, float* running_mean,
                           float* running_var, int64_t* num_batches_tracked, float momentum, void* stream) {
  RH_REQUIRE(h && stats && scratch, RH_ERR_INVALID_ARG, "rh_colstats: NULL pointer");
  RH_REQUIRE(rows > 0 && cols > 0 && h_ld >= cols, RH_ERR_INVALID_ARG, "rh_colstats: bad sizes");
  const bool vec = cols % 4 == 0 && h_ld % 4 == 0 && al16(h);
  const int per_tile = vec ? 128 : 32;
  const int col_tiles = (cols + per_tile - 1) / per_tile;
  // aim at ~2 blocks per SM; at least 64 rows per block so the partial sums amortise the reduction
  int64_t want_chunks = ((int64_t)num_sms() * 2 + col_tiles - 1) / col_tiles;
  int64_t rpb = (rows + want_chunks - 1) / want_chunks;
  if (rpb < 64) rpb = 64;
  const int64_t chunks = (rows + rpb - 1) / rpb;
  RH_REQUIRE(chunks <= 65535, RH_ERR_UNSUPPORTED, "rh_colstats: too many row chunks");
  dim3 grid(col_tiles, (unsigned)chunks);
  if (vec) {
    colstats_kernel<4><<<grid, 256, 0, (cudaStream_t)stream>>>(h, h_ld, rows, cols, rpb, stats, scratch, running_mean, running_var,
                                                               reinterpret_cast<long long*>(num_batches_tracked), momentum);
  } else {
    colstats_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>(h, h_ld, rows, cols, rpb, stats, scratch, running_mean, running_var,
                                                               reinterpret_cast<long long*>(num_batches_tracked), momentum);
  }
  RH_LAUNCH_CHECK();
  return RH_OK;
}

// K steps of 32 lanes x VEC columns cover the row
#define RH_DISPATCH_BN(cols, vec, CALL)                                                   \
  do {                                                                                    \
    const int kk__ = ((cols) + ((vec) ? 128 : 32) - 1) / ((vec) ? 128 : 32);              \
    if (vec) {                                                                            \
      if (kk__ <= 1) { CALL(1, 4); }                                                      \
      else if (kk__ <= 2) { CALL(2, 4); }                                                 \
      else if (kk__ <= 4) { CALL(4, 4); }                                                 \
      else if (kk__ <= 8) { CALL(8, 4); }                                                 \
      else { set_error("bn_act: %d columns > 1024 not supported", (cols)); return RH_ERR_UNSUPPORTED; } \
    } else {                                                                              \
      if (kk__ <= 1) { CALL(1, 1); }                                                      \
      else if (kk__ <= 2) { CALL(2, 1); }                                                 \
      else if (kk__ <= 4) { CALL(4, 1); }                                                 \
      else if (kk__ <= 8) { CALL(8, 1); }                                                 \
      else if (kk__ <= 16) { CALL(16, 1); }                                               \
      else if (kk__ <= 32) { CALL(32, 1); }                                               \
      else { set_error("bn_act: %d columns > 1024 not supported", (cols)); return RH_ERR_UNSUPPORTED; } \
    }                                                                                     \
  } while (0)

extern "C" int rh_bn_act_fwd(const float* h, int64_t h_ld, int64_t rows, int cols, const float* mean, const float* var, float bn_eps,
                             const float* gamma, const float* beta, int act, const float* act_param, float dice_eps, float p_drop,
                             uint32_t dropout_seed, const float* dropout_counter, float* y, int64_t y_ld, void* stream) {
  RH_REQUIRE(h && y, RH_ERR_INVALID_ARG, "rh_bn_act_fwd: NULL pointer");
  RH_REQUIRE((mean == nullptr) == (var == nullptr), RH_ERR_INVALID_ARG, "rh_bn_act_fwd: mean and var go together");
  RH_REQUIRE(rows >= 0 && cols > 0 && h_ld >= cols && y_ld >= cols, RH_ERR_INVALID_ARG, "rh_bn_act_fwd: bad sizes");
  RH_REQUIRE(act >= 0 && act <= 5, RH_ERR_INVALID_ARG, "rh_bn_act_fwd: act %d unknown", act);
  RH_REQUIRE(!((act == ACT_DICE || act == ACT_PRELU) && act_param == nullptr), RH_ERR_INVALID_ARG, "rh_bn_act_fwd: Dice/PReLU need act_param");
  RH_REQUIRE(p_drop >= 0.f && p_drop < 1.f, RH_ERR_INVALID_ARG, "rh_bn_act_fwd: p_drop must be in [0,1)");
  if (rows == 0) return RH_OK;
  BnActP p;
  memset(&p, 0, sizeof(p));
  p.h = h; p.h_ld = h_ld; p.rows = rows; p.cols = cols; p.mean = mean; p.var = var; p.bn_eps = bn_eps;
  p.gamma = gamma; p.beta = beta; p.act = act; p.alpha = act_param; p.dice_eps = dice_eps;
  p.p_drop = p_drop; p.seed = dropout_seed; p.counter = dropout_counter; p.y = y; p.y_ld = y_ld;
  const bool vec = cols % 4 == 0 && h_ld % 4 == 0 && y_ld % 4 == 0 && al16(h) && al16(y) && (!mean || (al16(mean) && al16(var))) &&
                   (!gamma || al16(gamma)) && (!beta || al16(beta));
  const int grid = rows_grid((rows + 3) / 4, 8, 4);
#define RH_CALL(K, V) bn_act_fwd_kernel<K, V><<<grid, 256, 0, (cudaStream_t)stream>>>(p)
  RH_DISPATCH_BN(cols, vec, RH_CALL);
#undef RH_CALL
  RH_LAUNCH_CHECK();
  return RH_OK;
}

extern "C" int rh_bn_act_bwd(const float* h, int64_t h_ld, int64_t rows, int cols, const float* mean, const float* var, float bn_eps,
                             const float* gamma, const float* beta, int act, const float* act_param, float dice_eps, float p_drop,
                             uint32_t dropout_seed, const float* dropout_counter, const float* d_y, int64_t d_y_ld, int training,
                             float* d_h, int64_t d_h_ld, float* d_gamma, float* d_beta, float* d_act_param, void* stream) {
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
  p.p_drop = p_drop; p.seed = dropout_seed; p.counter = dropout_counter;
  p.d_y = d_y; p.d_y_ld = d_y_ld; p.training = bn_train ? 1 : 0;
  p.d_h = d_h; p.d_h_ld = d_h_ld; p.d_gamma = d_gamma; p.d_beta = d_beta; p.d_alpha = d_act_param;
  const bool vec = cols % 4 == 0 && h_ld % 4 == 0 && d_y_ld % 4 == 0 && d_h_ld % 4 == 0 && al16(h) && al16(d_y) && al16(d_h) &&
                   (!mean || (al16(mean) && al16(var))) && (!gamma || al16(gamma)) && (!beta || al16(beta));
  cudaStream_t st = (cudaStream_t)stream;
  const int grid1 = rows_grid((rows + 3) / 4, 8, 2);
  const size_t smem = (size_t)8 * 2 * cols * sizeof(float);
#define RH_CALL(K, V)                                                                                                              \
  do {                                                                                                                             \
    if (smem > 48 * 1024) cudaFuncSetAttribute(bn_act_bwd_kernel<K, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   \
    bn_act_bwd_kernel<K, V><<<grid1, 256, smem, st>>>(p);                                                                          \
  } while (0)
  RH_DISPATCH_BN(cols, vec, RH_CALL);
#undef RH_CALL
  RH_LAUNCH_CHECK();
  if (bn_train) {
    const int grid2 = rows_grid((rows + 3) / 4, 8, 4);
#define RH_CALL(K, V) bn_bwd_apply_kernel<K, V><<<grid2, 256, 0, st>>>(p)
    RH_DISPATCH_BN(cols, vec, RH_CALL);
#undef RH_CALL
    RH_LAUNCH_CHECK();
  }
  return RH_OK;
}

extern "C" int rh_dense_update(int n_tensors, float* const* params, const float* const* grads, float* const* state1, float* const* state2,
                               const int64_t* numel, int kind, const float* lr_dev, const float* bias_corr_dev, float beta1, float beta2,
                               float eps, float weight_decay, void* stream) {
  RH_REQUIRE(n_tensors >= 0 && (n_tensors == 0 || (params && grads && numel)), RH_ERR_INVALID_ARG, "rh_dense_update: NULL pointer");
  RH_REQUIRE(kind >= 0 && kind <= 2, RH_ERR_INVALID_ARG, "rh_dense_update: kind %d unknown", kind);
  RH_REQUIRE(lr_dev != nullptr && (kind != 1 || bias_corr_dev != nullptr), RH_ERR_INVALID_ARG, "rh_dense_update: lr/bias-correction NULL");
  RH_REQUIRE(kind == 0 || state1 != nullptr, RH_ERR_INVALID_ARG, "rh_dense_update: state1 required");
  RH_REQUIRE(kind != 1 || state2 != nullptr, RH_ERR_INVALID_ARG, "rh_dense_update: state2 required for Adam");
  static thread_local DenseOptP t;
  for (int i0 = 0; i0 < n_tensors; i0 += kMaxDenseTensors) {
    const int cnt = n_tensors - i0 < kMaxDenseTensors ? n_tensors - i0 : kMaxDenseTensors;
    memset(&t, 0, sizeof(t));
    int64_t biggest = 0;
    for (int i = 0; i < cnt; ++i) {
      RH_REQUIRE(params[i0 + i] && grads[i0 + i], RH_ERR_INVALID_ARG, "rh_dense_update: tensor %d NULL", i0 + i);
      RH_REQUIRE(numel[i0 + i] >= 0 && numel[i0 + i] < ((int64_t)1 << 31), RH_ERR_INVALID_ARG, "rh_dense_update: tensor %d too large", i0 + i);
      t.p[i] = params[i0 + i];
      t.g[i] = grads[i0 + i];
      t.s1[i] = state1 ? state1[i0 + i] : nullptr;
      t.s2[i] = state2 ? state2[i0 + i] : nullptr;
      RH_REQUIRE(kind == 0 || t.s1[i], RH_ERR_INVALID_ARG, "rh_dense_update: state1[%d] NULL", i0 + i);
      RH_REQUIRE(kind != 1 || t.s2[i], RH_ERR_INVALID_ARG, "rh_dense_update: state2[%d] NULL", i0 + i);
      t.n[i] = (int32_t)numel[i0 + i];
      if (numel[i0 + i] > biggest) biggest = numel[i0 + i];
    }
    int gx = (int)((biggest + 255) / 256);  // one element per thread for tower-sized tensors: the update is a latency chain, not a stream
    if (gx > 1024) gx = 1024;
    if (gx < 1) gx = 1;
    launch_k(dense_update_kernel, dim3(gx, cnt), dim3(256), 0, (cudaStream_t)stream, t, kind, beta1, beta2, eps, weight_decay, lr_dev, bias_corr_dev);
    RH_LAUNCH_CHECK();
  }
  return RH_OK;
}
