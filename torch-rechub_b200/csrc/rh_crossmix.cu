// rh_crossmix.cu — DCN-v2's mixture of low-rank cross experts (CrossNetMix) and the full-rank CrossNetV2 as engine kernels.
//
// Reference arithmetic replaced: CrossNetMix.forward (basic/layers.py:470-506) — per layer l and expert e
//     v = tanh(V_e^T x_l);  v = tanh(C_e v);  uv = U_e v;  out_e = x_0 * (uv + bias_l);   gate_e = <g_e, x_l>
//     x_{l+1} = sum_e softmax(gate)_e out_e + x_l
// — a Python double loop of ~60 launches per layer there.  Here a layer is three tensor-core GEMMs (rh_gemm_tf32x3) over
// PACKED operands and three small fused maps:
//     [A | G] = x_l [V_1 .. V_E | g_1 .. g_E]              one (B, W) x (W, E r + E) product: every expert's projection + the gates
//     t1 = tanh(A), s = softmax(G)                          rh_crossmix_mid1_fwd
//     P  = t1 blockdiag(C_e)^T                               one (B, E r) x (E r, E r) product
//     t2 = tanh(P), z = [s_e t2_e]                           rh_crossmix_mid2_fwd   (the gate-weighted sum moves into the K dimension
//     u  = z [U_1 .. U_E]^T                                                          of the last product: sum_e s_e = 1 keeps the bias whole)
//     x_{l+1} = x_0 * (u + bias_l) + x_l                     rh_crossmix_out_fwd    (also CrossNetV2's  x_0 * (W x_l) + b + x_l, :440-444)
// and the hand-derived backward mirrors it (six GEMMs + three maps per layer).  rh_crossmix_pack builds the packed operands of
// all layers from the reference's parameter layout (u_list / v_list: (E, W, r), c_list: (E, r, r), gating[e].weight: (1, W)) in one
// launch per step; rh_crossmix_unpack_grads scatters the packed gradients back (the gate modules are shared by all layers,
// basic/layers.py:466, so their gradients accumulate over layers).
// All maps are HBM/L2-bound streaming kernels over (B, <= 432) activations.
#include "rh_common.cuh"

namespace rh {

constexpr int kMaxCrossLayers = 8;
constexpr int kMaxExperts = 8;

struct CrossMixPtrs {
  const float* u[kMaxCrossLayers];
  const float* v[kMaxCrossLayers];
  const float* c[kMaxCrossLayers];
  const float* gate[kMaxExperts];
  float* wcat[kMaxCrossLayers];  // (E r + E, ld_w): rows (e, j) = V_e[:, j], rows E r + e = gate_e
  float* cbd[kMaxCrossLayers];   // (E r, E r): block diagonal of C_e
  float* ucat[kMaxCrossLayers];  // (W, E r): [w, (e, j)] = U_e[w, j]
};

__global__ void __launch_bounds__(256) crossmix_pack_kernel(const __grid_constant__ CrossMixPtrs p, int L, int E, int W, int r, int ld_w) {
  const int l = blockIdx.y;
  const int Er = E * r, N1 = Er + E;
  const int64_t n_w = (int64_t)N1 * ld_w, n_c = (int64_t)Er * Er, n_u = (int64_t)W * Er;
  const int64_t total = n_w + n_c + n_u;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    if (i < n_w) {
      const int n = (int)(i / ld_w), k = (int)(i - (int64_t)n * ld_w);
      float val = 0.f;
      if (k < W) {
        if (n < Er) {
          const int e = n / r, j = n - e * r;
          val = __ldg(p.v[l] + ((int64_t)e * W + k) * r + j);
        } else {
          val = __ldg(p.gate[n - Er] + k);
        }
      }
      p.wcat[l][i] = val;
    } else if (i < n_w + n_c) {
      const int64_t t = i - n_w;
      const int row = (int)(t / Er), col = (int)(t - (int64_t)row * Er);
      const int e = row / r, e2 = col / r;
      p.cbd[l][t] = e == e2 ? __ldg(p.c[l] + ((int64_t)e * r + (row - e * r)) * r + (col - e * r)) : 0.f;
    } else {
      const int64_t t = i - n_w - n_c;
      const int w = (int)(t / Er), n = (int)(t - (int64_t)w * Er);
      const int e = n / r, j = n - e * r;
      p.ucat[l][t] = __ldg(p.u[l] + ((int64_t)e * W + w) * r + j);
    }
  }
}

struct CrossMixGradPtrs {
  const float* d_wcat[kMaxCrossLayers];  // (E r + E, ld_dw)
  const float* d_cbd[kMaxCrossLayers];   // (E r, ld_dc)
  const float* d_ucat[kMaxCrossLayers];  // (W, ld_du)
  float* d_u[kMaxCrossLayers];
  float* d_v[kMaxCrossLayers];
  float* d_c[kMaxCrossLayers];
  float* d_gate[kMaxExperts];
};

__global__ void __launch_bounds__(256) crossmix_unpack_kernel(const __grid_constant__ CrossMixGradPtrs p, int L, int E, int W, int r, int ld_dw, int ld_dc,
                                                              int ld_du) {
  const int Er = E * r;
  const int64_t n_uv = (int64_t)E * W * r, n_c = (int64_t)E * r * r, n_g = (int64_t)E * W;
  const int l = blockIdx.y;  // blockIdx.y == L: the shared gates (sum over layers)
  if (l < L) {
    const int64_t total = 2 * n_uv + n_c;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
      if (i < n_uv) {  // d_v[e, w, j] = d_wcat[(e, j), w]
        const int e = (int)(i / ((int64_t)W * r)), rem = (int)(i - (int64_t)e * W * r), w = rem / r, j = rem - w * r;
        p.d_v[l][i] = __ldg(p.d_wcat[l] + (int64_t)(e * r + j) * ld_dw + w);
      } else if (i < 2 * n_uv) {  // d_u[e, w, j] = d_ucat[w, (e, j)]
        const int64_t t = i - n_uv;
        const int e = (int)(t / ((int64_t)W * r)), rem = (int)(t - (int64_t)e * W * r), w = rem / r, j = rem - w * r;
        p.d_u[l][t] = __ldg(p.d_ucat[l] + (int64_t)w * ld_du + e * r + j);
      } else {  // d_c[e, i, j] = d_cbd[(e, i), (e, j)]
        const int64_t t = i - 2 * n_uv;
        const int e = (int)(t / (r * r)), rem = (int)(t - (int64_t)e * r * r), a = rem / r, b = rem - a * r;
        p.d_c[l][t] = __ldg(p.d_cbd[l] + (int64_t)(e * r + a) * ld_dc + e * r + b);
      }
    }
  } else {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_g; i += (int64_t)gridDim.x * blockDim.x) {
      const int e = (int)(i / W), w = (int)(i - (int64_t)e * W);
      float t = 0.f;
      for (int ll = 0; ll < L; ++ll) t += __ldg(p.d_wcat[ll] + (int64_t)(Er + e) * ld_dw + w);
      p.d_gate[e][w] = t;
    }
  }
}

// t1 = tanh(ag[:, :Er]);  s = softmax(ag[:, Er:Er+E])
__global__ void __launch_bounds__(256) crossmix_mid1_fwd_kernel(const float* __restrict__ ag, int64_t ld_ag, int64_t B, int E, int Er,
                                                                float* __restrict__ t1, float* __restrict__ s) {
  const int64_t total = B * Er;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / Er;
    const int c = (int)(i - b * Er);
    t1[i] = tanhf(__ldg(ag + b * ld_ag + c));
    if (c == 0) {
      const float* g = ag + b * ld_ag + Er;
      float mx = -INFINITY;
      for (int e = 0; e < E; ++e) mx = fmaxf(mx, __ldg(g + e));
      float ex[kMaxExperts], sum = 0.f;
      for (int e = 0; e < E; ++e) {
        ex[e] = expf(__ldg(g + e) - mx);
        sum += ex[e];
      }
      for (int e = 0; e < E; ++e) s[b * E + e] = ex[e] / sum;
    }
  }
}

// t2 = tanh(P);  z[:, (e, j)] = s[:, e] * t2[:, (e, j)]
__global__ void __launch_bounds__(256) crossmix_mid2_fwd_kernel(const float* __restrict__ P, const float* __restrict__ s, int64_t B, int E, int r,
                                                                float* __restrict__ t2, float* __restrict__ z) {
  const int Er = E * r;
  const int64_t total = B * Er;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / Er;
    const int c = (int)(i - b * Er);
    const float t = tanhf(__ldg(P + i));
    t2[i] = t;
    z[i] = t * __ldg(s + b * E + c / r);
  }
}

// bias_outside == 0: out = x0 * (u + bias) + xl  (CrossNetMix);   != 0: out = x0 * u + bias + xl  (CrossNetV2).   bias may be NULL
__global__ void __launch_bounds__(256) crossmix_out_fwd_kernel(const float* __restrict__ x0, int64_t ld0, const float* __restrict__ xl, int64_t ldl,
                                                               const float* __restrict__ u, int64_t ldu, const float* __restrict__ bias, int bias_outside,
                                                               int64_t B, int W, float* __restrict__ out, int64_t ldo) {
  const int64_t total = B * W;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / W;
    const int w = (int)(i - b * W);
    const float bb = bias != nullptr ? __ldg(bias + w) : 0.f;
    const float uv = __ldg(u + b * ldu + w);
    out[b * ldo + w] = bias_outside ? fmaf(__ldg(x0 + b * ld0 + w), uv, bb + __ldg(xl + b * ldl + w)) : fmaf(__ldg(x0 + b * ld0 + w), uv + bb, __ldg(xl + b * ldl + w));
  }
}

// g = g1 (+ g2);  g_sum = g;  d_u = g * x0;  d_x0_acc += g * (u + bias);  d_bias[w] += sum_b d_u[b, w]
// (bias_outside: d_x0_acc += g * u;  d_bias[w] += sum_b g[b, w]).   block = 32 rows; thread = column w, w + 256, ...
__global__ void __launch_bounds__(256) crossmix_out_bwd_kernel(const float* __restrict__ g1, int64_t ldg1, const float* __restrict__ g2, int64_t ldg2,
                                                               const float* __restrict__ x0, int64_t ld0, const float* __restrict__ u, int64_t ldu,
                                                               const float* __restrict__ bias, int bias_outside, int64_t B, int W, float* __restrict__ g_sum,
                                                               int64_t ldgs,
                                                               float* __restrict__ d_u, int64_t lddu, float* __restrict__ d_x0_acc, int64_t ldx,
                                                               float* __restrict__ d_bias) {
  const int64_t r0 = (int64_t)blockIdx.x * 32;
  const int64_t r1 = r0 + 32 < B ? r0 + 32 : B;
  for (int w = threadIdx.x; w < W; w += blockDim.x) {
    const float bb = bias != nullptr ? __ldg(bias + w) : 0.f;
    float acc = 0.f;
    for (int64_t b = r0; b < r1; ++b) {
      float g = __ldg(g1 + b * ldg1 + w);
      if (g2 != nullptr) g += __ldg(g2 + b * ldg2 + w);
      if (g_sum != nullptr) g_sum[b * ldgs + w] = g;
      const float du = g * __ldg(x0 + b * ld0 + w);
      d_u[b * lddu + w] = du;
      acc += bias_outside ? g : du;
      d_x0_acc[b * ldx + w] += g * (__ldg(u + b * ldu + w) + (bias_outside ? 0.f : bb));
    }
    if (d_bias != nullptr) atomicAdd(d_bias + w, acc);
  }
}

// warp = row.  d_t2 = d_z * s_e;  d_P = d_t2 * (1 - t2^2);  d_s_e = sum_j d_z[(e, j)] t2[(e, j)];  d_gate = s * (d_s - <s, d_s>)
__global__ void __launch_bounds__(256) crossmix_mid2_bwd_kernel(const float* __restrict__ d_z, int64_t ld_dz, const float* __restrict__ s,
                                                                const float* __restrict__ t2, int64_t B, int E, int r, float* __restrict__ d_P,
                                                                float* __restrict__ d_ag, int64_t ld_dag) {
  const int lane = threadIdx.x & 31;
  const int Er = E * r;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; b < B; b += warps) {
    float ds[kMaxExperts], sv[kMaxExperts];
    float dot = 0.f;
    for (int e = 0; e < E; ++e) {
      sv[e] = __ldg(s + b * E + e);
      float part = 0.f;
      for (int j = lane; j < r; j += 32) {
        const int c = e * r + j;
        const float dz = __ldg(d_z + b * ld_dz + c), t = __ldg(t2 + b * Er + c);
        part = fmaf(dz, t, part);
        d_P[b * Er + c] = dz * sv[e] * (1.f - t * t);
      }
      ds[e] = warp_sum(part);
      dot = fmaf(sv[e], ds[e], dot);
    }
    if (lane < E) {
      float dsl = 0.f, sl = 0.f;
      for (int e = 0; e < E; ++e) {  // select without dynamic register indexing
        if (e == lane) {
          dsl = ds[e];
          sl = sv[e];
        }
      }
      d_ag[b * ld_dag + Er + lane] = sl * (dsl - dot);
    }
  }
}

// d_ag[:, :Er] = d_t1 * (1 - t1^2)
__global__ void __launch_bounds__(256) crossmix_mid1_bwd_kernel(const float* __restrict__ d_t1, int64_t ld_dt1, const float* __restrict__ t1, int64_t B,
                                                                int Er, float* __restrict__ d_ag, int64_t ld_dag) {
  const int64_t total = B * Er;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / Er;
    const int c = (int)(i - b * Er);
    const float t = __ldg(t1 + i);
    d_ag[b * ld_dag + c] = __ldg(d_t1 + b * ld_dt1 + c) * (1.f - t * t);
  }
}

// out = a + b + c (any of b, c may be NULL)
__global__ void __launch_bounds__(256) sum3_kernel(const float* __restrict__ a, int64_t lda, const float* __restrict__ b, int64_t ldb,
                                                   const float* __restrict__ c, int64_t ldc, int64_t B, int W, float* __restrict__ out, int64_t ldo) {
  const int64_t total = B * W;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / W;
    const int w = (int)(i - row * W);
    float v = __ldg(a + row * lda + w);
    if (b != nullptr) v += __ldg(b + row * ldb + w);
    if (c != nullptr) v += __ldg(c + row * ldc + w);
    out[row * ldo + w] = v;
  }
}

static int ew_grid(int64_t total) {
  int64_t g = (total + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 8;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace rh

using namespace rh;

extern "C" int rh_crossmix_pack(int n_layers, int n_experts, int width, int rank, const float* const* u, const float* const* v, const float* const* c,
                                const float* const* gate, float* const* wcat, int64_t ld_w, float* const* cbd, float* const* ucat, void* stream) {
  RH_REQUIRE(n_layers > 0 && n_layers <= kMaxCrossLayers && n_experts > 0 && n_experts <= kMaxExperts && width > 0 && rank > 0, RH_ERR_UNSUPPORTED,
             "rh_crossmix_pack: %d layers / %d experts outside [1,%d] / [1,%d]", n_layers, n_experts, kMaxCrossLayers, kMaxExperts);
  RH_REQUIRE(u && v && c && gate && wcat && cbd && ucat && ld_w >= width, RH_ERR_INVALID_ARG, "rh_crossmix_pack: NULL pointer or ld_w < width");
  static thread_local CrossMixPtrs p;
  memset(&p, 0, sizeof(p));
  for (int l = 0; l < n_layers; ++l) {
    RH_REQUIRE(u[l] && v[l] && c[l] && wcat[l] && cbd[l] && ucat[l], RH_ERR_INVALID_ARG, "rh_crossmix_pack: layer %d has a NULL pointer", l);
    p.u[l] = u[l]; p.v[l] = v[l]; p.c[l] = c[l]; p.wcat[l] = wcat[l]; p.cbd[l] = cbd[l]; p.ucat[l] = ucat[l];
  }
  for (int e = 0; e < n_experts; ++e) {
    RH_REQUIRE(gate[e] != nullptr, RH_ERR_INVALID_ARG, "rh_crossmix_pack: gate %d is NULL", e);
    p.gate[e] = gate[e];
  }
  const int64_t Er = (int64_t)n_experts * rank;
  const int64_t total = (Er + n_experts) * ld_w + Er * Er + (int64_t)width * Er;
  dim3 grid(ew_grid(total), n_layers);
  crossmix_pack_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p, n_layers, n_experts, width, rank, (int)ld_w);
  RH_LAUNCH_CHECK();
  return RH_OK;
}

extern "C" int rh_crossmix_unpack_grads(int n_layers, int n_experts, int width, int rank, const float* const* d_wcat, int64_t ld_dw,
                                        const float* const* d_cbd, int64_t ld_dc, const float* const* d_ucat, int64_t ld_du, float* const* d_u,
                                        float* const* d_v, float* const* d_c, float* const* d_gate, void* stream) {
  RH_REQUIRE(n_layers > 0 && n_layers <= kMaxCrossLayers && n_experts > 0 && n_experts <= kMaxExperts && width > 0 && rank > 0, RH_ERR_UNSUPPORTED,
             "rh_crossmix_unpack_grads: layers / experts out of range");
  RH_REQUIRE(d_wcat && d_cbd && d_ucat && d_u && d_v && d_c && d_gate, RH_ERR_INVALID_ARG, "rh_crossmix_unpack_grads: NULL pointer");
  static thread_local CrossMixGradPtrs p;
  memset(&p, 0, sizeof(p));
  for (int l = 0; l < n_layers; ++l) {
    RH_REQUIRE(d_wcat[l] && d_cbd[l] && d_ucat[l] && d_u[l] && d_v[l] && d_c[l], RH_ERR_INVALID_ARG, "rh_crossmix_unpack_grads: layer %d NULL", l);
    p.d_wcat[l] = d_wcat[l]; p.d_cbd[l] = d_cbd[l]; p.d_ucat[l] = d_ucat[l]; p.d_u[l] = d_u[l]; p.d_v[l] = d_v[l]; p.d_c[l] = d_c[l];
  }
  for (int e = 0; e < n_experts; ++e) {
    RH_REQUIRE(d_gate[e] != nullptr, RH_ERR_INVALID_ARG, "rh_crossmix_unpack_grads: gate %d NULL", e);
    p.d_gate[e] = d_gate[e];
  }
  const int64_t total = 2 * (int64_t)n_experts * width * rank + (int64_t)n_experts * rank * rank;
  dim3 grid(ew_grid(total), n_layers + 1);
  crossmix_unpack_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p, n_layers, n_experts, width, rank, (int)ld_dw, (int)ld_dc, (int)ld_du);
  RH_LAUNCH_CHECK();
  return RH_OK;
}

extern "C" int rh_crossmix_mid1_fwd(const float* ag, int64_t ld_ag, int64_t batch, int n_experts, int rank, float* t1, float* s, void* stream) {
  RH_REQUIRE(ag && t1 && s && n_experts > 0 && n_experts <= kMaxExperts && rank > 0 && ld_ag >= (int64_t)n_experts * rank + n_experts, RH_ERR_INVALID_ARG,
             "rh_crossmix_mid1_fwd: bad arguments");
  if (batch <= 0) return RH_OK;
  const int Er = n_experts * rank;
  crossmix_mid1_fwd_kernel<<<ew_grid(batch * Er), 256, 0, (cudaStream_t)stream>>>(ag, ld_ag, batch, n_experts, Er, t1, s);
  RH_LAUNCH_CHECK();
  return RH_OK;
}

extern "C" int rh_crossmix_mid2_fwd(const float* P, const float* s, int64_t batch, int n_experts, int rank, float* t2, float* z, void* stream) {
  RH_REQUIRE(P && s && t2 && z && n_experts > 0 && n_experts <= kMaxExperts && rank > 0, RH_ERR_INVALID_ARG, "rh_crossmix_mid2_fwd: bad arguments");
  if (batch <= 0) return RH_OK;
  crossmix_mid2_fwd_kernel<<<ew_grid(batch * n_experts * rank), 256, 0, (cudaStream_t)stream>>>(P, s, batch, n_experts, rank, t2, z);
  RH_LAUNCH_CHECK();
  return RH_OK;
}

extern "C" int rh_crossmix_out_fwd(const float* x0, int64_t ld0, const float* xl, int64_t ldl, const float* u, int64_t ldu, const float* bias,
                                   int bias_outside, int64_t batch, int width, float* out, int64_t ldo, void* stream) {
  RH_REQUIRE(x0 && xl && u && out && width > 0 && ld0 >= width && ldl >= width && ldu >= width && ldo >= width, RH_ERR_INVALID_ARG,
             "rh_crossmix_out_fwd: bad arguments");
  if (batch <= 0) return RH_OK;
  crossmix_out_fwd_kernel<<<ew_grid(batch * width), 256, 0, (cudaStream_t)stream>>>(x0, ld0, xl, ldl, u, ldu, bias, bias_outside, batch, width, out, ldo);
  RH_LAUNCH_CHECK();
  return RH_OK;
}

extern "C" int rh_crossmix_out_bwd(const float* g1, int64_t ldg1, const float* g2, int64_t ldg2, const float* x0, int64_t ld0, const float* u,
                                   int64_t ldu, const float* bias, int bias_outside, int64_t batch, int width, float* g_sum, int64_t ldgs, float* d_u,
                                   int64_t lddu, float* d_x0_acc, int64_t ldx, float* d_bias, void* stream) {
  RH_REQUIRE(g1 && x0 && u && d_u && d_x0_acc && width > 0, RH_ERR_INVALID_ARG, "rh_crossmix_out_bwd: NULL pointer");
  RH_REQUIRE(ldg1 >= width && ld0 >= width && ldu >= width && lddu >= width && ldx >= width && (g2 == nullptr || ldg2 >= width) &&
                 (g_sum == nullptr || ldgs >= width),
             RH_ERR_INVALID_ARG, "rh_crossmix_out_bwd: leading dimension < width");
  if (batch <= 0) return RH_OK;
  const int grid = (int)((batch + 31) / 32);
  crossmix_out_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(g1, ldg1, g2, ldg2, x0, ld0, u, ldu, bias, bias_outside, batch, width, g_sum, ldgs, d_u, lddu,
                                                                   d_x0_acc, ldx, d_bias);
  RH_LAUNCH_CHECK();
  return RH_OK;
}

extern "C" int rh_crossmix_mid2_bwd(const float* d_z, int64_t ld_dz, const float* s, const float* t2, int64_t batch, int n_experts, int rank, float* d_P,
                                    float* d_ag, int64_t ld_dag, void* stream) {
  RH_REQUIRE(d_z && s && t2 && d_P && d_ag && n_experts > 0 && n_experts <= kMaxExperts && rank > 0, RH_ERR_INVALID_ARG, "rh_crossmix_mid2_bwd: bad arguments");
  RH_REQUIRE(ld_dz >= (int64_t)n_experts * rank && ld_dag >= (int64_t)n_experts * rank + n_experts, RH_ERR_INVALID_ARG, "rh_crossmix_mid2_bwd: leading dimension");
  if (batch <= 0) return RH_OK;
  int64_t g = (batch + 7) / 8;
  const int64_t cap = (int64_t)num_sms() * 8;
  if (g > cap) g = cap;
  crossmix_mid2_bwd_kernel<<<(int)g, 256, 0, (cudaStream_t)stream>>>(d_z, ld_dz, s, t2, batch, n_experts, rank, d_P, d_ag, ld_dag);
  RH_LAUNCH_CHECK();
  return RH_OK;
}

extern "C" int rh_crossmix_mid1_bwd(const float* d_t1, int64_t ld_dt1, const float* t1, int64_t batch, int n_experts, int rank, float* d_ag, int64_t ld_dag,
                                    void* stream) {
  RH_REQUIRE(d_t1 && t1 && d_ag && n_experts > 0 && rank > 0 && ld_dt1 >= (int64_t)n_experts * rank && ld_dag >= (int64_t)n_experts * rank, RH_ERR_INVALID_ARG,
             "rh_crossmix_mid1_bwd: bad arguments");
  if (batch <= 0) return RH_OK;
  const int Er = n_experts * rank;
  crossmix_mid1_bwd_kernel<<<ew_grid(batch * Er), 256, 0, (cudaStream_t)stream>>>(d_t1, ld_dt1, t1, batch, Er, d_ag, ld_dag);
  RH_LAUNCH_CHECK();
  return RH_OK;
}

extern "C" int rh_sum3(const float* a, int64_t lda, const float* b, int64_t ldb, const float* c, int64_t ldc, int64_t rows, int cols, float* out,
                       int64_t ldo, void* stream) {
  RH_REQUIRE(a && out && cols > 0 && lda >= cols && ldo >= cols && (b == nullptr || ldb >= cols) && (c == nullptr || ldc >= cols), RH_ERR_INVALID_ARG,
             "rh_sum3: bad arguments");
  if (rows <= 0) return RH_OK;
  sum3_kernel<<<ew_grid(rows * cols), 256, 0, (cudaStream_t)stream>>>(a, lda, b, ldb, c, ldc, rows, cols, out, ldo);
  RH_LAUNCH_CHECK();
  return RH_OK;
}
