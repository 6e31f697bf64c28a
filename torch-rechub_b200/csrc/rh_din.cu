// rh_din.cu — DIN target attention around its MLP: behaviour-sequence gather fused with the
// [t, h, t-h, t*h] feature construction, the attention-weighted pooling (optionally softmaxed with
// warp shuffles), and their backward down to the scatter-add into the two tables.
//
// Reference arithmetic replaced: ActivationUnit.forward models/ranking/din.py:77-93 (everything but
// the attention MLP itself, which runs as tower GEMMs + rh_bn_act_*), and the history/target lookups
// of DIN.forward din.py:42-44 (EmbeddingLayer with pooling="concat", basic/layers.py:91-99).
#include "rh_common.cuh"

namespace rh {

// ---- att_in = [t, h, t-h, t*h] ------------------------------------------------------------------
// one lane per (b, l, 16-byte quarter) when dim % 4 == 0, else one lane per (b, l, d).
template <bool VEC>
__global__ void __launch_bounds__(256) din_attn_input_fwd_kernel(const float* __restrict__ hist_table, int hist_vocab,
                                                                 const float* __restrict__ tgt_table, int tgt_vocab, int dim,
                                                                 const void* __restrict__ hist_ids, const void* __restrict__ tgt_ids,
                                                                 bool is_i32, int64_t tgt_id_stride, int batch, int L,
                                                                 float* __restrict__ att_in, float* __restrict__ hist_out,
                                                                 float* __restrict__ tgt_out, int32_t* err) {
  const int lanes = VEC ? dim / 4 : dim;
  const int64_t total = (int64_t)batch * L * lanes;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / lanes;  // b*L + l
    const int q = (int)(i - r * lanes);
    const int b = (int)(r / L);
    const int l = (int)(r - (int64_t)b * L);
    const int64_t hid = load_id(hist_ids, r, is_i32);
    const int64_t tid = load_id(tgt_ids, (int64_t)b * tgt_id_stride, is_i32);
    const bool h_ok = (uint64_t)hid < (uint64_t)hist_vocab, t_ok = (uint64_t)tid < (uint64_t)tgt_vocab;
    if ((!h_ok || !t_ok) && q == 0 && err != nullptr) *err = 1;
    float* row = att_in + r * 4 * dim;
    if (VEC) {
      const float4 h = h_ok ? ldg_row16(hist_table + hid * dim + 4 * q) : f4_zero();
      const float4 t = t_ok ? ldg_row16(tgt_table + tid * dim + 4 * q) : f4_zero();
      stg_row16(row + 4 * q, t);
      stg_row16(row + dim + 4 * q, h);
      stg_row16(row + 2 * dim + 4 * q, make_float4(t.x - h.x, t.y - h.y, t.z - h.z, t.w - h.w));
      stg_row16(row + 3 * dim + 4 * q, make_float4(t.x * h.x, t.y * h.y, t.z * h.z, t.w * h.w));
      if (hist_out != nullptr) stg_row16(hist_out + r * dim + 4 * q, h);
      if (tgt_out != nullptr && l == 0) stg_row16(tgt_out + (int64_t)b * dim + 4 * q, t);
    } else {
      const float h = h_ok ? __ldg(hist_table + hid * dim + q) : 0.f;
      const float t = t_ok ? __ldg(tgt_table + tid * dim + q) : 0.f;
      row[q] = t;
      row[dim + q] = h;
      row[2 * dim + q] = t - h;
      row[3 * dim + q] = t * h;
      if (hist_out != nullptr) hist_out[r * dim + q] = h;
      if (tgt_out != nullptr && l == 0) tgt_out[(int64_t)b * dim + q] = t;
    }
  }
}

// ---- out[b,:] = sum_l w[b,l] * hist[b,l,:] ------------------------------------------------------
// one warp per sample; lane <-> sequence positions (l = lane, lane+32, ...); softmax via shuffles.
constexpr int kDinDChunk = 8;

__global__ void __launch_bounds__(256) din_weighted_sum_fwd_kernel(const float* __restrict__ att_w, const float* __restrict__ hist,
                                                                   int batch, int L, int dim, int use_softmax,
                                                                   float* __restrict__ w_used, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int b = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (b >= batch) return;
  const float* wb = att_w + (int64_t)b * L;
  float mx = -INFINITY, den = 1.f;
  if (use_softmax) {
    for (int l = lane; l < L; l += 32) mx = fmaxf(mx, __ldg(wb + l));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float s = 0.f;
    for (int l = lane; l < L; l += 32) s += expf(__ldg(wb + l) - mx);
    den = warp_sum(s);
  }
  for (int d0 = 0; d0 < dim; d0 += kDinDChunk) {
    float acc[kDinDChunk];
#pragma unroll
    for (int j = 0; j < kDinDChunk; ++j) acc[j] = 0.f;
    for (int l = lane; l < L; l += 32) {
      float wv = __ldg(wb + l);
      if (use_softmax) wv = expf(wv - mx) / den;
      if (d0 == 0 && w_used != nullptr) w_used[(int64_t)b * L + l] = wv;
      const float* hrow = hist + ((int64_t)b * L + l) * dim + d0;
#pragma unroll
      for (int j = 0; j < kDinDChunk; ++j)
        if (d0 + j < dim) acc[j] = fmaf(wv, __ldg(hrow + j), acc[j]);
    }
#pragma unroll
    for (int j = 0; j < kDinDChunk; ++j) {
      const float t = warp_sum(acc[j]);
      if (lane == 0 && d0 + j < dim) out[(int64_t)b * dim + d0 + j] = t;
    }
  }
}

// d_w[b,l] = <d_out[b], hist[b,l]> (then through the softmax), d_hist[b,l,:] = w[b,l] * d_out[b,:]
__global__ void __launch_bounds__(256) din_weighted_sum_bwd_kernel(const float* __restrict__ w_used, const float* __restrict__ hist,
                                                                   const float* __restrict__ d_out, int batch, int L, int dim,
                                                                   int use_softmax, float* __restrict__ d_att_w,
                                                                   float* __restrict__ d_hist) {
  const int lane = threadIdx.x & 31;
  const int b = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (b >= batch) return;
  const float* dob = d_out + (int64_t)b * dim;
  float dot_sum = 0.f;  // sum_l w*dw for the softmax Jacobian
  for (int l = lane; l < L; l += 32) {
    const float wv = __ldg(w_used + (int64_t)b * L + l);
    const float* hrow = hist + ((int64_t)b * L + l) * dim;
    float* dhrow = d_hist + ((int64_t)b * L + l) * dim;
    float dw = 0.f;
    for (int d = 0; d < dim; ++d) {
      const float g = __ldg(dob + d);
      dw = fmaf(g, __ldg(hrow + d), dw);
      dhrow[d] = wv * g;
    }
    d_att_w[(int64_t)b * L + l] = dw;  // raw; fixed up below when softmax
    dot_sum = fmaf(wv, dw, dot_sum);
  }
  if (use_softmax) {
    dot_sum = warp_sum(dot_sum);
    for (int l = lane; l < L; l += 32) {
      const float wv = __ldg(w_used + (int64_t)b * L + l);
      const int64_t o = (int64_t)b * L + l;
      d_att_w[o] = wv * (d_att_w[o] - dot_sum);
    }
  }
}

// backward of the feature construction + scatter-add.  One warp per sample; lanes <-> positions.
//   d_h[b,l] = d1 - d2 + d3*t + d_hist ;  d_t[b] = sum_l (d0 + d2 + d3*h) + d_tgt_extra[b]
__global__ void __launch_bounds__(256) din_attn_input_bwd_kernel(float* __restrict__ hist_grad, int hist_vocab, int hist_pad,
                                                                 float* __restrict__ tgt_grad, int tgt_vocab, int tgt_pad, int dim,
                                                                 const void* __restrict__ hist_ids, const void* __restrict__ tgt_ids,
                                                                 bool is_i32, int64_t tgt_id_stride, int batch, int L,
                                                                 const float* __restrict__ hist, const float* __restrict__ tgt,
                                                                 const float* __restrict__ d_att_in, const float* __restrict__ d_hist,
                                                                 const float* __restrict__ d_tgt_extra, int32_t* err) {
  const int lane = threadIdx.x & 31;
  const int b = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (b >= batch) return;
  const int64_t tid = load_id(tgt_ids, (int64_t)b * tgt_id_stride, is_i32);
  const bool t_ok = (uint64_t)tid < (uint64_t)tgt_vocab;
  if (!t_ok && lane == 0 && err != nullptr) *err = 1;
  for (int d0 = 0; d0 < dim; d0 += kDinDChunk) {
    float dt[kDinDChunk], tv[kDinDChunk];
#pragma unroll
    for (int j = 0; j < kDinDChunk; ++j) {
      dt[j] = 0.f;
      tv[j] = (d0 + j < dim) ? __ldg(tgt + (int64_t)b * dim + d0 + j) : 0.f;
    }
    for (int l = lane; l < L; l += 32) {
      const int64_t r = (int64_t)b * L + l;
      const int64_t hid = load_id(hist_ids, r, is_i32);
      const bool h_ok = (uint64_t)hid < (uint64_t)hist_vocab;
      if (!h_ok && d0 == 0 && err != nullptr) *err = 1;
      const float* din = d_att_in + r * 4 * dim + d0;
#pragma unroll
      for (int j = 0; j < kDinDChunk; ++j) {
        if (d0 + j >= dim) continue;
        const float g0 = __ldg(din + j), g1 = __ldg(din + dim + j), g2 = __ldg(din + 2 * dim + j), g3 = __ldg(din + 3 * dim + j);
        const float hv = __ldg(hist + r * dim + d0 + j);
        dt[j] += g0 + g2 + g3 * hv;
        float dh = g1 - g2 + g3 * tv[j];
        if (d_hist != nullptr) dh += __ldg(d_hist + r * dim + d0 + j);
        if (hist_grad != nullptr && h_ok && hid != hist_pad) atomicAdd(hist_grad + hid * dim + d0 + j, dh);
      }
    }
#pragma unroll
    for (int j = 0; j < kDinDChunk; ++j) {
      float t = warp_sum(dt[j]);
      if (lane == 0 && d0 + j < dim && tgt_grad != nullptr && t_ok && tid != tgt_pad) {
        if (d_tgt_extra != nullptr) t += __ldg(d_tgt_extra + (int64_t)b * dim + d0 + j);
        atomicAdd(tgt_grad + tid * dim + d0 + j, t);
      }
    }
  }
}

static inline int din_grid(int64_t threads_needed, int block, int cap_per_sm) {
  int64_t g = (threads_needed + block - 1) / block;
  const int64_t cap = (int64_t)num_sms() * cap_per_sm;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace rh

using namespace rh;

extern "C" int rh_din_attn_input_fwd(const float* hist_table, int hist_vocab, const float* tgt_table, int tgt_vocab, int dim,
                                     const void* hist_ids, const void* tgt_ids, int ids_are_i32, int64_t tgt_id_stride, int batch,
                                     int seq_len, float* att_in, float* hist_out, float* tgt_out, int32_t* err_flag, void* stream) {
  RH_REQUIRE(hist_table && tgt_table && hist_ids && tgt_ids && att_in, RH_ERR_INVALID_ARG, "rh_din_attn_input_fwd: NULL pointer");
  RH_REQUIRE(hist_vocab > 0 && tgt_vocab > 0 && dim > 0 && batch >= 0 && seq_len > 0, RH_ERR_INVALID_ARG, "rh_din_attn_input_fwd: bad sizes");
  if (batch == 0) return RH_OK;
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  const bool vec = dim % 4 == 0 && al(hist_table) && al(tgt_table) && al(att_in) && (!hist_out || al(hist_out)) && (!tgt_out || al(tgt_out));
  const int64_t lanes = (int64_t)batch * seq_len * (vec ? dim / 4 : dim);
  const int grid = din_grid(lanes, 256, 16);
  if (vec) {
    din_attn_input_fwd_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(hist_table, hist_vocab, tgt_table, tgt_vocab, dim, hist_ids,
                                                                            tgt_ids, ids_are_i32 != 0, tgt_id_stride, batch, seq_len,
                                                                            att_in, hist_out, tgt_out, err_flag);
  } else {
    din_attn_input_fwd_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(hist_table, hist_vocab, tgt_table, tgt_vocab, dim, hist_ids,
                                                                             tgt_ids, ids_are_i32 != 0, tgt_id_stride, batch, seq_len,
                                                                             att_in, hist_out, tgt_out, err_flag);
  }
  RH_LAUNCH_CHECK();
  return RH_OK;
}

extern "C" int rh_din_weighted_sum_fwd(const float* att_w, const float* hist, int batch, int seq_len, int dim, int use_softmax,
                                       float* w_used, float* out, void* stream) {
  RH_REQUIRE(att_w && hist && out, RH_ERR_INVALID_ARG, "rh_din_weighted_sum_fwd: NULL pointer");
  RH_REQUIRE(batch >= 0 && seq_len > 0 && dim > 0, RH_ERR_INVALID_ARG, "rh_din_weighted_sum_fwd: bad sizes");
  RH_REQUIRE(!use_softmax || w_used != nullptr, RH_ERR_INVALID_ARG, "rh_din_weighted_sum_fwd: softmax needs w_used");
  if (batch == 0) return RH_OK;
  din_weighted_sum_fwd_kernel<<<(batch + 7) / 8, 256, 0, (cudaStream_t)stream>>>(att_w, hist, batch, seq_len, dim, use_softmax, w_used, out);
  RH_LAUNCH_CHECK();
  return RH_OK;
}

extern "C" int rh_din_weighted_sum_bwd(const float* w_used, const float* hist, const float* d_out, int batch, int seq_len, int dim,
                                       int use_softmax, float* d_att_w, float* d_hist, void* stream) {
  RH_REQUIRE(w_used && hist && d_out && d_att_w && d_hist, RH_ERR_INVALID_ARG, "rh_din_weighted_sum_bwd: NULL pointer");
  RH_REQUIRE(batch >= 0 && seq_len > 0 && dim > 0, RH_ERR_INVALID_ARG, "rh_din_weighted_sum_bwd: bad sizes");
  if (batch == 0) return RH_OK;
  din_weighted_sum_bwd_kernel<<<(batch + 7) / 8, 256, 0, (cudaStream_t)stream>>>(w_used, hist, d_out, batch, seq_len, dim, use_softmax,
                                                                                 d_att_w, d_hist);
  RH_LAUNCH_CHECK();
  return RH_OK;
}

extern "C" int rh_din_attn_input_bwd(float* hist_grad, int hist_vocab, int hist_padding_idx, float* tgt_grad, int tgt_vocab,
                                     int tgt_padding_idx, int dim, const void* hist_ids, const void* tgt_ids, int ids_are_i32,
                                     int64_t tgt_id_stride, int batch, int seq_len, const float* hist, const float* tgt,
                                     const float* d_att_in, const float* d_hist, const float* d_tgt_extra, int32_t* err_flag,
                                     void* stream) {
  RH_REQUIRE(hist_ids && tgt_ids && hist && tgt && d_att_in, RH_ERR_INVALID_ARG, "rh_din_attn_input_bwd: NULL pointer");
  RH_REQUIRE(hist_vocab > 0 && tgt_vocab > 0 && dim > 0 && batch >= 0 && seq_len > 0, RH_ERR_INVALID_ARG, "rh_din_attn_input_bwd: bad sizes");
  if (batch == 0) return RH_OK;
  din_attn_input_bwd_kernel<<<(batch + 7) / 8, 256, 0, (cudaStream_t)stream>>>(hist_grad, hist_vocab, hist_padding_idx, tgt_grad, tgt_vocab,
                                                                               tgt_padding_idx, dim, hist_ids, tgt_ids, ids_are_i32 != 0,
                                                                               tgt_id_stride, batch, seq_len, hist, tgt, d_att_in, d_hist,
                                                                               d_tgt_extra, err_flag);
  RH_LAUNCH_CHECK();
  return RH_OK;
}
