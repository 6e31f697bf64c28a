// rh_rows.cu — single-table row kernels: gather / scatter-add / masked sequence pooling / sparse
// zeroing / row-wise optimisers.  All HBM-bound; rows move as 16-byte lanes when dim % 4 == 0.
//
// Reference arithmetic replaced: nn.Embedding forward/backward as called from basic/layers.py:83-99,
// SumPooling/AveragePooling + InputMask (basic/layers.py:148-161, 209-251), the dense zero-fill of
// embedding gradients and the table part of torch.optim.* (ctr_trainer.py:99).
#ifndef RH_PDL_FAMILY  // (the trace tools include several of these files into one unit: the first one names the family)
#define RH_PDL_FAMILY 8  /* rh_set_pdl mask bit of this file's kernels */
#endif
#include "rh_common.cuh"

namespace rh {

static bool aligned16r(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

static inline int grid_for(int64_t threads_needed, int block, int max_blocks_per_sm = 16) {
  int64_t g = (threads_needed + block - 1) / block;
  const int64_t cap = (int64_t)num_sms() * max_blocks_per_sm;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

// ---- gather ------------------------------------------------------------------------------------
template <int UNROLL>
__global__ void __launch_bounds__(256) rows_gather_v4(const float* __restrict__ table, int vocab, int d4, const void* __restrict__ ids,
                                                      bool is_i32, int64_t n, float* __restrict__ out, int32_t* err) {
  // one 16-byte lane per (row, quarter); UNROLL independent rows in flight per thread
  const int64_t total = n * d4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += stride * UNROLL) {
    float4 v[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const int64_t i = i0 + u * stride;
      v[u] = f4_zero();
      if (i < total) {
        const int64_t r = i / d4;
        const int q = (int)(i - r * d4);
        const int64_t id = load_id(ids, r, is_i32);
        if ((uint64_t)id < (uint64_t)vocab) {
          v[u] = ldg_row16(table + (id * d4 + q) * 4);
        } else if (q == 0 && err != nullptr) {
          *err = 1;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const int64_t i = i0 + u * stride;
      if (i < total) stg_row16(out + i * 4, v[u]);
    }
  }
}

__global__ void __launch_bounds__(256) rows_gather_scalar(const float* __restrict__ table, int vocab, int dim, const void* __restrict__ ids,
                                                          bool is_i32, int64_t n, float* __restrict__ out, int32_t* err) {
  const int64_t total = n * dim;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / dim;
    const int d = (int)(i - r * dim);
    const int64_t id = load_id(ids, r, is_i32);
    float v = 0.f;
    if ((uint64_t)id < (uint64_t)vocab) {
      v = __ldg(table + id * dim + d);
    } else if (d == 0 && err != nullptr) {
      *err = 1;
    }
    out[i] = v;
  }
}

// ---- scatter-add -------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) rows_scatter_add_v4(float* __restrict__ grad, int vocab, int d4, int pad_idx,
                                                           const void* __restrict__ ids, bool is_i32, int64_t n,
                                                           const float* __restrict__ d_out, int32_t* err) {
  const int64_t total = n * d4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / d4;
    const int q = (int)(i - r * d4);
    const int64_t id = load_id(ids, r, is_i32);
    if ((uint64_t)id >= (uint64_t)vocab) {
      if (q == 0 && err != nullptr) *err = 1;
      continue;
    }
    if (id == pad_idx) continue;
    red_add_row16(grad + (id * d4 + q) * 4, ldg_row16(d_out + i * 4));
  }
}

__global__ void __launch_bounds__(256) rows_scatter_add_scalar(float* __restrict__ grad, int vocab, int dim, int pad_idx,
                                                               const void* __restrict__ ids, bool is_i32, int64_t n,
                                                               const float* __restrict__ d_out, int32_t* err) {
  const int64_t total = n * dim;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / dim;
    const int d = (int)(i - r * dim);
    const int64_t id = load_id(ids, r, is_i32);
    if ((uint64_t)id >= (uint64_t)vocab) {
      if (d == 0 && err != nullptr) *err = 1;
      continue;
    }
    if (id == pad_idx) continue;
    atomicAdd(grad + id * dim + d, __ldg(d_out + i));
  }
}

// ---- zero rows ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) rows_zero_kernel(float* __restrict__ grad, int vocab, int dim, const void* __restrict__ ids,
                                                        bool is_i32, int64_t n, bool vec) {
  if (vec) {
    const int d4 = dim / 4;
    const int64_t total = n * d4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
      const int64_t r = i / d4;
      const int q = (int)(i - r * d4);
      const int64_t id = load_id(ids, r, is_i32);
      if ((uint64_t)id < (uint64_t)vocab) *reinterpret_cast<float4*>(grad + (id * d4 + q) * 4) = f4_zero();
    }
  } else {
    const int64_t total = n * dim;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
      const int64_t r = i / dim;
      const int d = (int)(i - r * dim);
      const int64_t id = load_id(ids, r, is_i32);
      if ((uint64_t)id < (uint64_t)vocab) grad[id * dim + d] = 0.f;
    }
  }
}

// ---- masked sequence pooling -------------------------------------------------------------------
// one warp per sample; lane l strides over the dim (dim <= 32*?): lanes own columns, loop over L.
// For dim % 4 == 0 each lane owns a 16-byte quarter and several sequence positions run in flight.
template <int LPR>
__global__ void __launch_bounds__(128) seq_pool_fwd_v4(const float* __restrict__ table, int vocab, int dim, const void* __restrict__ ids,
                                                       bool is_i32, int batch, int L, int mode, int64_t mask_id,
                                                       float* __restrict__ out, int64_t out_ld, int32_t* err) {
  const int spb = blockDim.x / LPR;
  const int b = blockIdx.x * spb + (int)threadIdx.x / LPR;
  const int q = (int)threadIdx.x % LPR;
  if (b >= batch) return;
  const bool lane_on = 4 * q < dim;
  float4 acc = f4_zero();
  float cnt = 0.f;
  constexpr int U = 8;
  for (int l0 = 0; l0 < L; l0 += U) {
    int32_t rid[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      rid[u] = -1;
      if (l0 + u < L) {
        const int64_t id = load_id(ids, (int64_t)b * L + l0 + u, is_i32);
        if (id != mask_id) {
          if ((uint64_t)id < (uint64_t)vocab) {
            rid[u] = (int32_t)id;
            cnt += 1.f;
          } else if (q == 0 && err != nullptr) {
            *err = 1;
          }
        }
      }
    }
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      v[u] = f4_zero();
      if (rid[u] >= 0 && lane_on) v[u] = ldg_row16(table + (int64_t)rid[u] * dim + 4 * q);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) acc = f4_add(acc, v[u]);
  }
  if (mode == 2) {
    const float den = cnt + 1e-16f;
    acc = make_float4(acc.x / den, acc.y / den, acc.z / den, acc.w / den);
  }
  if (lane_on) *reinterpret_cast<float4*>(out + (int64_t)b * out_ld + 4 * q) = acc;
}

__global__ void __launch_bounds__(128) seq_pool_fwd_scalar(const float* __restrict__ table, int vocab, int dim, const void* __restrict__ ids,
                                                           bool is_i32, int batch, int L, int mode, int64_t mask_id,
                                                           float* __restrict__ out, int64_t out_ld, int32_t* err) {
  // one warp per sample, lanes stride over columns
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= batch) return;
  const int b = warp;
  float cnt = 0.f;
  for (int d0 = 0; d0 < dim; d0 += 32) {
    const int d = d0 + lane;
    float acc = 0.f;
    cnt = 0.f;
    for (int l = 0; l < L; ++l) {
      const int64_t id = load_id(ids, (int64_t)b * L + l, is_i32);
      if (id == mask_id) continue;
      if ((uint64_t)id >= (uint64_t)vocab) {
        if (lane == 0 && err != nullptr) *err = 1;
        continue;
      }
      cnt += 1.f;
      if (d < dim) acc += __ldg(table + id * dim + d);
    }
    if (mode == 2) acc = acc / (cnt + 1e-16f);
    if (d < dim) out[(int64_t)b * out_ld + d] = acc;
  }
}

__global__ void __launch_bounds__(256) seq_pool_bwd_kernel(float* __restrict__ grad, int vocab, int dim, int pad_idx,
                                                           const void* __restrict__ ids, bool is_i32, int batch, int L, int mode,
                                                           int64_t mask_id, const float* __restrict__ d_out, int64_t d_out_ld, int32_t* err, bool vec) {
  // thread per (b, l, lane); the per-sample count is recomputed by a short loop (L is tens).
  const int lanes = vec ? dim / 4 : dim;
  const int64_t total = (int64_t)batch * L * lanes;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / lanes;  // b*L + l
    const int q = (int)(i - r * lanes);
    const int b = (int)(r / L);
    const int64_t id = load_id(ids, r, is_i32);
    if (id == mask_id) continue;
    if ((uint64_t)id >= (uint64_t)vocab) {
      if (q == 0 && err != nullptr) *err = 1;
      continue;
    }
    if (id == pad_idx) continue;
    float scale = 1.f;
    if (mode == 2) {
      float cnt = 0.f;
      for (int l = 0; l < L; ++l) {
        const int64_t other = load_id(ids, (int64_t)b * L + l, is_i32);
        if (other != mask_id && (uint64_t)other < (uint64_t)vocab) cnt += 1.f;
      }
      scale = 1.f / (cnt + 1e-16f);
    }
    if (vec) {
      float4 g = ldg_row16(d_out + (int64_t)b * d_out_ld + 4 * q);
      red_add_row16(grad + id * dim + 4 * q, f4_scale(g, scale));
    } else {
      atomicAdd(grad + id * dim + q, __ldg(d_out + (int64_t)b * d_out_ld + q) * scale);
    }
  }
}

// ---- row-wise optimisers -----------------------------------------------------------------------
struct OptArgs {
  int kind;
  float beta1, beta2, eps, wd;
};

__device__ __forceinline__ float opt_update(float w, float g, float* s1, float* s2, const OptArgs& a, float lr, float bc1, float bc2_sqrt) {
  if (a.kind == 0) {  // SGD
    g = fmaf(a.wd, w, g);
    return w - lr * g;
  } else if (a.kind == 1) {  // Adam (torch.optim.Adam arithmetic on the touched row)
    g = fmaf(a.wd, w, g);
    const float m = a.beta1 * (*s1) + (1.f - a.beta1) * g;
    const float v = a.beta2 * (*s2) + (1.f - a.beta2) * g * g;
    *s1 = m;
    *s2 = v;
    const float denom = sqrtf(v) / bc2_sqrt + a.eps;
    return w - (lr / bc1) * (m / denom);
  } else {  // Adagrad
    g = fmaf(a.wd, w, g);
    const float acc = *s1 + g * g;
    *s1 = acc;
    return w - lr * g / (sqrtf(acc) + a.eps);
  }
}

__global__ void __launch_bounds__(256) rowwise_update_kernel(float* __restrict__ table, float* __restrict__ grad, float* __restrict__ st1,
                                                             float* __restrict__ st2, int32_t* __restrict__ stamp, int vocab, int dim,
                                                             const void* __restrict__ ids, bool is_i32, int64_t n, OptArgs a,
                                                             const int32_t* __restrict__ step_dev, const float* __restrict__ lr_dev,
                                                             const float* __restrict__ bc_dev, bool vec, int G, int64_t sstride) {
  // G (power of two <= 32) consecutive threads own one entry of `ids`.  The group's lane 0 claims
  // the row through stamp[id] (first claimant of this step wins: duplicates of an id are skipped)
  // and broadcasts the verdict; the owner group then updates the row's `lanes` 16-byte (or scalar) slots.
  const int step = *step_dev;
  const float lr = *lr_dev;
  // Adam bias corrections 1 - beta1^t and sqrt(1 - beta2^t): computed ONCE per step by rh_opt_advance (fp64)
  const float bc1 = a.kind == 1 ? bc_dev[0] : 1.f, bc2s = a.kind == 1 ? bc_dev[1] : 1.f;
  const int lanes = vec ? dim / 4 : dim;
  const int g = (int)threadIdx.x & (G - 1);
  const int lane = (int)threadIdx.x & 31;
  const int64_t gthreads = (int64_t)gridDim.x * blockDim.x;
  for (int64_t base = (int64_t)blockIdx.x * blockDim.x; base < n * G; base += gthreads) {  // block-uniform trip count
    const int64_t r = (base + threadIdx.x) / G;
    int64_t id = -1;
    if (r < n) id = load_id(ids, r, is_i32);
    const bool valid = (uint64_t)id < (uint64_t)vocab;
    int old = step;
    if (valid && g == 0) old = atomicExch(stamp + id, step);
    old = __shfl_sync(0xffffffffu, old, lane & ~(G - 1));
    if (!valid || old == step) continue;
    for (int q = g; q < lanes; q += G) {
      if (vec) {
        float* gp = grad + id * dim + 4 * q;
        float* wp = table + id * dim + 4 * q;
        float4 gr = *reinterpret_cast<float4*>(gp);
        *reinterpret_cast<float4*>(gp) = f4_zero();
        float4 w = *reinterpret_cast<float4*>(wp);
        float4 m = f4_zero(), v = f4_zero();
        if (a.kind != 0) m = *reinterpret_cast<float4*>(st1 + id * sstride + 4 * q);
        if (a.kind == 1) v = *reinterpret_cast<float4*>(st2 + id * sstride + 4 * q);
        w.x = opt_update(w.x, gr.x, &m.x, &v.x, a, lr, bc1, bc2s);
        w.y = opt_update(w.y, gr.y, &m.y, &v.y, a, lr, bc1, bc2s);
        w.z = opt_update(w.z, gr.z, &m.z, &v.z, a, lr, bc1, bc2s);
        w.w = opt_update(w.w, gr.w, &m.w, &v.w, a, lr, bc1, bc2s);
        *reinterpret_cast<float4*>(wp) = w;
        if (a.kind != 0) *reinterpret_cast<float4*>(st1 + id * sstride + 4 * q) = m;
        if (a.kind == 1) *reinterpret_cast<float4*>(st2 + id * sstride + 4 * q) = v;
      } else {
        const int64_t o = id * dim + q, so = id * sstride + q;
        const float gr = grad[o];
        grad[o] = 0.f;
        float m = a.kind != 0 ? st1[so] : 0.f;
        float v = a.kind == 1 ? st2[so] : 0.f;
        table[o] = opt_update(table[o], gr, &m, &v, a, lr, bc1, bc2s);
        if (a.kind != 0) st1[so] = m;
        if (a.kind == 1) st2[so] = v;
      }
    }
  }
}

// ---- all tables of one batch in ONE launch: grid.y = field ------------------------------------------
struct MultiP {
  float* table[RH_MAX_FIELDS];
  float* grad[RH_MAX_FIELDS];
  float* st1[RH_MAX_FIELDS];
  float* st2[RH_MAX_FIELDS];
  int32_t* stamp[RH_MAX_FIELDS];
  const void* ids[RH_MAX_FIELDS];
  int32_t id_stride[RH_MAX_FIELDS];
  int32_t vocab[RH_MAX_FIELDS];
  uint8_t is_i32[RH_MAX_FIELDS];
  int32_t batch, dim;
};

__global__ void __launch_bounds__(128) fields_rowwise_update_kernel(const __grid_constant__ MultiP p, OptArgs a,
                                                                    const int32_t* __restrict__ step_dev,
                                                                    const float* __restrict__ lr_dev,
                                                                    const float* __restrict__ bc_dev, int G, int64_t state_stride) {
  // same claim protocol as rowwise_update_kernel; 16-byte lanes only (dim % 4 == 0).
  // Latency shape: ids -> {claim atomic, g, w, m, v loads} -> stores.  The row loads are issued BEFORE the claim's verdict
  // is known (duplicates are rare and a wasted read is harmless), which removes one dependent DRAM round trip.
  pdl_wait();
  const int f = blockIdx.y;
  const int step = *step_dev;
  const float lr = *lr_dev;
  // Adam bias corrections 1 - beta1^t and sqrt(1 - beta2^t): computed ONCE per step by rh_opt_advance (fp64)
  const float bc1 = a.kind == 1 ? bc_dev[0] : 1.f, bc2s = a.kind == 1 ? bc_dev[1] : 1.f;
  const int dim = p.dim, lanes = dim / 4;
  const int g = (int)threadIdx.x & (G - 1);
  const int lane = (int)threadIdx.x & 31;
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
  int64_t id = -1;
  if (r < p.batch) id = load_id(p.ids[f], r * p.id_stride[f], p.is_i32[f] != 0);
  const bool valid = (uint64_t)id < (uint64_t)p.vocab[f];
  const bool lane_on = valid && g < lanes;  // G == pow2_ceil(lanes) <= 32: one 16-byte slot per lane
  float* gp = p.grad[f] + id * dim + 4 * g;
  float* wp = p.table[f] + id * dim + 4 * g;
  float* s1p = a.kind != 0 ? p.st1[f] + id * state_stride + 4 * g : nullptr;
  float* s2p = a.kind == 1 ? p.st2[f] + id * state_stride + 4 * g : nullptr;
  float4 gr = f4_zero(), w = f4_zero(), m = f4_zero(), v = f4_zero();
  if (lane_on) {
    gr = *reinterpret_cast<const float4*>(gp);
    w = *reinterpret_cast<const float4*>(wp);
    if (a.kind != 0) m = *reinterpret_cast<const float4*>(s1p);
    if (a.kind == 1) v = *reinterpret_cast<const float4*>(s2p);
  }
  int old = step;
  if (valid && g == 0) old = atomicExch(p.stamp[f] + id, step);
  old = __shfl_sync(0xffffffffu, old, lane & ~(G - 1));
  if (!lane_on || old == step) return;
  *reinterpret_cast<float4*>(gp) = f4_zero();
  w.x = opt_update(w.x, gr.x, &m.x, &v.x, a, lr, bc1, bc2s);
  w.y = opt_update(w.y, gr.y, &m.y, &v.y, a, lr, bc1, bc2s);
  w.z = opt_update(w.z, gr.z, &m.z, &v.z, a, lr, bc1, bc2s);
  w.w = opt_update(w.w, gr.w, &m.w, &v.w, a, lr, bc1, bc2s);
  *reinterpret_cast<float4*>(wp) = w;
  if (a.kind != 0) *reinterpret_cast<float4*>(s1p) = m;
  if (a.kind == 1) *reinterpret_cast<float4*>(s2p) = v;
}

__global__ void __launch_bounds__(128) fields_zero_kernel(const __grid_constant__ MultiP p) {
  const int f = blockIdx.y;
  const int lanes = p.dim / 4;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t r = i / lanes;
  const int q = (int)(i - r * lanes);
  if (r >= p.batch) return;
  const int64_t id = load_id(p.ids[f], r * p.id_stride[f], p.is_i32[f] != 0);
  if ((uint64_t)id < (uint64_t)p.vocab[f]) *reinterpret_cast<float4*>(p.grad[f] + id * p.dim + 4 * q) = f4_zero();
}

// L2 prefetch of the rows a FUTURE batch will touch (tables, gradient rows, optimiser records): one thread per (sample, array).
// Issued one step ahead on the copy stream, it turns the latency-bound random accesses of the next step's gather, scatter and
// row-wise update into L2 hits.  Nothing is read into registers and nothing is written: a wrong or stale id costs bandwidth only.
__global__ void __launch_bounds__(128) fields_prefetch_kernel(const __grid_constant__ MultiP p, int64_t state_row_bytes) {
  const int f = blockIdx.y;
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= p.batch) return;
  const int64_t id = load_id(p.ids[f], r * p.id_stride[f], p.is_i32[f] != 0);
  if ((uint64_t)id >= (uint64_t)p.vocab[f]) return;
  const int64_t row_bytes = (int64_t)p.dim * 4;
  const char* bases[3] = {reinterpret_cast<const char*>(p.table[f]), reinterpret_cast<const char*>(p.grad[f]), reinterpret_cast<const char*>(p.st1[f])};
  const int64_t strides[3] = {row_bytes, row_bytes, state_row_bytes};
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    if (bases[a] == nullptr) continue;
    const char* row = bases[a] + id * strides[a];
    for (int64_t off = 0; off < strides[a]; off += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(row + off));
  }
}

__global__ void opt_advance_kernel(int32_t* step_dev, float* bc_dev, float beta1, float beta2) {
  pdl_wait();
  const int step = *step_dev + 1;
  *step_dev = step;
  if (bc_dev != nullptr) {
    bc_dev[0] = (float)(1.0 - pow((double)beta1, (double)step));
    bc_dev[1] = (float)sqrt(1.0 - pow((double)beta2, (double)step));
  }
}

}  // namespace rh

using namespace rh;

extern "C" int rh_rows_gather(const float* table, int vocab, int dim, const void* ids, int ids_are_i32, int64_t n, float* out,
                              int32_t* err_flag, void* stream) {
  RH_REQUIRE(table && ids && out, RH_ERR_INVALID_ARG, "rh_rows_gather: NULL pointer");
  RH_REQUIRE(vocab > 0 && dim > 0 && n >= 0, RH_ERR_INVALID_ARG, "rh_rows_gather: bad sizes");
  if (n == 0) return RH_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (dim % 4 == 0 && aligned16r(table) && aligned16r(out)) {
    const int d4 = dim / 4;
    const int64_t lanes = n * d4;
    // small problems: one lane per thread so everything is in flight at once; large: 4 rows / thread
    if (lanes <= (int64_t)num_sms() * 2048) {
      rows_gather_v4<1><<<grid_for(lanes, 256, 8), 256, 0, st>>>(table, vocab, d4, ids, ids_are_i32 != 0, n, out, err_flag);
    } else {
      rows_gather_v4<4><<<grid_for((lanes + 3) / 4, 256, 8), 256, 0, st>>>(table, vocab, d4, ids, ids_are_i32 != 0, n, out, err_flag);
    }
  } else {
    rows_gather_scalar<<<grid_for(n * dim, 256), 256, 0, st>>>(table, vocab, dim, ids, ids_are_i32 != 0, n, out, err_flag);
  }
  RH_LAUNCH_CHECK();
  return RH_OK;
}

extern "C" int rh_rows_scatter_add(float* table_grad, int vocab, int dim, int padding_idx, const void* ids, int ids_are_i32, int64_t n,
                                   const float* d_out, int32_t* err_flag, void* stream) {
  RH_REQUIRE(table_grad && ids && d_out, RH_ERR_INVALID_ARG, "rh_rows_scatter_add: NULL pointer");
  RH_REQUIRE(vocab > 0 && dim > 0 && n >= 0, RH_ERR_INVALID_ARG, "rh_rows_scatter_add: bad sizes");
  if (n == 0) return RH_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (dim % 4 == 0 && aligned16r(table_grad) && aligned16r(d_out)) {
    rows_scatter_add_v4<<<grid_for(n * (dim / 4), 256), 256, 0, st>>>(table_grad, vocab, dim / 4, padding_idx, ids, ids_are_i32 != 0, n, d_out,
                                                                      err_flag);
  } else {
    rows_scatter_add_scalar<<<grid_for(n * dim, 256), 256, 0, st>>>(table_grad, vocab, dim, padding_idx, ids, ids_are_i32 != 0, n, d_out,
                                                                    err_flag);
  }
  RH_LAUNCH_CHECK();
  return RH_OK;
}

extern "C" int rh_rows_zero(float* table_grad, int vocab, int dim, const void* ids, int ids_are_i32, int64_t n, void* stream) {
  RH_REQUIRE(table_grad && ids, RH_ERR_INVALID_ARG, "rh_rows_zero: NULL pointer");
  RH_REQUIRE(vocab > 0 && dim > 0 && n >= 0, RH_ERR_INVALID_ARG, "rh_rows_zero: bad sizes");
  if (n == 0) return RH_OK;
  const bool vec = dim % 4 == 0 && aligned16r(table_grad);
  const int64_t lanes = n * (vec ? dim / 4 : dim);
  rows_zero_kernel<<<grid_for(lanes, 256), 256, 0, (cudaStream_t)stream>>>(table_grad, vocab, dim, ids, ids_are_i32 != 0, n, vec);
  RH_LAUNCH_CHECK();
  return RH_OK;
}

extern "C" int rh_seq_pool_fwd(const float* table, int vocab, int dim, const void* ids, int ids_are_i32, int batch, int seq_len, int mode,
                               int64_t mask_id, float* out, int64_t out_ld, int32_t* err_flag, void* stream) {
  RH_REQUIRE(table && ids && out, RH_ERR_INVALID_ARG, "rh_seq_pool_fwd: NULL pointer");
  RH_REQUIRE(vocab > 0 && dim > 0 && batch >= 0 && seq_len > 0, RH_ERR_INVALID_ARG, "rh_seq_pool_fwd: bad sizes");
  RH_REQUIRE(mode == 1 || mode == 2, RH_ERR_INVALID_ARG, "rh_seq_pool_fwd: mode must be 1 (sum) or 2 (mean)");
  if (batch == 0) return RH_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const bool i32 = ids_are_i32 != 0;
  if (dim % 4 == 0 && dim <= 128 && aligned16r(table) && aligned16r(out) && out_ld % 4 == 0) {
    const int lpr = pow2_ceil(dim / 4);
    const int threads = 128;
    const int grid = (batch + threads / lpr - 1) / (threads / lpr);
    switch (lpr) {
      case 1: seq_pool_fwd_v4<1><<<grid, threads, 0, st>>>(table, vocab, dim, ids, i32, batch, seq_len, mode, mask_id, out, out_ld, err_flag); break;
      case 2: seq_pool_fwd_v4<2><<<grid, threads, 0, st>>>(table, vocab, dim, ids, i32, batch, seq_len, mode, mask_id, out, out_ld, err_flag); break;
      case 4: seq_pool_fwd_v4<4><<<grid, threads, 0, st>>>(table, vocab, dim, ids, i32, batch, seq_len, mode, mask_id, out, out_ld, err_flag); break;
      case 8: seq_pool_fwd_v4<8><<<grid, threads, 0, st>>>(table, vocab, dim, ids, i32, batch, seq_len, mode, mask_id, out, out_ld, err_flag); break;
      case 16: seq_pool_fwd_v4<16><<<grid, threads, 0, st>>>(table, vocab, dim, ids, i32, batch, seq_len, mode, mask_id, out, out_ld, err_flag); break;
      default: seq_pool_fwd_v4<32><<<grid, threads, 0, st>>>(table, vocab, dim, ids, i32, batch, seq_len, mode, mask_id, out, out_ld, err_flag); break;
    }
  } else {
    const int threads = 128;
    const int grid = (batch + 3) / 4;
    seq_pool_fwd_scalar<<<grid, threads, 0, st>>>(table, vocab, dim, ids, i32, batch, seq_len, mode, mask_id, out, out_ld, err_flag);
  }
  RH_LAUNCH_CHECK();
  return RH_OK;
}

extern "C" int rh_seq_pool_bwd(float* table_grad, int vocab, int dim, int padding_idx, const void* ids, int ids_are_i32, int batch,
                               int seq_len, int mode, int64_t mask_id, const float* d_out, int64_t d_out_ld, int32_t* err_flag, void* stream) {
  RH_REQUIRE(table_grad && ids && d_out, RH_ERR_INVALID_ARG, "rh_seq_pool_bwd: NULL pointer");
  RH_REQUIRE(vocab > 0 && dim > 0 && batch >= 0 && seq_len > 0, RH_ERR_INVALID_ARG, "rh_seq_pool_bwd: bad sizes");
  RH_REQUIRE(mode == 1 || mode == 2, RH_ERR_INVALID_ARG, "rh_seq_pool_bwd: mode must be 1 (sum) or 2 (mean)");
  if (batch == 0) return RH_OK;
  const bool vec = dim % 4 == 0 && aligned16r(table_grad) && aligned16r(d_out) && d_out_ld % 4 == 0;
  const int64_t lanes = (int64_t)batch * seq_len * (vec ? dim / 4 : dim);
  seq_pool_bwd_kernel<<<grid_for(lanes, 256), 256, 0, (cudaStream_t)stream>>>(table_grad, vocab, dim, padding_idx, ids, ids_are_i32 != 0, batch,
                                                                              seq_len, mode, mask_id, d_out, d_out_ld, err_flag, vec);
  RH_LAUNCH_CHECK();
  return RH_OK;
}

extern "C" int rh_rowwise_update(float* table, float* table_grad, float* state1, float* state2, int32_t* stamp, int vocab, int dim,
                                 const void* ids, int ids_are_i32, int64_t n, int kind, const int32_t* step_dev, const float* lr_dev,
                                 const float* bias_corr_dev, int64_t state_row_stride, float beta1, float beta2, float eps, float weight_decay,
                                 void* stream) {
  RH_REQUIRE(table && table_grad && stamp && ids && step_dev && lr_dev, RH_ERR_INVALID_ARG, "rh_rowwise_update: NULL pointer");
  RH_REQUIRE(state_row_stride >= dim, RH_ERR_INVALID_ARG, "rh_rowwise_update: state_row_stride < dim");
  RH_REQUIRE(kind != 1 || bias_corr_dev != nullptr, RH_ERR_INVALID_ARG, "rh_rowwise_update: Adam needs bias_corr_dev");
  RH_REQUIRE(kind >= 0 && kind <= 2, RH_ERR_INVALID_ARG, "rh_rowwise_update: kind %d unknown", kind);
  RH_REQUIRE(kind == 0 || state1 != nullptr, RH_ERR_INVALID_ARG, "rh_rowwise_update: state1 required");
  RH_REQUIRE(kind != 1 || state2 != nullptr, RH_ERR_INVALID_ARG, "rh_rowwise_update: state2 required for Adam");
  RH_REQUIRE(vocab > 0 && dim > 0 && n >= 0, RH_ERR_INVALID_ARG, "rh_rowwise_update: bad sizes");
  if (n == 0) return RH_OK;
  const bool vec = dim % 4 == 0 && state_row_stride % 4 == 0 && aligned16r(table) && aligned16r(table_grad) &&
                   (state1 == nullptr || aligned16r(state1)) && (state2 == nullptr || aligned16r(state2));
  OptArgs a{kind, beta1, beta2, eps, weight_decay};
  int G = pow2_ceil(vec ? dim / 4 : dim);
  if (G > 32) G = 32;
  rowwise_update_kernel<<<grid_for(n * G, 256), 256, 0, (cudaStream_t)stream>>>(table, table_grad, state1, state2, stamp, vocab, dim, ids,
                                                                                ids_are_i32 != 0, n, a, step_dev, lr_dev, bias_corr_dev, vec, G, state_row_stride);
  RH_LAUNCH_CHECK();
  return RH_OK;
}

extern "C" int rh_opt_advance(int32_t* step_dev, float* bias_corr_dev, float beta1, float beta2, void* stream) {
  RH_REQUIRE(step_dev != nullptr, RH_ERR_INVALID_ARG, "rh_opt_advance: NULL");
  launch_k(opt_advance_kernel, dim3(1), dim3(1), 0, (cudaStream_t)stream, step_dev, bias_corr_dev, beta1, beta2);
  RH_LAUNCH_CHECK();
  return RH_OK;
}

static int pack_multi(rh::MultiP& p, const rh_field* fields, int n_fields, int dim, int batch, float* const* tables, float* const* state1,
                      float* const* state2, int32_t* const* stamp, bool need_opt) {
  RH_REQUIRE(n_fields > 0 && n_fields <= RH_MAX_FIELDS && fields != nullptr, RH_ERR_INVALID_ARG, "n_fields %d not in [1,%d]", n_fields,
             RH_MAX_FIELDS);
  RH_REQUIRE(dim > 0 && dim % 4 == 0, RH_ERR_UNSUPPORTED, "multi-table kernels need dim %% 4 == 0 (dim=%d)", dim);
  memset(&p, 0, sizeof(p));
  for (int i = 0; i < n_fields; ++i) {
    const rh_field& s = fields[i];
    RH_REQUIRE(s.table_grad != nullptr && s.ids != nullptr && s.vocab > 0, RH_ERR_INVALID_ARG, "field %d: grad/ids NULL", i);
    RH_REQUIRE(aligned16r(s.table_grad), RH_ERR_UNSUPPORTED, "field %d: gradient buffer not 16-byte aligned", i);
    RH_REQUIRE(s.id_stride >= 0 && s.id_stride < (int64_t)1 << 31, RH_ERR_INVALID_ARG, "field %d: id_stride out of range", i);
    p.grad[i] = s.table_grad;
    p.ids[i] = s.ids;
    p.id_stride[i] = (int32_t)s.id_stride;
    p.vocab[i] = s.vocab;
    p.is_i32[i] = (uint8_t)(s.ids_are_i32 != 0);
    if (need_opt) {
      RH_REQUIRE(tables && tables[i] && stamp && stamp[i], RH_ERR_INVALID_ARG, "field %d: table/stamp NULL", i);
      RH_REQUIRE(aligned16r(tables[i]), RH_ERR_UNSUPPORTED, "field %d: table not 16-byte aligned", i);
      p.table[i] = tables[i];
      p.stamp[i] = stamp[i];
      p.st1[i] = state1 ? state1[i] : nullptr;
      p.st2[i] = state2 ? state2[i] : nullptr;
    }
  }
  p.batch = batch;
  p.dim = dim;
  return RH_OK;
}

extern "C" int rh_fields_rowwise_update(const rh_field* fields, int n_fields, int dim, int batch, float* const* tables,
                                        float* const* state1, float* const* state2, int32_t* const* stamp, int kind,
                                        const int32_t* step_dev, const float* lr_dev, const float* bias_corr_dev, int64_t state_row_stride,
                                        float beta1, float beta2, float eps, float weight_decay, void* stream) {
  RH_REQUIRE(state_row_stride >= dim && state_row_stride % 4 == 0, RH_ERR_INVALID_ARG, "rh_fields_rowwise_update: state_row_stride %lld", (long long)state_row_stride);
  RH_REQUIRE(kind != 1 || bias_corr_dev != nullptr, RH_ERR_INVALID_ARG, "rh_fields_rowwise_update: Adam needs bias_corr_dev");
  RH_REQUIRE(kind >= 0 && kind <= 2, RH_ERR_INVALID_ARG, "rh_fields_rowwise_update: kind %d unknown", kind);
  RH_REQUIRE(step_dev && lr_dev, RH_ERR_INVALID_ARG, "rh_fields_rowwise_update: step/lr NULL");
  RH_REQUIRE(kind == 0 || state1 != nullptr, RH_ERR_INVALID_ARG, "rh_fields_rowwise_update: state1 required");
  RH_REQUIRE(kind != 1 || state2 != nullptr, RH_ERR_INVALID_ARG, "rh_fields_rowwise_update: state2 required for Adam");
  if (batch <= 0) return RH_OK;
  static thread_local rh::MultiP p;
  int rc = pack_multi(p, fields, n_fields, dim, batch, tables, state1, state2, stamp, true);
  if (rc != RH_OK) return rc;
  for (int i = 0; i < n_fields; ++i) {
    RH_REQUIRE(kind == 0 || p.st1[i] != nullptr, RH_ERR_INVALID_ARG, "field %d: state1 NULL", i);
    RH_REQUIRE(kind != 1 || p.st2[i] != nullptr, RH_ERR_INVALID_ARG, "field %d: state2 NULL", i);
  }
  int G = pow2_ceil(dim / 4);
  if (G > 32) G = 32;
  OptArgs a{kind, beta1, beta2, eps, weight_decay};
  const int threads = 128;
  dim3 grid((unsigned)(((int64_t)batch * G + threads - 1) / threads), n_fields);
  launch_k(fields_rowwise_update_kernel, grid, dim3(threads), 0, (cudaStream_t)stream, p, a, step_dev, lr_dev, bias_corr_dev, G, state_row_stride);
  RH_LAUNCH_CHECK();
  return RH_OK;
}

extern "C" int rh_fields_prefetch(const rh_field* fields, int n_fields, int dim, int batch, float* const* state, int64_t state_row_stride,
                                  void* stream) {
  RH_REQUIRE(n_fields > 0 && n_fields <= RH_MAX_FIELDS && fields != nullptr, RH_ERR_INVALID_ARG, "rh_fields_prefetch: n_fields %d not in [1,%d]",
             n_fields, RH_MAX_FIELDS);
  RH_REQUIRE(dim > 0 && state_row_stride >= 0, RH_ERR_INVALID_ARG, "rh_fields_prefetch: bad dim / state_row_stride");
  if (batch <= 0) return RH_OK;
  static thread_local rh::MultiP p;
  memset(&p, 0, sizeof(p));
  for (int i = 0; i < n_fields; ++i) {
    const rh_field& s = fields[i];
    RH_REQUIRE(s.table != nullptr && s.ids != nullptr && s.vocab > 0, RH_ERR_INVALID_ARG, "rh_fields_prefetch: field %d: table/ids NULL", i);
    RH_REQUIRE(s.id_stride >= 0 && s.id_stride < (int64_t)1 << 31, RH_ERR_INVALID_ARG, "rh_fields_prefetch: field %d: id_stride out of range", i);
    p.table[i] = const_cast<float*>(s.table);
    p.grad[i] = s.table_grad;
    p.st1[i] = state != nullptr ? state[i] : nullptr;
    p.ids[i] = s.ids;
    p.id_stride[i] = (int32_t)s.id_stride;
    p.vocab[i] = s.vocab;
    p.is_i32[i] = (uint8_t)(s.ids_are_i32 != 0);
  }
  p.batch = batch;
  p.dim = dim;
  dim3 grid((unsigned)((batch + 127) / 128), n_fields);
  fields_prefetch_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(p, state_row_stride * 4);
  RH_LAUNCH_CHECK();
  return RH_OK;
}

extern "C" int rh_fields_zero(const rh_field* fields, int n_fields, int dim, int batch, void* stream) {
  if (batch <= 0) return RH_OK;
  static thread_local rh::MultiP p;
  int rc = pack_multi(p, fields, n_fields, dim, batch, nullptr, nullptr, nullptr, nullptr, false);
  if (rc != RH_OK) return rc;
  const int threads = 128;
  dim3 grid((unsigned)(((int64_t)batch * (dim / 4) + threads - 1) / threads), n_fields);
  fields_zero_kernel<<<grid, threads, 0, (cudaStream_t)stream>>>(p);
  RH_LAUNCH_CHECK();
  return RH_OK;
}
