// rh_interact.cu — interaction layers on materialised tiles: FM (stand-alone) and CrossNetwork.
//
// Reference arithmetic replaced: FM.forward basic/layers.py:313-319; CrossNetwork.forward
// basic/layers.py:412-420 (x_{l+1} = x0 * <w_l, x_l> + b_l + x_l).  Both are HBM-bound row-local
// maps: one warp owns one sample, the row lives in registers across ALL cross layers, so the
// (batch, width) tile is read once and written once whatever the depth.
#include "rh_common.cuh"

namespace rh {

// ---- FM on (batch, n_fields, dim) ---------------------------------------------------------------
__global__ void __launch_bounds__(256) fm_fwd_kernel(const float* __restrict__ x, int batch, int n_fields, int dim, int reduce_sum,
                                                     float* __restrict__ y) {
  const int warp = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (warp >= batch) return;
  const float* xb = x + (int64_t)warp * n_fields * dim;
  float total = 0.f;
  for (int d0 = 0; d0 < dim; d0 += 32) {
    const int d = d0 + lane;
    float s = 0.f, ss = 0.f;
    if (d < dim) {
      for (int f = 0; f < n_fields; ++f) {
        const float v = __ldg(xb + f * dim + d);
        s += v;
        ss = fmaf(v, v, ss);
      }
    }
    const float ix = s * s - ss;
    if (reduce_sum) {
      total += ix;
    } else if (d < dim) {
      y[(int64_t)warp * dim + d] = 0.5f * ix;
    }
  }
  if (reduce_sum) {
    total = warp_sum(total);
    if (lane == 0) y[warp] = 0.5f * total;
  }
}

__global__ void __launch_bounds__(256) fm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ d_y, int batch, int n_fields,
                                                     int dim, int reduce_sum, float* __restrict__ d_x) {
  const int warp = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (warp >= batch) return;
  const float* xb = x + (int64_t)warp * n_fields * dim;
  float* dxb = d_x + (int64_t)warp * n_fields * dim;
  for (int d0 = 0; d0 < dim; d0 += 32) {
    const int d = d0 + lane;
    if (d >= dim) continue;
    float s = 0.f;
    for (int f = 0; f < n_fields; ++f) s += __ldg(xb + f * dim + d);
    const float g = reduce_sum ? __ldg(d_y + warp) : __ldg(d_y + (int64_t)warp * dim + d);
    for (int f = 0; f < n_fields; ++f) dxb[f * dim + d] = g * (s - __ldg(xb + f * dim + d));
  }
}

// ---- FM + LR on an already materialised tile ------------------------------------------------------
// The sharded (multi-GPU) front end receives its rows over NVLink instead of gathering them, so FM / LR run on the
// (batch, n_fields*dim) tile: same arithmetic as the fused gather kernel (rh_fields_fwd), rows read from the tile.
// lane = (sample, 16-byte quarter); each lane walks the fields of its sample.
template <int LPR>
__global__ void __launch_bounds__(128) tile_fm_lr_fwd_kernel(const float* __restrict__ tile, int64_t ld, int batch, int n_fields, int dim,
                                                             const float* __restrict__ lrw, const float* __restrict__ lrb,
                                                             float* __restrict__ yfm, float* __restrict__ ylr, float* __restrict__ fsum) {
  const int spb = blockDim.x / LPR;
  const int b = blockIdx.x * spb + (int)threadIdx.x / LPR;
  const int q = (int)threadIdx.x % LPR;
  const bool live = b < batch, lane_on = live && 4 * q < dim;
  float4 s = f4_zero();
  float ss = 0.f, lr = 0.f;
  if (lane_on) {
    const float* row = tile + (int64_t)b * ld + 4 * q;
#pragma unroll 4
    for (int f = 0; f < n_fields; ++f) {
      const float4 v = ldg_row16(row + (int64_t)f * dim);
      s = f4_add(s, v);
      ss += f4_dot(v, v);
      if (lrw != nullptr) lr += f4_dot(v, __ldg(reinterpret_cast<const float4*>(lrw + (int64_t)f * dim + 4 * q)));
    }
  }
  float t = (s.x * s.x + s.y * s.y) + (s.z * s.z + s.w * s.w) - ss;
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) {
    t += __shfl_xor_sync(0xffffffffu, t, o);
    lr += __shfl_xor_sync(0xffffffffu, lr, o);
  }
  if (live && q == 0) {
    if (yfm != nullptr) yfm[b] = 0.5f * t;
    if (ylr != nullptr) ylr[b] = lr + (lrb != nullptr ? __ldg(lrb) : 0.f);
  }
  if (fsum != nullptr && lane_on) *reinterpret_cast<float4*>(fsum + (int64_t)b * dim + 4 * q) = s;
}

// d_tile[b, f] (+)= d_yfm[b] * (S[b] - e) + d_ylr[b] * w[f];   d_lrw[f] += sum_b d_ylr[b] * e;   d_lrb += sum_b d_ylr[b]
// grid = (sample chunks, fields): the LR weight gradient of a field reduces inside the block.
template <int LPR>
__global__ void __launch_bounds__(128) tile_fm_lr_bwd_kernel(const float* __restrict__ tile, int64_t ld, int batch, int n_fields, int dim,
                                                             const float* __restrict__ lrw, const float* __restrict__ fsum,
                                                             const float* __restrict__ dyfm, const float* __restrict__ dylr,
                                                             float* __restrict__ d_tile, int64_t d_ld, int accumulate,
                                                             float* __restrict__ d_lrw, float* __restrict__ d_lrb) {
  __shared__ float4 sm_dw[4][32];
  __shared__ float sm_db[4];
  constexpr int ITER = 4;
  const int f = blockIdx.y;
  const int spi = blockDim.x / LPR;
  const int q = (int)threadIdx.x % LPR, sl = (int)threadIdx.x / LPR;
  const bool lane_on = 4 * q < dim;
  const int base = blockIdx.x * spi * ITER;
  float4 w = f4_zero();
  if (lrw != nullptr && dylr != nullptr && lane_on) w = __ldg(reinterpret_cast<const float4*>(lrw + (int64_t)f * dim + 4 * q));
  float4 dw = f4_zero();
  float db = 0.f;
#pragma unroll
  for (int i = 0; i < ITER; ++i) {
    const int b = base + i * spi + sl;
    if (b < batch && lane_on) {
      const float4 e = ldg_row16(tile + (int64_t)b * ld + (int64_t)f * dim + 4 * q);
      const float gf = dyfm != nullptr ? __ldg(dyfm + b) : 0.f;
      const float gl = dylr != nullptr ? __ldg(dylr + b) : 0.f;
      float4 g = f4_zero();
      if (dyfm != nullptr) {
        const float4 S = ldg_row16(fsum + (int64_t)b * dim + 4 * q);
        g = make_float4(gf * (S.x - e.x), gf * (S.y - e.y), gf * (S.z - e.z), gf * (S.w - e.w));
      }
      g = f4_fma(w, make_float4(gl, gl, gl, gl), g);
      float* dst = d_tile + (int64_t)b * d_ld + (int64_t)f * dim + 4 * q;
      if (accumulate) g = f4_add(g, *reinterpret_cast<const float4*>(dst));
      *reinterpret_cast<float4*>(dst) = g;
      dw = f4_fma(e, make_float4(gl, gl, gl, gl), dw);
      if (q == 0) db += gl;
    }
  }
  const bool want_dw = d_lrw != nullptr && dylr != nullptr;
  const bool want_db = f == 0 && d_lrb != nullptr && dylr != nullptr;
  if (want_dw || want_db) {
#pragma unroll
    for (int o = 16; o >= LPR; o >>= 1) {
      dw.x += __shfl_xor_sync(0xffffffffu, dw.x, o);
      dw.y += __shfl_xor_sync(0xffffffffu, dw.y, o);
      dw.z += __shfl_xor_sync(0xffffffffu, dw.z, o);
      dw.w += __shfl_xor_sync(0xffffffffu, dw.w, o);
    }
    db = warp_sum(db);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane < LPR) sm_dw[warp][lane] = dw;
    if (lane == 0) sm_db[warp] = db;
    __syncthreads();
    const int nwarp = blockDim.x >> 5;
    if (warp == 0) {
      if (want_dw && lane < LPR && 4 * lane < dim) {
        float4 t = sm_dw[0][lane];
        for (int k = 1; k < nwarp; ++k) t = f4_add(t, sm_dw[k][lane]);
        red_add_row16(d_lrw + (int64_t)f * dim + 4 * lane, t);
      }
      if (want_db && lane == 0) {
        float t = sm_db[0];
        for (int k = 1; k < nwarp; ++k) t += sm_db[k];
        atomicAdd(d_lrb, t);
      }
    }
  }
}

// ---- CrossNetwork ------------------------------------------------------------------------------
// Per-layer parameters stay where the reference keeps them (one Linear(width,1) weight and one bias
// vector per layer, basic/layers.py:409-410): the kernels take pointer tables, not stacked copies.
constexpr int kCrossMaxLayers = 16;
struct CrossPtrs {
  const float* w[kCrossMaxLayers];
  const float* b[kCrossMaxLayers];
  float* dw[kCrossMaxLayers];
  float* db[kCrossMaxLayers];
};

// KMAX registers per lane hold the row: column c = k*32 + lane.
template <int KMAX>
__global__ void __launch_bounds__(256) cross_fwd_kernel(const float* __restrict__ x0p, int64_t x_ld, int batch, int width, int n_layers,
                                                        const __grid_constant__ CrossPtrs cp,
                                                        float* __restrict__ out, int64_t out_ld, float* __restrict__ xw_saved) {
  const int lane = threadIdx.x & 31;
  const int warps_total = (int)((gridDim.x * (int64_t)blockDim.x) >> 5);
  for (int b = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5); b < batch; b += warps_total) {
    float x0[KMAX], x[KMAX];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      const int c = k * 32 + lane;
      x0[k] = c < width ? __ldg(x0p + (int64_t)b * x_ld + c) : 0.f;
      x[k] = x0[k];
    }
    for (int l = 0; l < n_layers; ++l) {
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < KMAX; ++k) {
        const int c = k * 32 + lane;
        if (c < width) s = fmaf(__ldg(cp.w[l] + c), x[k], s);
      }
      s = warp_sum(s);
      if (lane == 0 && xw_saved != nullptr) xw_saved[(int64_t)l * batch + b] = s;
#pragma unroll
      for (int k = 0; k < KMAX; ++k) {
        const int c = k * 32 + lane;
        if (c < width) x[k] = fmaf(x0[k], s, __ldg(cp.b[l] + c)) + x[k];
      }
    }
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      const int c = k * 32 + lane;
      if (c < width) out[(int64_t)b * out_ld + c] = x[k];
    }
  }
}

// backward: recomputes x_l exactly as the forward did (from x0 and the saved scalars) and walks the
// layers in reverse.  Batch reductions for d_w / d_b: registers over the SPW samples a warp holds ->
// the warp's PRIVATE shared-memory slab (plain read-modify-write, no atomics) -> one pass over the
// block's slabs -> 16-byte vector RED per block.
constexpr int kCrossSPW = 2;  // samples a warp keeps in registers at once

template <int KMAX>
__global__ void __launch_bounds__(256) cross_bwd_kernel(const float* __restrict__ x0p, int64_t x_ld, int batch, int width, int n_layers,
                                                        const __grid_constant__ CrossPtrs cp,
                                                        const float* __restrict__ xw_saved, const float* __restrict__ d_out,
                                                        int64_t d_out_ld, float* __restrict__ d_x0, int64_t d_x0_ld) {
  extern __shared__ float sm[];  // [warps][2][n_layers][width]
  constexpr int SPW = kCrossSPW;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int wpb = blockDim.x >> 5;
  const int lw = n_layers * width;
  const int slab = 2 * lw;
  for (int i = threadIdx.x; i < wpb * slab; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  float* my_dw = sm + (int64_t)warp * slab;
  float* my_db = my_dw + lw;

  const int warps_total = gridDim.x * wpb;
  const int n_chunks = (batch + SPW - 1) / SPW;
  for (int chunk = blockIdx.x * wpb + warp; chunk < n_chunks; chunk += warps_total) {
    float x0[SPW][KMAX], g[SPW][KMAX], gx0[SPW][KMAX];
#pragma unroll
    for (int s = 0; s < SPW; ++s) {
      const int b = chunk * SPW + s;
#pragma unroll
      for (int k = 0; k < KMAX; ++k) {
        const int c = k * 32 + lane;
        const bool on = c < width && b < batch;
        x0[s][k] = on ? __ldg(x0p + (int64_t)b * x_ld + c) : 0.f;
        g[s][k] = on ? __ldg(d_out + (int64_t)b * d_out_ld + c) : 0.f;
        gx0[s][k] = 0.f;
      }
    }
    for (int l = n_layers - 1; l >= 0; --l) {
      float adw[KMAX], adb[KMAX];
#pragma unroll
      for (int k = 0; k < KMAX; ++k) adw[k] = adb[k] = 0.f;
#pragma unroll
      for (int s = 0; s < SPW; ++s) {
        const int b = chunk * SPW + s;
        if (b >= batch) continue;  // warp-uniform
        float x[KMAX];
#pragma unroll
        for (int k = 0; k < KMAX; ++k) x[k] = x0[s][k];
        for (int j = 0; j < l; ++j) {  // x_l: l forward steps, bit-identical to the forward pass
          const float sj = __ldg(xw_saved + (int64_t)j * batch + b);
#pragma unroll
          for (int k = 0; k < KMAX; ++k) {
            const int c = k * 32 + lane;
            if (c < width) x[k] = fmaf(x0[s][k], sj, __ldg(cp.b[j] + c)) + x[k];
          }
        }
        const float sl = __ldg(xw_saved + (int64_t)l * batch + b);
        float t = 0.f;  // dL/ds_l = <g_{l+1}, x0>
#pragma unroll
        for (int k = 0; k < KMAX; ++k) t = fmaf(g[s][k], x0[s][k], t);
        t = warp_sum(t);
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
          const int c = k * 32 + lane;
          if (c < width) {
            adb[k] += g[s][k];
            adw[k] = fmaf(t, x[k], adw[k]);
            gx0[s][k] = fmaf(g[s][k], sl, gx0[s][k]);
            g[s][k] = fmaf(t, __ldg(cp.w[l] + c), g[s][k]);
          }
        }
      }
#pragma unroll
      for (int k = 0; k < KMAX; ++k) {
        const int c = k * 32 + lane;
        if (c < width) {
          my_dw[l * width + c] += adw[k];
          my_db[l * width + c] += adb[k];
        }
      }
    }
#pragma unroll
    for (int s = 0; s < SPW; ++s) {
      const int b = chunk * SPW + s;
#pragma unroll
      for (int k = 0; k < KMAX; ++k) {
        const int c = k * 32 + lane;
        if (c < width && b < batch) d_x0[(int64_t)b * d_x0_ld + c] = g[s][k] + gx0[s][k];
      }
    }
  }
  __syncthreads();
  // block reduction over the warps' slabs, then one RED per element per block
  for (int i = threadIdx.x; i < slab; i += blockDim.x) {
    float t = 0.f;
    for (int wv = 0; wv < wpb; ++wv) t += sm[(int64_t)wv * slab + i];
    const int j = i < lw ? i : i - lw;
    const int l = j / width, c = j - l * width;
    if (i < lw) atomicAdd(cp.dw[l] + c, t);
    else atomicAdd(cp.db[l] + c, t);
  }
}

template <int KMAX>
static int launch_cross(bool fwd, const float* x0, int64_t x_ld, int batch, int width, int n_layers, const CrossPtrs& cp,
                        float* out, int64_t out_ld, float* xw_saved, const float* d_out, int64_t d_out_ld, float* d_x0, int64_t d_x0_ld,
                        cudaStream_t st) {
  const int threads = 256;
  const int wpb = threads / 32;
  int grid = (batch + wpb - 1) / wpb;
  if (fwd) {
    const int cap = num_sms() * 8;
    if (grid > cap) grid = cap;
    cross_fwd_kernel<KMAX><<<grid, threads, 0, st>>>(x0, x_ld, batch, width, n_layers, cp, out, out_ld, xw_saved);
  } else {
    const int cap = num_sms();  // one block per SM: one RED set per SM for d_w / d_b
    grid = (batch + wpb * kCrossSPW - 1) / (wpb * kCrossSPW);
    if (grid > cap) grid = cap;
    const size_t smem = (size_t)wpb * 2 * n_layers * width * sizeof(float);
    if (smem > 200 * 1024) {
      set_error("rh_cross_bwd: n_layers*width = %d needs %zu bytes of shared memory", n_layers * width, smem);
      return RH_ERR_UNSUPPORTED;
    }
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(cross_bwd_kernel<KMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) {
        set_error("rh_cross_bwd: %zu bytes of shared memory not available: %s", smem, cudaGetErrorString(e));
        return RH_ERR_UNSUPPORTED;
      }
    }
    cross_bwd_kernel<KMAX><<<grid, threads, smem, st>>>(x0, x_ld, batch, width, n_layers, cp, xw_saved, d_out, d_out_ld, d_x0, d_x0_ld);
  }
  RH_LAUNCH_CHECK();
  return RH_OK;
}

static int dispatch_cross(bool fwd, const float* x0, int64_t x_ld, int batch, int width, int n_layers, const CrossPtrs& cp,
                          float* out, int64_t out_ld, float* xw_saved, const float* d_out, int64_t d_out_ld, float* d_x0,
                          int64_t d_x0_ld, cudaStream_t st) {
  const int k = (width + 31) / 32;
#define RH_CROSS(K) return launch_cross<K>(fwd, x0, x_ld, batch, width, n_layers, cp, out, out_ld, xw_saved, d_out, d_out_ld, d_x0, d_x0_ld, st)
  if (k <= 4) RH_CROSS(4);
  if (k <= 8) RH_CROSS(8);
  if (k <= 16) RH_CROSS(16);
  if (k <= 32) RH_CROSS(32);
  if (k <= 64) RH_CROSS(64);
#undef RH_CROSS
  set_error("cross network width %d > 2048 is not supported by the register-resident kernel", width);
  return RH_ERR_UNSUPPORTED;
}

}  // namespace rh

using namespace rh;

extern "C" int rh_fm_fwd(const float* x, int batch, int n_fields, int dim, int reduce_sum, float* y, void* stream) {
  RH_REQUIRE(x && y, RH_ERR_INVALID_ARG, "rh_fm_fwd: NULL pointer");
  RH_REQUIRE(batch >= 0 && n_fields > 0 && dim > 0, RH_ERR_INVALID_ARG, "rh_fm_fwd: bad sizes");
  if (batch == 0) return RH_OK;
  const int wpb = 8;
  fm_fwd_kernel<<<(batch + wpb - 1) / wpb, wpb * 32, 0, (cudaStream_t)stream>>>(x, batch, n_fields, dim, reduce_sum, y);
  RH_LAUNCH_CHECK();
  return RH_OK;
}

extern "C" int rh_fm_bwd(const float* x, const float* d_y, int batch, int n_fields, int dim, int reduce_sum, float* d_x, void* stream) {
  RH_REQUIRE(x && d_y && d_x, RH_ERR_INVALID_ARG, "rh_fm_bwd: NULL pointer");
  RH_REQUIRE(batch >= 0 && n_fields > 0 && dim > 0, RH_ERR_INVALID_ARG, "rh_fm_bwd: bad sizes");
  if (batch == 0) return RH_OK;
  const int wpb = 8;
  fm_bwd_kernel<<<(batch + wpb - 1) / wpb, wpb * 32, 0, (cudaStream_t)stream>>>(x, d_y, batch, n_fields, dim, reduce_sum, d_x);
  RH_LAUNCH_CHECK();
  return RH_OK;
}

static int pack_cross(CrossPtrs& cp, int n_layers, const float* const* w, const float* const* b, float* const* d_w, float* const* d_b) {
  RH_REQUIRE(n_layers >= 0 && n_layers <= kCrossMaxLayers, RH_ERR_UNSUPPORTED, "cross network: %d layers > %d", n_layers, kCrossMaxLayers);
  memset(&cp, 0, sizeof(cp));
  for (int l = 0; l < n_layers; ++l) {
    RH_REQUIRE(w && b && w[l] && b[l], RH_ERR_INVALID_ARG, "cross network: layer %d weight/bias NULL", l);
    cp.w[l] = w[l];
    cp.b[l] = b[l];
    if (d_w != nullptr) {
      RH_REQUIRE(d_b && d_w[l] && d_b[l], RH_ERR_INVALID_ARG, "cross network: layer %d gradient buffer NULL", l);
      cp.dw[l] = d_w[l];
      cp.db[l] = d_b[l];
    }
  }
  return RH_OK;
}

extern "C" int rh_cross_fwd(const float* x0, int64_t x_ld, int batch, int width, int n_layers, const float* const* w, const float* const* b,
                            float* out, int64_t out_ld, float* xw_saved, void* stream) {
  RH_REQUIRE(x0 && out, RH_ERR_INVALID_ARG, "rh_cross_fwd: NULL pointer");
  RH_REQUIRE(batch >= 0 && width > 0 && x_ld >= width && out_ld >= width, RH_ERR_INVALID_ARG, "rh_cross_fwd: bad sizes");
  CrossPtrs cp;
  int rc = pack_cross(cp, n_layers, w, b, nullptr, nullptr);
  if (rc != RH_OK) return rc;
  if (batch == 0) return RH_OK;
  return dispatch_cross(true, x0, x_ld, batch, width, n_layers, cp, out, out_ld, xw_saved, nullptr, 0, nullptr, 0, (cudaStream_t)stream);
}

extern "C" int rh_cross_bwd(const float* x0, int64_t x_ld, int batch, int width, int n_layers, const float* const* w, const float* const* b,
                            const float* xw_saved, const float* d_out, int64_t d_out_ld, float* d_x0, int64_t d_x0_ld, float* const* d_w,
                            float* const* d_b, void* stream) {
  RH_REQUIRE(x0 && d_out && d_x0, RH_ERR_INVALID_ARG, "rh_cross_bwd: NULL pointer");
  RH_REQUIRE(batch >= 0 && width > 0, RH_ERR_INVALID_ARG, "rh_cross_bwd: bad sizes");
  RH_REQUIRE(n_layers == 0 || (xw_saved && d_w && d_b), RH_ERR_INVALID_ARG, "rh_cross_bwd: NULL pointer");
  CrossPtrs cp;
  int rc = pack_cross(cp, n_layers, w, b, d_w, d_b);
  if (rc != RH_OK) return rc;
  if (batch == 0) return RH_OK;
  if (n_layers == 0) {  // identity network: d_x0 = d_out
    cudaError_t e = cudaMemcpy2DAsync(d_x0, d_x0_ld * sizeof(float), d_out, d_out_ld * sizeof(float), (size_t)width * sizeof(float),
                                      (size_t)batch, cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
    RH_REQUIRE(e == cudaSuccess, RH_ERR_CUDA, "rh_cross_bwd: copy failed: %s", cudaGetErrorString(e));
    return RH_OK;
  }
  return dispatch_cross(false, x0, x_ld, batch, width, n_layers, cp, nullptr, 0, const_cast<float*>(xw_saved), d_out, d_out_ld, d_x0,
                        d_x0_ld, (cudaStream_t)stream);
}

static bool al16i(const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

extern "C" int rh_tile_fm_lr_fwd(const float* tile, int64_t tile_ld, int batch, int n_fields, int dim, const float* lr_weight,
                                 const float* lr_bias, float* y_fm, float* y_lr, float* field_sum, void* stream) {
  RH_REQUIRE(tile != nullptr && batch >= 0 && n_fields > 0 && dim > 0, RH_ERR_INVALID_ARG, "rh_tile_fm_lr_fwd: bad arguments");
  RH_REQUIRE(dim % 4 == 0 && dim <= 128 && tile_ld % 4 == 0 && al16i(tile) && al16i(lr_weight) && al16i(field_sum), RH_ERR_UNSUPPORTED,
             "rh_tile_fm_lr_fwd needs dim %% 4 == 0 (<= 128) and 16-byte aligned rows (dim=%d ld=%lld)", dim, (long long)tile_ld);
  if (batch == 0) return RH_OK;
  const int lpr = pow2_ceil(dim / 4), threads = 128;
  const int grid = (batch + threads / lpr - 1) / (threads / lpr);
  cudaStream_t st = (cudaStream_t)stream;
#define RH_TF(L) tile_fm_lr_fwd_kernel<L><<<grid, threads, 0, st>>>(tile, tile_ld, batch, n_fields, dim, lr_weight, lr_bias, y_fm, y_lr, field_sum)
  switch (lpr) {
    case 1: RH_TF(1); break;
    case 2: RH_TF(2); break;
    case 4: RH_TF(4); break;
    case 8: RH_TF(8); break;
    case 16: RH_TF(16); break;
    default: RH_TF(32); break;
  }
#undef RH_TF
  RH_LAUNCH_CHECK();
  return RH_OK;
}

extern "C" int rh_tile_fm_lr_bwd(const float* tile, int64_t tile_ld, int batch, int n_fields, int dim, const float* lr_weight,
                                 const float* field_sum, const float* d_y_fm, const float* d_y_lr, float* d_tile, int64_t d_tile_ld,
                                 int accumulate, float* d_lr_weight, float* d_lr_bias, void* stream) {
  RH_REQUIRE(tile != nullptr && d_tile != nullptr && batch >= 0 && n_fields > 0 && dim > 0, RH_ERR_INVALID_ARG, "rh_tile_fm_lr_bwd: bad arguments");
  RH_REQUIRE(d_y_fm == nullptr || field_sum != nullptr, RH_ERR_INVALID_ARG, "rh_tile_fm_lr_bwd: d_y_fm needs field_sum");
  RH_REQUIRE(d_y_lr == nullptr || lr_weight != nullptr, RH_ERR_INVALID_ARG, "rh_tile_fm_lr_bwd: d_y_lr needs lr_weight");
  RH_REQUIRE(dim % 4 == 0 && dim <= 128 && tile_ld % 4 == 0 && d_tile_ld % 4 == 0 && al16i(tile) && al16i(d_tile) && al16i(lr_weight) &&
                 al16i(field_sum) && al16i(d_lr_weight),
             RH_ERR_UNSUPPORTED, "rh_tile_fm_lr_bwd needs dim %% 4 == 0 (<= 128) and 16-byte aligned rows");
  if (batch == 0) return RH_OK;
  const int lpr = pow2_ceil(dim / 4), threads = 128;
  const int spb = threads / lpr * 4;
  dim3 grid((batch + spb - 1) / spb, n_fields);
  cudaStream_t st = (cudaStream_t)stream;
#define RH_TB(L)                                                                                                                      \
  tile_fm_lr_bwd_kernel<L><<<grid, threads, 0, st>>>(tile, tile_ld, batch, n_fields, dim, lr_weight, field_sum, d_y_fm, d_y_lr, d_tile, \
                                                     d_tile_ld, accumulate, d_lr_weight, d_lr_bias)
  switch (lpr) {
    case 1: RH_TB(1); break;
    case 2: RH_TB(2); break;
    case 4: RH_TB(4); break;
    case 8: RH_TB(8); break;
    case 16: RH_TB(16); break;
    default: RH_TB(32); break;
  }
#undef RH_TB
  RH_LAUNCH_CHECK();
  return RH_OK;
}
