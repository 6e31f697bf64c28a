// rh_bnfuse.cu — training-mode BatchNorm1d + activation + dropout of a tower layer as ONE launch each way, optionally
// fused with the tower's output head (Linear(cols, 1) + per-sample side terms + sigmoid).
//
// Reference arithmetic replaced: the [BatchNorm1d -> activation -> Dropout] part of MLP.forward (basic/layers.py:282-285),
// Dice.forward (basic/activation.py:15-25) and, in head mode, the MLP's output layer + the model's tail
// (basic/layers.py:279-280, models/ranking/deepfm.py:41-43) — and their backward.
//
// Why one launch: training-mode BatchNorm needs the column statistics of ALL rows before any element can be normalised
// (forward), and the column sums of dz and dz * xhat before any d_h element is final (backward).  The two-kernel forms
// (rh_colstats + rh_bn_act_fwd; rh_bn_act_bwd's two passes) cost 2 x ~6-9 us at batch 4096 — launch ramps and a second trip
// through L2 for a 2-4 MB activation whose streaming time is under 2 us.  Here every CTA keeps ITS rows in registers across a
// grid-wide barrier: phase 1 loads the rows once and publishes per-column partial sums (block reduction in shared memory, one
// atomicAdd per column per CTA), the barrier waits until every CTA has published, phase 2 finishes from registers.  The
// activation `y` of the last hidden layer never exists in HBM in head mode: the head's dot product consumes it in registers.
//
// Grid barrier: grid <= number of SMs, one 512-thread CTA each, so all CTAs are co-resident whatever else runs (a spinning
// CTA can only wait for CTAs of its own grid that are resident or about to be placed; nothing queued behind this kernel in
// its stream can start before it ends).  The scratch holds a launch counter, a never-reset arrival counter and TWO sum buffers
// used alternately: launch g accumulates into buffer g & 1 and CTA 0 zeroes the other one for launch g + 1, so no CTA has to
// be "the last one out" and graph replays need no memset.
//
// Mapping: warp = row (RPW rows per warp, interleaved by 16), lane = 4 consecutive columns per step, KMAX steps per row:
// RPW * KMAX * 4 floats of h per thread (x2 in backward).  Shapes beyond that budget (rows > SMs * 16 * RPW) report
// rh_bn_fused_supported() == 0 and callers stay on the two-kernel route (DIN's 204 800-row attention MLP).
// Tuning history (tools/bnfuse_trace.cu, ncu): the first version (256 threads, run-time activation switch inside the unrolled
// element loops, last-CTA finalisation) spent 8 us in its apply phase at 10 % issue-active — 90 instructions per element, 8 warps
// per SM — and 2.8 us in the serial tail; 16 warps, a compile-time activation and the parity buffers address exactly those.
#ifndef RH_PDL_FAMILY  // (the trace tools include several of these files into one unit: the first one names the family)
#define RH_PDL_FAMILY 2  /* rh_set_pdl mask bit of this file's kernels */
#endif
#include "rh_bn_common.cuh"

namespace rh {

struct BnFuseP {
  const float* h;
  int64_t h_ld;
  int64_t rows;
  int cols;
  float bn_eps;
  const float* gamma;
  const float* beta;
  int act;
  const float* alpha;
  float dice_eps;
  float p_drop;
  uint32_t seed;
  float* running_mean;
  float* running_var;
  long long* nbt;
  float momentum;
  float* stats;    // (2 cols + 1): mean | biased var | step-counter bits      (forward: out; backward: in)
  float* scratch;  // rh_bn_fused_scratch_floats(cols): generation + two parity buffers (see FuseScratch)
  float* y;
  int64_t y_ld;
  // head
  const float* head_w;
  const float* head_b;
  const float* e0;
  const float* e1;
  int apply_sigmoid;
  float* head_out;
  // backward
  const float* d_y;
  int64_t d_y_ld;
  const float* d_head_out;
  float* d_h;
  int64_t d_h_ld;
  float* d_gamma;
  float* d_beta;
  float* d_alpha;
  float* d_head_w;
  float* d_head_b;
  float* d_extra;
  float* d_lin_bias;
#ifdef RH_BN_TRACE
  unsigned long long* trace;  // tools/bnfuse_trace.cu: [cta][8] clock64 stamps
#endif
};

#ifdef RH_BN_TRACE
#define RH_BT(ev)                                                                        \
  do {                                                                                   \
    if (p.trace != nullptr && threadIdx.x == 0) p.trace[(size_t)blockIdx.x * 8 + (ev)] = clock64(); \
  } while (0)
#else
#define RH_BT(ev) \
  do {            \
  } while (0)
#endif

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

constexpr int kFuseThreads = 512;  // 16 warps: at 8 warps / SM the kernel was bound by instruction latency (ncu: 10 % issue-active)
constexpr int kFuseWarps = kFuseThreads / 32;

// Scratch layout (floats): [0] generation (launches so far, uint), [1..31] spare | two parity buffers of kFuseHdr + 3 cols floats
// (rounded up to 32: every buffer starts on its own 128-byte line): [0] arrivals of THIS launch (uint), [1] alpha, [2] head bias,
// [32..) sums | sums | sums.  Launch g uses buffer g & 1 (all zero on entry) and, after its barrier, CTA 0 zeroes buffer
// (g + 1) & 1 — dirty since launch g - 1, whose CTAs have all exited — and stores g + 1.  Nothing has to wait for "the last CTA
// out", and launches of different grid sizes can share one scratch (the arrival count starts at zero for every launch; a
// never-reset counter compared against (g + 1) * n_ctas deadlocked as soon as a shorter last batch changed the grid).
constexpr int kFuseHdr = 32;
struct FuseScratch {
  unsigned gen;
  unsigned* arrive;
  float* extra;  // [1] alpha, [2] head bias
  float* sums;   // this launch's buffer
  float* other;  // the whole buffer the next launch will use (header included)
  int other_floats;
};
__host__ __device__ __forceinline__ int fuse_buffer_floats(int cols) { return kFuseHdr + ((3 * cols + 31) / 32) * 32; }
__device__ __forceinline__ FuseScratch fuse_scratch(float* scratch, int cols) {
  FuseScratch f;
  unsigned* hdr = reinterpret_cast<unsigned*>(scratch);
  f.gen = hdr[0];
  const int per = fuse_buffer_floats(cols);
  float* mine = scratch + 32 + (size_t)(f.gen & 1u) * per;
  f.arrive = reinterpret_cast<unsigned*>(mine);
  f.extra = mine;
  f.sums = mine + kFuseHdr;
  f.other = scratch + 32 + (size_t)((f.gen + 1u) & 1u) * per;
  f.other_floats = per;
  return f;
}

// every CTA of the grid has executed everything before this call once any CTA returns from it
__device__ __forceinline__ void grid_barrier(const FuseScratch& f, unsigned n_ctas) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(f.arrive, 1u);
    while (ld_acquire_u32(f.arrive) < n_ctas) __nanosleep(20);
    __threadfence();
  }
  __syncthreads();
}

template <int ACT>
__device__ __forceinline__ float act_value(int act_rt, float z, float alpha, float ps) {
  const int act = ACT >= 0 ? ACT : act_rt;
  switch (act) {
    case ACT_RELU: return fmaxf(z, 0.f);
    case ACT_DICE: return ps * z + (1.f - ps) * alpha * z;
    case ACT_PRELU: return z > 0.f ? z : alpha * z;
    case ACT_SIGMOID: return sigmoidf_precise(z);
    case ACT_LEAKY: return z > 0.f ? z : 0.01f * z;
    default: return z;
  }
}

template <int KMAX>
__device__ __forceinline__ void load_colconst(const float* src, int lane, int cols, float (&out)[KMAX][4], float fill) {
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    const int c0 = (k * 32 + lane) * 4;
    out[k][0] = out[k][1] = out[k][2] = out[k][3] = fill;
    if (c0 < cols && src != nullptr) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(src + c0));
      out[k][0] = v.x; out[k][1] = v.y; out[k][2] = v.z; out[k][3] = v.w;
    }
  }
}

// Dice row statistics over the `cols` valid entries of z (padding lanes hold 0 and are skipped in the variance)
template <int KMAX>
__device__ __forceinline__ void dice_row_stats(const float (&z)[KMAX][4], int lane, int cols, float dice_eps, float& m, float& inv_s) {
  float t = 0.f;
#pragma unroll
  for (int k = 0; k < KMAX; ++k)
#pragma unroll
    for (int j = 0; j < 4; ++j) t += z[k][j];
  m = warp_sum(t) * (1.f / (float)cols);
  float qv = 0.f;
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    if ((k * 32 + lane) * 4 < cols) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float d = z[k][j] - m;
        qv += fmaf(d, d, dice_eps);
      }
    }
  }
  inv_s = 1.f / sqrtf(warp_sum(qv));
}

// =====================================================================================================
// forward.   ACT: compile-time activation (ACT_RELU / ACT_DICE), or -1 = read p.act (the rarer ones share one instantiation:
// with the activation a run-time switch inside the fully unrolled element loops the kernel executed ~90 instructions per element)
// =====================================================================================================
template <int KMAX, int RPW, int ACT, bool HEAD>
__global__ void __launch_bounds__(kFuseThreads) bn_fused_fwd_kernel(const BnFuseP p) {
  extern __shared__ float smem[];  // [16 warps][2 cols]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cols = p.cols;
  const int act = ACT >= 0 ? ACT : p.act;
  const int64_t row0 = (int64_t)blockIdx.x * kFuseWarps * RPW;
  pdl_wait();
  const FuseScratch fs = fuse_scratch(p.scratch, cols);
  // dropout stream id of this forward = num_batches_tracked + 1 (CTA 0 stores the incremented value after the barrier)
  const long long count = p.nbt != nullptr ? *p.nbt + 1 : 0;
  const uint32_t counter = (uint32_t)(count & 0x7fffffff);
  RH_BT(0);

  // ---- phase 1: rows -> registers, shifted column sums ----
  float hv[RPW][KMAX][4];
  float sh[KMAX][4];
  load_colconst<KMAX>(p.h, lane, cols, sh, 0.f);  // shift = row 0 (keeps fp32 accurate when |mean| >> std)
#pragma unroll
  for (int i = 0; i < RPW; ++i) {
    const int64_t row = row0 + warp + kFuseWarps * i;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      const int c0 = (k * 32 + lane) * 4;
      hv[i][k][0] = hv[i][k][1] = hv[i][k][2] = hv[i][k][3] = 0.f;
      if (row < p.rows && c0 < cols) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(p.h + row * p.h_ld + c0));
        hv[i][k][0] = v.x; hv[i][k][1] = v.y; hv[i][k][2] = v.z; hv[i][k][3] = v.w;
      }
    }
  }
  {
    float* my = smem + (int64_t)warp * 2 * cols;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      const int c0 = (k * 32 + lane) * 4;
      if (c0 < cols) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float s1 = 0.f, s2 = 0.f;
#pragma unroll
          for (int i = 0; i < RPW; ++i) {
            if (row0 + warp + kFuseWarps * i < p.rows) {
              const float d = hv[i][k][j] - sh[k][j];
              s1 += d;
              s2 = fmaf(d, d, s2);
            }
          }
          my[c0 + j] = s1;
          my[cols + c0 + j] = s2;
        }
      }
    }
  }
  __syncthreads();
  RH_BT(1);
  for (int i = threadIdx.x; i < 2 * cols; i += blockDim.x) {
    float t = 0.f;
#pragma unroll
    for (int wv = 0; wv < kFuseWarps; ++wv) t += smem[(int64_t)wv * 2 * cols + i];
    atomicAdd(fs.sums + i, t);
  }
  RH_BT(2);
  grid_barrier(fs, gridDim.x);
  RH_BT(3);

  // ---- phase 2: statistics -> per-column mean / scale, apply from registers ----
  // ONE thread per column turns the grid's sums into (mean, gamma * rstd) and shares them through shared memory.  (Every thread
  // deriving them for its own 4 * KMAX columns — 5 IEEE divisions / square roots per column, 16x redundant across the CTA's warps,
  // all of them hitting the same few L2 lines right behind the atomics — was ~6 us of an 11 us kernel, tools/bnfuse_trace.cu.)
  const float n = (float)p.rows;
  float* s_mu = smem;
  float* s_sc = smem + cols;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) {
    const float t1 = __ldcg(fs.sums + c), t2 = __ldcg(fs.sums + cols + c);
    const float mean = __ldg(p.h + c) + t1 / n;
    float v = (t2 - t1 * t1 / n) / n;
    if (v < 0.f) v = 0.f;
    s_mu[c] = mean;
    s_sc[c] = (p.gamma != nullptr ? __ldg(p.gamma + c) : 1.f) * (1.f / sqrtf(v + p.bn_eps));
    if (blockIdx.x == 0) {  // CTA 0 publishes the statistics (concurrently with everybody's apply phase)
      p.stats[c] = mean;
      p.stats[cols + c] = v;
      if (p.running_mean != nullptr) p.running_mean[c] = (1.f - p.momentum) * p.running_mean[c] + p.momentum * mean;
      if (p.running_var != nullptr) {
        const float unbiased = p.rows > 1 ? v * (n / (n - 1.f)) : v;
        p.running_var[c] = (1.f - p.momentum) * p.running_var[c] + p.momentum * unbiased;
      }
    }
  }
  __syncthreads();
  float mu[KMAX][4], sc[KMAX][4], bt[KMAX][4], hw[KMAX][4];
  load_colconst<KMAX>(p.beta, lane, cols, bt, 0.f);
  if (HEAD) load_colconst<KMAX>(p.head_w, lane, cols, hw, 0.f);
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    const int c0 = (k * 32 + lane) * 4;
    float4 m4 = f4_zero(), s4 = f4_zero();
    if (c0 < cols) {
      m4 = *reinterpret_cast<const float4*>(s_mu + c0);
      s4 = *reinterpret_cast<const float4*>(s_sc + c0);
    }
    mu[k][0] = m4.x; mu[k][1] = m4.y; mu[k][2] = m4.z; mu[k][3] = m4.w;
    sc[k][0] = s4.x; sc[k][1] = s4.y; sc[k][2] = s4.z; sc[k][3] = s4.w;
  }
  const float alpha = p.alpha != nullptr ? __ldg(p.alpha) : 0.f;
  const bool drop = p.p_drop > 0.f;
  const float keep_scale = drop ? 1.f / (1.f - p.p_drop) : 1.f;
  const float hb = (HEAD && p.head_b != nullptr) ? __ldg(p.head_b) : 0.f;
#pragma unroll
  for (int i = 0; i < RPW; ++i) {
    const int64_t row = row0 + warp + kFuseWarps * i;
    if (row >= p.rows) continue;  // warp-uniform
    float z[KMAX][4];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      const bool on = (k * 32 + lane) * 4 < cols;
#pragma unroll
      for (int j = 0; j < 4; ++j) z[k][j] = on ? fmaf(hv[i][k][j] - mu[k][j], sc[k][j], bt[k][j]) : 0.f;
    }
    float m = 0.f, inv_s = 0.f;
    if (act == ACT_DICE) dice_row_stats<KMAX>(z, lane, cols, p.dice_eps, m, inv_s);
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      const int c0 = (k * 32 + lane) * 4;
      if (c0 >= cols) continue;
      float o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float zz = z[k][j];
        const float ps = act == ACT_DICE ? sigmoidf_precise((zz - m) * inv_s) : 0.f;
        float v = act_value<ACT>(p.act, zz, alpha, ps);
        if (drop) v = dropout_keep(p.seed, counter, (uint64_t)row * cols + c0 + j, p.p_drop) ? v * keep_scale : 0.f;
        o[j] = v;
        if (HEAD) acc = fmaf(v, hw[k][j], acc);
      }
      if (!HEAD || p.y != nullptr) *reinterpret_cast<float4*>(p.y + row * p.y_ld + c0) = make_float4(o[0], o[1], o[2], o[3]);
    }
    if (HEAD) {
      acc = warp_sum(acc);
      if (lane == 0) {
        float yv = acc + hb;
        if (p.e0 != nullptr) yv += __ldg(p.e0 + row);
        if (p.e1 != nullptr) yv += __ldg(p.e1 + row);
        p.head_out[row] = p.apply_sigmoid ? 1.f / (1.f + expf(-yv)) : yv;
      }
    }
  }
  RH_BT(4);

  // ---- CTA 0 prepares the scratch of the next launch ----
  if (blockIdx.x != 0) return;
  for (int c = threadIdx.x; c < fs.other_floats; c += blockDim.x) fs.other[c] = 0.f;
  if (threadIdx.x == 0) {
    if (p.nbt != nullptr) *p.nbt = count;
    p.stats[2 * cols] = __int_as_float((int)counter);
    reinterpret_cast<unsigned*>(p.scratch)[0] = fs.gen + 1u;
  }
  RH_BT(5);
}

// =====================================================================================================
// backward
// =====================================================================================================
template <int KMAX, int RPW, int ACT, bool HEAD>
__global__ void __launch_bounds__(kFuseThreads) bn_fused_bwd_kernel(const BnFuseP p) {
  extern __shared__ float smem[];  // [16 warps][NS cols], NS = 3 in head mode else 2
  __shared__ float sm_alpha[kFuseWarps], sm_hb[kFuseWarps];
  constexpr int NS = HEAD ? 3 : 2;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cols = p.cols;
  const int act = ACT >= 0 ? ACT : p.act;
  const int64_t row0 = (int64_t)blockIdx.x * kFuseWarps * RPW;
  pdl_wait();
  const FuseScratch fs = fuse_scratch(p.scratch, cols);
  const uint32_t counter = (uint32_t)__float_as_int(__ldg(p.stats + 2 * cols));

  float mu[KMAX][4], rstd[KMAX][4], gam[KMAX][4], bet[KMAX][4], hw[KMAX][4];
  load_colconst<KMAX>(p.stats, lane, cols, mu, 0.f);
  load_colconst<KMAX>(p.gamma, lane, cols, gam, 1.f);
  load_colconst<KMAX>(p.beta, lane, cols, bet, 0.f);
  if (HEAD) load_colconst<KMAX>(p.head_w, lane, cols, hw, 0.f);
  // rstd: one thread per column (an IEEE division + square root per column per THREAD was 16x redundant), shared through smem
  for (int c = threadIdx.x; c < cols; c += blockDim.x) smem[c] = 1.f / sqrtf(__ldg(p.stats + cols + c) + p.bn_eps);
  __syncthreads();
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    const int c0 = (k * 32 + lane) * 4;
    float4 r4 = make_float4(1.f, 1.f, 1.f, 1.f);
    if (c0 < cols) r4 = *reinterpret_cast<const float4*>(smem + c0);
    rstd[k][0] = r4.x; rstd[k][1] = r4.y; rstd[k][2] = r4.z; rstd[k][3] = r4.w;
  }
  __syncthreads();  // the per-warp slabs below reuse this shared memory
  const float alpha = p.alpha != nullptr ? __ldg(p.alpha) : 0.f;
  const bool drop = p.p_drop > 0.f;
  const float keep_scale = drop ? 1.f / (1.f - p.p_drop) : 1.f;
  const float inv_n = 1.f / (float)cols;

  // ---- phase 1: dz and xhat of this CTA's rows in registers; column sums of dz, dz*xhat (and g*y in head mode) ----
  float xh[RPW][KMAX][4], dz[RPW][KMAX][4];
  float acc_b[KMAX][4], acc_g[KMAX][4], acc_w[KMAX][4];
#pragma unroll
  for (int k = 0; k < KMAX; ++k)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc_b[k][j] = acc_g[k][j] = acc_w[k][j] = 0.f;
  float acc_alpha = 0.f, acc_hb = 0.f;
  // all loads first
#pragma unroll
  for (int i = 0; i < RPW; ++i) {
    const int64_t row = row0 + warp + kFuseWarps * i;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      const int c0 = (k * 32 + lane) * 4;
      xh[i][k][0] = xh[i][k][1] = xh[i][k][2] = xh[i][k][3] = 0.f;
      dz[i][k][0] = dz[i][k][1] = dz[i][k][2] = dz[i][k][3] = 0.f;
      if (row < p.rows && c0 < cols) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(p.h + row * p.h_ld + c0));
        xh[i][k][0] = v.x; xh[i][k][1] = v.y; xh[i][k][2] = v.z; xh[i][k][3] = v.w;
        if (!HEAD) {
          const float4 g = __ldg(reinterpret_cast<const float4*>(p.d_y + row * p.d_y_ld + c0));
          dz[i][k][0] = g.x; dz[i][k][1] = g.y; dz[i][k][2] = g.z; dz[i][k][3] = g.w;
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < RPW; ++i) {
    const int64_t row = row0 + warp + kFuseWarps * i;
    if (row >= p.rows) continue;  // warp-uniform
    float g_row = 0.f;
    if (HEAD) {
      g_row = __ldg(p.d_head_out + row);
      if (p.apply_sigmoid) {
        const float pr = __ldg(p.head_out + row);
        g_row = g_row * (1.f - pr) * pr;
      }
      if (lane == 0) {
        acc_hb += g_row;
        if (p.d_extra != nullptr) p.d_extra[row] = g_row;
      }
    }
    float z[KMAX][4];
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      const bool on = (k * 32 + lane) * 4 < cols;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float x = on ? (xh[i][k][j] - mu[k][j]) * rstd[k][j] : 0.f;
        xh[i][k][j] = x;
        z[k][j] = on ? fmaf(x, gam[k][j], bet[k][j]) : 0.f;
      }
    }
    float m = 0.f, inv_s = 0.f;
    if (act == ACT_DICE) dice_row_stats<KMAX>(z, lane, cols, p.dice_eps, m, inv_s);
    // upstream gradient of the activation output (through dropout)
    float da[KMAX][4], pj[KMAX][4];
    float a1 = 0.f, a2 = 0.f;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      const int c0 = (k * 32 + lane) * 4;
      const bool on = c0 < cols;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        da[k][j] = pj[k][j] = 0.f;
        if (!on) continue;
        float g = HEAD ? g_row * hw[k][j] : dz[i][k][j];
        const bool keep = !drop || dropout_keep(p.seed, counter, (uint64_t)row * cols + c0 + j, p.p_drop);
        g = keep ? g * keep_scale : 0.f;
        da[k][j] = g;
        const float zz = z[k][j];
        if (act == ACT_DICE) pj[k][j] = sigmoidf_precise((zz - m) * inv_s);
        if (HEAD) {  // d_w3 += g_row * y, y = dropout(act(z))
          const float yv = keep ? act_value<ACT>(p.act, zz, alpha, pj[k][j]) * keep_scale : 0.f;
          acc_w[k][j] = fmaf(g_row, yv, acc_w[k][j]);
        }
        if (act == ACT_DICE) {
          const float aj = g * zz * (1.f - alpha) * pj[k][j] * (1.f - pj[k][j]);
          a1 += aj;
          a2 = fmaf(aj, zz - m, a2);
          acc_alpha = fmaf(g * zz, 1.f - pj[k][j], acc_alpha);
        }
      }
    }
    float k1 = 0.f, k2 = 0.f;
    if (act == ACT_DICE) {
      a1 = warp_sum(a1);
      a2 = warp_sum(a2);
      k1 = a1 * inv_n * inv_s;
      k2 = a2 * inv_s * inv_s * inv_s;
    }
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      const bool on = (k * 32 + lane) * 4 < cols;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float zz = z[k][j], g = da[k][j];
        float d;
        switch (act) {
          case ACT_RELU: d = zz > 0.f ? g : 0.f; break;
          case ACT_DICE: {
            const float aj = g * zz * (1.f - alpha) * pj[k][j] * (1.f - pj[k][j]);
            d = g * (alpha + (1.f - alpha) * pj[k][j]) + aj * inv_s - k1 - (zz - m) * k2;
          } break;
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
        if (!on) d = 0.f;
        dz[i][k][j] = d;
        acc_b[k][j] += d;
        acc_g[k][j] = fmaf(d, xh[i][k][j], acc_g[k][j]);
      }
    }
  }
  {
    float* my = smem + (int64_t)warp * NS * cols;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      const int c0 = (k * 32 + lane) * 4;
      if (c0 < cols) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          my[c0 + j] = acc_b[k][j];
          my[cols + c0 + j] = acc_g[k][j];
          if (HEAD) my[2 * cols + c0 + j] = acc_w[k][j];
        }
      }
    }
    acc_alpha = warp_sum(acc_alpha);
    if (lane == 0) {
      sm_alpha[warp] = acc_alpha;
      sm_hb[warp] = acc_hb;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < NS * cols; i += blockDim.x) {
    float t = 0.f;
#pragma unroll
    for (int wv = 0; wv < kFuseWarps; ++wv) t += smem[(int64_t)wv * NS * cols + i];
    atomicAdd(fs.sums + i, t);
  }
  if (threadIdx.x == 0) {
    float ta = 0.f, tb = 0.f;
#pragma unroll
    for (int wv = 0; wv < kFuseWarps; ++wv) {
      ta += sm_alpha[wv];
      tb += sm_hb[wv];
    }
    if (act == ACT_DICE || act == ACT_PRELU) atomicAdd(fs.extra + 1, ta);
    if (HEAD) atomicAdd(fs.extra + 2, tb);
  }
  grid_barrier(fs, gridDim.x);

  // ---- phase 2: d_h = gamma * rstd * (dz - mean(dz) - xhat * mean(dz * xhat)) from registers ----
  const float inv_rows = 1.f / (float)p.rows;
  float mb[KMAX][4], mg[KMAX][4];
  for (int i = threadIdx.x; i < (NS * cols) / 4; i += blockDim.x) reinterpret_cast<float4*>(smem)[i] = __ldcg(reinterpret_cast<const float4*>(fs.sums) + i);
  __syncthreads();
#pragma unroll
  for (int k = 0; k < KMAX; ++k) {
    const int c0 = (k * 32 + lane) * 4;
    float4 t1 = f4_zero(), t2 = f4_zero();
    if (c0 < cols) {
      t1 = *reinterpret_cast<const float4*>(smem + c0);
      t2 = *reinterpret_cast<const float4*>(smem + cols + c0);
    }
    mb[k][0] = t1.x * inv_rows; mb[k][1] = t1.y * inv_rows; mb[k][2] = t1.z * inv_rows; mb[k][3] = t1.w * inv_rows;
    mg[k][0] = t2.x * inv_rows; mg[k][1] = t2.y * inv_rows; mg[k][2] = t2.z * inv_rows; mg[k][3] = t2.w * inv_rows;
  }
#pragma unroll
  for (int i = 0; i < RPW; ++i) {
    const int64_t row = row0 + warp + kFuseWarps * i;
    if (row >= p.rows) continue;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      const int c0 = (k * 32 + lane) * 4;
      if (c0 >= cols) continue;
      float o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) o[j] = gam[k][j] * rstd[k][j] * (dz[i][k][j] - mb[k][j] - xh[i][k][j] * mg[k][j]);
      *reinterpret_cast<float4*>(p.d_h + row * p.d_h_ld + c0) = make_float4(o[0], o[1], o[2], o[3]);
    }
  }

  // ---- CTA 0 writes the parameter gradients and prepares the scratch of the next launch ----
  if (blockIdx.x != 0) return;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) {
    if (p.d_beta != nullptr) p.d_beta[c] = smem[c];
    if (p.d_gamma != nullptr) p.d_gamma[c] = smem[cols + c];
    if (HEAD && p.d_head_w != nullptr) p.d_head_w[c] = smem[2 * cols + c];
    if (p.d_lin_bias != nullptr) p.d_lin_bias[c] = 0.f;  // a Linear bias in front of a batch-statistics BatchNorm: gradient exactly 0
  }
  for (int c = threadIdx.x; c < fs.other_floats; c += blockDim.x) fs.other[c] = 0.f;
  if (threadIdx.x == 0) {
    if (p.d_alpha != nullptr) *p.d_alpha = __ldcg(fs.extra + 1);
    if (HEAD && p.d_head_b != nullptr) *p.d_head_b = __ldcg(fs.extra + 2);
    reinterpret_cast<unsigned*>(p.scratch)[0] = fs.gen + 1u;
  }
}

// ---- shape -> (KMAX, RPW, grid) --------------------------------------------------------------------------------------
struct FusePlan {
  int kmax, rpw, grid;
};

// instantiated (KMAX, RPW): (1, 2) (1, 4) (2, 2) (4, 1) — RPW * KMAX * 8 floats of row data per thread in backward
static bool plan_for(int64_t rows, int cols, bool head, FusePlan* out) {
  if (rows <= 0 || cols <= 0 || cols % 4 != 0) return false;
  const int steps = (cols + 127) / 128;
  const int kmax = steps <= 1 ? 1 : (steps <= 2 ? 2 : (steps <= 4 ? 4 : 0));
  if (kmax == 0) return false;
  if (head && cols > 256) return false;  // shared-memory slab of the backward reduction: 16 x 3 x cols floats
  const int sms = num_sms();
  const int choices[3] = {kmax == 4 ? 1 : 2, kmax == 1 ? 4 : 0, 0};
  for (int t = 0; t < 3 && choices[t] > 0; ++t) {
    const int rpw = choices[t];
    const int64_t grid = (rows + kFuseWarps * rpw - 1) / (kFuseWarps * rpw);
    if (grid <= sms) {
      out->kmax = kmax;
      out->rpw = rpw;
      out->grid = (int)grid;
      return true;
    }
  }
  return false;
}

static bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace rh

using namespace rh;

extern "C" int64_t rh_bn_fused_scratch_floats(int cols) { return 32 + 2 * (int64_t)fuse_buffer_floats(cols); }

extern "C" int rh_bn_fused_supported(int64_t rows, int cols, int head) {
  FusePlan pl;
  return plan_for(rows, cols, head != 0, &pl) ? 1 : 0;
}

#ifdef RH_BN_TRACE
unsigned long long* g_bn_trace = nullptr;
#endif

#define RH_FUSE_SHAPES(KERNEL, ACTV, HEADV)                                                                  \
  do {                                                                                                       \
    const int key__ = pl.kmax * 16 + pl.rpw;                                                                 \
    switch (key__) {                                                                                         \
      case 1 * 16 + 2: launch_k(KERNEL<1, 2, ACTV, HEADV>, dim3(pl.grid), dim3(kFuseThreads), smem, st, p); break;             \
      case 1 * 16 + 4: launch_k(KERNEL<1, 4, ACTV, HEADV>, dim3(pl.grid), dim3(kFuseThreads), smem, st, p); break;             \
      case 2 * 16 + 2: launch_k(KERNEL<2, 2, ACTV, HEADV>, dim3(pl.grid), dim3(kFuseThreads), smem, st, p); break;             \
      case 4 * 16 + 1:                                                                                                         \
        if (smem > 48 * 1024) cudaFuncSetAttribute(KERNEL<4, 1, ACTV, HEADV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024); \
        launch_k(KERNEL<4, 1, ACTV, HEADV>, dim3(pl.grid), dim3(kFuseThreads), smem, st, p);                                   \
        break;                                                                                                                 \
      default: set_error("bn_fused: no instantiation for kmax %d rpw %d", pl.kmax, pl.rpw); return RH_ERR_UNSUPPORTED; \
    }                                                                                                        \
  } while (0)

#define RH_FUSE_DISPATCH(KERNEL, HEADV)                                      \
  do {                                                                       \
    if (smem > 64 * 1024 || (smem > 48 * 1024 && pl.kmax != 4)) {            \
      set_error("bn_fused: %zu bytes of shared memory", smem);               \
      return RH_ERR_UNSUPPORTED;                                             \
    }                                                                        \
    if (act == ACT_RELU) RH_FUSE_SHAPES(KERNEL, ACT_RELU, HEADV);            \
    else if (act == ACT_DICE) RH_FUSE_SHAPES(KERNEL, ACT_DICE, HEADV);       \
    else RH_FUSE_SHAPES(KERNEL, -1, HEADV);                                  \
  } while (0)

extern "C" int rh_bn_act_fused_fwd(const float* h, int64_t h_ld, int64_t rows, int cols, float bn_eps, const float* gamma, const float* beta, int act,
                                   const float* act_param, float dice_eps, float p_drop, uint32_t dropout_seed, float* running_mean,
                                   float* running_var, int64_t* num_batches_tracked, float momentum, float* stats, float* scratch, float* y,
                                   int64_t y_ld, const float* head_w, const float* head_b, const float* extra0, const float* extra1,
                                   int apply_sigmoid, float* head_out, void* stream) {
  RH_REQUIRE(h && stats && scratch, RH_ERR_INVALID_ARG, "rh_bn_act_fused_fwd: NULL pointer");
  RH_REQUIRE(act >= 0 && act <= 5, RH_ERR_INVALID_ARG, "rh_bn_act_fused_fwd: act %d unknown", act);
  RH_REQUIRE(!((act == ACT_DICE || act == ACT_PRELU) && act_param == nullptr), RH_ERR_INVALID_ARG, "rh_bn_act_fused_fwd: Dice/PReLU need act_param");
  RH_REQUIRE(p_drop >= 0.f && p_drop < 1.f, RH_ERR_INVALID_ARG, "rh_bn_act_fused_fwd: p_drop must be in [0,1)");
  const bool head = head_w != nullptr;
  RH_REQUIRE(head ? head_out != nullptr : y != nullptr, RH_ERR_INVALID_ARG, "rh_bn_act_fused_fwd: output pointer is NULL");
  FusePlan pl;
  RH_REQUIRE(plan_for(rows, cols, head, &pl), RH_ERR_UNSUPPORTED, "rh_bn_act_fused_fwd: shape (%lld, %d) outside the fused kernel (see rh_bn_fused_supported)",
             (long long)rows, cols);
  RH_REQUIRE(h_ld >= cols && h_ld % 4 == 0 && al16(h) && (!gamma || al16(gamma)) && (!beta || al16(beta)) && (!head_w || al16(head_w)) &&
                 (y == nullptr || (al16(y) && y_ld % 4 == 0 && y_ld >= cols)),
             RH_ERR_UNSUPPORTED, "rh_bn_act_fused_fwd: operands must be 16-byte aligned with row strides that are multiples of 4 floats");
  BnFuseP p;
  memset(&p, 0, sizeof(p));
  p.h = h; p.h_ld = h_ld; p.rows = rows; p.cols = cols; p.bn_eps = bn_eps; p.gamma = gamma; p.beta = beta; p.act = act; p.alpha = act_param;
  p.dice_eps = dice_eps; p.p_drop = p_drop; p.seed = dropout_seed; p.running_mean = running_mean; p.running_var = running_var;
  p.nbt = reinterpret_cast<long long*>(num_batches_tracked); p.momentum = momentum; p.stats = stats; p.scratch = scratch; p.y = y; p.y_ld = y_ld;
  p.head_w = head_w; p.head_b = head_b; p.e0 = extra0; p.e1 = extra1; p.apply_sigmoid = apply_sigmoid; p.head_out = head_out;
#ifdef RH_BN_TRACE
  p.trace = g_bn_trace;
#endif
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = (size_t)kFuseWarps * 2 * cols * sizeof(float);
  if (head) RH_FUSE_DISPATCH(bn_fused_fwd_kernel, true);
  else RH_FUSE_DISPATCH(bn_fused_fwd_kernel, false);
  RH_LAUNCH_CHECK();
  return RH_OK;
}

extern "C" int rh_bn_act_fused_bwd(const float* h, int64_t h_ld, int64_t rows, int cols, const float* stats, float bn_eps, const float* gamma,
                                   const float* beta, int act, const float* act_param, float dice_eps, float p_drop, uint32_t dropout_seed,
                                   const float* d_y, int64_t d_y_ld, const float* head_w, const float* head_out, const float* d_head_out,
                                   int apply_sigmoid, float* scratch, float* d_h, int64_t d_h_ld, float* d_gamma, float* d_beta,
                                   float* d_act_param, float* d_head_w, float* d_head_b, float* d_extra, float* d_lin_bias, void* stream) {
  RH_REQUIRE(h && stats && scratch && d_h, RH_ERR_INVALID_ARG, "rh_bn_act_fused_bwd: NULL pointer");
  RH_REQUIRE(act >= 0 && act <= 5, RH_ERR_INVALID_ARG, "rh_bn_act_fused_bwd: act %d unknown", act);
  RH_REQUIRE(!((act == ACT_DICE || act == ACT_PRELU) && act_param == nullptr), RH_ERR_INVALID_ARG, "rh_bn_act_fused_bwd: Dice/PReLU need act_param");
  const bool head = head_w != nullptr;
  RH_REQUIRE(head ? (d_head_out != nullptr && (!apply_sigmoid || head_out != nullptr)) : d_y != nullptr, RH_ERR_INVALID_ARG,
             "rh_bn_act_fused_bwd: upstream gradient is NULL");
  FusePlan pl;
  RH_REQUIRE(plan_for(rows, cols, head, &pl), RH_ERR_UNSUPPORTED, "rh_bn_act_fused_bwd: shape (%lld, %d) outside the fused kernel", (long long)rows, cols);
  RH_REQUIRE(h_ld >= cols && h_ld % 4 == 0 && d_h_ld >= cols && d_h_ld % 4 == 0 && al16(h) && al16(d_h) && al16(stats) && (cols % 4 == 0) &&
                 (!gamma || al16(gamma)) && (!beta || al16(beta)) && (!head_w || al16(head_w)) && (d_y == nullptr || (al16(d_y) && d_y_ld % 4 == 0 && d_y_ld >= cols)),
             RH_ERR_UNSUPPORTED, "rh_bn_act_fused_bwd: operands must be 16-byte aligned with row strides that are multiples of 4 floats");
  BnFuseP p;
  memset(&p, 0, sizeof(p));
  p.h = h; p.h_ld = h_ld; p.rows = rows; p.cols = cols; p.bn_eps = bn_eps; p.gamma = gamma; p.beta = beta; p.act = act; p.alpha = act_param;
  p.dice_eps = dice_eps; p.p_drop = p_drop; p.seed = dropout_seed; p.stats = const_cast<float*>(stats); p.scratch = scratch;
  p.head_w = head_w; p.head_out = const_cast<float*>(head_out); p.d_head_out = d_head_out; p.apply_sigmoid = apply_sigmoid;
  p.d_y = d_y; p.d_y_ld = d_y_ld; p.d_h = d_h; p.d_h_ld = d_h_ld; p.d_gamma = d_gamma; p.d_beta = d_beta; p.d_alpha = d_act_param;
  p.d_head_w = d_head_w; p.d_head_b = d_head_b; p.d_extra = d_extra; p.d_lin_bias = d_lin_bias;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t smem = (size_t)kFuseWarps * (head ? 3 : 2) * cols * sizeof(float);
  if (head) RH_FUSE_DISPATCH(bn_fused_bwd_kernel, true);
  else RH_FUSE_DISPATCH(bn_fused_bwd_kernel, false);
  RH_LAUNCH_CHECK();
  return RH_OK;
}
