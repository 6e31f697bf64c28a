// rh_dist.cu — the dense half of the multi-GPU step without a collective library: all-reduce of the replicated parameters'
// gradients over NVLink peer memory FUSED with the optimiser update.
//
// Reference semantics replaced: nn.DataParallel's gradient reduction to device 0 + optimizer.step() (trainers/ctr_trainer.py:53-55,
// 97-99).  The sharded engine (b200/dist.py) needs, once per step, the SUM over ranks of ~0.6 MB of tower gradients (+ the loss
// scalar).  NCCL's ring all-reduce spends 21.6 us on it at 2 GPUs — pure latency (profiles/r01c_warm_kernel_times_n2_direct.txt).
// Here:
//   rh_dense_pack_signal        every rank PUSHES its gradients (and extra device scalars) into slot [rank] of EVERY rank's
//                               peer-mapped staging buffer (posted NVLink stores, no round trip; double-buffered by step parity),
//                               then the last CTA to finish stores the step number into its flag slot on every peer
//                               (st.release.sys after ONE __threadfence_system).
//   rh_dense_reduce_update      waits until every peer's flag shows this step (ld.acquire.sys spins on LOCAL memory), then each
//                               element is the sum over the W LOCAL slots IN RANK ORDER — identical bits on every rank — and goes
//                               straight into the SGD / Adam / Adagrad update of the parameter (rh_dense_update's arithmetic).
//                               The summed extras are written out for the host.
// One-shot: (W - 1) x 0.6 MB outbound per GPU — 4 MB at 8 ranks — and no second exchange.  (The first version PULLED: every rank
// read the peers' buffers inside the update kernel — 14-16 us against 4.5 us for the single-GPU update: NVLink read round trips.)
// Step numbers only grow, so flags never need resetting; a slot of parity p is rewritten two steps later, after its writer has
// seen every peer's NEXT signal, which a peer issues after its reads of step p completed (stream order).
// Everything is static-shaped and reads its step number from device memory: CUDA-graph capturable.
#ifndef RH_PDL_FAMILY  // (the trace tools include several of these files into one unit: the first one names the family)
#define RH_PDL_FAMILY 64  /* rh_set_pdl mask bit of this file's kernels */
#endif
#include "rh_common.cuh"

namespace rh {

constexpr int kMaxPackTensors = 96;
constexpr int kMaxRanks = 8;

struct PackP {
  const float* g[kMaxPackTensors];   // gradient of tensor i or NULL (contributes zeros)
  int32_t n[kMaxPackTensors];
  int32_t off[kMaxPackTensors];      // float offset inside a staging slot
  const float* extra[4];
  int32_t n_tensors, n_extra, total;  // total = floats per slot (tensors + extras), padded to 4
  float* peer_stage[kMaxRanks];       // staging buffer of every rank (peer-mapped): 2 parities x world slots of `total` floats; I write slot [rank] of each
  int32_t* peer_flags[kMaxRanks];     // flags array (kMaxRanks ints) of every rank (peer-mapped); I write element [rank]
  int32_t rank, world;
  const int32_t* epoch;               // completed all-reduces so far (device)
  unsigned* ticket;                   // zero on entry, left zero
};


// One system-scope fence per LAUNCH, not per block: every block makes its copies visible device-wide (__threadfence) before it takes
// its ticket; the last block — which therefore observes all of them — issues the only __threadfence_system() and publishes the flag
// (fences are cumulative).  The first version ran ~2500 blocks with a system fence each: 25 us for a 600 KB copy.
__global__ void __launch_bounds__(256) dense_pack_signal_kernel(const __grid_constant__ PackP p) {
  __shared__ int is_last;
  pdl_wait();
  const int e = *p.epoch + 1;
  const size_t slot = ((size_t)(e & 1) * p.world + p.rank) * p.total;  // my slot in everybody's buffer
  for (int ti = blockIdx.y; ti < p.n_tensors; ti += gridDim.y) {
    const float* g = p.g[ti];
    const size_t dst = slot + p.off[ti];
    const int n = p.n[ti];
    const int t0 = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
    int done = 0;
    if (g != nullptr && (reinterpret_cast<uintptr_t>(g) & 15u) == 0) {  // slot offsets are multiples of 4 floats
      const int n4 = n >> 2;
      for (int i = t0; i < n4; i += nt) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(g) + i);
#pragma unroll
        for (int s = 0; s < kMaxRanks; ++s)
          if (s < p.world) reinterpret_cast<float4*>(p.peer_stage[s] + dst)[i] = v;  // posted NVLink stores: no round trip
      }
      done = n4 << 2;
    }
    for (int i = done + t0; i < n; i += nt) {
      const float v = g != nullptr ? g[i] : 0.f;
      for (int s = 0; s < p.world; ++s) p.peer_stage[s][dst + i] = v;
    }
  }
  if (blockIdx.x == 0 && blockIdx.y == 0 && (int)threadIdx.x < p.n_extra) {
    const float v = *p.extra[threadIdx.x];
    for (int s = 0; s < p.world; ++s) p.peer_stage[s][slot + p.total - 4 + threadIdx.x] = v;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(p.ticket, 1u) == gridDim.x * gridDim.y - 1) ? 1 : 0;
  __syncthreads();
  if (!is_last) return;
  if ((int)threadIdx.x < p.world) {
    __threadfence_system();
    st_release_sys(p.peer_flags[threadIdx.x] + p.rank, e);
  }
  if (threadIdx.x == 0) *p.ticket = 0u;
}

struct ReduceP {
  float* w[kMaxPackTensors];
  float* s1[kMaxPackTensors];
  float* s2[kMaxPackTensors];
  int32_t n[kMaxPackTensors];
  int32_t off[kMaxPackTensors];
  int32_t n_tensors, n_extra, total;
  const float* stage;                  // MY staging buffer: 2 parities x world slots of `total` floats, slot s written by rank s
  const int32_t* flags;                // MY flags array: flags[s] = last step rank s has published
  int32_t rank, world;
  int32_t* epoch;
  unsigned* ticket;
  float* extra_out;                    // (n_extra) summed extras
  int kind;
  float beta1, beta2, eps, wd;
  const float* lr_dev;
  const float* bc_dev;
};

__global__ void __launch_bounds__(256) dense_reduce_update_kernel(const __grid_constant__ ReduceP p) {
  __shared__ int is_last;
  pdl_wait();
  const int e = *p.epoch + 1;
  if ((int)threadIdx.x < p.world) {
    while (ld_acquire_sys(p.flags + threadIdx.x) < e) __nanosleep(64);
  }
  __syncthreads();
  const size_t par = (size_t)(e & 1) * p.world * p.total;  // this step's parity: world slots of `total` floats, slot s from rank s
  const float lr = *p.lr_dev;
  const float bc1 = p.kind == 1 ? p.bc_dev[0] : 1.f, bc2s = p.kind == 1 ? p.bc_dev[1] : 1.f;
  for (int ti = blockIdx.y; ti < p.n_tensors; ti += gridDim.y) {
    float* w = p.w[ti];
    float* s1 = p.s1[ti];
    float* s2 = p.s2[ti];
    const size_t base = par + p.off[ti];
    const int n = p.n[ti];
    const int t0 = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
    int done = 0;
    // four elements per thread, all ranks' 16-byte loads in flight together; the slots are LOCAL (the peers pushed them)
    if (p.kind == 1 && ((reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(s1) | reinterpret_cast<uintptr_t>(s2)) & 15u) == 0) {
      const int n4 = n >> 2;
      for (int i = t0; i < n4; i += nt) {
        float4 gs[kMaxRanks];
#pragma unroll
        for (int s = 0; s < kMaxRanks; ++s)
          if (s < p.world) gs[s] = __ldcg(reinterpret_cast<const float4*>(p.stage + base + (size_t)s * p.total) + i);
        float4 wv = reinterpret_cast<float4*>(w)[i], m4 = reinterpret_cast<float4*>(s1)[i], v4 = reinterpret_cast<float4*>(s2)[i];
        float g[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int s = 0; s < kMaxRanks; ++s)
          if (s < p.world) {  // rank order: identical sums everywhere
            g[0] += gs[s].x; g[1] += gs[s].y; g[2] += gs[s].z; g[3] += gs[s].w;
          }
        float wa[4] = {wv.x, wv.y, wv.z, wv.w}, ma[4] = {m4.x, m4.y, m4.z, m4.w}, va[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float gr = fmaf(p.wd, wa[j], g[j]);
          ma[j] = p.beta1 * ma[j] + (1.f - p.beta1) * gr;
          va[j] = p.beta2 * va[j] + (1.f - p.beta2) * gr * gr;
          wa[j] -= (lr / bc1) * (ma[j] / (sqrtf(va[j]) / bc2s + p.eps));
        }
        reinterpret_cast<float4*>(s1)[i] = make_float4(ma[0], ma[1], ma[2], ma[3]);
        reinterpret_cast<float4*>(s2)[i] = make_float4(va[0], va[1], va[2], va[3]);
        reinterpret_cast<float4*>(w)[i] = make_float4(wa[0], wa[1], wa[2], wa[3]);
      }
      done = n4 << 2;
    }
    for (int i = done + t0; i < n; i += nt) {
      float g = 0.f;
      for (int s = 0; s < p.world; ++s) g += __ldcg(p.stage + base + (size_t)s * p.total + i);  // rank order: identical sums everywhere
      float wv = w[i];
      float gr = fmaf(p.wd, wv, g);
      if (p.kind == 0) {
        wv -= lr * gr;
      } else if (p.kind == 1) {
        const float m = p.beta1 * s1[i] + (1.f - p.beta1) * gr;
        const float v = p.beta2 * s2[i] + (1.f - p.beta2) * gr * gr;
        s1[i] = m;
        s2[i] = v;
        wv -= (lr / bc1) * (m / (sqrtf(v) / bc2s + p.eps));
      } else {
        const float acc = s1[i] + gr * gr;
        s1[i] = acc;
        wv -= lr * gr / (sqrtf(acc) + p.eps);
      }
      w[i] = wv;
    }
  }
  if (blockIdx.x == 0 && blockIdx.y == 0 && (int)threadIdx.x < p.n_extra) {
    float t = 0.f;
    for (int s = 0; s < p.world; ++s) t += __ldcg(p.stage + par + (size_t)s * p.total + p.total - 4 + threadIdx.x);
    p.extra_out[threadIdx.x] = t;
  }
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(p.ticket, 1u) == gridDim.x * gridDim.y - 1) ? 1 : 0;
  __syncthreads();
  if (is_last && threadIdx.x == 0) {
    *p.epoch = e;  // every CTA of this launch has read the old value (it is past its loops)
    *p.ticket = 0u;
  }
}

// Cross-GPU barrier of the sharded exchange: one warp publishes this rank's next step number into its slot of every peer's flags
// (st.release.sys after __threadfence_system: everything this GPU wrote before — rows stored into the peers' tiles, gradient REDs —
// is visible to a peer that acquires the flag) and spins on its own flags until every rank has published.  ~2 us; replaces a
// library barrier kernel of 4.9 us at three places per step.
struct PeerFlagPtrs {
  int32_t* p[kMaxRanks];
};
__global__ void __launch_bounds__(32) peer_barrier_kernel_v(const __grid_constant__ PeerFlagPtrs peers, const int32_t* my_flags, int rank, int world, int32_t* epoch) {
  pdl_wait();
  const int e = *epoch + 1;
  __syncwarp();
  if ((int)threadIdx.x < world) {
    __threadfence_system();
    st_release_sys(peers.p[threadIdx.x] + rank, e);
    while (ld_acquire_sys(my_flags + threadIdx.x) < e) __nanosleep(40);
  }
  __syncwarp();
  if (threadIdx.x == 0) *epoch = e;
}

// Wait only: the flags were (or will be) published by the peers' rh_dense_pack_signal launches of this step — which sit behind
// their backward kernels in stream order and behind a system fence — so "every rank has published step e" also means "every rank's
// row-gradient REDs have landed".  Replaces the third barrier of the step (no signal round of its own).
__global__ void __launch_bounds__(32) peer_wait_kernel(const int32_t* my_flags, int world, const int32_t* epoch) {
  pdl_wait();
  const int e = *epoch + 1;
  if ((int)threadIdx.x < world) {
    while (ld_acquire_sys(my_flags + threadIdx.x) < e) __nanosleep(40);
  }
}

}  // namespace rh

using namespace rh;

extern "C" int rh_peer_wait(const int32_t* my_flags, int world, const int32_t* epoch_dev, void* stream) {
  RH_REQUIRE(my_flags && epoch_dev && world >= 1 && world <= kMaxRanks, RH_ERR_INVALID_ARG, "rh_peer_wait: bad arguments");
  launch_k(peer_wait_kernel, dim3(1), dim3(32), 0, (cudaStream_t)stream, my_flags, world, epoch_dev);
  RH_LAUNCH_CHECK();
  return RH_OK;
}

extern "C" int rh_peer_barrier(int32_t* const* peer_flags, const int32_t* my_flags, int rank, int world, int32_t* epoch_dev, void* stream) {
  RH_REQUIRE(peer_flags && my_flags && epoch_dev && world >= 1 && world <= kMaxRanks && rank >= 0 && rank < world, RH_ERR_INVALID_ARG, "rh_peer_barrier: bad arguments");
  PeerFlagPtrs pf;
  memset(&pf, 0, sizeof(pf));
  for (int s = 0; s < world; ++s) {
    RH_REQUIRE(peer_flags[s] != nullptr, RH_ERR_INVALID_ARG, "rh_peer_barrier: flags of rank %d NULL", s);
    pf.p[s] = peer_flags[s];
  }
  launch_k(peer_barrier_kernel_v, dim3(1), dim3(32), 0, (cudaStream_t)stream, pf, my_flags, rank, world, epoch_dev);
  RH_LAUNCH_CHECK();
  return RH_OK;
}

extern "C" int64_t rh_dense_stage_floats(int n_tensors, const int64_t* numel) {
  int64_t t = 0;
  for (int i = 0; i < n_tensors; ++i) t += (numel[i] + 3) / 4 * 4;
  return t + 4;  // + up to 4 extra scalars
}

static int layout(int n_tensors, const int64_t* numel, int32_t* n, int32_t* off, int32_t* total) {
  RH_REQUIRE(n_tensors > 0 && n_tensors <= kMaxPackTensors, RH_ERR_UNSUPPORTED, "rh_dense_*: %d tensors not in [1,%d]", n_tensors, kMaxPackTensors);
  int64_t t = 0;
  for (int i = 0; i < n_tensors; ++i) {
    RH_REQUIRE(numel[i] >= 0 && numel[i] < ((int64_t)1 << 30), RH_ERR_INVALID_ARG, "rh_dense_*: tensor %d too large", i);
    n[i] = (int32_t)numel[i];
    off[i] = (int32_t)t;
    t += (numel[i] + 3) / 4 * 4;
  }
  RH_REQUIRE(t + 4 < ((int64_t)1 << 31), RH_ERR_UNSUPPORTED, "rh_dense_*: staging slot too large");
  *total = (int32_t)(t + 4);
  return RH_OK;
}

extern "C" int rh_dense_pack_signal(int n_tensors, const float* const* grads, const int64_t* numel, const float* const* extra, int n_extra,
                                    float* const* peer_stage, int32_t* const* peer_flags, int rank, int world, const int32_t* epoch_dev, int32_t* ticket_dev,
                                    void* stream) {
  RH_REQUIRE(grads && numel && peer_stage && peer_flags && epoch_dev && ticket_dev, RH_ERR_INVALID_ARG, "rh_dense_pack_signal: NULL pointer");
  RH_REQUIRE(world >= 1 && world <= kMaxRanks && rank >= 0 && rank < world && n_extra >= 0 && n_extra <= 4, RH_ERR_INVALID_ARG, "rh_dense_pack_signal: bad rank/world/extras");
  static thread_local PackP p;
  memset(&p, 0, sizeof(p));
  int rc = layout(n_tensors, numel, p.n, p.off, &p.total);
  if (rc != RH_OK) return rc;
  int64_t biggest = 1;
  for (int i = 0; i < n_tensors; ++i) {
    p.g[i] = grads[i];
    if (numel[i] > biggest) biggest = numel[i];
  }
  for (int i = 0; i < n_extra; ++i) {
    RH_REQUIRE(extra != nullptr && extra[i] != nullptr, RH_ERR_INVALID_ARG, "rh_dense_pack_signal: extra %d NULL", i);
    p.extra[i] = extra[i];
  }
  for (int s = 0; s < world; ++s) {
    RH_REQUIRE(peer_flags[s] != nullptr && peer_stage[s] != nullptr, RH_ERR_INVALID_ARG, "rh_dense_pack_signal: flags / staging buffer of rank %d NULL", s);
    p.peer_flags[s] = peer_flags[s];
    p.peer_stage[s] = peer_stage[s];
  }
  p.n_tensors = n_tensors; p.n_extra = n_extra; p.rank = rank; p.world = world; p.epoch = epoch_dev;
  p.ticket = reinterpret_cast<unsigned*>(ticket_dev);
  const int gy = n_tensors < 32 ? n_tensors : 32;
  int gx = (int)((biggest / 4 + 255) / 256);  // one float4 per thread for the largest tensor ...
  const int cap = (2 * num_sms() + gy - 1) / gy;  // ... but no more blocks than ~2 per SM in total (each ends with a fence + a ticket)
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  launch_k(dense_pack_signal_kernel, dim3(gx, gy), dim3(256), 0, (cudaStream_t)stream, p);
  RH_LAUNCH_CHECK();
  return RH_OK;
}

extern "C" int rh_dense_reduce_update(int n_tensors, float* const* params, float* const* state1, float* const* state2, const int64_t* numel, int n_extra,
                                      float* extra_out, const float* stage, const int32_t* flags, int rank, int world, int32_t* epoch_dev,
                                      int32_t* ticket_dev, int kind, const float* lr_dev, const float* bias_corr_dev, float beta1, float beta2, float eps,
                                      float weight_decay, void* stream) {
  RH_REQUIRE(params && numel && stage && flags && epoch_dev && ticket_dev && lr_dev, RH_ERR_INVALID_ARG, "rh_dense_reduce_update: NULL pointer");
  RH_REQUIRE(world >= 1 && world <= kMaxRanks && rank >= 0 && rank < world && n_extra >= 0 && n_extra <= 4 && (n_extra == 0 || extra_out), RH_ERR_INVALID_ARG,
             "rh_dense_reduce_update: bad rank/world/extras");
  RH_REQUIRE(kind >= 0 && kind <= 2 && (kind != 1 || bias_corr_dev != nullptr) && (kind == 0 || state1 != nullptr) && (kind != 1 || state2 != nullptr), RH_ERR_INVALID_ARG,
             "rh_dense_reduce_update: optimiser kind / state");
  static thread_local ReduceP p;
  memset(&p, 0, sizeof(p));
  int rc = layout(n_tensors, numel, p.n, p.off, &p.total);
  if (rc != RH_OK) return rc;
  int64_t biggest = 1;
  for (int i = 0; i < n_tensors; ++i) {
    RH_REQUIRE(params[i] != nullptr, RH_ERR_INVALID_ARG, "rh_dense_reduce_update: param %d NULL", i);
    p.w[i] = params[i];
    p.s1[i] = state1 ? state1[i] : nullptr;
    p.s2[i] = state2 ? state2[i] : nullptr;
    RH_REQUIRE(kind == 0 || p.s1[i], RH_ERR_INVALID_ARG, "rh_dense_reduce_update: state1[%d] NULL", i);
    RH_REQUIRE(kind != 1 || p.s2[i], RH_ERR_INVALID_ARG, "rh_dense_reduce_update: state2[%d] NULL", i);
    if (numel[i] > biggest) biggest = numel[i];
  }
  p.stage = stage;
  p.n_tensors = n_tensors; p.n_extra = n_extra; p.flags = flags; p.rank = rank; p.world = world; p.epoch = epoch_dev;
  p.ticket = reinterpret_cast<unsigned*>(ticket_dev); p.extra_out = extra_out; p.kind = kind; p.beta1 = beta1; p.beta2 = beta2; p.eps = eps;
  p.wd = weight_decay; p.lr_dev = lr_dev; p.bc_dev = bias_corr_dev;
  const int gy = n_tensors < 32 ? n_tensors : 32;
  int gx = (int)((biggest / 4 + 255) / 256);  // four elements per thread (Adam); every block first polls the peers' flags
  if (gx > 512) gx = 512;
  if (gx < 1) gx = 1;
  launch_k(dense_reduce_update_kernel, dim3(gx, gy), dim3(256), 0, (cudaStream_t)stream, p);
  RH_LAUNCH_CHECK();
  return RH_OK;
}
