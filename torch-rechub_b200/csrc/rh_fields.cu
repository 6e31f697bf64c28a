// rh_fields.cu — fused multi-field embedding gather + FM + LR + flattened tile, and its backward
// (sparse-gradient scatter-add).  The hot kernel of the DeepFM / DCN / DCNv2 front end.
//
// Reference arithmetic replaced (torch-rechub v0.8.0): basic/layers.py:77-127 (EmbeddingLayer.forward),
// :313-319 (FM.forward), :183-189 (LR.forward); backward = aten::embedding_dense_backward per lookup.
//
// Bound: HBM latency/bandwidth (arithmetic intensity < 1 FLOP/B).  Design:
//   forward : a warp owns one sample (dim 16: 8 field slots x 4 quarter lanes); each lane issues its
//             row loads (up to CH in flight) before consuming any, so a 4096-sample batch has its whole
//             6.8 MB of rows outstanding at once (Little's law: ~6 MB @ ~0.8 us).  FM and LR are reduced
//             in registers + xor-shuffles; the tile row is written with 16-byte streaming stores,
//             512 contiguous bytes per warp and pass.
//   backward: one (field, sample-chunk) per block so the LR weight gradient reduces in registers;
//             per-row gradient goes out as ONE 16-byte vector RED per lane (REDG.E.ADD.F32x4).
#ifndef RH_PDL_FAMILY  // (the trace tools include several of these files into one unit: the first one names the family)
#define RH_PDL_FAMILY 4  /* rh_set_pdl mask bit of this file's kernels */
#endif
#include "rh_common.cuh"

namespace rh {

struct FieldDev {
  const float* table;
  float* grad;
  const void* ids;
  int32_t id_stride;
  int32_t vocab;
  int32_t pad_idx;
  int32_t tile_col;
  int32_t fm_slot;
  int32_t is_i32;
};

struct DenseDev {
  const void* src;
  int32_t stride;
  int32_t dtype;
  int32_t width;
  int32_t tile_col;
};

#ifdef RH_FIELDS_TRACE
unsigned long long* g_fields_trace = nullptr;
#endif

// Cross-GPU hand-overs folded into a forward launch (include/rechub_b200.h: rh_sync); all NULL / 0 on one GPU.
struct SyncDev {
  const int32_t* wait_flags;
  const int32_t* step;
  int32_t* sig_flags[8];
  unsigned* ticket;
  int32_t wait_mask, sig_world, sig_rank, pad_;
  int64_t snap_delta;
};

struct FwdParams {
  FieldDev f[RH_MAX_FIELDS];
  DenseDev d[RH_MAX_DENSE];
  int32_t n_fields, n_dense, batch, dim;
  int32_t tile_vec, dest_rows;  // tile_vec: 16-byte aligned tile rows -> vector stores; dest_rows > 0: rows go to dest[b / dest_rows]
  float* dest[8];               // peer-memory tiles (one per destination GPU) for the fused gather + all-to-all
  float* tile;
  int64_t tile_ld;
  const float* lrw;
  const float* lrb;
  float* yfm;
  float* ylr;
  float* fsum;
  int32_t* err;
  SyncDev sync;
#ifdef RH_FIELDS_TRACE
  unsigned long long* trace;  // tools/fields_trace.cu: [block][16]: clock64 stamps of thread 0 (0..6), globaltimer at entry (8) and exit (9)
#endif
};

#ifdef RH_FIELDS_TRACE
// stamp after `dep` is available (the unused asm input keeps the clock read behind the load it depends on)
#define RH_FT(ev, dep)                                                                                  \
  do {                                                                                                  \
    if (p.trace != nullptr && threadIdx.x == 0 && threadIdx.y == 0) {                                   \
      unsigned long long t__;                                                                           \
      asm volatile("mov.u64 %0, %%clock64;" : "=l"(t__) : "r"(__float_as_int((float)(dep))) : "memory"); \
      p.trace[(size_t)blockIdx.x * 16 + (ev)] = t__;                                                    \
      if ((ev) == 0 || (ev) == 6) {                                                                     \
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t__));                                          \
        p.trace[(size_t)blockIdx.x * 16 + ((ev) == 0 ? 8 : 9)] = t__;                                   \
      }                                                                                                 \
    }                                                                                                   \
  } while (0)
#else
#define RH_FT(ev, dep) \
  do {                 \
  } while (0)
#endif

struct BwdParams {
  FieldDev f[RH_MAX_FIELDS];
  int32_t n_fields, batch, dim, tile_vec;  // tile_vec / dtile_vec: rows are 16-byte aligned
  int32_t dtile_vec, pad1_;
  const float* tile;
  const float* dtile;
  int64_t tile_ld;
  int64_t dtile_ld;
  const float* dyfm;
  const float* dylr;
  const float* lrw;
  const float* fsum;
  float* dlrw;
  float* dlrb;
  int32_t* err;
};

__device__ __forceinline__ float load_dense_value(const void* src, int64_t idx, int dtype) {
  switch (dtype) {
    case 0: return __ldg(reinterpret_cast<const float*>(src) + idx);
    case 1: return (float)__ldg(reinterpret_cast<const double*>(src) + idx);
    case 2: return (float)__ldg(reinterpret_cast<const long long*>(src) + idx);
    default: return (float)__ldg(reinterpret_cast<const int32_t*>(src) + idx);
  }
}

__device__ __forceinline__ float4 ld_tile4(const float* p, bool vec) {
  if (vec) return ldg_row16(p);
  return make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), __ldg(p + 3));
}
__device__ __forceinline__ void st_tile4(float* p, const float4& v, bool vec) {
  if (vec) {
    stg_row16(p, v);
  } else {
    p[0] = v.x;
    p[1] = v.y;
    p[2] = v.z;
    p[3] = v.w;
  }
}


// ---- hand-overs inside the launch (multi-GPU exchange) ----
// consumer: every CTA waits until all awaited ranks have published this step, BEFORE its first read of peer-written memory
// (bounded: a peer that never publishes — a crashed rank — ends in the launch's error flag after ~2^26 polls, tens of seconds, instead of
// a kernel that spins until the box is reset; _lib.check_errors turns RH_ERRFLAG_SYNC_TIMEOUT into a RuntimeError)
__device__ __forceinline__ void sync_wait_head(const SyncDev& sy, int32_t* err) {
  if (sy.wait_flags == nullptr) return;  // grid-uniform
  const int t = threadIdx.y * blockDim.x + threadIdx.x;
  if (t < 8 && ((sy.wait_mask >> t) & 1)) {
    const int e = *sy.step;
    unsigned polls = 0;
    while (ld_acquire_sys(sy.wait_flags + t) < e) {
      __nanosleep(32);
      if (++polls > (1u << 26)) {
        if (err != nullptr) *err = RH_ERRFLAG_SYNC_TIMEOUT;
        break;
      }
    }
  }
  __syncthreads();
}
// producer: the grid's last CTA publishes the step to every destination (ONE system fence per launch; the CTAs' own fences +
// the ticket make their peer stores precede it — fences are cumulative)
__device__ __forceinline__ void sync_signal_tail(const SyncDev& sy, unsigned n_ctas) {
  __shared__ int is_last;
  __threadfence();
  __syncthreads();
  const int t = threadIdx.y * blockDim.x + threadIdx.x;
  if (t == 0) is_last = (atomicAdd(sy.ticket, 1u) == n_ctas - 1) ? 1 : 0;
  __syncthreads();
  if (!is_last) return;
  if (t < sy.sig_world) {
    __threadfence_system();
    st_release_sys(sy.sig_flags[t] + sy.sig_rank, *sy.step);
  }
  if (t == 0) *sy.ticket = 0u;
}

// ------------------------------------------------------------------------------------------------
// forward, 16-byte lanes.
//   lane  = (sample slot, 16-byte quarter): LPR lanes per sample (LPR = pow2 >= dim/4), 32/LPR samples per warp
//   warp  = (sample group sw of 8/NG, field group g of NG): it gathers fields g, g+NG, g+2NG, ... for its 32/LPR samples;
//           blocks are always 8 warps (with few fields — an owner's 3-4 of 26 at 8 ranks — one-warp blocks ran at 20.9 us for
//           131 k rows: 4096 blocks of 32 threads, half the warps an SM can hold)
// A 4096-sample, 26-field batch is 4096 warps of ~4 row loads each instead of 512 warps of 26: the first
// version was bound by the instruction latency of ONE warp per scheduler (ncu r01: 5 % warps active,
// 37 k cycles for 3.5 k instructions per warp), not by memory.  The per-field descriptor index is
// warp-uniform (constant-bank loads).  FM / LR partials meet in shared memory; warp 0 finishes the sample.
// Tried and measured slower (round 2, B200): warp = one sample with 8 field slots x 4 quarter lanes, shuffle reductions and a
// contiguous 512-byte tile piece per pass ("v5"): 10.6 us with the descriptors staged in shared memory, 12.8 us with per-lane
// indexed parameter reads (8 replays per constant load), against 9.5 us here; at B = 262144 it reached 1.89 TB/s against 3.16.
// ------------------------------------------------------------------------------------------------
template <int LPR, int NG>
__global__ void __launch_bounds__(256) fields_fwd_v4(const __grid_constant__ FwdParams p) {
  constexpr int SPB = 32 / LPR;  // samples per warp row
  constexpr int SW = 8 / NG;     // warp rows (sample groups) per block: always 8 warps per block
  constexpr int CH = 4;          // row loads in flight per lane per chunk
  __shared__ float4 sm_s[SW][NG][32];
  __shared__ float sm_ss[SW][NG][32];
  __shared__ float sm_lr[SW][NG][32];

  pdl_wait();
  sync_wait_head(p.sync, p.err);
  const int lane = threadIdx.x & 31;
  const int g = threadIdx.x >> 5;
  const int q = lane % LPR;
  const int sw = threadIdx.y;
  const int b = (blockIdx.x * SW + sw) * SPB + lane / LPR;
  const int dim = p.dim;
  const bool live = b < p.batch;
  const bool lane_on = live && (4 * q < dim);
  const bool want_fm = p.yfm != nullptr || p.ylr != nullptr || p.fsum != nullptr;

  float4 s = f4_zero();
  float ss = 0.f, lr = 0.f;
  RH_FT(0, 0);

  // numeric columns, part 1: this warp's first two columns are loaded NOW so that their latency hides behind the gather
  // (loaded after it they were a third dependent memory round trip, tools/fields_trace.cu)
  float dv[2] = {0.f, 0.f};
#pragma unroll
  for (int t = 0; t < 2; ++t) {
    const int j = g + NG * t;
    if (live && p.tile != nullptr && j < p.n_dense && q < p.d[j].width) dv[t] = load_dense_value(p.d[j].src, (int64_t)b * p.d[j].stride + q, p.d[j].dtype);
  }
  const bool want_lr = want_fm && p.lrw != nullptr;

  for (int f0 = g; f0 < p.n_fields; f0 += NG * CH) {
    int32_t rid[CH];
    float4 w[CH];
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      const int f = f0 + j * NG;  // warp-uniform
      rid[j] = -1;
      w[j] = f4_zero();
      if (f < p.n_fields && live) {
        const FieldDev& fd = p.f[f];
        const int64_t id = load_id(fd.ids, (int64_t)b * fd.id_stride, fd.is_i32 != 0);
        if (p.sync.snap_delta != 0 && q == 0) {  // owner-side gather: the ids it used stay behind in local memory (int64 id buffers)
          *reinterpret_cast<long long*>(reinterpret_cast<char*>(const_cast<void*>(fd.ids)) + (int64_t)b * fd.id_stride * 8 + p.sync.snap_delta) = id;
        }
        // The LR weights of the chunk's fields do not depend on the ids: they travel with them.  (Loaded in the consume loop below
        // they sat behind the volatile tile stores — four SERIAL L2 round trips, 2.7 us of a 7.9 us launch in tools/fields_trace.cu.)
        if (want_lr && fd.fm_slot >= 0 && lane_on) w[j] = __ldg(reinterpret_cast<const float4*>(p.lrw + (int64_t)fd.fm_slot * dim + 4 * q));
        if ((uint64_t)id < (uint64_t)fd.vocab) {
          rid[j] = (int32_t)id;
        } else if (q == 0 && p.err != nullptr) {
          *p.err = 1 + f;
        }
      }
    }
    RH_FT(1, rid[0] + rid[1] + rid[2] + rid[3]);
    float4 v[CH];
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      v[j] = f4_zero();
      if (rid[j] >= 0 && lane_on) v[j] = ldg_row16(p.f[f0 + j * NG].table + (int64_t)rid[j] * dim + 4 * q);
    }
    RH_FT(2, v[0].x + v[1].x + v[2].x + v[3].x);
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      const int f = f0 + j * NG;
      if (f < p.n_fields && lane_on) {
        const FieldDev& fd = p.f[f];
        if (fd.tile_col >= 0 && p.tile != nullptr) {
          float* trow = p.dest_rows > 0 ? p.dest[b / p.dest_rows] + (int64_t)(b % p.dest_rows) * p.tile_ld : p.tile + (int64_t)b * p.tile_ld;
          st_tile4(trow + fd.tile_col + 4 * q, v[j], p.tile_vec != 0);
        }
        if (fd.fm_slot >= 0 && want_fm) {
          s = f4_add(s, v[j]);
          ss += f4_dot(v[j], v[j]);
          lr += f4_dot(v[j], w[j]);
        }
      }
    }
  }

  RH_FT(3, ss);
  // numeric columns, part 2: store the preloaded values; anything beyond them (more than 2 NG columns, widths above LPR) the slow way
  if (live && p.tile != nullptr) {
    float* drow = p.tile + (int64_t)b * p.tile_ld;
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const int j = g + NG * t;
      if (j < p.n_dense && q < p.d[j].width) drow[p.d[j].tile_col + q] = dv[t];
    }
    for (int j = g; j < p.n_dense; j += NG) {
      const DenseDev& dd = p.d[j];
      for (int k = q + (j < g + 2 * NG ? LPR : 0); k < dd.width; k += LPR) drow[dd.tile_col + k] = load_dense_value(dd.src, (int64_t)b * dd.stride + k, dd.dtype);
    }
  }

  RH_FT(4, 0);
  if (p.sync.ticket != nullptr) sync_signal_tail(p.sync, gridDim.x);  // grid-uniform; every thread of the block is still here
  if (!want_fm) return;  // block-uniform
  if (NG > 1) {
    sm_s[sw][g][lane] = s;
    sm_ss[sw][g][lane] = ss;
    sm_lr[sw][g][lane] = lr;
    __syncthreads();
    RH_FT(5, 0);
    if (g != 0) return;
#pragma unroll
    for (int k = 1; k < NG; ++k) {
      s = f4_add(s, sm_s[sw][k][lane]);
      ss += sm_ss[sw][k][lane];
      lr += sm_lr[sw][k][lane];
    }
  }
  float t = (s.x * s.x + s.y * s.y) + (s.z * s.z + s.w * s.w) - ss;
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) {
    t += __shfl_xor_sync(0xffffffffu, t, o);
    lr += __shfl_xor_sync(0xffffffffu, lr, o);
  }
  if (live && q == 0) {
    if (p.yfm != nullptr) p.yfm[b] = 0.5f * t;
    if (p.ylr != nullptr) p.ylr[b] = lr + (p.lrb != nullptr ? __ldg(p.lrb) : 0.f);
  }
  if (p.fsum != nullptr && lane_on) {
    *reinterpret_cast<float4*>(p.fsum + (int64_t)b * dim + 4 * q) = s;
  }
  RH_FT(6, t);
}


// ------------------------------------------------------------------------------------------------
// forward, 16-byte lanes, v6 (opt-in, RECHUB_B200_FIELDS_FWD=6; measured SLOWER than v4, see use_fwd_v4): the v4 mapping with the
// descriptor reads taken off the dependent chain.
// What v4's trace + SASS showed (tools/fields_trace.cu, profiles/r02_fields_trace.txt): the ids arrived 1.95 us after block entry
// and the four tile stores + FM updates took 2.3 us AFTER their rows had landed.  Neither is memory time: every use of p.f[f]
// / p.d[j] is an INDEXED constant-bank read (LDC c[0x0][R+off]; the field index depends on the warp), the compiler re-reads the
// descriptor word at each use, and the uses are separated by branches — 3-4 serial LDCs per field (tile_col -> dest_rows ->
// dest[] -> fm_slot), the first touch of each 128-byte constant line per SM a miss to L2.  Here:
//   * the block copies the used descriptors (n_fields x 48 B + n_dense x 24 B) into shared memory ONCE, one 16-byte piece per
//     thread, all constant-bank misses in flight together; every later descriptor read is an LDS broadcast;
//   * everything addressed by the sample alone (tile row pointer — in the peer-memory variant the destination GPU's tile and
//     the integer division that picks it) is computed once per lane, not once per field;
//   * the peer-memory route is a template flag, so the single-GPU kernel carries no division at all.
// ------------------------------------------------------------------------------------------------
template <int LPR, int NG, bool P2P>
__global__ void __launch_bounds__(256, 4) fields_fwd_v6(const __grid_constant__ FwdParams p) {
  constexpr int SPB = 32 / LPR;  // samples per warp row
  constexpr int SW = 8 / NG;     // warp rows (sample groups) per block: always 8 warps per block
  constexpr int CH = 4;          // row loads in flight per lane per chunk
  __shared__ __align__(16) FieldDev s_f[RH_MAX_FIELDS];
  __shared__ __align__(16) DenseDev s_d[RH_MAX_DENSE];
  __shared__ float4 sm_s[SW][NG][32];
  __shared__ float sm_ss[SW][NG][32];
  __shared__ float sm_lr[SW][NG][32];
  static_assert(sizeof(FieldDev) % 16 == 0 && (sizeof(DenseDev) * RH_MAX_DENSE) % 16 == 0, "descriptor arrays are copied in 16-byte pieces");

  pdl_wait();
  RH_FT(0, 0);
  const int lane = threadIdx.x & 31;
  const int g = threadIdx.x >> 5;
  const int q = lane % LPR;
  const int sw = threadIdx.y;
  const int n_fields = p.n_fields, n_dense = p.n_dense;
  {
    const int tid = sw * (NG * 32) + (int)threadIdx.x;
    const uint4* src = reinterpret_cast<const uint4*>(p.f);
    uint4* dst = reinterpret_cast<uint4*>(s_f);
    const int nf16 = n_fields * (int)(sizeof(FieldDev) / 16);
    for (int i = tid; i < nf16; i += 256) dst[i] = src[i];
    const uint4* srcd = reinterpret_cast<const uint4*>(p.d);
    uint4* dstd = reinterpret_cast<uint4*>(s_d);
    const int nd16 = (n_dense * (int)sizeof(DenseDev) + 15) / 16;  // rounds up inside the fixed-size array
    for (int i = tid; i < nd16; i += 256) dstd[i] = srcd[i];
  }
  const int b = (blockIdx.x * SW + sw) * SPB + lane / LPR;
  const int dim = p.dim;
  const bool live = b < p.batch;
  const bool lane_on = live && (4 * q < dim);
  const bool want_fm = p.yfm != nullptr || p.ylr != nullptr || p.fsum != nullptr;
  const bool want_lr = want_fm && p.lrw != nullptr;
  const bool tile_vec = p.tile_vec != 0;
  const float* const lrw = p.lrw;
  // the sample's tile row (peer-memory variant: inside the tile of the GPU that holds the sample)
  float* trow = nullptr;
  if (live && p.tile != nullptr) {
    if (P2P) {
      const int dr = p.dest_rows;
      const int which = b / dr;
      trow = p.dest[which] + (int64_t)(b - which * dr) * p.tile_ld;
    } else {
      trow = p.tile + (int64_t)b * p.tile_ld;
    }
  }
  __syncthreads();

  float4 s = f4_zero();
  float ss = 0.f, lr = 0.f;
  // numeric columns, part 1: this warp's first two columns are loaded NOW so that their latency hides behind the gather
  float dv[2] = {0.f, 0.f};
  if (!P2P) {
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const int j = g + NG * t;
      if (trow != nullptr && j < n_dense) {
        const DenseDev dd = s_d[j];
        if (q < dd.width) dv[t] = load_dense_value(dd.src, (int64_t)b * dd.stride + q, dd.dtype);
      }
    }
  }

  for (int f0 = g; f0 < n_fields; f0 += NG * CH) {
    int32_t rid[CH];
    float4 w[CH];
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      const int f = f0 + j * NG;  // warp-uniform
      rid[j] = -1;
      w[j] = f4_zero();
      if (f < n_fields) {
        const FieldDev& fd = s_f[f];  // LDS broadcasts (descriptor words are re-read from shared memory where they are used:
        if (live) {                    //  ~25 cycles each and independent across j — holding them cost 18 registers = one block/SM)
          const int64_t id = load_id(fd.ids, (int64_t)b * fd.id_stride, fd.is_i32 != 0);
          // the LR weights of the chunk's fields do not depend on the ids: they travel with them
          if (want_lr && fd.fm_slot >= 0 && lane_on) w[j] = __ldg(reinterpret_cast<const float4*>(lrw + (int64_t)fd.fm_slot * dim + 4 * q));
          if ((uint64_t)id < (uint64_t)fd.vocab) {
            rid[j] = (int32_t)id;
          } else if (q == 0 && p.err != nullptr) {
            *p.err = 1 + f;
          }
        }
      }
    }
    RH_FT(1, rid[0] + rid[1] + rid[2] + rid[3]);
    float4 v[CH];
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      v[j] = f4_zero();
      if (rid[j] >= 0 && lane_on) v[j] = ldg_row16(s_f[f0 + j * NG].table + (int64_t)rid[j] * dim + 4 * q);
    }
    RH_FT(2, v[0].x + v[1].x + v[2].x + v[3].x);
#pragma unroll
    for (int j = 0; j < CH; ++j) {
      if (f0 + j * NG < n_fields && lane_on) {
        const int tcol = s_f[f0 + j * NG].tile_col, slot = s_f[f0 + j * NG].fm_slot;
        if (tcol >= 0 && trow != nullptr) st_tile4(trow + tcol + 4 * q, v[j], tile_vec);
        if (slot >= 0 && want_fm) {
          s = f4_add(s, v[j]);
          ss += f4_dot(v[j], v[j]);
          lr += f4_dot(v[j], w[j]);
        }
      }
    }
  }

  RH_FT(3, ss);
  // numeric columns, part 2: store the preloaded values; anything beyond them (more than 2 NG columns, widths above LPR) the slow way
  if (!P2P && trow != nullptr) {
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const int j = g + NG * t;
      if (j < n_dense) {
        const DenseDev dd = s_d[j];
        if (q < dd.width) trow[dd.tile_col + q] = dv[t];
      }
    }
    for (int j = g; j < n_dense; j += NG) {
      const DenseDev dd = s_d[j];
      for (int k = q + (j < g + 2 * NG ? LPR : 0); k < dd.width; k += LPR) trow[dd.tile_col + k] = load_dense_value(dd.src, (int64_t)b * dd.stride + k, dd.dtype);
    }
  }

  RH_FT(4, 0);
  if (!want_fm) return;  // block-uniform
  if (NG > 1) {
    sm_s[sw][g][lane] = s;
    sm_ss[sw][g][lane] = ss;
    sm_lr[sw][g][lane] = lr;
    __syncthreads();
    RH_FT(5, 0);
    if (g != 0) return;
#pragma unroll
    for (int k = 1; k < NG; ++k) {
      s = f4_add(s, sm_s[sw][k][lane]);
      ss += sm_ss[sw][k][lane];
      lr += sm_lr[sw][k][lane];
    }
  }
  float t = (s.x * s.x + s.y * s.y) + (s.z * s.z + s.w * s.w) - ss;
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) {
    t += __shfl_xor_sync(0xffffffffu, t, o);
    lr += __shfl_xor_sync(0xffffffffu, lr, o);
  }
  if (live && q == 0) {
    if (p.yfm != nullptr) p.yfm[b] = 0.5f * t;
    if (p.ylr != nullptr) p.ylr[b] = lr + (p.lrb != nullptr ? __ldg(p.lrb) : 0.f);
  }
  if (p.fsum != nullptr && lane_on) {
    *reinterpret_cast<float4*>(p.fsum + (int64_t)b * dim + 4 * q) = s;
  }
  RH_FT(6, t);
}


// forward, scalar lanes: any dim / any alignment, tile emission only (no FM/LR).
__global__ void __launch_bounds__(256) fields_fwd_scalar(const __grid_constant__ FwdParams p) {
  const int dim = p.dim;
  const int64_t per_sample = (int64_t)p.n_fields * dim;
  const int64_t total = (int64_t)p.batch * per_sample;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / per_sample);
    const int r = (int)(i - (int64_t)b * per_sample);
    const int f = r / dim, d = r - f * dim;
    const FieldDev& fd = p.f[f];
    const int64_t id = load_id(fd.ids, (int64_t)b * fd.id_stride, fd.is_i32 != 0);
    float v = 0.f;
    if ((uint64_t)id < (uint64_t)fd.vocab) {
      v = __ldg(fd.table + id * dim + d);
    } else if (d == 0 && p.err != nullptr) {
      *p.err = 1 + f;
    }
    if (fd.tile_col >= 0) p.tile[(int64_t)b * p.tile_ld + fd.tile_col + d] = v;
  }
  // dense columns
  int dense_w = 0;
  for (int j = 0; j < p.n_dense; ++j) dense_w += p.d[j].width;
  const int64_t dtotal = (int64_t)p.batch * dense_w;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < dtotal; i += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / dense_w);
    int r = (int)(i - (int64_t)b * dense_w);
    int j = 0;
    while (r >= p.d[j].width) {
      r -= p.d[j].width;
      ++j;
    }
    const DenseDev& dd = p.d[j];
    p.tile[(int64_t)b * p.tile_ld + dd.tile_col + r] = load_dense_value(dd.src, (int64_t)b * dd.stride + r, dd.dtype);
  }
}

// ------------------------------------------------------------------------------------------------
// backward, 16-byte lanes.  grid = (sample chunks, fields); ITER samples per lane.
// ------------------------------------------------------------------------------------------------
template <int LPR, int ITER>
__global__ void __launch_bounds__(128) fields_bwd_v4(const __grid_constant__ BwdParams p) {
  __shared__ float4 sm_dw[4][32];
  __shared__ float sm_db[4];
  pdl_wait();

  const int f = blockIdx.y;
  const FieldDev& fd = p.f[f];
  const int spi = blockDim.x / LPR;  // samples per iteration
  const int q = (int)threadIdx.x % LPR;
  const int sl = (int)threadIdx.x / LPR;
  const int dim = p.dim;
  const bool lane_on = 4 * q < dim;
  const int base = blockIdx.x * spi * ITER;
  const bool fm = fd.fm_slot >= 0 && (p.dyfm != nullptr || p.dylr != nullptr);
  const bool from_tile = p.tile != nullptr && fd.tile_col >= 0;
  const bool has_dtile = p.dtile != nullptr && fd.tile_col >= 0;

  float4 w = f4_zero();
  if (fm && p.lrw != nullptr && p.dylr != nullptr && lane_on) {
    w = __ldg(reinterpret_cast<const float4*>(p.lrw + (int64_t)fd.fm_slot * dim + 4 * q));
  }

  // Everything a sample needs except the table row itself is addressed by the sample, not by its id: those loads are issued
  // together with the id loads (behind the id check they were a second dependent round trip; an out-of-range id only costs the
  // loads of that sample, and nothing is written for it).
  int64_t idv[ITER];
#pragma unroll
  for (int i = 0; i < ITER; ++i) {
    const int b = base + i * spi + sl;
    idv[i] = -1;
    if (b < p.batch) idv[i] = load_id(fd.ids, (int64_t)b * fd.id_stride, fd.is_i32 != 0);
  }
  float4 g[ITER], e[ITER], S[ITER];
  float dyf[ITER], dyl[ITER];
#pragma unroll
  for (int i = 0; i < ITER; ++i) {
    const int b = base + i * spi + sl;
    g[i] = f4_zero();
    e[i] = f4_zero();
    S[i] = f4_zero();
    dyf[i] = 0.f;
    dyl[i] = 0.f;
    if (b < p.batch && lane_on) {
      if (has_dtile) g[i] = ld_tile4(p.dtile + (int64_t)b * p.dtile_ld + fd.tile_col + 4 * q, p.dtile_vec != 0);
      if (fm) {
        if (from_tile) e[i] = ld_tile4(p.tile + (int64_t)b * p.tile_ld + fd.tile_col + 4 * q, p.tile_vec != 0);
        if (p.dyfm != nullptr) {
          S[i] = ldg_row16(p.fsum + (int64_t)b * dim + 4 * q);
          dyf[i] = __ldg(p.dyfm + b);
        }
        if (p.dylr != nullptr) dyl[i] = __ldg(p.dylr + b);
      }
    }
  }
  int32_t rid[ITER];
#pragma unroll
  for (int i = 0; i < ITER; ++i) {
    const int b = base + i * spi + sl;
    rid[i] = -1;
    if (b < p.batch) {
      if ((uint64_t)idv[i] < (uint64_t)fd.vocab) {
        rid[i] = (int32_t)idv[i];
      } else if (q == 0 && p.err != nullptr) {
        *p.err = 1 + f;
      }
    }
    if (fm && !from_tile && rid[i] >= 0 && lane_on) e[i] = ldg_row16(fd.table + (int64_t)rid[i] * dim + 4 * q);  // no saved tile: re-gather
  }

  float4 dw = f4_zero();
  float db = 0.f;
#pragma unroll
  for (int i = 0; i < ITER; ++i) {
    if (rid[i] >= 0 && lane_on) {
      if (fm) {
        // d/d e of 0.5*sum_d[(sum_f e)^2 - sum_f e^2] = S - e ;  d/d e of <e, w> = w
        g[i].x += dyf[i] * (S[i].x - e[i].x) + dyl[i] * w.x;
        g[i].y += dyf[i] * (S[i].y - e[i].y) + dyl[i] * w.y;
        g[i].z += dyf[i] * (S[i].z - e[i].z) + dyl[i] * w.z;
        g[i].w += dyf[i] * (S[i].w - e[i].w) + dyl[i] * w.w;
        dw = f4_fma(e[i], make_float4(dyl[i], dyl[i], dyl[i], dyl[i]), dw);
        if (q == 0) db += dyl[i];
      }
      if (fd.grad != nullptr && rid[i] != fd.pad_idx) {
        red_add_row16(fd.grad + (int64_t)rid[i] * dim + 4 * q, g[i]);
      }
    }
  }

  // LR weight / bias gradient: reduce over the samples of this block, one RED set per block.
  const bool want_dw = fm && p.dlrw != nullptr && p.dylr != nullptr;
  const bool want_db = fd.fm_slot == 0 && p.dlrb != nullptr && p.dylr != nullptr && fm;
  if (want_dw || want_db) {  // block-uniform
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
        red_add_row16(p.dlrw + (int64_t)fd.fm_slot * dim + 4 * lane, t);
      }
      if (want_db && lane == 0) {
        float t = sm_db[0];
        for (int k = 1; k < nwarp; ++k) t += sm_db[k];
        atomicAdd(p.dlrb, t);
      }
    }
  }
}

// backward, scalar lanes: d_tile -> table gradients only.
__global__ void __launch_bounds__(256) fields_bwd_scalar(const __grid_constant__ BwdParams p) {
  const int dim = p.dim;
  const int64_t per_sample = (int64_t)p.n_fields * dim;
  const int64_t total = (int64_t)p.batch * per_sample;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / per_sample);
    const int r = (int)(i - (int64_t)b * per_sample);
    const int f = r / dim, d = r - f * dim;
    const FieldDev& fd = p.f[f];
    if (fd.tile_col < 0 || fd.grad == nullptr) continue;
    const int64_t id = load_id(fd.ids, (int64_t)b * fd.id_stride, fd.is_i32 != 0);
    if ((uint64_t)id >= (uint64_t)fd.vocab) {
      if (d == 0 && p.err != nullptr) *p.err = 1 + f;
      continue;
    }
    if (id == fd.pad_idx) continue;
    atomicAdd(fd.grad + id * dim + d, __ldg(p.dtile + (int64_t)b * p.dtile_ld + fd.tile_col + d));
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

static int pack_fields(const rh_field* fields, int n_fields, FieldDev* out, bool need_grad) {
  for (int i = 0; i < n_fields; ++i) {
    const rh_field& s = fields[i];
    RH_REQUIRE(s.table != nullptr && s.ids != nullptr, RH_ERR_INVALID_ARG, "field %d: table/ids is NULL", i);
    RH_REQUIRE(s.vocab > 0, RH_ERR_INVALID_ARG, "field %d: vocab must be > 0", i);
    RH_REQUIRE(s.id_stride >= 0 && s.id_stride < (int64_t)1 << 31, RH_ERR_INVALID_ARG, "field %d: id_stride %lld out of range", i,
               (long long)s.id_stride);
    FieldDev& d = out[i];
    d.table = s.table;
    d.grad = need_grad ? s.table_grad : nullptr;
    d.ids = s.ids;
    d.id_stride = (int32_t)s.id_stride;
    d.vocab = s.vocab;
    d.pad_idx = s.padding_idx;
    d.tile_col = s.tile_col;
    d.fm_slot = s.fm_slot;
    d.is_i32 = s.ids_are_i32;
  }
  return RH_OK;
}

// RECHUB_B200_FIELDS_FWD=6 selects the variant with the descriptors staged in shared memory (fields_fwd_v6).  Measured on B200
// (tools/fields_trace.cu, profiles/r02c_fields_trace_v{4,6}.txt): the staging itself — indexed constant-bank reads with a different
// address per lane, then a block barrier — delays the id loads from +1.93 us to +3.13 us after block entry, more than the
// LDS-fed phases behind them win back (rows +1.02 vs +0.79 us, tile + FM +1.68 vs +2.27 us): 8.37 vs 7.73 us per launch,
// 10.9 vs 10.3 us in bench.py's cold-row replay.  v4 stays the default.
static bool use_fwd_v4() {
  static const int v = [] {
    const char* e = getenv("RECHUB_B200_FIELDS_FWD");
    return (e != nullptr && e[0] == '6') ? 0 : 1;
  }();
  return v != 0;
}

template <int LPR>
static void launch_fwd_v4(const FwdParams& p, cudaStream_t st) {
  constexpr int SPB = 32 / LPR;
  // enough field groups that every warp has at most ~4 row loads; small field counts put more sample groups into the block instead
  const int work = p.n_fields > p.n_dense ? p.n_fields : (p.n_dense + 3) / 4;
  const int ng = work <= 4 ? 1 : (work <= 8 ? 2 : (work <= 16 ? 4 : 8));
  const int sw = 8 / ng;
  const int grid = (p.batch + SPB * sw - 1) / (SPB * sw);
  const dim3 block(ng * 32, sw);
  const bool has_sync = p.sync.wait_flags != nullptr || p.sync.ticket != nullptr || p.sync.snap_delta != 0;  // implemented by v4 only
  if (use_fwd_v4() || has_sync) {
    switch (ng) {
      case 1: launch_k(fields_fwd_v4<LPR, 1>, dim3(grid), block, 0, st, p); break;
      case 2: launch_k(fields_fwd_v4<LPR, 2>, dim3(grid), block, 0, st, p); break;
      case 4: launch_k(fields_fwd_v4<LPR, 4>, dim3(grid), block, 0, st, p); break;
      default: launch_k(fields_fwd_v4<LPR, 8>, dim3(grid), block, 0, st, p); break;
    }
    return;
  }
  const bool p2p = p.dest_rows > 0;
#define RH_FWD6(NGV)                                                                        \
  do {                                                                                      \
    if (p2p) launch_k(fields_fwd_v6<LPR, NGV, true>, dim3(grid), block, 0, st, p);          \
    else launch_k(fields_fwd_v6<LPR, NGV, false>, dim3(grid), block, 0, st, p);             \
  } while (0)
  switch (ng) {
    case 1: RH_FWD6(1); break;
    case 2: RH_FWD6(2); break;
    case 4: RH_FWD6(4); break;
    default: RH_FWD6(8); break;
  }
#undef RH_FWD6
}

template <int LPR>
static void launch_bwd_v4(const BwdParams& p, cudaStream_t st) {
  constexpr int ITER = 4;
  const int threads = 128;
  const int spb = threads / LPR * ITER;
  dim3 grid((p.batch + spb - 1) / spb, p.n_fields);
  launch_k(fields_bwd_v4<LPR, ITER>, grid, dim3(threads), 0, st, p);
}

}  // namespace rh

using namespace rh;

static int fields_fwd_impl(const rh_field* fields, int n_fields, int dim, const rh_dense* dense, int n_dense, int batch, float* tile,
                           int64_t tile_ld, const float* lr_weight, const float* lr_bias, float* y_fm, float* y_lr, float* field_sum,
                           int32_t* err_flag, void* stream, float* const* dest, int n_dest, int dest_rows, const rh_sync* sync = nullptr) {
  RH_REQUIRE(n_fields >= 0 && n_fields <= RH_MAX_FIELDS, RH_ERR_INVALID_ARG, "n_fields %d not in [0,%d]", n_fields, RH_MAX_FIELDS);
  RH_REQUIRE(n_dense >= 0 && n_dense <= RH_MAX_DENSE, RH_ERR_INVALID_ARG, "n_dense %d not in [0,%d]", n_dense, RH_MAX_DENSE);
  RH_REQUIRE(n_fields == 0 || fields != nullptr, RH_ERR_INVALID_ARG, "fields is NULL");
  RH_REQUIRE(n_dense == 0 || dense != nullptr, RH_ERR_INVALID_ARG, "dense is NULL");
  RH_REQUIRE(batch >= 0, RH_ERR_INVALID_ARG, "negative batch");
  RH_REQUIRE(n_fields == 0 || dim > 0, RH_ERR_INVALID_ARG, "dim must be > 0");
  RH_REQUIRE(n_dense == 0 || tile != nullptr, RH_ERR_INVALID_ARG, "dense columns need a tile");
  if (batch == 0 || (n_fields == 0 && n_dense == 0)) return RH_OK;
  cudaStream_t st = (cudaStream_t)stream;

  static thread_local FwdParams p;
  memset(&p, 0, sizeof(p));
  int rc = pack_fields(fields, n_fields, p.f, false);
  if (rc != RH_OK) return rc;
  bool any_tile = false, any_fm = false, vec_ok = (dim % 4 == 0) && dim <= 128, tile_vec = true;
  for (int i = 0; i < n_fields; ++i) {
    any_tile |= fields[i].tile_col >= 0;
    any_fm |= fields[i].fm_slot >= 0;
    vec_ok &= aligned16(fields[i].table);
    if (fields[i].tile_col >= 0) tile_vec &= (fields[i].tile_col % 4 == 0);
  }
  RH_REQUIRE(!any_tile || tile != nullptr, RH_ERR_INVALID_ARG, "a field has tile_col >= 0 but tile is NULL");
  if (tile != nullptr) tile_vec &= aligned16(tile) && (tile_ld % 4 == 0);
  const bool want_fm = (y_fm != nullptr || y_lr != nullptr || field_sum != nullptr) && any_fm;
  if (want_fm) {
    RH_REQUIRE(y_lr == nullptr || lr_weight != nullptr, RH_ERR_INVALID_ARG, "y_lr requested without lr_weight");
    vec_ok &= (lr_weight == nullptr || aligned16(lr_weight)) && (field_sum == nullptr || aligned16(field_sum));
    RH_REQUIRE(vec_ok, RH_ERR_UNSUPPORTED,
               "fused FM/LR needs dim %% 4 == 0, dim <= 128 and 16-byte aligned tables (dim=%d)", dim);
  }
  for (int j = 0; j < n_dense; ++j) {
    RH_REQUIRE(dense[j].values != nullptr && dense[j].width > 0 && dense[j].tile_col >= 0, RH_ERR_INVALID_ARG, "dense %d invalid", j);
    RH_REQUIRE(dense[j].dtype >= 0 && dense[j].dtype <= 3, RH_ERR_INVALID_ARG, "dense %d: dtype %d unknown", j, dense[j].dtype);
    RH_REQUIRE(dense[j].stride >= 0 && dense[j].stride < (int64_t)1 << 31, RH_ERR_INVALID_ARG, "dense %d: stride out of range", j);
    p.d[j].src = dense[j].values;
    p.d[j].stride = (int32_t)dense[j].stride;
    p.d[j].dtype = dense[j].dtype;
    p.d[j].width = dense[j].width;
    p.d[j].tile_col = dense[j].tile_col;
  }
#ifdef RH_FIELDS_TRACE
  p.trace = g_fields_trace;
#endif
  p.n_fields = n_fields;
  p.n_dense = n_dense;
  p.batch = batch;
  p.dim = n_fields > 0 ? dim : 4;
  p.tile = tile;
  p.tile_ld = tile_ld;
  p.tile_vec = tile_vec ? 1 : 0;
  p.dest_rows = 0;
  if (n_dest > 0) {
    RH_REQUIRE(vec_ok && n_dense == 0 && n_dest <= 8 && dest_rows > 0 && dest != nullptr, RH_ERR_UNSUPPORTED,
               "peer-tile gather needs the 16-byte-lane kernel, no dense columns and <= 8 destinations");
    for (int i = 0; i < n_dest; ++i) {
      RH_REQUIRE(dest[i] != nullptr && aligned16(dest[i]), RH_ERR_INVALID_ARG, "peer tile %d NULL or misaligned", i);
      p.dest[i] = dest[i];
    }
    RH_REQUIRE((int64_t)n_dest * dest_rows >= batch, RH_ERR_INVALID_ARG, "peer tiles cover %d x %d rows < batch %d", n_dest, dest_rows, batch);
    p.dest_rows = dest_rows;
  }
  p.lrw = want_fm ? lr_weight : nullptr;
  p.lrb = want_fm ? lr_bias : nullptr;
  p.yfm = want_fm ? y_fm : nullptr;
  p.ylr = want_fm ? y_lr : nullptr;
  p.fsum = want_fm ? field_sum : nullptr;
  p.err = err_flag;
  if (sync != nullptr && (sync->wait_flags != nullptr || sync->sig_flags != nullptr || sync->id_snapshot_delta != 0)) {
    RH_REQUIRE(vec_ok, RH_ERR_UNSUPPORTED, "rh_fields_fwd_sync needs the 16-byte-lane kernel (dim %% 4 == 0, aligned tables)");
    RH_REQUIRE(sync->wait_flags == nullptr || sync->step != nullptr, RH_ERR_INVALID_ARG, "rh_fields_fwd_sync: wait_flags without step");
    p.sync.wait_flags = sync->wait_flags;
    p.sync.wait_mask = sync->wait_flags != nullptr ? (sync->wait_mask & 0xff) : 0;
    p.sync.step = sync->step;
    if (sync->sig_flags != nullptr) {
      RH_REQUIRE(sync->sig_world >= 1 && sync->sig_world <= 8 && sync->sig_rank >= 0 && sync->sig_rank < 8 && sync->ticket != nullptr && sync->step != nullptr,
                 RH_ERR_INVALID_ARG, "rh_fields_fwd_sync: bad signal side (world %d rank %d)", sync->sig_world, sync->sig_rank);
      for (int i = 0; i < sync->sig_world; ++i) {
        RH_REQUIRE(sync->sig_flags[i] != nullptr, RH_ERR_INVALID_ARG, "rh_fields_fwd_sync: flags of rank %d NULL", i);
        p.sync.sig_flags[i] = sync->sig_flags[i];
      }
      p.sync.sig_world = sync->sig_world;
      p.sync.sig_rank = sync->sig_rank;
      p.sync.ticket = reinterpret_cast<unsigned*>(sync->ticket);
    }
    if (sync->id_snapshot_delta != 0) {
      for (int i = 0; i < n_fields; ++i)
        RH_REQUIRE(fields[i].ids_are_i32 == 0, RH_ERR_UNSUPPORTED, "rh_fields_fwd_sync: the id snapshot needs int64 id buffers (field %d)", i);
      RH_REQUIRE(sync->id_snapshot_delta % 8 == 0, RH_ERR_INVALID_ARG, "rh_fields_fwd_sync: id_snapshot_delta must be a multiple of 8 bytes");
      p.sync.snap_delta = sync->id_snapshot_delta;
    }
  }

  if (vec_ok) {
    const int lpr = n_fields > 0 ? pow2_ceil(dim / 4) : 4;
    switch (lpr) {
      case 1: launch_fwd_v4<1>(p, st); break;
      case 2: launch_fwd_v4<2>(p, st); break;
      case 4: launch_fwd_v4<4>(p, st); break;
      case 8: launch_fwd_v4<8>(p, st); break;
      case 16: launch_fwd_v4<16>(p, st); break;
      default: launch_fwd_v4<32>(p, st); break;
    }
  } else {
    const int64_t work = (int64_t)batch * ((int64_t)n_fields * dim + 1);
    int grid = (int)((work + 255) / 256);
    const int cap = num_sms() * 8;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    fields_fwd_scalar<<<grid, 256, 0, st>>>(p);
  }
  RH_LAUNCH_CHECK();
  return RH_OK;
}

extern "C" int rh_fields_bwd(const rh_field* fields, int n_fields, int dim, int batch, const float* tile, int64_t tile_ld,
                             const float* d_tile, int64_t d_tile_ld, const float* d_y_fm, const float* d_y_lr, const float* lr_weight,
                             const float* field_sum, float* d_lr_weight, float* d_lr_bias, int32_t* err_flag, void* stream) {
  RH_REQUIRE(n_fields >= 0 && n_fields <= RH_MAX_FIELDS, RH_ERR_INVALID_ARG, "n_fields %d not in [0,%d]", n_fields, RH_MAX_FIELDS);
  RH_REQUIRE(n_fields == 0 || fields != nullptr, RH_ERR_INVALID_ARG, "fields is NULL");
  RH_REQUIRE(batch >= 0 && dim > 0, RH_ERR_INVALID_ARG, "bad batch/dim");
  if (batch == 0 || n_fields == 0) return RH_OK;
  cudaStream_t st = (cudaStream_t)stream;

  static thread_local BwdParams p;
  memset(&p, 0, sizeof(p));
  int rc = pack_fields(fields, n_fields, p.f, true);
  if (rc != RH_OK) return rc;
  bool any_fm = false, vec_ok = (dim % 4 == 0) && dim <= 128, tile_vec = true, dtile_vec = true;
  for (int i = 0; i < n_fields; ++i) {
    any_fm |= fields[i].fm_slot >= 0;
    vec_ok &= aligned16(fields[i].table) && (fields[i].table_grad == nullptr || aligned16(fields[i].table_grad));
    if (fields[i].tile_col >= 0) tile_vec &= (fields[i].tile_col % 4 == 0);
  }
  const bool want_fm = any_fm && (d_y_fm != nullptr || d_y_lr != nullptr);
  dtile_vec = tile_vec;  // column alignment is shared
  if (tile != nullptr) tile_vec &= aligned16(tile) && (tile_ld % 4 == 0);
  if (d_tile != nullptr) dtile_vec &= aligned16(d_tile) && (d_tile_ld % 4 == 0);
  if (want_fm) {
    RH_REQUIRE(d_y_fm == nullptr || field_sum != nullptr, RH_ERR_INVALID_ARG, "d_y_fm needs field_sum");
    RH_REQUIRE(d_y_lr == nullptr || lr_weight != nullptr, RH_ERR_INVALID_ARG, "d_y_lr needs lr_weight");
    vec_ok &= (lr_weight == nullptr || aligned16(lr_weight)) && (field_sum == nullptr || aligned16(field_sum)) &&
              (d_lr_weight == nullptr || aligned16(d_lr_weight));
    RH_REQUIRE(vec_ok, RH_ERR_UNSUPPORTED, "fused FM/LR backward needs dim %% 4 == 0, dim <= 128 and 16-byte alignment (dim=%d)", dim);
  }
  p.n_fields = n_fields;
  p.batch = batch;
  p.dim = dim;
  p.tile = tile;
  p.dtile = d_tile;
  p.tile_ld = tile_ld;
  p.dtile_ld = d_tile_ld;
  p.tile_vec = tile_vec ? 1 : 0;
  p.dtile_vec = dtile_vec ? 1 : 0;
  p.dyfm = want_fm ? d_y_fm : nullptr;
  p.dylr = want_fm ? d_y_lr : nullptr;
  p.lrw = lr_weight;
  p.fsum = field_sum;
  p.dlrw = d_lr_weight;
  p.dlrb = d_lr_bias;
  p.err = err_flag;

  if (vec_ok) {
    const int lpr = pow2_ceil(dim / 4);
    switch (lpr) {
      case 1: launch_bwd_v4<1>(p, st); break;
      case 2: launch_bwd_v4<2>(p, st); break;
      case 4: launch_bwd_v4<4>(p, st); break;
      case 8: launch_bwd_v4<8>(p, st); break;
      case 16: launch_bwd_v4<16>(p, st); break;
      default: launch_bwd_v4<32>(p, st); break;
    }
  } else {
    if (d_tile == nullptr) return RH_OK;  // nothing flows into the tables
    const int64_t work = (int64_t)batch * n_fields * dim;
    int grid = (int)((work + 255) / 256);
    const int cap = num_sms() * 8;
    if (grid > cap) grid = cap;
    fields_bwd_scalar<<<grid, 256, 0, st>>>(p);
  }
  RH_LAUNCH_CHECK();
  return RH_OK;
}

extern "C" int rh_fields_fwd(const rh_field* fields, int n_fields, int dim, const rh_dense* dense, int n_dense, int batch, float* tile,
                             int64_t tile_ld, const float* lr_weight, const float* lr_bias, float* y_fm, float* y_lr, float* field_sum,
                             int32_t* err_flag, void* stream) {
  return fields_fwd_impl(fields, n_fields, dim, dense, n_dense, batch, tile, tile_ld, lr_weight, lr_bias, y_fm, y_lr, field_sum, err_flag, stream,
                         nullptr, 0, 0);
}

extern "C" int rh_fields_fwd_p2p(const rh_field* fields, int n_fields, int dim, int batch, float* const* dest_tiles, int n_dest,
                                 int rows_per_dest, int64_t tile_ld, int32_t* err_flag, void* stream) {
  RH_REQUIRE(dest_tiles != nullptr && n_dest > 0, RH_ERR_INVALID_ARG, "rh_fields_fwd_p2p: no destinations");
  return fields_fwd_impl(fields, n_fields, dim, nullptr, 0, batch, dest_tiles[0], tile_ld, nullptr, nullptr, nullptr, nullptr, nullptr, err_flag, stream,
                         dest_tiles, n_dest, rows_per_dest);
}

extern "C" int rh_fields_fwd_sync(const rh_field* fields, int n_fields, int dim, const rh_dense* dense, int n_dense, int batch, float* tile,
                                  int64_t tile_ld, const float* lr_weight, const float* lr_bias, float* y_fm, float* y_lr, float* field_sum,
                                  float* const* dest_tiles, int n_dest, int rows_per_dest, const rh_sync* sync, int32_t* err_flag, void* stream) {
  if (dest_tiles != nullptr) {
    RH_REQUIRE(n_dest > 0 && dense == nullptr && n_dense == 0 && lr_weight == nullptr && y_fm == nullptr && y_lr == nullptr && field_sum == nullptr,
               RH_ERR_INVALID_ARG, "rh_fields_fwd_sync: the peer-memory gather takes no dense columns and no FM / LR outputs");
    return fields_fwd_impl(fields, n_fields, dim, nullptr, 0, batch, dest_tiles[0], tile_ld, nullptr, nullptr, nullptr, nullptr, nullptr, err_flag, stream,
                           dest_tiles, n_dest, rows_per_dest, sync);
  }
  return fields_fwd_impl(fields, n_fields, dim, dense, n_dense, batch, tile, tile_ld, lr_weight, lr_bias, y_fm, y_lr, field_sum, err_flag, stream, nullptr, 0,
                         0, sync);
}

// ids of my samples -> the owners' id buffers (peer memory).  One thread per (destination, sample); the id of the
// destination's k-th field lands at  dst[k * slot_stride + sample * sample_stride].  Field-major buffers
// (sample_stride 1) make every warp store one contiguous 256-byte run over NVLink, and give the owner's gather unit-stride ids.
namespace rh {
struct IdScatterP {
  const void* src[RH_MAX_FIELDS];   // columns sorted by (destination, slot)
  int32_t stride[RH_MAX_FIELDS];
  uint8_t is_i32[RH_MAX_FIELDS];
  long long* dst[8];                 // per destination: base of MY block in its (W, b, fmax) id buffer
  int32_t first[8], count[8];        // columns of destination d: [first[d], first[d] + count[d])
  int32_t batch, fmax;
  long long slot_stride, sample_stride;
  // rh_ids_scatter_signal: publish *step + 1 to the owners' id-phase flags once every CTA's ids are stored (NULL: plain scatter)
  int32_t* sig_flags[8];
  int32_t* step;
  unsigned* ticket;
  int32_t sig_rank, sig_world;
};
__global__ void __launch_bounds__(256) ids_scatter_kernel(const __grid_constant__ IdScatterP p) {
  __shared__ int is_last;
  const int e = p.step != nullptr ? *p.step + 1 : 0;  // the last CTA overwrites *step only after every CTA has taken its ticket
  const int d = blockIdx.y;
  const int first = p.first[d], count = p.count[d];
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < p.batch; j += gridDim.x * blockDim.x) {
    long long* out = p.dst[d] + (int64_t)j * p.sample_stride;
#pragma unroll 4
    for (int k = 0; k < count; ++k)
      out[(int64_t)k * p.slot_stride] = (long long)load_id(p.src[first + k], (int64_t)j * p.stride[first + k], p.is_i32[first + k] != 0);
  }
  if (p.ticket == nullptr) return;  // grid-uniform
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(p.ticket, 1u) == gridDim.x * gridDim.y - 1) ? 1 : 0;
  __syncthreads();
  if (!is_last) return;
  if ((int)threadIdx.x < p.sig_world) {
    __threadfence_system();
    st_release_sys(p.sig_flags[threadIdx.x] + p.sig_rank, e);
  }
  if (threadIdx.x == 0) {
    *p.ticket = 0u;
    *p.step = e;
  }
}
}  // namespace rh

static int ids_scatter_impl(const rh_field* cols, int n_cols, const int32_t* col_dest, int batch, int64_t* const* dest_base, int n_dest, int fmax,
                            int64_t slot_stride, int64_t sample_stride, int32_t* const* peer_flags, int rank, int world, int32_t* step_dev,
                            int32_t* ticket_dev, void* stream) {
  RH_REQUIRE(cols != nullptr && col_dest != nullptr && dest_base != nullptr && n_cols > 0 && n_cols <= RH_MAX_FIELDS && batch >= 0 && fmax > 0 &&
                 n_dest > 0 && n_dest <= 8 && slot_stride > 0 && sample_stride > 0,
             RH_ERR_INVALID_ARG, "rh_ids_scatter: bad arguments");
  RH_REQUIRE(batch > 0 || peer_flags == nullptr, RH_ERR_INVALID_ARG, "rh_ids_scatter_signal: an empty batch cannot publish a step");
  if (batch == 0) return RH_OK;
  static thread_local rh::IdScatterP p;
  memset(&p, 0, sizeof(p));
  if (peer_flags != nullptr) {
    RH_REQUIRE(world >= 1 && world <= 8 && rank >= 0 && rank < world && step_dev != nullptr && ticket_dev != nullptr, RH_ERR_INVALID_ARG,
               "rh_ids_scatter_signal: bad rank / world / step / ticket");
    for (int i = 0; i < world; ++i) {
      RH_REQUIRE(peer_flags[i] != nullptr, RH_ERR_INVALID_ARG, "rh_ids_scatter_signal: flags of rank %d NULL", i);
      p.sig_flags[i] = peer_flags[i];
    }
    p.sig_rank = rank;
    p.sig_world = world;
    p.step = step_dev;
    p.ticket = reinterpret_cast<unsigned*>(ticket_dev);
  }
  int prev = -1;
  for (int i = 0; i < n_cols; ++i) {
    const int d = col_dest[i];
    RH_REQUIRE(cols[i].ids != nullptr && d >= 0 && d < n_dest && d >= prev, RH_ERR_INVALID_ARG, "rh_ids_scatter: column %d (columns must be sorted by destination)", i);
    RH_REQUIRE(cols[i].id_stride >= 0 && cols[i].id_stride < ((int64_t)1 << 31), RH_ERR_INVALID_ARG, "rh_ids_scatter: stride out of range");
    if (d != prev) p.first[d] = i;
    p.count[d] += 1;
    prev = d;
    p.src[i] = cols[i].ids;
    p.stride[i] = (int32_t)cols[i].id_stride;
    p.is_i32[i] = (uint8_t)(cols[i].ids_are_i32 != 0);
  }
  for (int d = 0; d < n_dest; ++d) {
    RH_REQUIRE(dest_base[d] != nullptr && p.count[d] <= fmax, RH_ERR_INVALID_ARG, "rh_ids_scatter: destination %d", d);
    p.dst[d] = reinterpret_cast<long long*>(dest_base[d]);
  }
  p.batch = batch;
  p.fmax = fmax;
  p.slot_stride = slot_stride;
  p.sample_stride = sample_stride;
  dim3 grid((batch + 255) / 256, n_dest);
  rh::ids_scatter_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p);
  RH_LAUNCH_CHECK();
  return RH_OK;
}

extern "C" int rh_ids_scatter(const rh_field* cols, int n_cols, const int32_t* col_dest, int batch, int64_t* const* dest_base, int n_dest, int fmax,
                              int64_t slot_stride, int64_t sample_stride, void* stream) {
  return ids_scatter_impl(cols, n_cols, col_dest, batch, dest_base, n_dest, fmax, slot_stride, sample_stride, nullptr, 0, 0, nullptr, nullptr, stream);
}

extern "C" int rh_ids_scatter_signal(const rh_field* cols, int n_cols, const int32_t* col_dest, int batch, int64_t* const* dest_base, int n_dest, int fmax,
                                     int64_t slot_stride, int64_t sample_stride, int32_t* const* peer_flags, int rank, int world, int32_t* step_dev,
                                     int32_t* ticket_dev, void* stream) {
  RH_REQUIRE(peer_flags != nullptr, RH_ERR_INVALID_ARG, "rh_ids_scatter_signal: peer_flags is NULL");
  return ids_scatter_impl(cols, n_cols, col_dest, batch, dest_base, n_dest, fmax, slot_stride, sample_stride, peer_flags, rank, world, step_dev, ticket_dev,
                          stream);
}
