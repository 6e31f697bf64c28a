// rh_gemm.cu — fp32-accurate tower GEMM on the 5th-generation tensor cores: C[M,N] = A[M,K] * B[N,K]^T (+ bias[N]).
//
// Why not TF32 straight: the parity bar is |d logit| <= 1e-4 |ref| + 1e-6 against an fp32 reference; one TF32 product
// carries ~5e-4 relative error.  3xTF32 (split every fp32 operand x = hi + lo, hi = top 19 bits, and accumulate
// hi*hi + hi*lo + lo*hi in fp32 TMEM accumulators) restores ~2^-22 per product — fp32-level — at 3 tensor-core MMAs
// per k-step, still ~10x the fp32 SIMT rate the reference-equivalent cuBLAS sgemm runs at on this part.
//
// Structure (one 128 x BN output tile per CTA, BN = 128 or — when 128-wide tiles would leave more than half of the SMs idle — 64;
// BLOCK_K = 32 floats = one 128-byte swizzle row, 3-stage ring):
//   warp 0      TMA producer: cp.async.bulk.tensor.2d of the raw fp32 A / B tiles (SWIZZLE_128B), mbarrier tx bytes
//   warps 2..7  splitters: read the raw tiles, write hi (truncated) back in place and lo into the twin tiles,
//               fence.proxy.async, arrive on `ready`
//   warp 1      MMA issuer: one thread issues the tcgen05.mma.kind::tf32 (M128 K8) of a k-block, tcgen05.commit frees the stage;
//               also owns the TMEM allocation (256 columns: main + cross-term accumulators).  With the B twins laid out back
//               to back (concat_b, default) hi*hi and hi*lo are ONE MMA of width 2 BN (A_hi read once): 8 MMAs per k-block
//               instead of 12, 5 operand-tile reads instead of 6
//   all warps   epilogue (default): tcgen05.ld 32x32b.x32 of both accumulators -> registers -> (+bias) -> a 32 x 32 box in the
//               idle stage ring (128-byte swizzle) -> ONE cp.async.bulk.tensor store (cp.reduce...add for split-K) per box.
//               Warp w owns TMEM lane quadrant w % 4 and every second 32-column chunk.  (Round 2 trace: the lane-per-row
//               st.global.v4 epilogue cost 7.2 k cycles per 128 x 128 tile — 32 scattered 16-byte stores per instruction, after
//               TMEM reads that could not overlap them; kept as the fallback for unaligned C and for the STATS instantiation.)
// Operands may be K-major (row-major [rows, K]) or MN-major (row-major [K, rows], i.e. the transposed use of a stored
// matrix): dX = dH * W and dW = dH^T * X read W, dH and X as stored — no transposed copies.
//
// Replaces the Linear GEMMs of MLP.forward/backward (reference basic/layers.py:281-292; cuBLAS sgemm via ATen there).
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#ifndef RH_PDL_FAMILY  // (the trace tools include several of these files into one unit: the first one names the family)
#define RH_PDL_FAMILY 1  /* rh_set_pdl mask bit of this file's kernels */
#endif
#include "rh_common.cuh"

namespace rh {

constexpr int kBM = 128, kBN = 128, kBK = 32, kStages = 3;
constexpr int kTileBytes = kBM * kBK * 4;         // 16 KB: one operand tile (hi or lo)
constexpr int kStageBytes = 4 * kTileBytes;       // A_hi, B_hi, A_lo, B_lo
constexpr int kSplitWarps = 6;                    // warps 2..7
constexpr int kGemmThreads = 256;
constexpr size_t kGemmSmem = (size_t)kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
               "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_c, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_c),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ float rna_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
        "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
        "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t smem_addr, int c0, int c1, bool reduce_add) {
  if (reduce_add) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(smem_addr), "r"(c0), "r"(c1)
                 : "memory");
  } else {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(smem_addr), "r"(c0), "r"(c1) : "memory");
  }
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Shared-memory matrix descriptors (cute/arch/mma_sm100_desc.hpp: SmemDescriptor), version 1.
//   K-major : SWIZZLE_128B (layout type 2): rows of 128 B, 8-row groups 1024 B apart (SBO); one MMA k-step (8 tf32) = +32 B
//             on the start address inside the swizzle atom.
//   MN-major: tf32 MN-major operands only exist with SWIZZLE_128B_BASE32B (layout type 1; TMA SWIZZLE_128B_ATOM_32B):
//             atom = 32 MN elements (128 B) x 4 k rows (512 B), 32-B chunks XOR-ed with (row & 3).  Our tile = four TMA
//             boxes {32 MN, 32 K} of 4096 B: MN groups LBO = 4096 B apart, 4-k groups SBO = 512 B apart; one MMA k-step
//             (8 k rows) = +1024 B.
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, bool mn_major) {
  uint64_t lbo = 1u, sbo = 1024u >> 4, type = 2;
  if (mn_major) {
    type = 1;
    lbo = 4096u >> 4;
    sbo = 512u >> 4;
  }
  return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (lbo << 16) | (sbo << 32) | (1ull << 46) | (type << 61);
}

struct GemmP {
  float* C;
  int64_t ldc;
  const float* bias;
  int M, N, K;
  int a_mn, b_mn;   // operand majors: 0 = K-major, 1 = MN-major
  int kblocks_per_split;
  int reduce;       // 1: red.global.add into C (split-K), 0: plain stores
  int bn;           // output tile width: 128 or 64 (B tile = bn x 32 floats; the stage layout keeps its 16 KB slots)
  int epi;          // 1: epilogue through shared memory + TMA stores (map_c), by all 8 warps; 0: lane-per-row stores by warps 4..7
  int bcat;         // 1: stage = [A_hi | A_lo | B_hi | B_lo] with B_lo right behind B_hi: hi*hi and hi*lo are one MMA of width 2 bn
  // STATS instantiation only (BatchNorm column statistics of C in the epilogue):
  float* stats;            // (2N + 1): mean | biased variance | step counter bits
  float* partial;          // [n_tile][m_tile][2][128]: per-tile column mean and M2
  unsigned* tickets;       // [n_tiles], zero on entry, left zero
  float* running_mean;
  float* running_var;
  long long* num_batches_tracked;
  float momentum;
#ifdef RH_GEMM_TRACE
  unsigned long long* trace;  // tools/gemm_trace.cu: [cta][16] clock64 stamps of the pipeline's milestones
#endif
};

#ifdef RH_GEMM_TRACE
#define RH_TR(ev)                                                                                                              \
  do {                                                                                                                         \
    if (p.trace != nullptr) p.trace[((size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 16 + (ev)] = clock64(); \
  } while (0)
#else
#define RH_TR(ev) \
  do {            \
  } while (0)
#endif

// Transpose-reduce across a warp: every lane holds 32 values v[j]; afterwards lane j holds sum over lanes of v[j] in v[0].
// 31 shuffles (16 + 8 + 4 + 2 + 1) instead of 32 x 5 for column-wise warp sums.
template <int S>
__device__ __forceinline__ void butterfly_step(float (&v)[32], int lane) {
  const bool up = (lane & S) != 0;
#pragma unroll
  for (int i = 0; i < S; ++i) {
    const float keep = up ? v[i + S] : v[i];
    const float send = up ? v[i] : v[i + S];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, S);
  }
}
__device__ __forceinline__ void warp_transpose_sum(float (&v)[32], int lane) {
  butterfly_step<16>(v, lane);
  butterfly_step<8>(v, lane);
  butterfly_step<4>(v, lane);
  butterfly_step<2>(v, lane);
  butterfly_step<1>(v, lane);
}

// Chan's pairwise update of (count, mean, M2) with another group's (nb, mean_b, M2_b).
__device__ __forceinline__ void welford_merge(float& n, float& mean, float& m2, float nb, float mean_b, float m2_b) {
  if (nb <= 0.f) return;
  const float tot = n + nb;
  const float delta = mean_b - mean;
  mean += delta * (nb / tot);
  m2 += m2_b + delta * delta * (n * nb / tot);
  n = tot;
}

template <bool STATS>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const __grid_constant__ CUtensorMap map_c,
                   const GemmP p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)kStages * kStageBytes);
  uint64_t* full = bars;                 // [kStages]
  uint64_t* ready = bars + kStages;      // [kStages]
  uint64_t* empty = bars + 2 * kStages;  // [kStages]
  uint64_t* tmem_full = bars + 3 * kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * kStages + 1);

  pdl_trigger();  // the successor may be scheduled as soon as every CTA of this grid is running (its pdl_wait still waits for our completion)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bn = STATS ? kBN : p.bn;
  const int m0 = blockIdx.x * kBM, n0 = blockIdx.y * bn;
  const int total_kb = (p.K + kBK - 1) / kBK;
  const int kb0 = blockIdx.z * p.kblocks_per_split;
  int kb1 = kb0 + p.kblocks_per_split;
  if (kb1 > total_kb) kb1 = total_kb;
  const int n_iter = kb1 - kb0;  // >= 1 by construction of the grid
  // stage layout (byte offsets inside a 64 KB stage): legacy [A_hi][B_hi][A_lo][B_lo] in 16 KB slots; concat_b [A_hi][A_lo][B_hi][B_lo]
  // with B_lo RIGHT behind the bn x 128 bytes of B_hi (one 2 bn-row operand for the MMA)
  const bool bcat = !STATS && p.bcat != 0;
  const uint32_t off_alo = bcat ? (uint32_t)kTileBytes : 2u * kTileBytes;
  const uint32_t off_bhi = bcat ? 2u * kTileBytes : (uint32_t)kTileBytes;
  const uint32_t off_blo = bcat ? 2u * kTileBytes + (uint32_t)bn * 128u : 3u * kTileBytes;
  const uint32_t cross_col = bcat ? (uint32_t)bn : 128u;  // TMEM column of the cross-term accumulator
  if (threadIdx.x == 0) RH_TR(0);
  if (threadIdx.x == 32) {  // the descriptors' first use costs a fetch (~800 cycles before the first TMA issue, tools/gemm_trace.cu): start it now
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_a)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_b)) : "memory");
    if (!STATS && p.epi) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&map_c)) : "memory");
  }

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&ready[s], kSplitWarps);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // TMEM: 2 x 128 fp32 accumulator columns (main + cross terms)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) RH_TR(1);
  pdl_wait();  // barriers, TMEM and the descriptor prefetch are set up; the operands (and C, for split-K) belong to the predecessor until here

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (int it = 0; it < n_iter; ++it) {
        const int s = it % kStages;
        const uint32_t ph = (it / kStages) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        uint8_t* st = smem + (size_t)s * kStageBytes;
        mbar_arrive_expect_tx(&full[s], kTileBytes + bn * kBK * 4);
        const int k0 = (kb0 + it) * kBK;
        if (p.a_mn) {
          for (int i = 0; i < 4; ++i) tma_load_2d(st + i * 4096, &map_a, &full[s], m0 + 32 * i, k0);
        } else {
          tma_load_2d(st, &map_a, &full[s], k0, m0);
        }
        if (p.b_mn) {
          for (int i = 0; i < bn / 32; ++i) tma_load_2d(st + off_bhi + i * 4096, &map_b, &full[s], n0 + 32 * i, k0);
        } else {
          tma_load_2d(st + off_bhi, &map_b, &full[s], k0, n0);
        }
        if (it == 0) RH_TR(2);
        if (it == n_iter - 1) RH_TR(3);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer =====
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)p.a_mn << 15) | ((uint32_t)p.b_mn << 16) | ((uint32_t)(bn >> 3) << 17) |
                           ((uint32_t)(kBM >> 4) << 24);
    const uint32_t idesc2 = (idesc & ~(0x3Fu << 17)) | ((uint32_t)((2 * bn) >> 3) << 17);  // the same MMA, N = 2 bn
    const uint32_t a_step = p.a_mn ? (1024u >> 4) : (32u >> 4);
    const uint32_t b_step = p.b_mn ? (1024u >> 4) : (32u >> 4);
    for (int it = 0; it < n_iter; ++it) {
      const int s = it % kStages;
      const uint32_t ph = (it / kStages) & 1;
      mbar_wait(&ready[s], ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (lane == 0) {
        if (it == 0) RH_TR(6);
        const uint32_t base = smem_u32(smem + (size_t)s * kStageBytes);
        const uint64_t a_hi = make_desc(base, p.a_mn != 0), b_hi = make_desc(base + off_bhi, p.b_mn != 0);
        const uint64_t a_lo = make_desc(base + off_alo, p.a_mn != 0), b_lo = make_desc(base + off_blo, p.b_mn != 0);
#pragma unroll
        for (int kk = 0; kk < kBK / 8; ++kk) {
          const uint64_t da = (uint64_t)(kk * a_step), db = (uint64_t)(kk * b_step);
          // The tensor core accumulates with round-toward-zero: every MMA into a large accumulator adds a -2^-24-relative
          // bias (measured: 160 MMAs into one accumulator = 11x the fp32 error).  The small cross terms therefore get their
          // OWN accumulator; the main one sees one MMA per k-step.  The epilogue adds the two in fp32.
          const uint32_t first = (it > 0 || kk > 0) ? 1u : 0u;
          if (bcat) {
            // [B_hi ; B_lo] is one operand of 2 bn rows: columns [0, bn) of the accumulator block receive A_hi B_hi^T (main),
            // columns [bn, 2 bn) A_hi B_lo^T (cross) — the same sums as the two MMAs below, with A_hi read once
            umma_tf32(tmem_base, a_hi + da, b_hi + db, idesc2, first);
            umma_tf32(tmem_base + cross_col, a_lo + da, b_hi + db, idesc, 1u);
          } else {
            umma_tf32(tmem_base, a_hi + da, b_hi + db, idesc, first);
            umma_tf32(tmem_base + 128u, a_hi + da, b_lo + db, idesc, first);
            umma_tf32(tmem_base + 128u, a_lo + da, b_hi + db, idesc, 1u);
          }
        }
        umma_commit(&empty[s]);                       // stage free once these MMAs have read it
        if (it == n_iter - 1) umma_commit(tmem_full);  // accumulator complete
        if (it == n_iter - 1) RH_TR(7);
      }
      __syncwarp();
    }
  } else {
    // ===== splitters (warps 2..7): hi = top 19 bits (in place), lo = x - hi (twin tile, same swizzled offsets) =====
    const int t = threadIdx.x - 64;  // 0..191
    for (int it = 0; it < n_iter; ++it) {
      const int s = it % kStages;
      const uint32_t ph = (it / kStages) & 1;
      mbar_wait(&full[s], ph);
      if (t == 0 && it == 0) RH_TR(4);
      uint8_t* st = smem + (size_t)s * kStageBytes;
      for (int i = t; i < (kTileBytes + bn * kBK * 4) / 16; i += kSplitWarps * 32) {  // the A tile's 1024 16-byte pieces, then the bn x 32 B tile's
        const bool is_a = i < kTileBytes / 16;
        const uint32_t so = is_a ? (uint32_t)i * 16u : off_bhi + (uint32_t)(i - kTileBytes / 16) * 16u;
        float4* src = reinterpret_cast<float4*>(st + so);
        float4* dst = reinterpret_cast<float4*>(st + so + (is_a ? off_alo : off_blo - off_bhi));
        const float4 x = *src;
        // lo = x - (top 19 bits of x): exact in fp32 (the 13 dropped mantissa bits); the tensor core keeps its top 11 bits.
        // (Rounding lo to tf32 with cvt.rna was measured to change nothing — the residual error is the accumulator's
        // round-toward-zero, handled by the separate cross-term accumulator — and cost 8 k conversions per k-block.)
        float4 l;
        l.x = x.x - __uint_as_float(__float_as_uint(x.x) & 0xffffe000u);
        l.y = x.y - __uint_as_float(__float_as_uint(x.y) & 0xffffe000u);
        l.z = x.z - __uint_as_float(__float_as_uint(x.z) & 0xffffe000u);
        l.w = x.w - __uint_as_float(__float_as_uint(x.w) & 0xffffe000u);
        *dst = l;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor-core (async) proxy
      __syncwarp();
      if (lane == 0) mbar_arrive(&ready[s]);
      if (t == 0 && it == 0) RH_TR(5);
      if (t == 0 && it == n_iter - 1) RH_TR(8);
    }
    // ===== lane-per-row epilogue (warps 4..7 own TMEM lane quadrants 0..3): STATS instantiation and unaligned C =====
    if (warp >= 4 && (STATS || !p.epi)) {
      const int q = warp - 4;
      mbar_wait(tmem_full, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (q == 0 && lane == 0) RH_TR(9);
      const int m = m0 + q * 32 + lane;
      const bool add_bias = p.bias != nullptr && blockIdx.z == 0;
      const bool vec_ok = (p.ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15u) == 0);
#pragma unroll 1
      for (int c = 0; c < bn / 32; ++c) {
        uint32_t r[32], x[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32);
        tmem_ld32(taddr, r);
        tmem_ld32(taddr + cross_col, x);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(x[j]));
        if constexpr (STATS) {
          // column statistics of this warp's 32 rows (bias included: BatchNorm sees h = xW^T + b), two-pass inside the warp:
          // column means by a transpose-reduce, then the squared deviations from the warp's own mean
          const int nbs = n0 + c * 32;
          const bool row_ok = m < p.M;
          int cnt = p.M - (m0 + q * 32);
          cnt = cnt < 0 ? 0 : (cnt > 32 ? 32 : cnt);
          float t[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float v = __uint_as_float(r[j]);
            if (add_bias && nbs + j < p.N) v += __ldg(p.bias + nbs + j);
            t[j] = row_ok ? v : 0.f;
          }
          float dsq[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) dsq[j] = t[j];
          warp_transpose_sum(dsq, lane);
          const float wmean = cnt > 0 ? dsq[0] / (float)cnt : 0.f;  // lane j: mean of column nbs + j over the warp's rows
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float d = t[j] - __shfl_sync(0xffffffffu, wmean, j);
            dsq[j] = row_ok ? d * d : 0.f;
          }
          warp_transpose_sum(dsq, lane);
          float* sm_stats = reinterpret_cast<float*>(smem);  // [4 warps][2][128]; the stage ring is idle by now
          sm_stats[(q * 2 + 0) * 128 + c * 32 + lane] = wmean;
          sm_stats[(q * 2 + 1) * 128 + c * 32 + lane] = dsq[0];
        }
        // ---- write-out: lane = row, 16-byte stores.  The bias of the chunk is read into registers AHEAD of the stores: with a load
        // ---- behind every store (same loop body) the compiler must keep them in order and each load waits out its latency —
        // ---- measured 15.7 k cycles for this epilogue with bias against 4.7 k without (tools/gemm_trace.cu).  (Staging the block
        // ---- through shared memory for row-contiguous stores was measured too: 5.6 k cycles — the TMEM reads, not the store
        // ---- pattern, set the pace: 2 x 64 KB per CTA at 64 B/cycle.)
        if (m < p.M) {
          const int nb = n0 + c * 32;
          float* crow = p.C + (int64_t)m * p.ldc + nb;
          if (add_bias) {
            float bv[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) bv[j] = (nb + j < p.N) ? __ldg(p.bias + nb + j) : 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + bv[j]);
          }
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 v = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
            if (vec_ok && nb + j + 3 < p.N) {
              if (p.reduce) red_add_row16(crow + j, v);
              else *reinterpret_cast<float4*>(crow + j) = v;
            } else {
              const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                if (nb + j + e < p.N) {
                  if (p.reduce) atomicAdd(crow + j + e, vv[e]);
                  else crow[j + e] = vv[e];
                }
              }
            }
          }
        }
      }
      if (q == 0 && lane == 0) RH_TR(10);
      if constexpr (STATS) {
        // ---- the four epilogue warps merge their row groups, publish the tile's (mean, M2) per column, and the last m-tile of
        // ---- this column block merges all tiles in a fixed order (deterministic) and finalises BatchNorm's statistics
        float* sm_stats = reinterpret_cast<float*>(smem);
        int* sm_flag = reinterpret_cast<int*>(smem + 4096);
        asm volatile("bar.sync 1, 128;" ::: "memory");
        const int tcol = threadIdx.x - 128;  // 0..127: one column of the tile per epilogue thread
        const int m_tiles = gridDim.x;
        {
          float n = 0.f, mean = 0.f, m2 = 0.f;
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            int cnt = p.M - (m0 + w * 32);
            cnt = cnt < 0 ? 0 : (cnt > 32 ? 32 : cnt);
            welford_merge(n, mean, m2, (float)cnt, sm_stats[(w * 2 + 0) * 128 + tcol], sm_stats[(w * 2 + 1) * 128 + tcol]);
          }
          float* out = p.partial + ((size_t)blockIdx.y * m_tiles + blockIdx.x) * 256;
          __stcg(out + tcol, mean);
          __stcg(out + 128 + tcol, m2);
        }
        __threadfence();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (tcol == 0) {
          const unsigned ticket = atomicAdd(p.tickets + blockIdx.y, 1u);
          *sm_flag = (ticket == (unsigned)(m_tiles - 1)) ? 1 : 0;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (*sm_flag) {
          __threadfence();
          const int col = n0 + tcol;
          if (col < p.N) {
            float n = 0.f, mean = 0.f, m2 = 0.f;
            for (int mt = 0; mt < m_tiles; ++mt) {
              int cnt = p.M - mt * kBM;
              cnt = cnt > kBM ? kBM : cnt;
              const float* in = p.partial + ((size_t)blockIdx.y * m_tiles + mt) * 256;
              welford_merge(n, mean, m2, (float)cnt, __ldcg(in + tcol), __ldcg(in + 128 + tcol));
            }
            float var = m2 / n;
            if (var < 0.f) var = 0.f;
            p.stats[col] = mean;
            p.stats[p.N + col] = var;
            if (p.running_mean != nullptr) p.running_mean[col] = (1.f - p.momentum) * p.running_mean[col] + p.momentum * mean;
            if (p.running_var != nullptr) {
              const float unbiased = p.M > 1 ? var * (n / (n - 1.f)) : var;
              p.running_var[col] = (1.f - p.momentum) * p.running_var[col] + p.momentum * unbiased;
            }
          }
          if (tcol == 0) {
            p.tickets[blockIdx.y] = 0u;  // ready for the next launch (graph replays included)
            if (blockIdx.y == 0) {      // once per launch: the step counter = dropout stream id of this forward (as rh_colstats)
              long long count = 0;
              if (p.num_batches_tracked != nullptr) {
                count = *p.num_batches_tracked + 1;
                *p.num_batches_tracked = count;
              }
              p.stats[2 * p.N] = __int_as_float((int)(count & 0x7fffffff));
            }
          }
        }
      }
    }
  }
  if (!STATS && p.epi) {
    // ===== epilogue through shared memory + TMA stores, all 8 warps =====
    // Warp w reads TMEM lane quadrant w % 4 (the hardware's access rule) = rows m0 + 32 (w % 4) ..+31 of the tile, and every second
    // 32-column chunk (chunk c belongs to the warps with (c & 1) == w / 4).  Per chunk: both accumulators -> registers (lane = row),
    // summed (+ bias), written as a 32 x 32 fp32 box into the idle stage ring in TMA's 128-byte swizzle (16-byte piece j of row r at
    // position j ^ (r & 7): conflict-free wavefronts), fence.proxy.async, then ONE bulk tensor store — or, for split-K, one bulk
    // reduce-add: L2 adds whole 16-byte vectors, rows and columns beyond (M, N) are clipped by the tensor map.  The stores drain
    // while the warp reads its next chunk; 8 KB of staging per warp.
    __syncwarp();
    const int q = warp & 3, half = warp >> 2;
    mbar_wait(tmem_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp == 4 && lane == 0) RH_TR(9);
    const bool add_bias = p.bias != nullptr && blockIdx.z == 0;
    const int row0 = m0 + q * 32;
    uint8_t* my_stage = smem + (size_t)warp * 8192;  // the ring's MMAs have all completed (tmem_full) and its TMA loads were consumed
    int slot = 0;
#pragma unroll 1
    for (int c = half; c < bn / 32; c += 2, ++slot) {
      const int nb = n0 + c * 32;
      if (row0 >= p.M || nb >= p.N) continue;  // warp-uniform: nothing of this box lies inside C
      uint32_t r[32], x[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32);
      tmem_ld32(taddr, r);
      tmem_ld32(taddr + cross_col, x);
      float bv[32];
      if (add_bias) {
#pragma unroll
        for (int j = 0; j < 32; ++j) bv[j] = (nb + j < p.N) ? __ldg(p.bias + nb + j) : 0.f;
      }
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(x[j]));
      if (add_bias) {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + bv[j]);
      }
      const uint32_t box = smem_u32(my_stage + (slot & 1) * 4096);
      const uint32_t rowaddr = box + (uint32_t)lane * 128u;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowaddr + (uint32_t)((j ^ (lane & 7)) << 4)), "r"(r[4 * j]), "r"(r[4 * j + 1]),
                     "r"(r[4 * j + 2]), "r"(r[4 * j + 3])
                     : "memory");
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the TMA engine
      __syncwarp();
      if (lane == 0) tma_store_2d(&map_c, box, nb, row0, p.reduce != 0);
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the boxes have been read: shared memory may go
    if (warp == 4 && lane == 0) RH_TR(10);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) RH_TR(11);
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
  }
}

// ---- host ---------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
  }
  return fn;
}

// rows x K fp32 matrix used as a [rows, K] operand.  K-major: stored [rows][ld] -> box {32 (K), tile_rows}.
// MN-major: stored [K][ld] with `rows` contiguous -> box {32 (rows), 32 (K)} (tile_rows / 32 boxes per tile).
static int make_map(CUtensorMap* map, const float* base, int64_t ld, int rows, int K, bool mn_major, int tile_rows) {
  EncodeTiledFn fn = encode_fn();
  RH_REQUIRE(fn != nullptr, RH_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t gdim[2], gstride[1];
  cuuint32_t box[2], estr[2] = {1, 1};
  if (mn_major) {
    gdim[0] = (cuuint64_t)rows;
    gdim[1] = (cuuint64_t)K;
    box[0] = 32;
    box[1] = 32;
  } else {
    gdim[0] = (cuuint64_t)K;
    gdim[1] = (cuuint64_t)rows;
    box[0] = 32;
    box[1] = (cuuint32_t)tile_rows;
  }
  gstride[0] = (cuuint64_t)ld * sizeof(float);
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  RH_REQUIRE(r == CUDA_SUCCESS, RH_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d): base %p ld %lld rows %d K %d mn %d", (int)r, (const void*)base,
             (long long)ld, rows, K, (int)mn_major);
  return RH_OK;
}

}  // namespace rh

using namespace rh;

struct StatsArgs {
  float* stats;
  float* scratch;
  float* running_mean;
  float* running_var;
  int64_t* num_batches_tracked;
  float momentum;
};

#ifdef RH_GEMM_TRACE
unsigned long long* g_gemm_trace = nullptr;  // set by tools/gemm_trace.cu
#endif
static int g_tile_n = 0;  // 0 = choose per problem, 64 / 128 = forced (rh_gemm_tile_n: A/B runs and tests)

extern "C" int rh_gemm_tile_n(int set) {
  if (set == 0 || set == 64 || set == 128) g_tile_n = set;
  return g_tile_n;
}

// Kernel variants (A/B runs and tests): bit 0 = epilogue through shared memory + TMA stores, bit 1 = concatenated B twins.
// Defaults: both on; RECHUB_B200_GEMM_TMA_EPILOGUE=0 / RECHUB_B200_GEMM_CONCAT_B=0 switch them off for the process.
static int g_gemm_opts = -1;
static int gemm_opts() {
  if (g_gemm_opts < 0) {
    const char* e = getenv("RECHUB_B200_GEMM_TMA_EPILOGUE");
    const char* c = getenv("RECHUB_B200_GEMM_CONCAT_B");
    g_gemm_opts = ((e == nullptr || e[0] != '0') ? 1 : 0) | ((c == nullptr || c[0] != '0') ? 2 : 0);
  }
  return g_gemm_opts;
}
extern "C" int rh_gemm_options(int tma_epilogue, int concat_b) {
  int o = gemm_opts();
  if (tma_epilogue >= 0) o = (o & ~1) | (tma_epilogue ? 1 : 0);
  if (concat_b >= 0) o = (o & ~2) | (concat_b ? 2 : 0);
  g_gemm_opts = o;
  return o;
}

// C (M x N, row stride ldc) as 32 x 32 boxes for the epilogue's bulk stores / reduce-adds.
static int make_map_c(CUtensorMap* map, float* base, int64_t ldc, int M, int N) {
  EncodeTiledFn fn = encode_fn();
  RH_REQUIRE(fn != nullptr, RH_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t gdim[2] = {(cuuint64_t)N, (cuuint64_t)M}, gstride[1] = {(cuuint64_t)ldc * sizeof(float)};
  cuuint32_t box[2] = {32, 32}, estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  RH_REQUIRE(r == CUDA_SUCCESS, RH_ERR_CUDA, "cuTensorMapEncodeTiled (C) failed (%d): base %p ldc %lld M %d N %d", (int)r, (void*)base, (long long)ldc, M, N);
  return RH_OK;
}

static int gemm_impl(const float* A, int64_t lda, int a_mn_major, const float* B, int64_t ldb, int b_mn_major, float* C, int64_t ldc, int M, int N,
                     int K, const float* bias, int split_k, void* stream, const StatsArgs* st) {
  RH_REQUIRE(A && B && C, RH_ERR_INVALID_ARG, "rh_gemm_tf32x3: NULL pointer");
  RH_REQUIRE(M > 0 && N > 0 && K > 0 && ldc >= N, RH_ERR_INVALID_ARG, "rh_gemm_tf32x3: bad sizes");
  RH_REQUIRE(lda % 4 == 0 && ldb % 4 == 0, RH_ERR_UNSUPPORTED, "rh_gemm_tf32x3: TMA needs 16-byte row strides (lda=%lld ldb=%lld)", (long long)lda,
             (long long)ldb);
  RH_REQUIRE(((uintptr_t)A & 15u) == 0 && ((uintptr_t)B & 15u) == 0, RH_ERR_UNSUPPORTED, "rh_gemm_tf32x3: operands must be 16-byte aligned");
  RH_REQUIRE(lda >= (a_mn_major ? M : K) && ldb >= (b_mn_major ? N : K), RH_ERR_INVALID_ARG, "rh_gemm_tf32x3: leading dimension too small");
  const int total_kb = (K + kBK - 1) / kBK;
  if (split_k < 1) split_k = 1;
  if (split_k > total_kb) split_k = total_kb;
  const int per = (total_kb + split_k - 1) / split_k;
  split_k = (total_kb + per - 1) / per;  // no empty splits

  // Tile width: the tower's GEMMs are small (M = 4096, N <= 429): 128-wide tiles give 32-64 CTAs for 148 SMs, and the k-loop of a
  // CTA is bound by its own shared-memory traffic (~1400 cycles per k-block, tools/gemm_trace.cu), so halving the B tile both
  // doubles the CTAs and shortens each k-block.  64 only while the grid still fits one wave.
  const int m_tiles = (M + kBM - 1) / kBM;
  int bn = kBN;
  if (st == nullptr && g_tile_n != 128) {
    const int64_t ctas128 = (int64_t)m_tiles * ((N + 127) / 128) * split_k, ctas64 = (int64_t)m_tiles * ((N + 63) / 64) * split_k;
    if (g_tile_n == 64 || (ctas128 * 2 <= num_sms() && ctas64 <= num_sms())) bn = 64;
  }
  CUtensorMap map_a, map_b, map_c;
  int rc = make_map(&map_a, A, lda, M, K, a_mn_major != 0, kBM);
  if (rc != RH_OK) return rc;
  rc = make_map(&map_b, B, ldb, N, K, b_mn_major != 0, bn);
  if (rc != RH_OK) return rc;
  const int opts = gemm_opts();
  // TMA needs 16-byte rows of C; its clipping works on 16-byte pieces, so the row must own the whole last piece (header contract)
  const bool tma_epi = st == nullptr && (opts & 1) != 0 && ldc % 4 == 0 && ((uintptr_t)C & 15u) == 0 && ldc >= (int64_t)((N + 3) / 4) * 4;
  if (tma_epi) {
    rc = make_map_c(&map_c, C, ldc, M, N);
    if (rc != RH_OK) return rc;
  } else {
    map_c = map_a;  // unused by the kernel
  }

  GemmP p;
  memset(&p, 0, sizeof(p));
  p.C = C;
  p.ldc = ldc;
  p.bias = bias;
  p.M = M;
  p.N = N;
  p.K = K;
  p.a_mn = a_mn_major != 0;
  p.b_mn = b_mn_major != 0;
  p.kblocks_per_split = per;
  p.reduce = split_k > 1 ? 1 : 0;
  p.bn = bn;
  p.epi = tma_epi ? 1 : 0;
  p.bcat = (st == nullptr && (opts & 2) != 0) ? 1 : 0;
#ifdef RH_GEMM_TRACE
  p.trace = g_gemm_trace;
#endif
  dim3 grid(m_tiles, (N + bn - 1) / bn, split_k);

  static bool configured[2] = {false, false};
  const int which = st != nullptr ? 1 : 0;
  if (!configured[which]) {
    cudaError_t e = which ? cudaFuncSetAttribute(gemm_tf32x3_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGemmSmem)
                          : cudaFuncSetAttribute(gemm_tf32x3_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGemmSmem);
    RH_REQUIRE(e == cudaSuccess, RH_ERR_CUDA, "rh_gemm_tf32x3: cannot reserve %zu bytes of shared memory: %s", kGemmSmem, cudaGetErrorString(e));
    configured[which] = true;
  }
  if (st != nullptr) {
    RH_REQUIRE(split_k == 1, RH_ERR_INVALID_ARG, "rh_gemm_tf32x3_stats: column statistics need the whole K range in one CTA (split_k = 1)");
    RH_REQUIRE(st->stats != nullptr && st->scratch != nullptr, RH_ERR_INVALID_ARG, "rh_gemm_tf32x3_stats: stats / scratch is NULL");
    p.stats = st->stats;
    p.partial = st->scratch;
    p.tickets = reinterpret_cast<unsigned*>(st->scratch + (size_t)grid.x * grid.y * 256);
    p.running_mean = st->running_mean;
    p.running_var = st->running_var;
    p.num_batches_tracked = reinterpret_cast<long long*>(st->num_batches_tracked);
    p.momentum = st->momentum;
    launch_k(gemm_tf32x3_kernel<true>, grid, dim3(kGemmThreads), kGemmSmem, (cudaStream_t)stream, map_a, map_b, map_c, p);
  } else {
    launch_k(gemm_tf32x3_kernel<false>, grid, dim3(kGemmThreads), kGemmSmem, (cudaStream_t)stream, map_a, map_b, map_c, p);
  }
  RH_LAUNCH_CHECK();
  return RH_OK;
}

extern "C" int rh_gemm_tf32x3(const float* A, int64_t lda, int a_mn_major, const float* B, int64_t ldb, int b_mn_major, float* C, int64_t ldc,
                              int M, int N, int K, const float* bias, int split_k, void* stream) {
  return gemm_impl(A, lda, a_mn_major, B, ldb, b_mn_major, C, ldc, M, N, K, bias, split_k, stream, nullptr);
}

extern "C" int64_t rh_gemm_stats_scratch_floats(int M, int N) {
  const int64_t mt = (M + kBM - 1) / kBM, nt = (N + kBN - 1) / kBN;
  return mt * nt * 256 + nt;
}

extern "C" int rh_gemm_tf32x3_stats(const float* A, int64_t lda, int a_mn_major, const float* B, int64_t ldb, int b_mn_major, float* C, int64_t ldc,
                                    int M, int N, int K, const float* bias, float* stats, float* scratch, float* running_mean,
                                    float* running_var, int64_t* num_batches_tracked, float momentum, void* stream) {
  StatsArgs st = {stats, scratch, running_mean, running_var, num_batches_tracked, momentum};
  return gemm_impl(A, lda, a_mn_major, B, ldb, b_mn_major, C, ldc, M, N, K, bias, 1, stream, &st);
}
