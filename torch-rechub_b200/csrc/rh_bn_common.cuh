// rh_bn_common.cuh — pieces shared by the BatchNorm + activation + dropout kernels (rh_mlp.cu, rh_bnfuse.cu).
#pragma once
#include "rh_common.cuh"

namespace rh {

enum { ACT_NONE = 0, ACT_RELU = 1, ACT_DICE = 2, ACT_PRELU = 3, ACT_SIGMOID = 4, ACT_LEAKY = 5 };

__device__ __forceinline__ float sigmoidf_precise(float u) { return 1.f / (1.f + expf(-u)); }

// counter-based uniform bits for dropout: (seed, stream counter, element index) -> 32 bits.
__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x7feb352dU;
  x ^= x >> 15;
  x *= 0x846ca68bU;
  x ^= x >> 16;
  return x;
}
__device__ __forceinline__ bool dropout_keep(uint32_t seed, uint32_t counter, uint64_t idx, float p_drop) {
  const uint32_t h = mix32(mix32((uint32_t)idx ^ seed) + mix32(counter * 0x9E3779B9U + (uint32_t)(idx >> 32)));
  return (float)(h >> 8) * (1.0f / 16777216.0f) >= p_drop;
}

}  // namespace rh
