// torch-compatible counter-based RNG.  The reference draws its randomness with many tiny torch
// calls (torch.rand(C) once per node per sweep, MCPG.py:140-141; randint + rand per Metropolis
// iteration, MCPG.py:106-111; rand_like per column, L2A/transformer.py:348-352).  torch's CUDA
// generators are Philox4x32-10 keyed by (seed, offset): element `li` of one such call gets
//
//     idx = li % T, k = li / T, Philox(counter = offset/4 + k/4, subsequence = idx, key = seed)[k % 4]
//
// with T = 256 * grid threads (ATen/native/cuda/DistributionTemplates.h: calc_execution_policy and
// distribution_elementwise_grid_stride_kernel; curand_init(seed, idx, offset) + curand4), and
// every call advances the generator offset by 4 * ceil(numel / (4 T)).  Because the stream is
// counter-based, a fused kernel can compute exactly the numbers the reference's call sequence
// would have produced -- same seed, same flip sequence -- without launching those calls.  The
// Philox rounds and the uint32 -> float maps are curand's own inline device functions, the same
// ones torch compiles.
#pragma once
#include <curand_kernel.h>
#include <stdint.h>

namespace rlsb {

struct TorchRng {
  uint64_t seed;
  uint64_t offset4;         // generator offset / 4 at the first call
  uint32_t threads;         // T of one call (256 * grid.x)
  uint32_t iters_per_call;  // ceil(numel / (4 T)) = counter increments one call consumes
  // Optional device-resident generator state {seed, offset} (CUDA-graph replays: the state advances on
  // the device, rlsb_rng_cursor_advance).  When set it replaces `seed`, and `offset4` counts from it.
  const uint64_t* dev;
};

struct PhiloxKey {
  uint64_t seed, offset4;
};
__device__ __forceinline__ PhiloxKey philox_key(const TorchRng& r) {
  PhiloxKey k{r.seed, r.offset4};
  if (r.dev) k.seed = __ldg(r.dev), k.offset4 += __ldg(r.dev + 1) >> 2;
  return k;
}

// Philox4x32-10 with the ten round keys precomputed (curand_Philox4x32_10 bumps the key inside every round;
// in a loop over counters the compiler re-derives the warp-uniform key schedule each trip on the uniform
// datapath, ~14 issue slots per call).  The keys are pinned in ordinary registers.
struct PhiloxKeys {
  uint32_t a[10], b[10];
};
__device__ __forceinline__ PhiloxKeys philox_keys(uint64_t seed) {
  PhiloxKeys k;
  uint32_t a = (uint32_t)seed, b = (uint32_t)(seed >> 32);
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    k.a[i] = a, k.b[i] = b;
    asm volatile("" : "+r"(k.a[i]), "+r"(k.b[i]));      // opaque: keep them as per-thread registers
    a += PHILOX_W32_0, b += PHILOX_W32_1;
  }
  return k;
}
__device__ __forceinline__ uint4 philox10(uint4 c, const PhiloxKeys& k) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    const uint32_t hi0 = __umulhi(PHILOX_M4x32_0, c.x), lo0 = PHILOX_M4x32_0 * c.x;
    const uint32_t hi1 = __umulhi(PHILOX_M4x32_1, c.z), lo1 = PHILOX_M4x32_1 * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.a[i], lo1, hi0 ^ c.w ^ k.b[i], lo0);
  }
  return c;
}

// the uint32 element `li` of consecutive call number `call` receives
__device__ __forceinline__ uint32_t torch_philox_u32(const TorchRng& r, uint64_t call, uint32_t li) {
  const uint32_t idx = li % r.threads, k = li / r.threads;
  const uint64_t ctr = r.offset4 + call * r.iters_per_call + (k >> 2);
  const uint4 o = curand_Philox4x32_10(make_uint4((uint32_t)ctr, (uint32_t)(ctr >> 32), idx, 0u),
                                       make_uint2((uint32_t)r.seed, (uint32_t)(r.seed >> 32)));
  const uint32_t c = k & 3u;
  return c == 0 ? o.x : c == 1 ? o.y : c == 2 ? o.z : o.w;
}

// torch.rand / rand_like (float32): curand_uniform then the (0,1] -> [0,1) bound reversal
__device__ __forceinline__ float torch_uniform_from_u32(uint32_t x) {
  const float u = _curand_uniform(x);
  return u == 1.0f ? 0.0f : u;
}

}  // namespace rlsb
