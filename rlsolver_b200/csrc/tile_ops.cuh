// Per-tile building blocks.  A "tile" is 32 environments x all nodes, held as one
// uint32 word per node (bit b = env 32*tile+b).  Every hot kernel gives one CTA one
// tile in shared memory and works on whole words, so one LOP/POPC serves 32 envs.
#pragma once
#include "common.cuh"
#include "vcount.cuh"

namespace rlsb {

// ---- bool rows -> packed words -------------------------------------------------
// One warp converts a strip of 32*VEC consecutive nodes for the 32 envs of a tile.
// Each load is one coalesced 32*VEC-byte row segment; the bit gather is lane-local
// (lane l ends up owning nodes node0 .. node0+VEC-1 of all 32 envs), so no shuffles.
template <int VEC>
__device__ __forceinline__ void pack_strip(const uint8_t* __restrict__ xs, int64_t num_envs, int32_t n,
                                           int64_t env0, int32_t node0, uint32_t (&w)[VEC]) {
#pragma unroll
  for (int b = 0; b < VEC; ++b) w[b] = 0;
  if (node0 >= n) return;
  const int valid = (int)min((int64_t)kTileEnvs, num_envs - env0);
  const uint8_t* base = xs + env0 * (int64_t)n + node0;
  if (VEC == 4) {
    uint32_t v[kTileEnvs];
#pragma unroll
    for (int e = 0; e < kTileEnvs; ++e)
      v[e] = (e < valid) ? __ldg(reinterpret_cast<const uint32_t*>(base + (int64_t)e * n)) : 0u;
#pragma unroll
    for (int e = 0; e < kTileEnvs; ++e) {
      const uint32_t t = __vcmpne4(v[e], 0u);   // 0xff per non-zero byte
#pragma unroll
      for (int b = 0; b < VEC; ++b) w[b] |= ((t >> (8 * b)) & 1u) << e;
    }
  } else {
#pragma unroll
    for (int e = 0; e < kTileEnvs; ++e) {
      if (e < valid) {
#pragma unroll
        for (int b = 0; b < VEC; ++b)
          if (node0 + b < n) w[b] |= (uint32_t)(__ldg(base + (int64_t)e * n + b) != 0) << e;
      }
    }
  }
}

// whole tile into shared memory (sP has np words; padding nodes become 0)
template <int VEC>
__device__ __forceinline__ void pack_tile_to_smem(const uint8_t* __restrict__ xs, int64_t num_envs, int32_t n,
                                                  int32_t np, int64_t tile, uint32_t* sP) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int strips = (np + 32 * VEC - 1) / (32 * VEC);
  for (int s = warp; s < strips; s += nwarps) {
    const int node0 = s * 32 * VEC + lane * VEC;
    uint32_t w[VEC];
    pack_strip<VEC>(xs, num_envs, n, tile * kTileEnvs, node0, w);
#pragma unroll
    for (int b = 0; b < VEC; ++b)
      if (node0 + b < np) sP[node0 + b] = w[b];
  }
}

// ---- packed words -> bool rows ---------------------------------------------------
template <int VEC>
__device__ __forceinline__ void unpack_strip(const uint32_t (&w)[VEC], uint8_t* __restrict__ xs, int64_t num_envs,
                                             int32_t n, int64_t env0, int32_t node0) {
  if (node0 >= n) return;
  const int valid = (int)min((int64_t)kTileEnvs, num_envs - env0);
  uint8_t* base = xs + env0 * (int64_t)n + node0;
#pragma unroll
  for (int e = 0; e < kTileEnvs; ++e) {
    if (e < valid) {
      if (VEC == 4) {
        uint32_t v = 0;
#pragma unroll
        for (int b = 0; b < VEC; ++b) v |= ((w[b] >> e) & 1u) << (8 * b);
        *reinterpret_cast<uint32_t*>(base + (int64_t)e * n) = v;
      } else {
#pragma unroll
        for (int b = 0; b < VEC; ++b)
          if (node0 + b < n) base[(int64_t)e * n + b] = (uint8_t)((w[b] >> e) & 1u);
      }
    }
  }
}

template <int VEC>
__device__ __forceinline__ void unpack_tile_from_smem(const uint32_t* sP, uint8_t* __restrict__ xs,
                                                      int64_t num_envs, int32_t n, int32_t np, int64_t tile,
                                                      int nwarps = 0) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (nwarps == 0) nwarps = blockDim.x >> 5;   // nwarps: the warps that call this (all of the CTA by default)
  const int strips = (np + 32 * VEC - 1) / (32 * VEC);
  for (int s = warp; s < strips; s += nwarps) {
    const int node0 = s * 32 * VEC + lane * VEC;
    uint32_t w[VEC];
#pragma unroll
    for (int b = 0; b < VEC; ++b) w[b] = (node0 + b < np) ? sP[node0 + b] : 0u;
    unpack_strip<VEC>(w, xs, num_envs, n, tile * kTileEnvs, node0);
  }
}

// bool rows can be read/written as 4-byte words iff every row start is 4-aligned
__host__ __device__ inline bool rows_vec4_ok(const void* p, int32_t n) {
  return (n % 4 == 0) && ((reinterpret_cast<uintptr_t>(p) & 3u) == 0);
}

}  // namespace rlsb
