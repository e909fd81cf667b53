// Bit-sliced ("vertical") counters.  A packed word carries one bit per environment (bit b =
// env b of the tile), so counting per environment means adding 1-bit-per-env words into P
// bit planes: plane p holds bit p of all 32 per-env counts.  One LOP3 therefore works on 32
// environments at once and a lane can own a whole node (or edge) instead of a warp owning it.
// Carry-save adders (Harley-Seal blocks of 8) bring the cost to ~2.5 LOP3 per added word.
#pragma once
#include "common.cuh"

namespace rlsb {

// full adder on 32 bit positions at once: l = a ^ b ^ c, h = majority(a, b, c)  (2 LOP3)
__device__ __forceinline__ void csa(uint32_t& h, uint32_t& l, uint32_t a, uint32_t b, uint32_t c) {
  const uint32_t u = a ^ b;
  h = (a & b) | (u & c);
  l = u ^ c;
}

template <int P>
struct VCount {
  uint32_t c[P];

  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int p = 0; p < P; ++p) c[p] = 0;
  }

  // add eight 1-bit-per-env words (the running total must stay below 2^P)
  __device__ __forceinline__ void add8(uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t x4,
                                       uint32_t x5, uint32_t x6, uint32_t x7) {
    uint32_t ta, tb, fa, fb, e;
    csa(ta, c[0], c[0], x0, x1);
    csa(tb, c[0], c[0], x2, x3);
    csa(fa, c[1], c[1], ta, tb);
    csa(ta, c[0], c[0], x4, x5);
    csa(tb, c[0], c[0], x6, x7);
    csa(fb, c[1], c[1], ta, tb);
    csa(e, c[2], c[2], fa, fb);
#pragma unroll
    for (int p = 3; p < P; ++p) {   // ripple the weight-8 carry upwards
      const uint32_t t = c[p] & e;
      c[p] ^= e;
      e = t;
    }
  }

  // per-env total over the 32 lanes of the warp: lane e returns the count of env e
  __device__ __forceinline__ int flush_warp(int lane) const {
    int cnt = 0;
#pragma unroll
    for (int p = 0; p < P; ++p) cnt += __popc(transpose32(c[p], lane)) << p;
    return cnt;
  }

  // mask of envs whose count is <= h (h is a per-lane scalar)
  __device__ __forceinline__ uint32_t le(uint32_t h) const {
    uint32_t lt = 0, eq = kFull;
#pragma unroll
    for (int p = P - 1; p >= 0; --p) {
      const uint32_t hm = ((h >> p) & 1u) ? kFull : 0u;
      lt |= eq & ~c[p] & hm;
      eq &= ~(c[p] ^ hm);
    }
    return lt | eq;
  }

  // max / min of the per-env counts over the envs selected by `valid` (valid != 0)
  __device__ __forceinline__ uint32_t max_over(uint32_t valid) const {
    uint32_t m = valid, v = 0;
#pragma unroll
    for (int p = P - 1; p >= 0; --p) {
      const uint32_t t = m & c[p];
      if (t) m = t, v |= 1u << p;
    }
    return v;
  }
  __device__ __forceinline__ uint32_t min_over(uint32_t valid) const {
    uint32_t m = valid, v = 0;
#pragma unroll
    for (int p = P - 1; p >= 0; --p) {
      const uint32_t t = m & ~c[p];
      if (t) m = t; else v |= 1u << p;
    }
    return v;
  }

  // counts of envs 4q..4q+3 as four bytes (P <= 8)
  __device__ __forceinline__ uint32_t bytes4(int q) const {
    uint32_t out = 0;
#pragma unroll
    for (int p = 0; p < P; ++p)
      out += ((((c[p] >> (4 * q)) & 0xFu) * 0x00204081u) & 0x01010101u) << p;
    return out;
  }
  // counts of envs 2q, 2q+1 as two halfwords (P <= 16)
  __device__ __forceinline__ uint32_t halves2(int q) const {
    uint32_t out = 0;
#pragma unroll
    for (int p = 0; p < P; ++p)
      out += ((((c[p] >> (2 * q)) & 0x3u) * 0x8001u) & 0x00010001u) << p;
    return out;
  }
};

// Count, for every slot of one SELL slice, the neighbours whose word differs from the
// slot's own word -- per environment, bit-sliced.  `self` = the slot's word.  SMEM = the SELL
// arrays live in shared memory (staged copy) instead of global memory.
template <int P, bool SMEM>
__device__ __forceinline__ void sell_cross(const SellDev& s, int slice, int lane, const uint32_t* sP, uint32_t self,
                                           VCount<P>& vc) {
  const int gb = SMEM ? s.off[slice] : __ldg(s.off + slice);
  const int nb = (SMEM ? s.off[slice + 1] : __ldg(s.off + slice + 1)) - gb;
  const uint2* col = reinterpret_cast<const uint2*>(s.col) + (int64_t)gb * 32 + lane;   // 4 ids per lane per block
  auto ld = [&](int b) { return SMEM ? col[b * 32] : __ldg(col + b * 32); };
  vc.clear();
  int b = 0;
  for (; b + 4 <= nb; b += 4) {          // 16 neighbours: four id loads in flight, then eight word loads twice
    const uint2 i0 = ld(b), i1 = ld(b + 1), i2 = ld(b + 2), i3 = ld(b + 3);
    vc.add8(sP[i0.x & 0xffffu] ^ self, sP[i0.x >> 16] ^ self, sP[i0.y & 0xffffu] ^ self, sP[i0.y >> 16] ^ self,
            sP[i1.x & 0xffffu] ^ self, sP[i1.x >> 16] ^ self, sP[i1.y & 0xffffu] ^ self, sP[i1.y >> 16] ^ self);
    vc.add8(sP[i2.x & 0xffffu] ^ self, sP[i2.x >> 16] ^ self, sP[i2.y & 0xffffu] ^ self, sP[i2.y >> 16] ^ self,
            sP[i3.x & 0xffffu] ^ self, sP[i3.x >> 16] ^ self, sP[i3.y & 0xffffu] ^ self, sP[i3.y >> 16] ^ self);
  }
  for (; b < nb; b += 2) {
    const uint2 i0 = ld(b);
    uint2 i1 = make_uint2(0, 0);
    const bool two = b + 1 < nb;
    if (two) i1 = ld(b + 1);
    const uint32_t x4 = two ? sP[i1.x & 0xffffu] ^ self : 0u, x5 = two ? sP[i1.x >> 16] ^ self : 0u;
    const uint32_t x6 = two ? sP[i1.y & 0xffffu] ^ self : 0u, x7 = two ? sP[i1.y >> 16] ^ self : 0u;
    vc.add8(sP[i0.x & 0xffffu] ^ self, sP[i0.x >> 16] ^ self, sP[i0.y & 0xffffu] ^ self, sP[i0.y >> 16] ^ self, x4, x5,
            x6, x7);
  }
}

// Cut of one tile: threads stream the original edge list (u | v << 16 per edge, four edges per
// 16-byte load, coalesced), XOR the two endpoint words from shared memory and add the result to
// a vertical counter; one 32x32 bit transpose per plane at the end turns "bit = env" into
// "lane = env".  Returns this warp's partial count for env == lane (0 for idle warps).
// Only the first `warps` warps of the CTA take edges (>= 64 edges per lane keeps the
// flush cost small against the streaming loop).
__device__ __forceinline__ uint32_t pair_xor(const uint32_t* sP, uint32_t pr) {
  return sP[pr & 0xffffu] ^ sP[pr >> 16];
}

// `warp` = this warp's index among the `warps` that take edges (default: its index in the CTA).
// `pairs` / `quads`: a zero-padded list of u | v << 16 words, four per 16-byte element.
__device__ __forceinline__ int tile_cut_range(const uint4* __restrict__ pairs, int quads, const uint32_t* sP, int warps,
                                              int warp = -1) {
  const int lane = threadIdx.x & 31;
  if (warp < 0) warp = threadIdx.x >> 5;
  if (warp >= warps) return 0;
  const int T = warps * 32;
  int total = 0;
  VCount<8> vc;
  vc.clear();
  int blocks = 0;
  // software pipeline: the four 16-byte edge loads of the next trip are in flight while this trip's
  // 32 shared-memory gathers and adds run (the list comes from L2: ~1000 cycles under load)
  auto load_trip = [&](int q0, uint4(&e)[4]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int q = q0 + lane + j * T;
      e[j] = q < quads ? __ldg(pairs + q) : make_uint4(0, 0, 0, 0);
    }
  };
  uint4 e[4], en[4];
  if (warp * 32 < quads) load_trip(warp * 32, e);
  for (int q0 = warp * 32; q0 < quads; q0 += 4 * T) {   // warp-uniform trip count; 16 edges per lane per trip
    const bool more = q0 + 4 * T < quads;
    if (more) load_trip(q0 + 4 * T, en);
    vc.add8(pair_xor(sP, e[0].x), pair_xor(sP, e[0].y), pair_xor(sP, e[0].z), pair_xor(sP, e[0].w),
            pair_xor(sP, e[1].x), pair_xor(sP, e[1].y), pair_xor(sP, e[1].z), pair_xor(sP, e[1].w));
    vc.add8(pair_xor(sP, e[2].x), pair_xor(sP, e[2].y), pair_xor(sP, e[2].z), pair_xor(sP, e[2].w),
            pair_xor(sP, e[3].x), pair_xor(sP, e[3].y), pair_xor(sP, e[3].z), pair_xor(sP, e[3].w));
    blocks += 2;
    if (blocks == 30) {              // 240 < 2^8: flush before the planes overflow
      total += vc.flush_warp(lane);
      vc.clear();
      blocks = 0;
    }
    if (more) {
#pragma unroll
      for (int j = 0; j < 4; ++j) e[j] = en[j];
    }
  }
  total += vc.flush_warp(lane);
  return total;
}

__device__ __forceinline__ int tile_cut_partial(const GraphDev& g, const uint32_t* sP, int warps, int warp = -1) {
  return tile_cut_range(reinterpret_cast<const uint4*>(g.edge_pair), (g.m + 3) >> 2, sP, warps, warp);
}

// Weighted cut of one tile: sum over the (bit of |w|, sign) buckets of scale * (cut edges of the bucket) -- a
// weighted popcount of the XORed words.  +-1 weights are two buckets, i.e. two passes of the unweighted loop.
__device__ __forceinline__ long long tile_cut_weighted_partial(const GraphDev& g, const uint32_t* sP, int warps,
                                                               int warp = -1) {
  long long total = 0;
  const uint4* base = reinterpret_cast<const uint4*>(g.wpair);
  for (int k = 0; k < g.wbuckets; ++k) {
    const int4 meta = __ldg(g.wmeta + k);                 // {first quad, quads, scale, edges}
    total += (long long)meta.z * tile_cut_range(base + meta.x, meta.y, sP, warps, warp);
  }
  return total;
}

}  // namespace rlsb
