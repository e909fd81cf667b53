// The reference's samplers, each a Python loop of tiny torch kernels there, each ONE launch here:
//
//   mcpg_sweeps    sampler_func                rlsolver/methods/MCPG.py:120-166  (num_ls * N node updates)
//   metro          metro_sampling              rlsolver/methods/MCPG.py:88-117   (up to 5*max_transfer iterations)
//   subset         sub_set_sampling            rlsolver/methods/L2A/transformer.py:335-353 (top_k columns)
//
// Chains are bit-packed 32 per word like the environments of the other kernels (tile = 32 chains x
// all nodes in shared memory).  Random numbers are either read from explicit arrays (replay of
// recorded draws) or computed in place from torch's Philox stream (philox.cuh), which reproduces
// the reference's call sequence bit for bit.
#include <algorithm>
#include <numeric>
#include <vector>

#include "graph_host.h"
#include "philox.cuh"
#include "tile_ops.cuh"

// ---------------------------------------------------------------------------- plan (host side)
struct rlsb_mcpg_plan {
  int32_t n = 0, np = 0, levels = 0, num_slices = 0, device = -1, max_deg = 0;
  std::vector<int32_t> order, level_slice;
  rlsb_graph::Sell earlier, later;
  std::vector<uint16_t> deg, nlater, pos;
  void* dev_blob = nullptr;
  // scratch owned by the plan (allocated on first use, grown on demand, freed with the plan): the tie-break words of
  // one rlsb_mcpg_sweeps call and the node degrees in visiting order
  mutable uint32_t* coin = nullptr;
  mutable size_t coin_words = 0;
  mutable uint16_t* degpos = nullptr;
  const rlsb_graph_t* graph = nullptr;
  struct Dev {
    int32_t levels, num_slices;
    const int32_t* level_slice;
    rlsb::SellDev earlier, later;     // node arrays shared
    const uint16_t *deg, *nlater, *pos;
  } dev{};
};

namespace rlsb {

// ---------------------------------------------------------------------------- bit-sliced helpers
// z += a << shift (z has Q planes, a has P planes)
template <int Q, int P, int SHIFT>
__device__ __forceinline__ void planes_add_shifted(uint32_t (&z)[Q], const VCount<P>& a) {
  uint32_t carry = 0;
#pragma unroll
  for (int p = 0; p < P; ++p) {
    const uint32_t zz = z[p + SHIFT], aa = a.c[p];
    const uint32_t u = zz ^ aa;
    z[p + SHIFT] = u ^ carry;
    carry = (zz & aa) | (u & carry);
  }
#pragma unroll
  for (int p = 0; p < Q; ++p) {
    if (p >= P + SHIFT) {
      const uint32_t zz = z[p];
      z[p] = zz ^ carry;
      carry &= zz;
    }
  }
}

// masks of bit positions whose Q-plane value is < k / == k
template <int Q>
__device__ __forceinline__ void planes_cmp(const uint32_t (&z)[Q], uint32_t k, uint32_t& lt, uint32_t& eq) {
  lt = 0, eq = kFull;
#pragma unroll
  for (int p = Q - 1; p >= 0; --p) {
    const uint32_t km = ((k >> p) & 1u) ? kFull : 0u;
    lt |= eq & ~z[p] & km;
    eq &= ~(z[p] ^ km);
  }
}

// Tie-breaks of sampler_func.  A node is set iff S + rand/4 < (deg + 1/4)/2 (MCPG.py:140-142): the random number only
// matters when S == deg/2 exactly, and then the outcome depends on nothing but (deg, rand).  Inside the level loop a
// lane met its ties one after the other -- a dozen dependent Philox blocks on the critical path of every level, two
// thirds of the kernel's stall samples.  Here the coin of every (sweep, node, chain) that can tie is decided by the
// whole GPU beforehand, in float32 exactly as the reference adds it, from the uniform torch's call number
// sweep*N + position returns; the sweep kernel ANDs its tie mask with the word.  coin[call][tile] bit c <-> chain
// 32*tile + c.  After the first sweep S is an integer, so odd degrees cannot tie (no Philox for them); in the first
// sweep the unvisited entries still hold {-0.5, 1.5} and every degree can.   grid = (tile groups, position groups, sweep).
constexpr int kCoinPos = 32;     // node positions per block of mcpg_coins_kernel
__global__ void __launch_bounds__(256) mcpg_coins_kernel(int n, int64_t num_chains, int64_t tiles,
                                                         const uint16_t* __restrict__ degpos,
                                                         const float* __restrict__ explicit_u, TorchRng rng,
                                                         uint32_t* __restrict__ coin) {
  const int lane = threadIdx.x & 31;
  const int64_t tile = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (tile >= tiles) return;
  const uint32_t sweep = blockIdx.z;
  const int64_t chain = tile * kTileEnvs + lane;
  const bool live = chain < num_chains;
  const bool direct = num_chains <= (int64_t)rng.threads;
  const uint2 pkey = make_uint2((uint32_t)rng.seed, (uint32_t)(rng.seed >> 32));
  for (uint32_t pos = blockIdx.y * kCoinPos; pos < min((uint32_t)n, (blockIdx.y + 1) * kCoinPos); ++pos) {
    const uint64_t call = (uint64_t)sweep * n + pos;
    const uint32_t deg = __ldg(degpos + pos);
    uint32_t word = 0;
    if (sweep == 0 || !(deg & 1u)) {                       // warp-uniform
      bool bit = false;
      if (live) {
        float u;
        if (explicit_u) {
          u = __ldg(explicit_u + call * num_chains + chain);
        } else if (direct) {                               // element `chain` of a call = output 0 of (its counter, subsequence chain)
          const uint64_t ctr = rng.offset4 + call * rng.iters_per_call;
          u = torch_uniform_from_u32(
              curand_Philox4x32_10(make_uint4((uint32_t)ctr, (uint32_t)(ctr >> 32), (uint32_t)chain, 0u), pkey).x);
        } else {
          u = torch_uniform_from_u32(torch_philox_u32(rng, call, (uint32_t)chain));
        }
        const float sf = 0.5f * (float)deg, tf = sf + 0.125f;
        bit = __fadd_rn(sf, __fmul_rn(u, 0.25f)) < tf;
      }
      word = __ballot_sync(kFull, bit);
    }
    if (lane == 0) coin[call * tiles + tile] = word;
  }
}

// Ones among a slot's neighbours with the first kPlanPre id blocks already in registers (ids of short rows point at
// the zero word `pad`): the loads left on the critical path are the state words themselves.
constexpr int kPlanPre = 6;
template <int P>
__device__ __forceinline__ void cross_prefetched(const uint2 (&id)[kPlanPre], int nb, const uint2* __restrict__ col,
                                                 const uint32_t* sP, uint32_t pad, VCount<P>& vc) {
  vc.clear();
  auto add_pair = [&](uint2 i0, uint2 i1) {
    vc.add8(sP[i0.x & 0xffffu], sP[i0.x >> 16], sP[i0.y & 0xffffu], sP[i0.y >> 16], sP[i1.x & 0xffffu], sP[i1.x >> 16],
            sP[i1.y & 0xffffu], sP[i1.y >> 16]);
  };
  const uint2 none = make_uint2(pad | (pad << 16), pad | (pad << 16));
#pragma unroll
  for (int b = 0; b < kPlanPre; b += 2)
    if (b < nb) add_pair(id[b], b + 1 < nb ? id[b + 1] : none);
  for (int b = kPlanPre; b < nb; b += 2) add_pair(__ldg(col + b * 32), b + 1 < nb ? __ldg(col + (b + 1) * 32) : none);
}

// ---------------------------------------------------------------------------- sampler_func sweeps
constexpr int kMcpgThreads = 128;

// xs: float32 [N][C] node-major (values 0/1), in place.  expected: float32 [C].
// COINS: the tie-breaks come from mcpg_coins_kernel (few tiles: a level is latency bound and its ties would sit on
// the critical path); otherwise a lane computes the coins of its ties in place (many tiles: other warps hide them,
// and only ~9 % of the coins are ever needed).
template <int P, bool COINS>
__global__ void __launch_bounds__(kMcpgThreads) mcpg_sweeps_kernel(GraphDev g, rlsb_mcpg_plan::Dev plan,
                                                                   float* __restrict__ xs, int64_t num_chains,
                                                                   int num_ls, const uint32_t* __restrict__ coin,
                                                                   const float* __restrict__ explicit_u, TorchRng rng,
                                                                   float* __restrict__ expected, int cut_warps) {
  extern __shared__ uint32_t sP[];               // np + 32 words; [np] stays zero (padding target)
  __shared__ int sCnt[kTileEnvs];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int64_t tiles = (num_chains + kTileEnvs - 1) / kTileEnvs;
  // The small static arrays of the plan live in shared memory for the whole kernel, so that fetching a level ahead
  // is ONE round trip to L2 (neighbour ids + coin word) instead of a chain of three (slice range -> offsets -> ids).
  const bool direct = num_chains <= (int64_t)rng.threads;
  const uint2 pkey = make_uint2((uint32_t)rng.seed, (uint32_t)(rng.seed >> 32));
  const int S = plan.num_slices;
  int32_t* sLvs = reinterpret_cast<int32_t*>(sP + g.np + 32);
  int32_t* sOffE = sLvs + plan.levels + 1;
  int32_t* sOffL = sOffE + S + 1;
  uint16_t* sNode = reinterpret_cast<uint16_t*>(sOffL + S + 1);
  uint16_t* sDeg = sNode + 32 * S;
  uint16_t* sNl = sDeg + 32 * S;
  uint16_t* sPos = sNl + 32 * S;
  for (int i = threadIdx.x; i <= plan.levels; i += blockDim.x) sLvs[i] = __ldg(plan.level_slice + i);
  for (int i = threadIdx.x; i <= S; i += blockDim.x) sOffE[i] = __ldg(plan.earlier.off + i), sOffL[i] = __ldg(plan.later.off + i);
  for (int i = threadIdx.x; i < 32 * S; i += blockDim.x) {
    sNode[i] = __ldg(plan.earlier.node + i), sDeg[i] = __ldg(plan.deg + i);
    sNl[i] = __ldg(plan.nlater + i), sPos[i] = __ldg(plan.pos + i);
  }
  __syncthreads();
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t c0 = tile * kTileEnvs;
    const int valid = (int)min((int64_t)kTileEnvs, num_chains - c0);
    const uint32_t vmask = valid == 32 ? kFull : ((1u << valid) - 1u);
    if (threadIdx.x < kTileEnvs) sCnt[threadIdx.x] = 0;
    // pack: one coalesced 128-byte row segment per node
    for (int i0 = warp * 16; i0 < g.np + 32; i0 += nwarps * 16) {      // 16 row loads in flight, then their ballots
      float v[16];
#pragma unroll
      for (int k = 0; k < 16; ++k)
        v[k] = (i0 + k < g.n && lane < valid) ? __ldg(xs + (int64_t)(i0 + k) * num_chains + c0 + lane) : 0.f;
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const uint32_t w = __ballot_sync(kFull, v[k] != 0.f);
        if (lane == 0 && i0 + k < g.np + 32) sP[i0 + k] = w;
      }
    }
    __syncthreads();
    // Everything a level needs that does not depend on the state -- its slice range, the slot's node / degree /
    // position and the first neighbour-id blocks of both lists -- is fetched one level ahead (the plan lives in
    // L2: four dependent round trips per level otherwise, and a level is only a handful of warps wide).
    struct Pre {
      int sb, se, nbe, nbl;
      uint32_t node, deg, nlater, coin;
      const uint2 *cole, *coll;
      uint2 ide[kPlanPre], idl[kPlanPre];
    };
    auto coin_of = [&](int sweep, int slot) {           // tie-break word of the slot's node in this sweep, this tile
      return COINS ? __ldg(coin + ((int64_t)sweep * g.n + sPos[slot]) * tiles + tile) : 0u;
    };
    auto fetch = [&](int l, int sweep, Pre& p) {
      p.sb = sLvs[l], p.se = sLvs[l + 1];
      const int s = p.sb + warp;
      p.nbe = p.nbl = 0, p.node = 0xFFFFu, p.deg = p.nlater = p.coin = 0u, p.cole = p.coll = nullptr;
      if (s < p.se) {
        p.node = sNode[s * 32 + lane];
        p.deg = sDeg[s * 32 + lane];
        p.nlater = sNl[s * 32 + lane];
        p.coin = p.node != 0xFFFFu ? coin_of(sweep, s * 32 + lane) : 0u;
        const int ge = sOffE[s], gl = sOffL[s];
        p.nbe = sOffE[s + 1] - ge, p.nbl = sOffL[s + 1] - gl;
        p.cole = reinterpret_cast<const uint2*>(plan.earlier.col) + (int64_t)ge * 32 + lane;
        p.coll = reinterpret_cast<const uint2*>(plan.later.col) + (int64_t)gl * 32 + lane;
#pragma unroll
        for (int b = 0; b < kPlanPre; ++b) {
          p.ide[b] = b < p.nbe ? __ldg(p.cole + b * 32) : make_uint2(0u, 0u);
          p.idl[b] = b < p.nbl ? __ldg(p.coll + b * 32) : make_uint2(0u, 0u);
        }
      }
    };
    Pre cur;
    if (num_ls > 0 && plan.levels > 0) fetch(0, 0, cur);
    for (int sweep = 0; sweep < num_ls; ++sweep) {
      for (int l = 0; l < plan.levels; ++l) {
        Pre nxt;
        const int ln = l + 1 < plan.levels ? l + 1 : 0;
        if (l + 1 < plan.levels || sweep + 1 < num_ls) fetch(ln, l + 1 < plan.levels ? sweep : sweep + 1, nxt);
        for (int s = cur.sb + warp; s < cur.se; s += nwarps) {
          const bool first = s == cur.sb + warp;
          const uint32_t node = first ? cur.node : (uint32_t)sNode[s * 32 + lane];
          const bool active = node != 0xFFFFu;
          VCount<P> a, b;
          if (first) {
            cross_prefetched<P>(cur.ide, cur.nbe, cur.cole, sP, (uint32_t)g.np, a);   // ones among already-visited neighbours
            cross_prefetched<P>(cur.idl, cur.nbl, cur.coll, sP, (uint32_t)g.np, b);   // ones among not-yet-visited neighbours
          } else {
            sell_cross<P, false>(plan.earlier, s, lane, sP, 0u, a);
            sell_cross<P, false>(plan.later, s, lane, sP, 0u, b);
          }
          const uint32_t deg = first ? cur.deg : (uint32_t)sDeg[s * 32 + lane];
          // 2*S = 2a + 4b - later  (first sweep: unvisited entries still hold {-0.5, 1.5}, MCPG.py:132-133)
          //     = 2a + 2b          (afterwards);   set iff S + rand/4 < (deg + 1/4) / 2
          uint32_t z[P + 3];
#pragma unroll
          for (int p = 0; p < P + 3; ++p) z[p] = 0;
          planes_add_shifted<P + 3, P, 1>(z, a);
          if (sweep == 0) planes_add_shifted<P + 3, P, 2>(z, b);
          else planes_add_shifted<P + 3, P, 1>(z, b);
          const uint32_t nlater = first ? cur.nlater : (uint32_t)sNl[s * 32 + lane];
          const uint32_t k = deg + (sweep == 0 ? nlater : 0u);
          uint32_t lt, eq;
          planes_cmp<P + 3>(z, k, lt, eq);
          // S == deg/2: the coin decides, in float32 exactly as the reference adds it
          uint32_t word = lt;
          if (COINS) {
            word |= eq & (first ? cur.coin : (active ? coin_of(sweep, s * 32 + lane) : 0u));
          } else {
            uint32_t ties = active ? (eq & vmask) : 0u;
            if (ties) {
              const float sf = 0.5f * (float)deg, tf = sf + 0.125f;
              const uint64_t call = (uint64_t)sweep * g.n + sPos[s * 32 + lane];
              while (ties) {
                const int bpos = __ffs(ties) - 1;
                ties &= ties - 1;
                const uint32_t chain = (uint32_t)(c0 + bpos);
                float u;
                if (explicit_u) {
                  u = __ldg(explicit_u + call * num_chains + chain);
                } else if (direct) {    // C <= T: element `chain` of a call is output 0 of (counter of the call, subsequence chain)
                  const uint64_t ctr = rng.offset4 + call * rng.iters_per_call;
                  u = torch_uniform_from_u32(
                      curand_Philox4x32_10(make_uint4((uint32_t)ctr, (uint32_t)(ctr >> 32), chain, 0u), pkey).x);
                } else {
                  u = torch_uniform_from_u32(torch_philox_u32(rng, call, chain));
                }
                if (__fadd_rn(sf, __fmul_rn(u, 0.25f)) < tf) word |= 1u << bpos;
              }
            }
          }
          if (active) sP[node] = word & vmask;
        }
        __syncthreads();
        cur = nxt;
      }
    }
    // expected_cut = sum_e (2x_u - 1)(2x_v - 1) = M - 2 * cut   (MCPG.py:148-154)
    const int cnt = tile_cut_partial(g, sP, cut_warps);
    if (cnt) atomicAdd(&sCnt[lane], cnt);
    __syncthreads();
    if (threadIdx.x < valid) expected[c0 + threadIdx.x] = (float)(g.m - 2 * sCnt[threadIdx.x]);
    for (int i = warp; i < g.n; i += nwarps) {
      const uint32_t w = sP[i];
      if (lane < valid) xs[(int64_t)i * num_chains + c0 + lane] = (float)((w >> lane) & 1u);
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------- metro_sampling
// One warp per tile of 32 chains, lane = chain.  count_only: accumulate the number of accepted
// moves of every iteration into acc[t] (the reference's global stop rule needs them);
// otherwise run *num_iters_dev iterations and write the final state.
constexpr int kMetroWarps = 4;

__global__ void __launch_bounds__(kMetroWarps * 32) metro_kernel(int n, int np, const float* __restrict__ probs,
                                                                  const float* __restrict__ start, float* __restrict__ out,
                                                                  int64_t num_chains, int max_iters,
                                                                  const int32_t* __restrict__ num_iters_dev,
                                                                  const int64_t* __restrict__ explicit_idx,
                                                                  const float* __restrict__ explicit_u, TorchRng rng,
                                                                  int32_t* __restrict__ acc, int count_only) {
  extern __shared__ uint32_t smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t* sP = smem + (size_t)warp * np;
  const int64_t tiles = (num_chains + kTileEnvs - 1) / kTileEnvs;
  const int iters = count_only ? max_iters : min(max_iters, *num_iters_dev);
  for (int64_t tile = (int64_t)blockIdx.x * kMetroWarps + warp; tile < tiles; tile += (int64_t)gridDim.x * kMetroWarps) {
    const int64_t c0 = tile * kTileEnvs;
    const bool live = c0 + lane < num_chains;
    const uint32_t chain = (uint32_t)(c0 + lane);
    for (int i = 0; i < n; ++i) {
      const float v = live ? start[(int64_t)i * num_chains + chain] : 0.f;
      const uint32_t w = __ballot_sync(kFull, v != 0.f);       // start_status.bool()
      if (lane == 0) sP[i] = w;
    }
    __syncwarp();
    for (int t = 0; t < iters; ++t) {
      uint32_t r;
      float u;
      if (explicit_idx) {
        r = live ? (uint32_t)explicit_idx[(int64_t)t * num_chains + chain] : 0u;
        u = live ? explicit_u[(int64_t)t * num_chains + chain] : 2.f;
      } else {
        r = torch_philox_u32(rng, 2 * (uint64_t)t, chain) % (uint32_t)n;           // torch.randint(0, N, [C])
        u = torch_uniform_from_u32(torch_philox_u32(rng, 2 * (uint64_t)t + 1, chain));   // torch.rand(C)
      }
      const bool bit = (sP[r] >> lane) & 1u;
      const float p = __ldg(probs + r);
      const float q = bit ? p : __fsub_rn(1.f, p);
      const float rate = __fdiv_rn(__fsub_rn(1.f, q), q);       // (1 - q) / q, IEEE division like torch
      const bool accept = live && (u < rate);
      __syncwarp();
      if (accept) atomicXor(&sP[r], 1u << lane);
      __syncwarp();
      if (count_only) {
        const int k = __popc(__ballot_sync(kFull, accept));
        if (lane == 0 && k) atomicAdd(acc + t, k);
      }
    }
    if (!count_only) {
      for (int i = 0; i < n; ++i)
        if (live) out[(int64_t)i * num_chains + chain] = (float)((sP[i] >> lane) & 1u);
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------- metro_sampling, split form
// metro_kernel above is a chain of up to 5*max_transfer dependent iterations per warp with two Philox blocks
// inside every one of them, run twice (count, then apply): 1.8 ms for 4096 chains x 1000 iterations, almost
// all of it latency.  What is sequential is only the state lookup; the draws are not.  The split form
//   1. metro_draws_kernel  every (iteration, chain) in parallel: the randint node and the uniform torch's
//                          calls 2t and 2t+1 would return                           (Philox bound, one pass)
//   2. metro_rates_kernel  (1-q)/q for both values of the bit, per node, with the float ops of MCPG.py:107-109
//   0. metro_pack_kernel   float32 [N][C] -> lane-private packed tiles (and metro_unpack_kernel back at the end)
//   3. metro_pass_kernel   the chain, once, over ALL iterations: state word lookup, rate lookup, compare,
//                          flip; logs the accept ballot of every (iteration, tile) and counts per iteration
//   4. metro_scan_kernel   the reference's stop rule (MCPG.py:101-103): number of executed iterations
//   5. metro_undo_kernel   flips commute: state(T*) = final state ^ the accepted flips of iterations >= T*,
//                          applied in parallel; writes the float32 [N, C] result
constexpr int kMetroPF = 8;    // iterations of draws in flight per lane

// draws of (iteration t, chain c) as one 8-byte record {node, bits of the uniform}; grid = (chain blocks, iterations).
// With C <= T (every call of the reference fits one round of torch's grid: always below 303104 chains) element c of a
// call is output 0 of the Philox block (counter of the call, subsequence c): no index arithmetic at all.
__global__ void __launch_bounds__(256) metro_draws_kernel(TorchRng rng, int iters, int64_t num_chains, uint32_t n,
                                                          uint2* __restrict__ draws) {
  const int64_t chain = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (chain >= num_chains) return;
  const bool direct = num_chains <= (int64_t)rng.threads;
  const uint2 key = make_uint2((uint32_t)rng.seed, (uint32_t)(rng.seed >> 32));
  for (int t = blockIdx.y; t < iters; t += gridDim.y) {
    uint32_t x0, x1;
    if (direct) {
      const uint64_t c0 = rng.offset4 + 2 * (uint64_t)t * rng.iters_per_call, c1 = c0 + rng.iters_per_call;
      x0 = curand_Philox4x32_10(make_uint4((uint32_t)c0, (uint32_t)(c0 >> 32), (uint32_t)chain, 0u), key).x;
      x1 = curand_Philox4x32_10(make_uint4((uint32_t)c1, (uint32_t)(c1 >> 32), (uint32_t)chain, 0u), key).x;
    } else {
      x0 = torch_philox_u32(rng, 2 * (uint64_t)t, (uint32_t)chain);
      x1 = torch_philox_u32(rng, 2 * (uint64_t)t + 1, (uint32_t)chain);
    }
    // torch.randint(0, N, [C]) = x % N, torch.rand(C) = curand_uniform with the (0,1] -> [0,1) reversal
    draws[(int64_t)t * num_chains + chain] = make_uint2(x0 % n, __float_as_uint(torch_uniform_from_u32(x1)));
  }
}

__global__ void metro_rates_kernel(int n, const float* __restrict__ probs, float* __restrict__ rates) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float p = __ldg(probs + i);
#pragma unroll
  for (int bit = 0; bit < 2; ++bit) {
    const float q = bit ? p : __fsub_rn(1.f, p);
    rates[2 * i + bit] = __fdiv_rn(__fsub_rn(1.f, q), q);       // (1 - q) / q, IEEE division like torch
  }
}

// float32 [N][C] (node-major, the reference's layout) <-> lane-private packed tiles [tile][N/32][32 lanes], by the
// whole GPU: thread (tile, b, lane) converts nodes 32b .. 32b+31 of chain 32*tile + lane; every warp access is one
// coalesced 128-byte row segment.  (The chain kernel itself runs on 4 warps per SM: converting there cost 350 us.)
__global__ void __launch_bounds__(256) metro_pack_kernel(int n, int np, const float* __restrict__ start, int64_t num_chains,
                                                         int64_t tiles, uint32_t* __restrict__ packed) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = (int)(gid & 31);
  const int nb = np >> 5;
  const int64_t tile = (gid >> 5) / nb;
  const int b = (int)((gid >> 5) - tile * nb);
  if (tile >= tiles) return;
  const int64_t chain = tile * kTileEnvs + lane;
  const bool live = chain < num_chains;
  uint32_t mine = 0;
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    const int i = 32 * b + k;
    const float v = (live && i < n) ? __ldg(start + (int64_t)i * num_chains + chain) : 0.f;
    mine |= (v != 0.f ? 1u : 0u) << k;                           // start_status.bool()
  }
  packed[gid] = mine;
}

__global__ void __launch_bounds__(256) metro_unpack_kernel(int n, int np, const uint32_t* __restrict__ packed,
                                                           int64_t num_chains, int64_t tiles, float* __restrict__ out) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = (int)(gid & 31);
  const int nb = np >> 5;
  const int64_t tile = (gid >> 5) / nb;
  const int b = (int)((gid >> 5) - tile * nb);
  if (tile >= tiles) return;
  const int64_t chain = tile * kTileEnvs + lane;
  if (chain >= num_chains) return;
  const uint32_t w = __ldg(packed + gid);
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    const int i = 32 * b + k;
    if (i < n) out[(int64_t)i * num_chains + chain] = (float)((w >> k) & 1u);
  }
}

// State of a tile in LANE-PRIVATE form: word [b][lane] holds nodes 32b .. 32b+31 of chain `lane`, so a chain
// reads and flips its bits with plain shared-memory loads / stores (no atomics, no bank conflicts, no ordering
// between lanes).  The chain loop is branch-free and padded to groups of 32 iterations; the accept ballots of
// a group are handed out one per lane and leave as one store + one counter add per lane and group.
__global__ void __launch_bounds__(kMetroWarps * 32) metro_pass_kernel(int n, int np, int64_t num_chains, int iters,
                                                                       const uint2* __restrict__ draws,
                                                                       const float* __restrict__ rates,
                                                                       uint32_t* __restrict__ bits, int32_t* __restrict__ acc,
                                                                       uint32_t* __restrict__ fin) {
  extern __shared__ uint32_t smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t* sL = smem + (size_t)warp * np;
  float* sRates = reinterpret_cast<float*>(smem + (size_t)kMetroWarps * np);      // [N][2], shared by the warps
  for (int i = threadIdx.x; i < 2 * n; i += kMetroWarps * 32) sRates[i] = __ldg(rates + i);
  __syncthreads();
  const int64_t tiles = (num_chains + kTileEnvs - 1) / kTileEnvs;
  for (int64_t tile = (int64_t)blockIdx.x * kMetroWarps + warp; tile < tiles; tile += (int64_t)gridDim.x * kMetroWarps) {
    const int64_t c0 = tile * kTileEnvs;
    const bool live = c0 + lane < num_chains;
    const int64_t chain = c0 + lane;
    for (int b = 0; b < np / 32; ++b) sL[b * 32 + lane] = fin[tile * np + b * 32 + lane];      // packed by metro_pack_kernel
    const uint2* dp = draws + chain;
    uint2 ring[kMetroPF];
#pragma unroll
    for (int k = 0; k < kMetroPF; ++k) {
      ring[k] = make_uint2(0u, 0x7F800000u);                    // u = +inf: never accepted (idle chain / past the end)
      if (live && k < iters) ring[k] = __ldg(dp + (int64_t)k * num_chains);
    }
    const uint2* dnext = dp + (int64_t)kMetroPF * num_chains;
    const int padded = (iters + 31) / 32 * 32;
    for (int t0 = 0; t0 < padded; t0 += 32) {
      uint32_t myword = 0;
#pragma unroll
      for (int g = 0; g < 32 / kMetroPF; ++g) {
#pragma unroll
        for (int k = 0; k < kMetroPF; ++k) {
          const int t = t0 + g * kMetroPF + k;
          const uint2 d = ring[k];
          ring[k] = make_uint2(0u, 0x7F800000u);
          if (live && t + kMetroPF < iters) ring[k] = __ldg(dnext);
          dnext += num_chains;
          const uint32_t r = d.x;
          uint32_t* wp = sL + (r >> 5) * 32 + lane;
          const uint32_t w = *wp, sh = r & 31u;
          const bool accept = __uint_as_float(d.y) < sRates[2 * r + ((w >> sh) & 1u)];
          if (accept) *wp = w ^ (1u << sh);
          const uint32_t word = __ballot_sync(kFull, accept);
          if (lane == g * kMetroPF + k) myword = word;
        }
      }
      const int t = t0 + lane;
      if (t < iters) {
        bits[(int64_t)t * tiles + tile] = myword;
        if (myword) atomicAdd(acc + t, __popc(myword));
      }
    }
    for (int b = 0; b < np / 32; ++b) fin[tile * np + b * 32 + lane] = sL[b * 32 + lane];
  }
}

// num_iters = #{t : accepted moves before iteration t < thresh}  (the reference tests the count BEFORE every iteration)
__global__ void metro_scan_kernel(const int32_t* __restrict__ acc, int iters, int64_t thresh, int32_t* __restrict__ num_iters) {
  const int lane = threadIdx.x;
  int64_t before = 0;
  int count = 0;
  for (int t0 = 0; t0 < iters; t0 += 32) {
    const int t = t0 + lane;
    int64_t v = t < iters ? (int64_t)acc[t] : 0;
    int64_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int64_t y = __shfl_up_sync(kFull, incl, o);
      if (lane >= o) incl += y;
    }
    const int64_t mine = before + incl - v;                     // accepted moves before iteration t
    count += __popc(__ballot_sync(kFull, t < iters && mine < thresh));
    before += __shfl_sync(kFull, incl, 31);
  }
  if (lane == 0) *num_iters = count;
}

// flips commute and every (iteration, chain) pair touches its own bit: the moves of the iterations that were not
// executed are taken back by the whole GPU, one thread per pair, with one atomic XOR on the packed tile in L2
__global__ void __launch_bounds__(256) metro_undo_kernel(int np, int64_t num_chains, int iters, int64_t tiles,
                                                         const int32_t* __restrict__ num_iters_dev,
                                                         const uint2* __restrict__ draws, const uint32_t* __restrict__ bits,
                                                         uint32_t* __restrict__ fin) {
  const int first = min(iters, *num_iters_dev);
  const int64_t total = (int64_t)(iters - first) * num_chains;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = (int64_t)first * num_chains + k;
    const int64_t t = e / num_chains, chain = e - t * num_chains;
    const int64_t tile = chain >> 5;
    const int lane = (int)(chain & 31);
    if ((__ldg(bits + t * tiles + tile) >> lane) & 1u) {
      const uint32_t r = __ldg(draws + e).x;
      atomicXor(fin + tile * np + (r >> 5) * 32 + lane, 1u << (r & 31u));
    }
  }
}

struct MetroWs {
  uint2* draws;
  uint32_t* bits;
  int32_t* acc;
  uint32_t* fin;
  float* rates;
  size_t bytes;
};
static MetroWs metro_carve(int n, int64_t num_chains, int iters, void* base) {
  const int np = (n + 31) / 32 * 32;
  const int64_t tiles = (num_chains + kTileEnvs - 1) / kTileEnvs;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* p = base ? static_cast<char*>(base) + off : nullptr;
    off += (bytes + 255) / 256 * 256;
    return p;
  };
  MetroWs w;
  w.draws = reinterpret_cast<uint2*>(take((size_t)iters * num_chains * 8));
  w.bits = reinterpret_cast<uint32_t*>(take((size_t)iters * tiles * 4));
  w.acc = reinterpret_cast<int32_t*>(take((size_t)iters * 4));
  w.fin = reinterpret_cast<uint32_t*>(take((size_t)tiles * np * 4));
  w.rates = reinterpret_cast<float*>(take((size_t)2 * n * 4));
  w.bytes = off + 256;
  return w;
}

// ---------------------------------------------------------------------------- sub_set_sampling
// xs[row][ids[row % S][k]] = rand_k[row] < vals[row % S][k]   for every (row, k); ids of one row are distinct
__global__ void subset_kernel(uint8_t* __restrict__ xs, int64_t rows, int n, int64_t num_sims, int top_k,
                              const int64_t* __restrict__ ids, const float* __restrict__ vals,
                              const float* __restrict__ explicit_u, TorchRng rng) {
  const int64_t task = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (task >= rows * top_k) return;
  const int64_t k = task / rows, row = task - k * rows;          // consecutive threads -> consecutive rows
  const int64_t sim = row % num_sims;
  const float u = explicit_u ? explicit_u[k * rows + row]
                             : torch_uniform_from_u32(torch_philox_u32(rng, (uint64_t)k, (uint32_t)row));
  xs[row * n + ids[sim * top_k + k]] = (uint8_t)(u < vals[sim * top_k + k]);
}

// ---------------------------------------------------------------------------- RNG test hooks
__global__ void torch_rand_kernel(TorchRng rng, int64_t calls, int64_t numel, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= calls * numel) return;
  out[i] = torch_uniform_from_u32(torch_philox_u32(rng, (uint64_t)(i / numel), (uint32_t)(i % numel)));
}
__global__ void torch_randint_kernel(TorchRng rng, int64_t calls, int64_t numel, uint32_t range,
                                     int64_t* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= calls * numel) return;
  out[i] = (int64_t)(torch_philox_u32(rng, (uint64_t)(i / numel), (uint32_t)(i % numel)) % range);
}

// ---- row-major independent-site Metropolis (metropolis_hastings_sampling_TNCO, rlsolver/envs/env_L2A.py:233-276).
// One round of the reference: ids = randperm(dim); for i in range(dim): column ids[i] of EVERY row proposes a flip,
// accepted when rand(num)[row] < (1 - q) / q with q = p if the bit is set else 1 - p; the accepted flips are counted and
// the loop stops after the first column at which the running count reaches the target.  Within a round every column is
// visited once and a decision reads nothing but its own bit, so all dim x num decisions are independent: (1) count the
// accepts per position with the uniform of torch call number i, (2) one CTA scans the counts for the stop position,
// (3) apply the flips of the positions up to it.  (1) and (3) recompute the same Philox numbers.
__device__ __forceinline__ bool mh_accept(const TorchRng& rng, const float* __restrict__ probs, const uint8_t* __restrict__ xs,
                                          int64_t row, int64_t src_row, int dim, int col, uint32_t call) {
  const float p0 = __ldg(probs + src_row * dim + col);
  const float q = xs[row * dim + col] ? p0 : __fsub_rn(1.f, p0);
  const float rate = __fdiv_rn(__fsub_rn(1.f, q), q);
  return torch_uniform_from_u32(torch_philox_u32(rng, call, (uint32_t)row)) < rate;
}

// grid = (row blocks, positions); counts[i] += accepts of position i
__global__ void __launch_bounds__(256) mh_rows_count_kernel(TorchRng rng, const float* __restrict__ probs,
                                                            const uint8_t* __restrict__ xs, const int64_t* __restrict__ perm,
                                                            int64_t num, int64_t num_src, int dim, int32_t* __restrict__ counts) {
  const int64_t row = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int i = blockIdx.y;
  bool acc = false;
  if (row < num) acc = mh_accept(rng, probs, xs, row, row % num_src, dim, (int)perm[i], (uint32_t)i);
  const int c = __popc(__ballot_sync(kFull, acc));
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(counts + i, c);
}

// state[0] = running accept count (int64, in/out), state[1] = positions of this round that were visited (out)
__global__ void __launch_bounds__(1024) mh_rows_stop_kernel(const int32_t* __restrict__ counts, int dim, int64_t target,
                                                            int64_t* __restrict__ state) {
  __shared__ int64_t sSum[32];
  __shared__ int64_t sBase;
  __shared__ int sStop;
  if (threadIdx.x == 0) sBase = state[0], sStop = dim;      // dim = never reached: every position is visited
  __syncthreads();
  for (int base = 0; base < dim && sStop == dim; base += 1024) {
    const int i = base + threadIdx.x;
    int64_t v = i < dim ? counts[i] : 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int64_t o = __shfl_up_sync(kFull, v, off);
      if (lane >= off) v += o;
    }
    if (lane == 31) sSum[warp] = v;
    __syncthreads();
    if (warp == 0) {
      int64_t w = sSum[lane];
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const int64_t o = __shfl_up_sync(kFull, w, off);
        if (lane >= off) w += o;
      }
      sSum[lane] = w;
    }
    __syncthreads();
    const int64_t incl = sBase + v + (warp ? sSum[warp - 1] : 0);
    if (i < dim && incl >= target) atomicMin(&sStop, i);
    __syncthreads();
    if (threadIdx.x == 1023) sBase = incl;                   // only used when no position of this chunk stopped
    __syncthreads();
  }
  // the count after the last visited position
  __shared__ int64_t sTotal;
  if (threadIdx.x == 0) sTotal = 0;
  __syncthreads();
  const int last = sStop < dim ? sStop : dim - 1;
  int64_t part = 0;
  for (int i = threadIdx.x; i <= last; i += 1024) part += counts[i];
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) part += __shfl_xor_sync(kFull, part, off);
  if ((threadIdx.x & 31) == 0 && part) atomicAdd(reinterpret_cast<unsigned long long*>(&sTotal), (unsigned long long)part);
  __syncthreads();
  if (threadIdx.x == 0) state[0] += sTotal, state[1] = last + 1;
}

__global__ void __launch_bounds__(256) mh_rows_apply_kernel(TorchRng rng, const float* __restrict__ probs,
                                                            uint8_t* __restrict__ xs, const int64_t* __restrict__ perm,
                                                            int64_t num, int64_t num_src, int dim,
                                                            const int64_t* __restrict__ state) {
  const int64_t row = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int i = blockIdx.y;
  if (row >= num || i >= (int)state[1]) return;
  const int col = (int)perm[i];
  if (mh_accept(rng, probs, xs, row, row % num_src, dim, col, (uint32_t)i)) xs[row * dim + col] ^= 1;
}

static TorchRng make_rng(uint64_t seed, uint64_t offset, uint32_t threads, uint32_t iters) {
  TorchRng r;
  r.seed = seed, r.offset4 = offset / 4, r.threads = threads ? threads : 1, r.iters_per_call = iters ? iters : 1;
  r.dev = nullptr;
  return r;
}

}  // namespace rlsb

extern "C" {

int rlsb_mcpg_plan_create(const rlsb_graph_t* g, const int32_t* h_order, rlsb_mcpg_plan_t** out) {
  using namespace rlsb;
  RLSB_REQUIRE(out != nullptr, RLSB_ERR_INVALID, "mcpg_plan_create: out is null");
  *out = nullptr;
  const GraphDev* gd;
  if (int rc = graph_check(g, &gd, "mcpg_plan_create")) return rc;
  RLSB_REQUIRE(h_order != nullptr || g->n == 0, RLSB_ERR_INVALID, "mcpg_plan_create: null order");
  const int32_t n = g->n;
  std::vector<int32_t> pos(n, -1);
  for (int32_t k = 0; k < n; ++k) {
    RLSB_REQUIRE(h_order[k] >= 0 && h_order[k] < n && pos[h_order[k]] < 0, RLSB_ERR_INVALID,
                 "mcpg_plan_create: order is not a permutation of 0..N-1 (entry %d)", k);
    pos[h_order[k]] = k;
  }
  for (int64_t k = 0; k < g->m; ++k)
    RLSB_REQUIRE(g->edge_u[k] != g->edge_v[k], RLSB_ERR_UNSUPPORTED,
                 "mcpg_plan_create: self loop at edge %lld (the reference adds a node to its own neighbour list twice)",
                 (long long)k);
  auto* p = new rlsb_mcpg_plan();
  p->n = n, p->np = g->np, p->device = g->device, p->graph = g, p->max_deg = g->max_full_deg;
  p->order.assign(h_order, h_order + n);
  // neighbours split by visiting order
  std::vector<int32_t> es, ed, ls, ld;
  for (int32_t i = 0; i < n; ++i)
    for (int32_t k = g->full_ptr[i]; k < g->full_ptr[i + 1]; ++k) {
      const int32_t j = g->full_col[k];
      if (pos[j] < pos[i]) es.push_back(i), ed.push_back(j);
      else ls.push_back(i), ld.push_back(j);
    }
  std::vector<int32_t> eptr, ecol, lptr, lcol;
  build_csr(n, es, ed, eptr, ecol);
  build_csr(n, ls, ld, lptr, lcol);
  // dependency levels in visiting order
  std::vector<int32_t> level(n, 0);
  int32_t nlev = n ? 1 : 0;
  for (int32_t k = 0; k < n; ++k) {
    const int32_t i = h_order[k];
    int32_t lv = 0;
    for (int32_t e = eptr[i]; e < eptr[i + 1]; ++e) lv = std::max(lv, level[ecol[e]] + 1);
    level[i] = lv;
    nlev = std::max(nlev, lv + 1);
  }
  p->levels = nlev;
  std::vector<int32_t> level_ptr(nlev + 1, 0), nodes(n);
  for (int32_t i = 0; i < n; ++i) level_ptr[level[i] + 1]++;
  for (int32_t l = 0; l < nlev; ++l) level_ptr[l + 1] += level_ptr[l];
  {
    std::vector<int32_t> fill(level_ptr.begin(), level_ptr.end() - 1);
    for (int32_t k = 0; k < n; ++k) nodes[fill[level[h_order[k]]]++] = h_order[k];
  }
  for (int32_t l = 0; l < nlev; ++l)
    std::stable_sort(nodes.begin() + level_ptr[l], nodes.begin() + level_ptr[l + 1], [&](int32_t a, int32_t b) {
      return g->full_ptr[a + 1] - g->full_ptr[a] > g->full_ptr[b + 1] - g->full_ptr[b];
    });
  std::vector<int32_t> ls_e, ls_l;
  build_sell(nodes, level_ptr, eptr, ecol, p->earlier, &ls_e, g->np);   // padding -> the zero word sP[np]
  build_sell(nodes, level_ptr, lptr, lcol, p->later, &ls_l, g->np);
  p->level_slice = ls_e;
  p->num_slices = int32_t(p->earlier.off.size()) - 1;
  p->deg.assign(size_t(p->num_slices) * 32, 0), p->nlater = p->deg, p->pos = p->deg;
  for (size_t slot = 0; slot < p->earlier.node.size(); ++slot) {
    const uint16_t i = p->earlier.node[slot];
    if (i == 0xFFFF) continue;
    p->deg[slot] = uint16_t(g->full_ptr[i + 1] - g->full_ptr[i]);
    p->nlater[slot] = uint16_t(lptr[i + 1] - lptr[i]);
    p->pos[slot] = uint16_t(pos[i]);
  }
  // device image
  int prev = 0;
  cudaError_t e = cudaGetDevice(&prev);
  if (e == cudaSuccess) e = cudaSetDevice(g->device);
  struct Part { const void* src; size_t bytes, off; };
  std::vector<Part> parts;
  size_t total = 0;
  auto add = [&](const void* src, size_t bytes) {
    parts.push_back({src, bytes, total});
    total += (bytes + 255) / 256 * 256 + 256;
    return int(parts.size()) - 1;
  };
  const int b_ls = add(p->level_slice.data(), p->level_slice.size() * 4);
  const int b_eoff = add(p->earlier.off.data(), p->earlier.off.size() * 4);
  const int b_loff = add(p->later.off.data(), p->later.off.size() * 4);
  const int b_node = add(p->earlier.node.data(), p->earlier.node.size() * 2);
  const int b_deg = add(p->deg.data(), p->deg.size() * 2), b_nl = add(p->nlater.data(), p->nlater.size() * 2);
  const int b_pos = add(p->pos.data(), p->pos.size() * 2);
  const int b_ecol = add(p->earlier.col.data(), p->earlier.col.size() * 2);
  const int b_lcol = add(p->later.col.data(), p->later.col.size() * 2);
  if (e == cudaSuccess) e = cudaMalloc(&p->dev_blob, total);
  for (size_t a = 0; a < parts.size() && e == cudaSuccess; ++a)
    if (parts[a].bytes)
      e = cudaMemcpy((char*)p->dev_blob + parts[a].off, parts[a].src, parts[a].bytes, cudaMemcpyHostToDevice);
  cudaSetDevice(prev);
  if (e != cudaSuccess) {
    set_error("mcpg_plan_create: CUDA error %s", cudaGetErrorString(e));
    if (p->dev_blob) cudaFree(p->dev_blob);
    delete p;
    return RLSB_ERR_CUDA;
  }
  auto i32 = [&](int b) { return reinterpret_cast<const int32_t*>((char*)p->dev_blob + parts[b].off); };
  auto u16 = [&](int b) { return reinterpret_cast<const uint16_t*>((char*)p->dev_blob + parts[b].off); };
  p->dev.levels = p->levels, p->dev.num_slices = p->num_slices, p->dev.level_slice = i32(b_ls);
  p->dev.earlier = SellDev{p->num_slices, i32(b_eoff), u16(b_node), u16(b_deg), u16(b_ecol)};
  p->dev.later = SellDev{p->num_slices, i32(b_loff), u16(b_node), u16(b_deg), u16(b_lcol)};
  p->dev.deg = u16(b_deg), p->dev.nlater = u16(b_nl), p->dev.pos = u16(b_pos);
  *out = p;
  return RLSB_OK;
}

int rlsb_mcpg_plan_destroy(rlsb_mcpg_plan_t* p) {
  if (!p) return RLSB_OK;
  if (p->dev_blob) cudaFree(p->dev_blob);
  if (p->coin) cudaFree(p->coin);
  if (p->degpos) cudaFree(p->degpos);
  delete p;
  return RLSB_OK;
}

int32_t rlsb_mcpg_plan_num_levels(const rlsb_mcpg_plan_t* p) { return p ? p->levels : 0; }

int rlsb_mcpg_sweeps(const rlsb_graph_t* gh, const rlsb_mcpg_plan_t* plan, float* xs, int64_t num_chains,
                     int32_t num_ls, const float* explicit_u, uint64_t seed, uint64_t offset, uint32_t rng_threads,
                     uint32_t rng_iters, float* expected, void* stream) {
  using namespace rlsb;
  const GraphDev* g;
  if (int rc = graph_check(gh, &g, "mcpg_sweeps")) return rc;
  RLSB_REQUIRE(plan != nullptr && plan->graph == gh, RLSB_ERR_INVALID, "mcpg_sweeps: plan does not belong to this graph");
  RLSB_REQUIRE(num_chains >= 0 && num_ls >= 0, RLSB_ERR_INVALID, "mcpg_sweeps: negative size");
  RLSB_REQUIRE(num_chains < (int64_t(1) << 31), RLSB_ERR_UNSUPPORTED, "mcpg_sweeps: more than 2^31 chains");
  if (num_chains == 0 || g->n == 0) return RLSB_OK;
  RLSB_REQUIRE(xs && expected, RLSB_ERR_INVALID, "mcpg_sweeps: null pointer");
  RLSB_REQUIRE(explicit_u || (rng_threads > 0 && rng_iters > 0), RLSB_ERR_INVALID, "mcpg_sweeps: no random source");
  const size_t smem = ((size_t)g->np + 32) * 4 + (size_t)(plan->levels + 1) * 4 + 2 * (size_t)(plan->num_slices + 1) * 4 +
                      4 * (size_t)plan->num_slices * 64 + 16;      // state tile + the plan's small static arrays
  RLSB_REQUIRE(smem <= 220 * 1024, RLSB_ERR_UNSUPPORTED, "mcpg_sweeps: %d nodes exceed the shared-memory tile", g->n);
  const int64_t tiles = (num_chains + kTileEnvs - 1) / kTileEnvs;
  const unsigned grid = (unsigned)(tiles < 16 * kNumSMs ? tiles : 16 * kNumSMs);
  auto st = static_cast<cudaStream_t>(stream);
  const TorchRng rng = make_rng(seed, offset, rng_threads, rng_iters);
  const int cw = cut_warps_for(g->m, kMcpgThreads / 32);
  RLSB_REQUIRE(g->n <= 65535 && num_ls <= 65535, RLSB_ERR_UNSUPPORTED, "mcpg_sweeps: more than 65535 nodes or sweeps");
  // few tiles: tie-breaks decided beforehand by the whole GPU, words [num_ls * N][tiles] in the plan's scratch
  const bool coins = num_ls > 0 && tiles <= 4 * kNumSMs;
  if (coins) {
    const size_t need = (size_t)num_ls * g->n * tiles;
    if (plan->coin_words < need) {
      if (plan->coin) RLSB_CUDA_OK(cudaFree(plan->coin));
      plan->coin = nullptr, plan->coin_words = 0;
      RLSB_CUDA_OK(cudaMalloc(&plan->coin, need * sizeof(uint32_t)));
      plan->coin_words = need;
    }
    if (!plan->degpos) {
      std::vector<uint16_t> degpos((size_t)g->n, 0);
      for (size_t slot = 0; slot < plan->deg.size(); ++slot)
        if (plan->earlier.node[slot] != 0xFFFFu) degpos[plan->pos[slot]] = plan->deg[slot];
      RLSB_CUDA_OK(cudaMalloc(&plan->degpos, degpos.size() * sizeof(uint16_t)));
      RLSB_CUDA_OK(cudaMemcpy(plan->degpos, degpos.data(), degpos.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
    }
    const dim3 cgrid((unsigned)((tiles + 7) / 8), (unsigned)((g->n + kCoinPos - 1) / kCoinPos), (unsigned)num_ls);
    mcpg_coins_kernel<<<cgrid, 256, 0, st>>>(g->n, num_chains, tiles, plan->degpos, explicit_u, rng, plan->coin);
    RLSB_LAUNCH_OK();
  }
#define RLSB_MCPG_K(P, C)                                                                                   \
  do {                                                                                                      \
    if (smem > 48 * 1024)                                                                                   \
      RLSB_CUDA_OK(cudaFuncSetAttribute(mcpg_sweeps_kernel<P, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                        (int)smem));                                                        \
    mcpg_sweeps_kernel<P, C><<<grid, kMcpgThreads, smem, st>>>(*g, plan->dev, xs, num_chains, num_ls, plan->coin, \
                                                               explicit_u, rng, expected, cw);              \
  } while (0)
#define RLSB_MCPG(P)          \
  do {                        \
    if (coins)                \
      RLSB_MCPG_K(P, true);   \
    else                      \
      RLSB_MCPG_K(P, false);  \
  } while (0)
  if (plan->max_deg <= 63) RLSB_MCPG(6);
  else if (plan->max_deg <= 255) RLSB_MCPG(8);
  else RLSB_MCPG(12);
#undef RLSB_MCPG
#undef RLSB_MCPG_K
  RLSB_LAUNCH_OK();
  return RLSB_OK;
}

int rlsb_metro_sampling(int32_t num_nodes, const float* probs, const float* start, float* out, int64_t num_chains,
                        int32_t max_iters, const int32_t* num_iters_dev, const int64_t* explicit_idx,
                        const float* explicit_u, uint64_t seed, uint64_t offset, uint32_t rng_threads,
                        uint32_t rng_iters, int32_t* acc, int32_t count_only, void* stream) {
  using namespace rlsb;
  RLSB_REQUIRE(num_nodes > 0 && num_chains >= 0 && max_iters >= 0, RLSB_ERR_INVALID, "metro_sampling: bad size");
  RLSB_REQUIRE(num_chains < (int64_t(1) << 31), RLSB_ERR_UNSUPPORTED, "metro_sampling: more than 2^31 chains");
  if (num_chains == 0) return RLSB_OK;
  RLSB_REQUIRE(probs && start && (count_only ? acc != nullptr : (out != nullptr && num_iters_dev != nullptr)),
               RLSB_ERR_INVALID, "metro_sampling: null pointer");
  RLSB_REQUIRE((explicit_idx == nullptr) == (explicit_u == nullptr), RLSB_ERR_INVALID,
               "metro_sampling: explicit_idx and explicit_u go together");
  RLSB_REQUIRE(explicit_u || (rng_threads > 0 && rng_iters > 0), RLSB_ERR_INVALID, "metro_sampling: no random source");
  const int np = (num_nodes + 31) / 32 * 32;
  const size_t smem = (size_t)kMetroWarps * np * 4;
  RLSB_REQUIRE(smem <= 220 * 1024, RLSB_ERR_UNSUPPORTED, "metro_sampling: %d nodes exceed the shared-memory tiles",
               num_nodes);
  if (smem > 48 * 1024)
    RLSB_CUDA_OK(cudaFuncSetAttribute(metro_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t tiles = (num_chains + kTileEnvs - 1) / kTileEnvs;
  const int64_t ctas = (tiles + kMetroWarps - 1) / kMetroWarps;
  const unsigned grid = (unsigned)(ctas < 32 * kNumSMs ? ctas : 32 * kNumSMs);
  metro_kernel<<<grid, kMetroWarps * 32, smem, static_cast<cudaStream_t>(stream)>>>(
      num_nodes, np, probs, start, out, num_chains, max_iters, num_iters_dev, explicit_idx, explicit_u,
      make_rng(seed, offset, rng_threads, rng_iters), acc, count_only);
  RLSB_LAUNCH_OK();
  return RLSB_OK;
}

int rlsb_subset_sampling(uint8_t* xs, int64_t rows, int32_t num_nodes, int64_t num_sims, int32_t top_k,
                         const int64_t* ids, const float* vals, const float* explicit_u, uint64_t seed,
                         uint64_t offset, uint32_t rng_threads, uint32_t rng_iters, void* stream) {
  using namespace rlsb;
  RLSB_REQUIRE(rows >= 0 && num_nodes > 0 && num_sims > 0 && top_k >= 0 && rows % num_sims == 0, RLSB_ERR_INVALID,
               "subset_sampling: bad shape (rows must be num_repeats * num_sims)");
  RLSB_REQUIRE(rows < (int64_t(1) << 31), RLSB_ERR_UNSUPPORTED, "subset_sampling: more than 2^31 rows");
  if (rows == 0 || top_k == 0) return RLSB_OK;
  RLSB_REQUIRE(xs && ids && vals, RLSB_ERR_INVALID, "subset_sampling: null pointer");
  RLSB_REQUIRE(explicit_u || (rng_threads > 0 && rng_iters > 0), RLSB_ERR_INVALID, "subset_sampling: no random source");
  const int64_t tasks = rows * top_k;
  subset_kernel<<<(unsigned)((tasks + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      xs, rows, num_nodes, num_sims, top_k, ids, vals, explicit_u, make_rng(seed, offset, rng_threads, rng_iters));
  RLSB_LAUNCH_OK();
  return RLSB_OK;
}

int rlsb_mh_rows_round(const float* probs, uint8_t* xs, const int64_t* perm, int64_t num_rows, int64_t num_sims,
                       int32_t dim, int64_t target, uint64_t seed, uint64_t offset, uint32_t rng_threads, uint32_t rng_iters,
                       int64_t* state, int32_t* counts, void* stream) {
  using namespace rlsb;
  RLSB_REQUIRE(num_rows >= 0 && num_sims > 0 && dim > 0 && dim <= 65535 && num_rows % num_sims == 0, RLSB_ERR_INVALID,
               "mh_rows_round: bad shape (rows must be num_repeats * num_sims, at most 65535 columns)");
  RLSB_REQUIRE(num_rows < (int64_t(1) << 31), RLSB_ERR_UNSUPPORTED, "mh_rows_round: more than 2^31 rows");
  if (num_rows == 0) return RLSB_OK;
  RLSB_REQUIRE(probs && xs && perm && state && counts, RLSB_ERR_INVALID, "mh_rows_round: null pointer");
  RLSB_REQUIRE(rng_threads > 0 && rng_iters > 0, RLSB_ERR_INVALID, "mh_rows_round: no random source");
  auto st = static_cast<cudaStream_t>(stream);
  const TorchRng rng = make_rng(seed, offset, rng_threads, rng_iters);
  RLSB_CUDA_OK(cudaMemsetAsync(counts, 0, (size_t)dim * sizeof(int32_t), st));
  const dim3 grid((unsigned)((num_rows + 255) / 256), (unsigned)dim);
  mh_rows_count_kernel<<<grid, 256, 0, st>>>(rng, probs, xs, perm, num_rows, num_sims, dim, counts);
  RLSB_LAUNCH_OK();
  mh_rows_stop_kernel<<<1, 1024, 0, st>>>(counts, dim, target, state);
  RLSB_LAUNCH_OK();
  mh_rows_apply_kernel<<<grid, 256, 0, st>>>(rng, probs, xs, perm, num_rows, num_sims, dim, state);
  RLSB_LAUNCH_OK();
  return RLSB_OK;
}

int rlsb_torch_rand(uint64_t seed, uint64_t offset, uint32_t rng_threads, uint32_t rng_iters, int64_t calls,
                    int64_t numel, float* out, void* stream) {
  using namespace rlsb;
  RLSB_REQUIRE(calls >= 0 && numel >= 0 && rng_threads > 0 && rng_iters > 0 && numel < (int64_t(1) << 31),
               RLSB_ERR_INVALID, "torch_rand: bad argument");
  if (calls * numel == 0) return RLSB_OK;
  RLSB_REQUIRE(out != nullptr, RLSB_ERR_INVALID, "torch_rand: null pointer");
  torch_rand_kernel<<<(unsigned)((calls * numel + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      make_rng(seed, offset, rng_threads, rng_iters), calls, numel, out);
  RLSB_LAUNCH_OK();
  return RLSB_OK;
}

int rlsb_torch_randint(uint64_t seed, uint64_t offset, uint32_t rng_threads, uint32_t rng_iters, int64_t calls,
                       int64_t numel, uint32_t range, int64_t* out, void* stream) {
  using namespace rlsb;
  RLSB_REQUIRE(calls >= 0 && numel >= 0 && rng_threads > 0 && rng_iters > 0 && range > 0 && range < (1u << 28) &&
                   numel < (int64_t(1) << 31),
               RLSB_ERR_INVALID, "torch_randint: bad argument (range must be below 2^28: torch's 32-bit path)");
  if (calls * numel == 0) return RLSB_OK;
  RLSB_REQUIRE(out != nullptr, RLSB_ERR_INVALID, "torch_randint: null pointer");
  torch_randint_kernel<<<(unsigned)((calls * numel + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      make_rng(seed, offset, rng_threads, rng_iters), calls, numel, range, out);
  RLSB_LAUNCH_OK();
  return RLSB_OK;
}

int64_t rlsb_metro_workspace_bytes(int32_t num_nodes, int64_t num_chains, int32_t max_iters) {
  if (num_nodes <= 0 || num_chains < 0 || max_iters < 0) return -1;
  return (int64_t)rlsb::metro_carve(num_nodes, num_chains, max_iters, nullptr).bytes;
}

int rlsb_metro_sampling_split(int32_t num_nodes, const float* probs, const float* start, float* out, int64_t num_chains,
                              int32_t max_iters, int64_t stop_count, uint64_t seed, uint64_t offset, uint32_t rng_threads,
                              uint32_t rng_iters, int32_t* num_iters_dev, void* workspace, void* stream) {
  using namespace rlsb;
  RLSB_REQUIRE(num_nodes > 0 && num_chains >= 0 && max_iters >= 0, RLSB_ERR_INVALID, "metro_sampling_split: bad size");
  RLSB_REQUIRE(num_chains < (int64_t(1) << 31), RLSB_ERR_UNSUPPORTED, "metro_sampling_split: more than 2^31 chains");
  if (num_chains == 0) return RLSB_OK;
  RLSB_REQUIRE(probs && start && out && num_iters_dev && workspace, RLSB_ERR_INVALID, "metro_sampling_split: null pointer");
  RLSB_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255u) == 0, RLSB_ERR_INVALID,
               "metro_sampling_split: workspace must be 256-byte aligned");
  RLSB_REQUIRE(rng_threads > 0 && rng_iters > 0, RLSB_ERR_INVALID, "metro_sampling_split: no random source");
  const int np = (num_nodes + 31) / 32 * 32;
  const size_t smem = (size_t)kMetroWarps * np * 4 + (size_t)2 * num_nodes * 4;      // state tiles + the rate table
  RLSB_REQUIRE(smem <= 220 * 1024, RLSB_ERR_UNSUPPORTED, "metro_sampling_split: %d nodes exceed the shared-memory tiles",
               num_nodes);
  auto st = static_cast<cudaStream_t>(stream);
  const MetroWs w = metro_carve(num_nodes, num_chains, max_iters, workspace);
  const TorchRng rng = make_rng(seed, offset, rng_threads, rng_iters);
  if (max_iters > 0) {
    const dim3 dgrid((unsigned)((num_chains + 255) / 256), (unsigned)(max_iters < 65535 ? max_iters : 65535));
    metro_draws_kernel<<<dgrid, 256, 0, st>>>(rng, max_iters, num_chains, (uint32_t)num_nodes, w.draws);
    RLSB_LAUNCH_OK();
    RLSB_CUDA_OK(cudaMemsetAsync(w.acc, 0, (size_t)max_iters * 4, st));
  }
  metro_rates_kernel<<<(num_nodes + 255) / 256, 256, 0, st>>>(num_nodes, probs, w.rates);
  RLSB_LAUNCH_OK();
  if (smem > 48 * 1024) {
    RLSB_CUDA_OK(cudaFuncSetAttribute(metro_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  const int64_t tiles = (num_chains + kTileEnvs - 1) / kTileEnvs;
  const int64_t ctas = (tiles + kMetroWarps - 1) / kMetroWarps;
  const unsigned grid = (unsigned)(ctas < 32 * kNumSMs ? ctas : 32 * kNumSMs);
  const int64_t words = tiles * np;
  metro_pack_kernel<<<(unsigned)((words + 255) / 256), 256, 0, st>>>(num_nodes, np, start, num_chains, tiles, w.fin);
  RLSB_LAUNCH_OK();
  metro_pass_kernel<<<grid, kMetroWarps * 32, smem, st>>>(num_nodes, np, num_chains, max_iters, w.draws, w.rates, w.bits,
                                                          w.acc, w.fin);
  RLSB_LAUNCH_OK();
  metro_scan_kernel<<<1, 32, 0, st>>>(w.acc, max_iters, stop_count, num_iters_dev);
  RLSB_LAUNCH_OK();
  if (max_iters > 0) {
    const int64_t total = (int64_t)max_iters * num_chains;
    const unsigned ugrid = (unsigned)((total + 255) / 256 < 32 * kNumSMs ? (total + 255) / 256 : 32 * kNumSMs);
    metro_undo_kernel<<<ugrid, 256, 0, st>>>(np, num_chains, max_iters, tiles, num_iters_dev, w.draws, w.bits, w.fin);
  }
  RLSB_LAUNCH_OK();
  metro_unpack_kernel<<<(unsigned)((words + 255) / 256), 256, 0, st>>>(num_nodes, np, w.fin, num_chains, tiles, out);
  RLSB_LAUNCH_OK();
  return RLSB_OK;
}

}  // extern "C"
