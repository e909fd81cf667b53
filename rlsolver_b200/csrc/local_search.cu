// Local search: the three phases of EnvMaxcut.local_search_inplace
// (rlsolver/envs/env_L2A.py:87-116) and LocalSearch.random_search
// (rlsolver/methods/LocalSearch.py:53-86) as three launches:
//
//   ls_begin   (cross_counts.cu prepare_kernel) pack + objective + per-node cross counts + their
//              max/min over the env batch (the `ws_std` coupling, env_L2A.py:92-93)
//   ls_rdstd   rd_std[i] = float(mult * (max_i - min_i)) * noise_std                  (N floats)
//   ls_run     kth-value threshold of the noise-perturbed weights (env_L2A.py:94-96), ALL noisy
//              multi-flip iterations (97-107), the exhaustive single-flip pass (110-115) and the
//              unpack -- one CTA per tile of 32 envs, state in shared memory.
//
// ls_run is a producer/consumer pipeline.  Everything that streams -- the float32 noise (4 B per
// env-node-pass, the HBM traffic of this path) and the 1-byte cross counts -- is moved by the TMA
// engine: one producer thread issues, per 128-node chunk, ONE tensor copy of the 32-env x 128-node
// noise box (3-D tensor map {16 floats, env rows, N/16 column groups}, 64-byte swizzle) plus one
// bulk copy of the chunk's cross counts into a shared-memory ring and runs ahead across passes,
// so the next iteration's noise lands while the 16 consumer warps evaluate the cut of the current
// candidate.  (Per-row 512-byte bulk copies were tried first: the copy engine retires only about
// one small copy per ~85 cycles per SM and the consumers starved.)  Consumers work lane = env:
// one conflict-free LDS.128 gives a lane 4 nodes of its env, the compare result of the 32 lanes
// is one BALLOT = the 32-env flip word of a node.
//
// Floating point: spin_rand = ws + noise * rd_std is evaluated exactly as the reference's torch
// kernels do -- one IEEE round-to-nearest multiply, one add, no FMA contraction; float(ws) is
// exact (small integer).
#include <cuda.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "ls_workspace.cuh"
#include "tile_ops.cuh"

namespace rlsb {

int prepare_tiles(const GraphDev& g, const uint8_t* xs, const uint32_t* packed_in, int64_t num_envs,
                  uint32_t* packed_out, void* cross, int cross_layout, uint8_t* cross_rows, int32_t* col_min,
                  int32_t* col_max, int64_t* vs, cudaStream_t st);

// order-preserving float -> uint32 key (so REDUX max works on floats)
__device__ __forceinline__ uint32_t float_key(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_float(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// Cross counts of the local-search workspace are stored per tile as [node/4][env][node%4]
// (kCrossTiled, cross_counts.cu): the 4 counts a lane (= env) needs for one node group are one
// word (uint8) / two words (uint16), and a warp reads 128 / 256 contiguous bytes.
template <typename CrossT>
__device__ __forceinline__ int cross_at(const CrossT* __restrict__ cross_tile, int i, int e) {
  return (int)__ldg(cross_tile + (((i >> 2) * kTileEnvs + e) << 2) + (i & 3));
}

template <typename CrossT>
struct Cross4 {
  uint32_t lo, hi;
  __device__ __forceinline__ void load_shared(const char* p) {   // p -> the lane's 4 counts
    if constexpr (sizeof(CrossT) == 1) {
      lo = *reinterpret_cast<const uint32_t*>(p), hi = 0;
    } else {
      const uint2 c = *reinterpret_cast<const uint2*>(p);
      lo = c.x, hi = c.y;
    }
  }
  __device__ __forceinline__ int get(int b) const {
    if constexpr (sizeof(CrossT) == 1) return (int)__byte_perm(lo, 0, 0x4440 + b);   // byte b, zero-extended
    return (int)(((b < 2 ? lo : hi) >> (16 * (b & 1))) & 0xffffu);
  }
};

__global__ void ls_rdstd_kernel(GraphDev g, const int32_t* __restrict__ col_min, const int32_t* __restrict__ col_max,
                                int mult, float noise_std, float* __restrict__ rd_std, int32_t* __restrict__ degm,
                                uint32_t* __restrict__ nd) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.np) return;
  const bool real = i < g.n;
  const float rd = real ? __fmul_rn((float)(mult * (col_max[i] - col_min[i])), noise_std) : 0.f;
  const int dm = (real ? g.listed_deg[i] : 0) + kMagicI;
  rd_std[i] = rd, degm[i] = dm;
  nd[(i >> 2) * 8 + (i & 3)] = __float_as_uint(rd);
  nd[(i >> 2) * 8 + 4 + (i & 3)] = (uint32_t)dm;
}

// sorted (descending) list of the KMAX largest values seen
template <int KMAX>
struct TopList {
  float top[KMAX > 0 ? KMAX : 1];
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int t = 0; t < KMAX; ++t) top[t] = -INFINITY;
  }
  __device__ __forceinline__ void push(float s) {
    if (s > top[KMAX - 1]) {
#pragma unroll
      for (int t = 0; t < KMAX; ++t) {
        const float hi = fmaxf(top[t], s);
        s = fminf(top[t], s);
        top[t] = hi;
      }
    }
  }
  __device__ __forceinline__ void pop() {
#pragma unroll
    for (int t = 0; t + 1 < KMAX; ++t) top[t] = top[t + 1];
    top[KMAX - 1] = -INFINITY;
  }
};

// ---------------------------------------------------------------- generic thresh (kthvalue)
// Fallback for rows that are not 16-byte aligned (N % 4 != 0).  One warp per environment; each
// lane keeps the KMAX largest values of its share; the lists are merged by K rounds of warp-max.
template <int KMAX, typename CrossT>
__global__ void __launch_bounds__(256) ls_thresh_kernel(GraphDev g, const CrossT* __restrict__ cross,
                                                        const float* __restrict__ rd_std,
                                                        const int32_t* __restrict__ degm, int mult,
                                                        const float* __restrict__ noise, int kth_big,
                                                        int64_t num_envs, float* __restrict__ thresh) {
  const int lane = threadIdx.x & 31;
  const int64_t env = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (env >= num_envs) return;
  TopList<KMAX> tl;
  tl.clear();
  const CrossT* ctile = cross + (env / kTileEnvs) * (int64_t)g.np * kTileEnvs;
  const int e = (int)(env % kTileEnvs);
  const float* nrow = noise + env * (int64_t)g.n;
  for (int i = lane; i < g.n; i += 32)
    tl.push(spin_rand(__ldg(degm + i), -mult, cross_at(ctile, i, e), ldg_stream(nrow + i), __ldg(rd_std + i)));
  uint32_t best = 0;
  for (int r = 0; r < kth_big; ++r) {
    const uint32_t head = float_key(tl.top[0]);
    best = __reduce_max_sync(kFull, head);
    const unsigned who = __ballot_sync(kFull, head == best);
    if (lane == __ffs(who) - 1) tl.pop();
  }
  if (lane == 0) thresh[env] = key_float(best);
}

// The threshold alone (the fused-RNG path: rlsb_ls_run with no noisy iteration and no finish).  One warp per
// environment over the row-major uint8 copy of the cross counts: 4096 independent warps instead of the pipelined
// kernel's 128 CTAs -- that kernel's strength is keeping a tile resident across several passes, which a lone threshold
// pass has no use for (25.9 us at G22 x 4096 against the ~7 us its 41 MB take at HBM speed).  A lane takes four
// consecutive nodes per trip: one 16-byte streaming load of noise, one word of counts, rd_std / degree words as
// vectors (L1 resident: every warp reads the same 16 KB).
template <int KMAX>
__global__ void __launch_bounds__(128) ls_thresh_rows_kernel(int n, int np, const uint8_t* __restrict__ cross_rows,
                                                             const float* __restrict__ rd_std,
                                                             const int32_t* __restrict__ degm, int mult,
                                                             const float* __restrict__ noise, int kth_big,
                                                             int64_t num_envs, float* __restrict__ thresh) {
  constexpr int kParkDepth = 8;
  __shared__ float sPark[4][kParkDepth * 32];
  const int lane = threadIdx.x & 31;
  const int64_t env = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (env >= num_envs) return;
  TopList<KMAX> tl;
  tl.clear();
  const uint32_t* crow = reinterpret_cast<const uint32_t*>(cross_rows + env * (int64_t)np);
  const float* nrow = noise + env * (int64_t)n;
  const int groups = n >> 2;
  auto take = [&](int q, const float4& z, uint32_t c) {
    const float4 rd = __ldg(reinterpret_cast<const float4*>(rd_std) + q);
    const int4 dm = __ldg(reinterpret_cast<const int4*>(degm) + q);
    tl.push(spin_rand(dm.x, -mult, (int)(c & 0xffu), z.x, rd.x));
    tl.push(spin_rand(dm.y, -mult, (int)((c >> 8) & 0xffu), z.y, rd.y));
    tl.push(spin_rand(dm.z, -mult, (int)((c >> 16) & 0xffu), z.z, rd.z));
    tl.push(spin_rand(dm.w, -mult, (int)(c >> 24), z.w, rd.w));
  };
  // Phase 1: the first two trips (256 nodes of the row) go into the lanes' lists unconditionally; the kth largest of
  // those is a lower bound of the row's kth largest.  Phase 2: only values above that bound can matter, and they are
  // rare (~k ln(N / 256) per row), so the 2 x KMAX-instruction insertion all but disappears from the loop (the
  // unconditional form spent 2800 instructions per warp, 11.4 M in all, and ran as long as the pipelined kernel).
  int q = lane;
  for (int trip = 0; trip < 2 && q < groups; ++trip, q += 32) take(q, ldg_stream4(nrow + 4 * q), __ldg(crow + q));
  float bound;
  {
    TopList<KMAX> probe = tl;
    uint32_t best = 0;
    for (int r = 0; r < kth_big; ++r) {
      const uint32_t head = float_key(probe.top[0]);
      best = __reduce_max_sync(kFull, head);
      const unsigned who = __ballot_sync(kFull, head == best);
      if (lane == __ffs(who) - 1) probe.pop();
    }
    bound = key_float(best);            // -inf while the row has fewer than kth_big values
  }
  // A value above the bound is only PARKED in a lane-private column of shared memory (a predicated store and an add, no
  // branch): some lane of the warp meets one in almost every trip, and the sorted insertion executed by the whole warp
  // for it was still 2/3 of this kernel's instructions.  A lane expects ~2 of them per row; if a column overflows
  // (kParkDepth) the row is redone with every value inserted.
  float* park = sPark[threadIdx.x >> 5] + lane;
  int parked = 0;
  auto keep = [&](float sv) {
    const bool hit = sv > bound;
    if (hit) park[32 * min(parked, kParkDepth - 1)] = sv;
    parked += hit ? 1 : 0;
  };
  auto take_above = [&](int qq, const float4& z, uint32_t c) {
    const float4 rd = __ldg(reinterpret_cast<const float4*>(rd_std) + qq);
    const int4 dm = __ldg(reinterpret_cast<const int4*>(degm) + qq);
    keep(spin_rand(dm.x, -mult, (int)(c & 0xffu), z.x, rd.x));
    keep(spin_rand(dm.y, -mult, (int)((c >> 8) & 0xffu), z.y, rd.y));
    keep(spin_rand(dm.z, -mult, (int)((c >> 16) & 0xffu), z.z, rd.z));
    keep(spin_rand(dm.w, -mult, (int)(c >> 24), z.w, rd.w));
  };
  for (; q + 96 < groups; q += 128) {                 // four trips in flight per lane
    const float4 z0 = ldg_stream4(nrow + 4 * q), z1 = ldg_stream4(nrow + 4 * (q + 32));
    const float4 z2 = ldg_stream4(nrow + 4 * (q + 64)), z3 = ldg_stream4(nrow + 4 * (q + 96));
    const uint32_t c0 = __ldg(crow + q), c1 = __ldg(crow + q + 32), c2 = __ldg(crow + q + 64), c3 = __ldg(crow + q + 96);
    take_above(q, z0, c0), take_above(q + 32, z1, c1), take_above(q + 64, z2, c2), take_above(q + 96, z3, c3);
  }
  for (; q < groups; q += 32) take_above(q, ldg_stream4(nrow + 4 * q), __ldg(crow + q));
  if (__any_sync(kFull, parked > kParkDepth)) {
    tl.clear();
    for (q = lane; q < groups; q += 32) take(q, ldg_stream4(nrow + 4 * q), __ldg(crow + q));
  } else {
    for (int k = 0; k < parked; ++k) tl.push(park[32 * k]);
  }
  uint32_t best = 0;
  for (int r = 0; r < kth_big; ++r) {
    const uint32_t head = float_key(tl.top[0]);
    best = __reduce_max_sync(kFull, head);
    const unsigned who = __ballot_sync(kFull, head == best);
    if (lane == __ffs(who) - 1) tl.pop();
  }
  if (lane == 0) thresh[env] = key_float(best);
}

// ---------------------------------------------------------------- fused search kernel
constexpr int kLSThreads = 512;                 // consumer threads
constexpr int kLSWarps = kLSThreads / 32;
constexpr int kPipeThreads = kLSThreads + 32;   // + the producer warp
constexpr int kLSMaxIters = 16;                 // noise tensors per launch (pointers travel by value)
constexpr int kChunkGroups = 32;                // node groups (of 4 nodes) per pipeline chunk
constexpr int kChunkG16 = kChunkGroups / 4;     // the same in 16-float column groups (tensor-map dim 2)
// noise box in shared memory: [column group of 16 floats][env row][64 B], 16-byte pieces swizzled
// by the TMA engine (SWIZZLE_64B: piece ^= (address >> 7) & 3) -- lane = env reads are conflict free
constexpr int kNoiseStageBytes = kChunkG16 * kTileEnvs * 64;  // 16384
constexpr int kMaxStages = 8;

struct TmapPack {
  CUtensorMap m[kLSMaxIters + 1];   // [0]: thresh noise, [1 + k]: noise of iteration k
  CUtensorMap cross, nd;            // cross counts (tiled); rd_std + degree words interleaved per node group
};

__device__ __forceinline__ void tma_load_1d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0) {
  asm volatile("cp.async.bulk.tensor.1d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3}], [%2];" ::"r"(
                   smem_u32(dst)),
               "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0)
               : "memory");
}

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

struct NoisePtrs {
  const float* p[kLSMaxIters];
};

struct LsArgs {
  uint32_t* packed;        // [W][Np] in/out (workspace)
  int64_t* vs;             // [E] in/out
  const void* cross;       // [W][Np/4][32][4] uint8 / uint16
  const float* rd_std;     // [Np]
  const int32_t* degm;     // [Np] listed degree + kMagicI
  const uint32_t* nd;      // [Np/4][rd_std x4 | degm x4]: the two arrays interleaved per node group (pipe kernel)
  float* thresh;           // [E]
  const float* thresh_noise;   // non-null: first compute thresh from this tensor
  NoisePtrs noise;
  int num_iters, mult, kth_big;
  int64_t num_envs;
  int finish;              // 1: run the single-flip pass and write the bool rows
  uint8_t* xs_out;         // [E][N] bool rows (finish)
  int unpack_vec4;
  int cut_warps;
  int stage_sweep;         // 1: the sweep structure is staged into shared memory
  int sweep_warps;         // warps that take part in the single-flip pass
  int negmult;             // -mult
  int stages, stage_bytes; // ring geometry (pipe kernel)
  int skip;                // diagnostic: see PipeCtx::skip
  int ring_off;            // byte offset of the ring in dynamic shared memory
  long long* times;        // debug: phase timestamps of CTA 0 (RLSB_LS_TIMES), else null
};

__device__ __forceinline__ void stamp(const LsArgs& a, int& k) {
  if (a.times && blockIdx.x == 0 && threadIdx.x == 0 && k < 64) a.times[k] = clock64();
  ++k;
}

// barrier over the consumer threads only (the producer warp never joins it)
__device__ __forceinline__ void ls_sync() {
  asm volatile("bar.sync 2, %0;" ::"n"(kLSThreads) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Exhaustive single-flip pass, Gauss-Seidel over nodes 0..N-1 with acceptance gain >= 0.
// Nodes of one dependency level (graph_store.cu) are pairwise non-adjacent: one lane per
// node decides them concurrently and a barrier separates levels, which reproduces the
// sequential order exactly.  gain = deg - 2*cross >= 0  <=>  cross <= floor(deg/2).
//
// The pass is a chain of `levels` short steps (46 for the G22 shape) with two or three busy warps each, so
// what counts is the latency of one step.  Everything that does not depend on the previous level's flips --
// the level's slice range, the slot's node / half / own word and its first 32 neighbour ids -- is fetched one
// level ahead, before the barrier; after the barrier only the neighbour-word gathers, the counter adds, the
// compare and the store remain.  That form is used when the structure is read from global memory / L2
// (rlsb_flip_sweep with many tiles: 116 -> 54 us at G22 x 4096 before it staged the structure for few tiles).
constexpr int kSweepPre = 8;     // neighbour-id blocks (of 4 ids) fetched ahead per slot

template <int P, bool SMEM>
__device__ __forceinline__ void sweep_tile(const GraphDev& g, const SweepView& sv, uint32_t* sP, int sweep_warps) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp >= sweep_warps) return;      // a level never has more slices than sweep_warps (or all warps take part)
  if constexpr (SMEM) {
    // Structure staged in shared memory: every load is ~30 cycles and the busy warps sit alone on their
    // schedulers, so the shortest instruction sequence wins (fetching ahead measured 40 % slower here).
    for (int l = 0; l < g.levels; ++l) {
      const int sb = sv.level_slice[l], se = sv.level_slice[l + 1];
      for (int s = sb + warp; s < se; s += sweep_warps) {
        const uint32_t node = sv.sell.node[s * 32 + lane];
        const uint32_t half = sv.sell.half[s * 32 + lane];
        const bool active = node != 0xFFFFu;
        const uint32_t self = active ? sP[node] : 0u;
        VCount<P> vc;
        sell_cross<P, true>(sv.sell, s, lane, sP, self, vc);
        const uint32_t flip = vc.le(half);
        if (active) sP[node] = self ^ flip;
      }
      asm volatile("bar.sync 1, %0;" ::"r"(sweep_warps * 32) : "memory");   // only the sweeping warps
    }
    return;
  }
  struct Pre {
    int sb, se, nb;
    uint32_t node, half, self;
    const uint2* col;
    uint2 id[kSweepPre];
  };
  auto fetch = [&](int l, Pre& p) {
    p.sb = SMEM ? sv.level_slice[l] : __ldg(sv.level_slice + l);
    p.se = SMEM ? sv.level_slice[l + 1] : __ldg(sv.level_slice + l + 1);
    const int s = p.sb + warp;
    p.nb = 0, p.node = 0xFFFFu, p.half = 0u, p.self = 0u, p.col = nullptr;
    if (s < p.se) {
      p.node = SMEM ? sv.sell.node[s * 32 + lane] : __ldg(sv.sell.node + s * 32 + lane);
      p.half = SMEM ? sv.sell.half[s * 32 + lane] : __ldg(sv.sell.half + s * 32 + lane);
      const int gb = SMEM ? sv.sell.off[s] : __ldg(sv.sell.off + s);
      p.nb = (SMEM ? sv.sell.off[s + 1] : __ldg(sv.sell.off + s + 1)) - gb;
      p.col = reinterpret_cast<const uint2*>(sv.sell.col) + (int64_t)gb * 32 + lane;
#pragma unroll
      for (int b = 0; b < kSweepPre; ++b)
        p.id[b] = b < p.nb ? (SMEM ? p.col[b * 32] : __ldg(p.col + b * 32)) : make_uint2(0u, 0u);
      if (p.node != 0xFFFFu) p.self = sP[p.node];       // a node's own word only changes in its own level
    }
  };
  Pre cur;
  if (g.levels > 0) fetch(0, cur);
  for (int l = 0; l < g.levels; ++l) {
    Pre nxt;
    if (l + 1 < g.levels) fetch(l + 1, nxt);
    if (cur.sb + warp < cur.se) {
      const bool active = cur.node != 0xFFFFu;
      const uint32_t self = cur.self, pad = active ? cur.node : 0u;
      VCount<P> vc;
      vc.clear();
      auto add_pair = [&](uint2 i0, uint2 i1) {
        vc.add8(sP[i0.x & 0xffffu] ^ self, sP[i0.x >> 16] ^ self, sP[i0.y & 0xffffu] ^ self, sP[i0.y >> 16] ^ self,
                sP[i1.x & 0xffffu] ^ self, sP[i1.x >> 16] ^ self, sP[i1.y & 0xffffu] ^ self, sP[i1.y >> 16] ^ self);
      };
      const uint2 own = make_uint2(pad | (pad << 16), pad | (pad << 16));     // word ^ word == 0
#pragma unroll
      for (int b = 0; b < kSweepPre; b += 2)
        if (b < cur.nb) add_pair(cur.id[b], b + 1 < cur.nb ? cur.id[b + 1] : own);
      for (int b = kSweepPre; b < cur.nb; b += 2) {
        const uint2 i0 = SMEM ? cur.col[b * 32] : __ldg(cur.col + b * 32);
        uint2 i1 = own;
        if (b + 1 < cur.nb) i1 = SMEM ? cur.col[(b + 1) * 32] : __ldg(cur.col + (b + 1) * 32);
        add_pair(i0, i1);
      }
      const uint32_t flip = vc.le(cur.half);
      if (active) sP[cur.node] = self ^ flip;
    }
    for (int s = cur.sb + warp + sweep_warps; s < cur.se; s += sweep_warps) {     // further slices of a wide level
      const uint32_t node = SMEM ? sv.sell.node[s * 32 + lane] : __ldg(sv.sell.node + s * 32 + lane);
      const uint32_t half = SMEM ? sv.sell.half[s * 32 + lane] : __ldg(sv.sell.half + s * 32 + lane);
      const bool active = node != 0xFFFFu;
      const uint32_t self = active ? sP[node] : 0u;
      VCount<P> vc;
      sell_cross<P, SMEM>(sv.sell, s, lane, sP, self, vc);
      const uint32_t flip = vc.le(half);
      if (active) sP[node] = self ^ flip;
    }
    asm volatile("bar.sync 1, %0;" ::"r"(sweep_warps * 32) : "memory");   // only the sweeping warps
    cur = nxt;
  }
}

// Shared tail of both search kernels (consumer threads only): evaluate the candidate in sX and
// keep rows that are not worse (vs1 >= vs0, util_read_data.py:199).  my_vs: warp 0, lane = env.
__device__ __forceinline__ void evaluate_and_accept(const GraphDev& g, const LsArgs& a, uint32_t* sP,
                                                    const uint32_t* sX, int* sCnt, uint32_t* sAccept, int valid,
                                                    int64_t& my_vs) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cnt = tile_cut_partial(g, sX, a.cut_warps);
  if (cnt) atomicAdd(&sCnt[lane], cnt);
  ls_sync();
  if (warp == 0) {
    const int64_t cand = sCnt[lane];
    const bool keep = lane < valid && cand >= my_vs;
    if (keep) my_vs = cand;
    const unsigned acc = __ballot_sync(kFull, keep);
    if (lane == 0) *sAccept = acc;
    sCnt[lane] = 0;
  }
  ls_sync();
  const uint32_t acc = *sAccept;
  for (int i = threadIdx.x; i < g.n; i += kLSThreads) sP[i] = (sX[i] & acc) | (sP[i] & ~acc);
  ls_sync();
}

// single-flip pass + final values + bool rows (consumer threads only; sCnt is zero on entry)
template <int P>
__device__ __forceinline__ void finish_tile(const GraphDev& g, const LsArgs& a, uint32_t* sP, const char* sSweep,
                                            uint64_t* sBar, int* sCnt, int64_t tile, int valid, int* tk = nullptr) {
  const int lane = threadIdx.x & 31;
  if (a.stage_sweep) {
    mbar_wait(sBar, 0);
    sweep_tile<P, true>(g, sweep_view(g, sSweep), sP, a.sweep_warps);
  } else {
    sweep_tile<P, false>(g, sweep_view(g, g.sweep_blob), sP, a.sweep_warps);
  }
  ls_sync();
  if (tk) stamp(a, *tk);
  const int cnt = tile_cut_partial(g, sP, a.cut_warps);
  if (cnt) atomicAdd(&sCnt[lane], cnt);
  ls_sync();
  if (tk) stamp(a, *tk);
  if (threadIdx.x < valid) a.vs[tile * kTileEnvs + threadIdx.x] = sCnt[threadIdx.x];
  if (!a.xs_out) return;        // packed output only (the caller reads the workspace's packed tiles)
  if (a.unpack_vec4)
    unpack_tile_from_smem<4>(sP, a.xs_out, a.num_envs, g.n, g.np, tile, kLSWarps);
  else
    unpack_tile_from_smem<1>(sP, a.xs_out, a.num_envs, g.n, g.np, tile, kLSWarps);
}

// one thread: TMA the sweep structure into shared memory (it lands while other work runs)
__device__ __forceinline__ void stage_sweep_blob(const GraphDev& g, char* dst, uint64_t* sBar) {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic reads of dst are done
  mbar_expect_tx(sBar, (uint32_t)g.sweep_blob_bytes);
  for (int off = 0; off < g.sweep_blob_bytes; off += 32768) {
    const int len = g.sweep_blob_bytes - off < 32768 ? g.sweep_blob_bytes - off : 32768;
    bulk_g2s(dst + off, g.sweep_blob + off, (uint32_t)len, sBar);
  }
}

// Threshold pass of the pipelined kernel.  With lane = env a per-lane sorted top-K list diverges on
// almost every value (some lane of the 32 always inserts; measured 26 us for G22 x 4096).  Instead
// every env has an append buffer in shared memory ([slot][env], conflict free): a value is
// appended iff it exceeds the env's current bound (predicated, no divergence), and when a buffer
// could overflow during the next chunk the CTA compacts all buffers to their kth_big largest
// values (two envs per warp, rounds of warp-max) and raises the bounds to the kth_big-th largest
// seen so far.  After the first chunk the bound is already tight: ~kth*ln(N/128) later appends.
constexpr int kSelCap = 192;                           // slots per env
constexpr int kSelBytes = kSelCap * kTileEnvs * 4;     // 24 KB of dynamic shared memory

// kth_big largest values of env e's buffer -> slots 0..kth_big-1 (descending); returns the last
__device__ __forceinline__ float select_env(float* sSel, int* sSelCnt, int e, int kth_big, int lane) {
  const int cnt = sSelCnt[e];
  float v[kSelCap / 32];
#pragma unroll
  for (int j = 0; j < kSelCap / 32; ++j) v[j] = (lane + 32 * j < cnt) ? sSel[(lane + 32 * j) * kTileEnvs + e] : -INFINITY;
  __syncwarp();
  float kth = -INFINITY;
  for (int r = 0; r < kth_big; ++r) {
    float m = v[0];
#pragma unroll
    for (int j = 1; j < kSelCap / 32; ++j) m = fmaxf(m, v[j]);
    const uint32_t key = float_key(m);
    const uint32_t best = __reduce_max_sync(kFull, key);
    const unsigned who = __ballot_sync(kFull, key == best);
    if (lane == __ffs(who) - 1) {          // drop ONE instance (ties are separate elements)
      bool done = false;
#pragma unroll
      for (int j = 0; j < kSelCap / 32; ++j)
        if (!done && v[j] == m) v[j] = -INFINITY, done = true;
    }
    kth = key_float(best);
    if (lane == 0) sSel[r * kTileEnvs + e] = kth;
  }
  __syncwarp();
  if (lane == 0) sSelCnt[e] = kth_big < cnt ? kth_big : cnt;
  return kth;
}

// ---- consumer side of the ring.  Per chunk a warp owns node groups `warp` and `warp + 16`; the
// lane's offsets inside a stage are the same for every chunk (noise: [group of 16 floats][env][64 B]
// with swizzled 16-byte pieces; cross: [node group][env][4]; then rd_std and degm of the chunk).
template <typename CrossT>
struct PipeCtx {
  const char* ring;
  int stage_bytes, stages;
  uint64_t *sFull, *sEmpty;
  int groups, nchunks, negmult;
  int s;            // ring cursor
  uint32_t phase;
  int skip;         // diagnostic (RLSB_LS_SKIP=1): consume the ring without evaluating anything
  static constexpr int kCrossBytes = kChunkGroups * kTileEnvs * 4 * (int)sizeof(CrossT);
  static constexpr int kNodeOff = kNoiseStageBytes + kCrossBytes;

  // spin_rand of the lane's env for the 4 nodes of group (warp + half * 16) of the stage
  __device__ __forceinline__ void group_values(const char* st, int half, float (&sr)[4]) const {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q = warp + half * kLSWarps;
    const float4 nz = *reinterpret_cast<const float4*>(st + (q >> 2) * (kTileEnvs * 64) + lane * 64 +
                                                       (((q & 3) ^ ((lane >> 1) & 3)) << 4));
    Cross4<CrossT> cr;
    cr.load_shared(st + kNoiseStageBytes + (q * kTileEnvs + lane) * (4 * (int)sizeof(CrossT)));
    const float4 rd = *reinterpret_cast<const float4*>(st + kNodeOff + q * 32);                      // broadcast
    const int4 dm = *reinterpret_cast<const int4*>(st + kNodeOff + q * 32 + 16);
    sr[0] = spin_rand(dm.x, negmult, cr.get(0), nz.x, rd.x);
    sr[1] = spin_rand(dm.y, negmult, cr.get(1), nz.y, rd.y);
    sr[2] = spin_rand(dm.z, negmult, cr.get(2), nz.z, rd.z);
    sr[3] = spin_rand(dm.w, negmult, cr.get(3), nz.w, rd.w);
  }
  __device__ __forceinline__ const char* acquire() const {
    mbar_wait(&sFull[s], phase);
    return ring + s * stage_bytes;
  }
  __device__ __forceinline__ void release() {
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(&sEmpty[s]);
    if (++s == stages) s = 0, phase ^= 1;
  }
};

// One noisy iteration's candidate: sX = sP ^ (spin_rand > thresh).  The ballot of the 32 lanes
// (= envs) is the 32-env flip word of a node.
template <typename CrossT>
__device__ __forceinline__ void mask_pass(PipeCtx<CrossT>& cx, float th, const uint32_t* sP, uint32_t* sX) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  auto emit = [&](const float(&sr)[4], int i0) {
    const uint32_t b0 = __ballot_sync(kFull, sr[0] > th), b1 = __ballot_sync(kFull, sr[1] > th);
    const uint32_t b2 = __ballot_sync(kFull, sr[2] > th), b3 = __ballot_sync(kFull, sr[3] > th);
    const uint4 pw = *reinterpret_cast<const uint4*>(sP + i0);
    if (lane == 0) *reinterpret_cast<uint4*>(sX + i0) = make_uint4(pw.x ^ b0, pw.y ^ b1, pw.z ^ b2, pw.w ^ b3);
  };
  for (int c = 0; c < cx.nchunks; ++c) {
    const char* st = cx.acquire();
    const int gc = cx.groups - c * kChunkGroups;           // node groups left (>= 32: a full chunk)
    const int i0 = (c * kChunkGroups + warp) * 4;
    float sa[4], sb[4];
    if (cx.skip) {
      // nothing: measures the pure streaming rate of the ring
    } else if (gc >= kChunkGroups) {
      cx.group_values(st, 0, sa);
      cx.group_values(st, 1, sb);
      emit(sa, i0);
      emit(sb, i0 + 4 * kLSWarps);
    } else {
      if (warp < gc) cx.group_values(st, 0, sa), emit(sa, i0);
      if (warp + kLSWarps < gc) cx.group_values(st, 1, sb), emit(sb, i0 + 4 * kLSWarps);
    }
    cx.release();
  }
}

// Threshold pass: returns the kth_big-th largest spin_rand of the lane's env.
template <typename CrossT>
__device__ __forceinline__ float thresh_pass(PipeCtx<CrossT>& cx, float* sSel, int* sSelCnt, int* sSelFull,
                                             float* sBound, int kth_big) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x < kTileEnvs) sSelCnt[threadIdx.x] = 0;
  if (threadIdx.x < 3) sSelFull[threadIdx.x] = 0;
  ls_sync();
  float bound = -INFINITY;
  int flag_slot = 0;
  // append the values above the env's bound to its buffer (lane = env; predicated, no divergence)
  auto append = [&](const float(&sr)[4]) {
    const bool q0 = sr[0] > bound, q1 = sr[1] > bound, q2 = sr[2] > bound, q3 = sr[3] > bound;
    const int cnt = (int)q0 + (int)q1 + (int)q2 + (int)q3;
    if (cnt) {
      int at = atomicAdd(&sSelCnt[lane], cnt);
      // a buffer must be able to take a whole further chunk (128 values): ask for a compaction
      // (the env's last appender of the chunk sees the full count)
      if (at + cnt + 4 * kChunkGroups > kSelCap) sSelFull[flag_slot] = 1;
      if (q0) sSel[(at++) * kTileEnvs + lane] = sr[0];
      if (q1) sSel[(at++) * kTileEnvs + lane] = sr[1];
      if (q2) sSel[(at++) * kTileEnvs + lane] = sr[2];
      if (q3) sSel[at * kTileEnvs + lane] = sr[3];
    }
  };
  for (int c = 0; c < cx.nchunks; ++c) {
    const char* st = cx.acquire();
    const int gc = cx.groups - c * kChunkGroups;
    float sa[4], sb[4];
    if (warp < gc) cx.group_values(st, 0, sa), append(sa);
    if (warp + kLSWarps < gc) cx.group_values(st, 1, sb), append(sb);
    cx.release();
    // Flag slots rotate over 3 chunks: slot (c+2)%3 was last read before this barrier and is next
    // written after the following one, so thread 0 can clear it in between.
    ls_sync();
    const bool compact = sSelFull[flag_slot] != 0 || c + 1 == cx.nchunks;
    flag_slot = flag_slot == 2 ? 0 : flag_slot + 1;
    if (threadIdx.x == 0) sSelFull[flag_slot == 2 ? 0 : flag_slot + 1] = 0;
    if (compact) {
      const float b0 = select_env(sSel, sSelCnt, 2 * warp, kth_big, lane);
      const float b1 = select_env(sSel, sSelCnt, 2 * warp + 1, kth_big, lane);
      if (lane == 0) {      // fewer than kth_big values so far: no bound yet
        sBound[2 * warp] = sSelCnt[2 * warp] >= kth_big ? b0 : -INFINITY;
        sBound[2 * warp + 1] = sSelCnt[2 * warp + 1] >= kth_big ? b1 : -INFINITY;
      }
      ls_sync();
      bound = sBound[lane];
    }
  }
  return bound;
}

// ---- pipelined kernel (rows made of whole 64-byte column groups).  THRESH: the threshold pass is
// compiled in.
template <int P, typename CrossT, bool THRESH>
__global__ void __launch_bounds__(kPipeThreads, 1) ls_pipe_kernel(GraphDev g, LsArgs a,
                                                                  const __grid_constant__ TmapPack maps) {
  extern __shared__ __align__(1024) uint32_t smem[];
  uint32_t* sP = smem;               // accepted state of the tile
  uint32_t* sX = smem + g.np;        // candidate state
  // noise/cross stages (1024-byte aligned for the swizzle); later the sweep structure
  char* ring = reinterpret_cast<char*>(smem) + a.ring_off;
  __shared__ int sCnt[kTileEnvs];
  __shared__ uint32_t sAccept;
  __shared__ __align__(8) uint64_t sFull[kMaxStages], sEmpty[kMaxStages], sBar;
  __shared__ int sSelCnt[kTileEnvs];
  __shared__ int sSelFull[3];        // "some buffer is past half" flag of chunk c lives in slot c % 3
  __shared__ float sBound[kTileEnvs];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t tile = blockIdx.x;
  const int64_t env0 = tile * kTileEnvs;
  const int valid = (int)min((int64_t)kTileEnvs, a.num_envs - env0);
  const bool has_thresh = THRESH && a.thresh_noise != nullptr;
  const int passes = (has_thresh ? 1 : 0) + a.num_iters;
  const int groups = g.n >> 2;                                   // N % 4 == 0 on this path
  const int nchunks = (groups + kChunkGroups - 1) / kChunkGroups;
  if (threadIdx.x == 0) {
    for (int s = 0; s < a.stages; ++s) mbar_init(&sFull[s], 1), mbar_init(&sEmpty[s], kLSWarps);
    mbar_init(&sBar, 1);
    if (a.finish && a.stage_sweep && passes == 0) stage_sweep_blob(g, ring, &sBar);
  }
  __syncthreads();

  if (warp == kLSWarps) {
    // ---------------- producer: one thread feeds the ring
    if (lane == 0) {
      int s = 0, round = 0;
      for (int p = 0; p < passes; ++p) {
        const CUtensorMap* map = &maps.m[has_thresh ? p : p + 1];
        for (int c = 0; c < nchunks; ++c) {
          if (round > 0) mbar_wait(&sEmpty[s], (round - 1) & 1);
          char* st = ring + s * a.stage_bytes;
          // Everything goes through TENSOR copies: plain cp.async.bulk copies stream at only ~6 B/clk per SM
          // (measured: the 4 KB cross-count copy per chunk capped the ring at 36 GB/s per SM).  Boxes are
          // always written whole (elements past the tensor read as zeros).
          mbar_expect_tx(&sFull[s], (uint32_t)a.stage_bytes);
          tma_load_3d(st, map, &sFull[s], 0, (int)env0, c * kChunkG16);
          tma_load_2d(st + kNoiseStageBytes, &maps.cross, &sFull[s], 0, (int)(tile * (g.np / 4)) + c * kChunkGroups);
          // the chunk's rd_std / degree words ride along (interleaved per node group: one copy), so the
          // consumers' loop has no global load at all
          tma_load_1d(st + kNoiseStageBytes + kChunkGroups * kTileEnvs * 4 * (int)sizeof(CrossT), &maps.nd, &sFull[s],
                      c * (kChunkGroups * 8));
          if (++s == a.stages) s = 0, ++round;
        }
      }
    }
    return;
  }

  // ---------------- consumer warps
  const bool thresh_only = has_thresh && a.num_iters == 0 && !a.finish;   // no candidate, no state change: skip the tile
  if (!thresh_only) {
    for (int i = threadIdx.x; i < g.np; i += kLSThreads) {
      const uint32_t w = a.packed[tile * g.np + i];
      sP[i] = w, sX[i] = w;        // padding nodes stay equal in both copies
    }
  }
  if (threadIdx.x < kTileEnvs) sCnt[threadIdx.x] = 0;
  int64_t my_vs = 0;   // warp 0: lane e owns env e's value
  if (warp == 0 && lane < valid) my_vs = a.vs[env0 + lane];
  float th = INFINITY;   // +inf for envs past the batch: never flips
  if (!has_thresh && a.num_iters > 0 && lane < valid) th = a.thresh[env0 + lane];
  ls_sync();

  PipeCtx<CrossT> cx;
  cx.ring = ring, cx.stage_bytes = a.stage_bytes, cx.stages = a.stages, cx.sFull = sFull, cx.sEmpty = sEmpty;
  cx.groups = groups, cx.nchunks = nchunks, cx.negmult = a.negmult, cx.s = 0, cx.phase = 0, cx.skip = a.skip;
  float* sSel = reinterpret_cast<float*>(ring + (size_t)a.stages * a.stage_bytes);   // [kSelCap][32] (THRESH)
  int tk = 0;
  stamp(a, tk);
  for (int p = 0; p < passes; ++p) {
    if (THRESH && has_thresh && p == 0) {
      const float kth = thresh_pass<CrossT>(cx, sSel, sSelCnt, sSelFull, sBound, a.kth_big);
      if (lane < valid) {
        th = kth;            // == the kth_big-th largest of the row (N > num_spin)
        if (warp == 0) a.thresh[env0 + lane] = kth;
      }
      stamp(a, tk);
      stamp(a, tk);
      continue;
    }
    mask_pass<CrossT>(cx, th, sP, sX);
    ls_sync();     // candidate complete
    stamp(a, tk);
    if (p == passes - 1 && a.finish && a.stage_sweep && threadIdx.x == 0) stage_sweep_blob(g, ring, &sBar);
    evaluate_and_accept(g, a, sP, sX, sCnt, &sAccept, valid, my_vs);
    stamp(a, tk);
  }
  if (has_thresh && passes == 1) {
    ls_sync();     // all reads of the ring are done
    if (a.finish && a.stage_sweep && threadIdx.x == 0) stage_sweep_blob(g, ring, &sBar);
  }
  const uint32_t vmask = valid == 32 ? kFull : ((1u << valid) - 1u);
  if (a.finish) {
    finish_tile<P>(g, a, sP, ring, &sBar, sCnt, tile, valid);
  } else {
    if (warp == 0 && lane < valid) a.vs[env0 + lane] = my_vs;
  }
  stamp(a, tk);
  if (!thresh_only)
    for (int i = threadIdx.x; i < g.np; i += kLSThreads) a.packed[tile * g.np + i] = sP[i] & vmask;
}

// ---- generic kernel (any N / alignment): direct global loads, thresh precomputed by
// ls_thresh_kernel.  Work item = one node x 8 consecutive envs (byte g of a word = envs 8g..8g+7).
template <typename CrossT>
__device__ __forceinline__ void noisy_candidate(const GraphDev& g, const LsArgs& a, const float* __restrict__ noise,
                                                int64_t tile, int valid, const float* sThresh, const uint32_t* sP,
                                                uint32_t* sX) {
  const float* __restrict__ nbase = noise + tile * kTileEnvs * (int64_t)g.n;
  const CrossT* __restrict__ ctile = static_cast<const CrossT*>(a.cross) + tile * (int64_t)g.np * kTileEnvs;
  const uint32_t n = (uint32_t)g.n;
  const uint8_t* sPb = reinterpret_cast<const uint8_t*>(sP);
  uint8_t* sXb = reinterpret_cast<uint8_t*>(sX);
  for (uint32_t task = threadIdx.x; task < n * 4; task += kLSThreads) {
    const uint32_t eg = task / n, i = task - eg * n;
    const uint32_t e0 = eg * 8;
    const float rd = __ldg(a.rd_std + i);
    const int dm = __ldg(a.degm + i);
    uint32_t bits = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if ((int)(e0 + j) < valid) {
        const float sr = spin_rand(dm, a.negmult, cross_at(ctile, (int)i, (int)(e0 + j)),
                                   ldg_stream(nbase + (e0 + j) * n + i), rd);
        if (sr > sThresh[e0 + j]) bits |= 1u << j;
      }
    }
    const uint32_t at = i * 4 + eg;
    sXb[at] = sPb[at] ^ (uint8_t)bits;
  }
}

template <int P, typename CrossT>
__global__ void __launch_bounds__(kLSThreads) ls_generic_kernel(GraphDev g, LsArgs a) {
  extern __shared__ __align__(1024) uint32_t smem[];
  uint32_t* sP = smem;
  uint32_t* sX = smem + g.np;
  char* sSweep = reinterpret_cast<char*>(smem + 2 * g.np);
  __shared__ float sThresh[kTileEnvs];
  __shared__ int sCnt[kTileEnvs];
  __shared__ uint32_t sAccept;
  __shared__ __align__(8) uint64_t sBar;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t tile = blockIdx.x;
  const int64_t env0 = tile * kTileEnvs;
  const int valid = (int)min((int64_t)kTileEnvs, a.num_envs - env0);
  if (threadIdx.x == 0) {
    mbar_init(&sBar, 1);
    if (a.finish && a.stage_sweep) stage_sweep_blob(g, sSweep, &sBar);
  }
  for (int i = threadIdx.x; i < g.np; i += kLSThreads) {
    const uint32_t w = a.packed[tile * g.np + i];
    sP[i] = w, sX[i] = w;
  }
  if (threadIdx.x < kTileEnvs) {
    sThresh[threadIdx.x] = (a.num_iters > 0 && threadIdx.x < valid) ? a.thresh[env0 + threadIdx.x] : INFINITY;
    sCnt[threadIdx.x] = 0;
  }
  int64_t my_vs = 0;
  if (warp == 0 && lane < valid) my_vs = a.vs[env0 + lane];
  ls_sync();
  for (int it = 0; it < a.num_iters; ++it) {
    noisy_candidate<CrossT>(g, a, a.noise.p[it], tile, valid, sThresh, sP, sX);
    ls_sync();
    evaluate_and_accept(g, a, sP, sX, sCnt, &sAccept, valid, my_vs);
  }
  const uint32_t vmask = valid == 32 ? kFull : ((1u << valid) - 1u);
  if (a.finish) {
    finish_tile<P>(g, a, sP, sSweep, &sBar, sCnt, tile, valid);
  } else {
    if (warp == 0 && lane < valid) a.vs[env0 + lane] = my_vs;
  }
  for (int i = threadIdx.x; i < g.np; i += kLSThreads) a.packed[tile * g.np + i] = sP[i] & vmask;
}

// ---- bit-mask kernel (rlsb_ls_run_masks): the flip masks of every iteration were produced by
// noise_masks.cu as flat bit arrays (bit e*N + n).  Lane = env fetches the 32 node bits of its row
// for one block of 32 nodes (two words + funnel shift: rows start at any bit), a 32x32 bit
// transpose turns them into the 32-env flip words of those nodes.
//
// A warp handles node blocks warp, warp + 16, ...; the words of four blocks are fetched together (8 loads in
// flight per lane) and the first four of the NEXT iteration are fetched before the current candidate is
// evaluated, so the L2 latency of the mask words stays off the critical path.
struct MaskChunk {
  uint32_t lo[4], hi[4];
};

// coherent: the words were written by a kernel that is still running (fused search): read them through L2
__device__ __forceinline__ MaskChunk mask_chunk_load(const uint32_t* __restrict__ mask, uint64_t row, int b0, int blocks,
                                                     bool live, bool coherent = false) {
  MaskChunk c;
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int b = b0 + u * kLSWarps;
    c.lo[u] = c.hi[u] = 0;
    if (live && b < blocks) {
      const uint64_t o = row + 32u * (uint32_t)b;
      if (coherent) c.lo[u] = __ldcg(mask + (o >> 5)), c.hi[u] = __ldcg(mask + (o >> 5) + 1);
      else c.lo[u] = __ldg(mask + (o >> 5)), c.hi[u] = __ldg(mask + (o >> 5) + 1);
    }
  }
  return c;
}

// sF / sNF (optional): the nodes whose flip word is not zero are appended to a list (ballot-compacted, one
// shared-memory atomic per block of 32 nodes) for the delta evaluation below.
__device__ __forceinline__ void mask_chunk_apply(const MaskChunk& c, uint64_t row, int b0, int n, const uint32_t* sP,
                                                 uint32_t* sX, uint16_t* sF, int* sNF) {
  const int lane = threadIdx.x & 31;
  // the four blocks of a chunk are independent: their transposes interleave, and ONE shared-memory atomic reserves the
  // list slots of all four (it was one per block, each with its round trip on the warp's critical path)
  uint32_t ball[4];
  bool flips[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int b = b0 + u * kLSWarps;
    flips[u] = false;
    if (32 * b < n) {                                               // warp-uniform
      const uint32_t word = __funnelshift_r(c.lo[u], c.hi[u], (uint32_t)(row + 32u * (uint32_t)b) & 31u);
      const uint32_t t = transpose32(word, lane);                   // lane = node 32b + lane, bit = env
      const int i = 32 * b + lane;
      flips[u] = i < n && t != 0u;
      if (i < n) sX[i] = sP[i] ^ t;
    }
  }
  if (sF) {
    int total = 0;
#pragma unroll
    for (int u = 0; u < 4; ++u) ball[u] = __ballot_sync(kFull, flips[u]), total += __popc(ball[u]);
    if (total) {                                                    // warp-uniform
      int base = 0;
      if (lane == 0) base = atomicAdd(sNF, total);
      base = __shfl_sync(kFull, base, 0);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (flips[u]) sF[base + __popc(ball[u] & ((1u << lane) - 1u))] = (uint16_t)(32 * (b0 + u * kLSWarps) + lane);
        base += __popc(ball[u]);
      }
    }
  }
}

// Delta evaluation of a candidate.  Only ~num_spin nodes per env flip in an iteration, so instead of
// re-counting all M edges (env_L2A.py:102 does a full calculate_obj_values) the value moves by the edges that
// toggle: for a flipped node i and a neighbour j that does NOT flip in the same env, edge (i, j) goes from
// cut to uncut (-1) or back (+1); edges whose two ends flip keep their state.  One lane per flipped node walks
// its row of the full-neighbour SELL structure (the sweep's, found through the node -> slot map sInv) and adds
// the words t = m_i & ~m_j split by the edge's current state into two vertical counters; cand = vs + U - C.
// Short rows are padded with the node's own id (t == 0).  Integer-exact, same decisions as the full count.
template <int P, bool SMEM>
__device__ __forceinline__ void delta_partial(const SweepView& sv, const uint16_t* sInv, const uint16_t* sF, int nf,
                                              const uint32_t* sP, const uint32_t* sX, int* sCnt) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int k0 = warp * 32; k0 < nf; k0 += kLSThreads) {            // warp-uniform
    const int k = k0 + lane;
    VCount<P> up, down;
    up.clear(), down.clear();
    if (k < nf) {
      const uint32_t i = sF[k];
      const uint32_t xi = sP[i], mi = xi ^ sX[i];
      const int slot = sInv[i], slice = slot >> 5;
      const int gb = SMEM ? sv.sell.off[slice] : __ldg(sv.sell.off + slice);
      const int nb = (SMEM ? sv.sell.off[slice + 1] : __ldg(sv.sell.off + slice + 1)) - gb;
      const uint2* col = reinterpret_cast<const uint2*>(sv.sell.col) + (int64_t)gb * 32 + (slot & 31);
      const uint32_t own = i | (i << 16);
      for (int b = 0; b < nb; b += 2) {
        const uint2 i0 = SMEM ? col[b * 32] : __ldg(col + b * 32);
        uint2 i1 = make_uint2(own, own);
        if (b + 1 < nb) i1 = SMEM ? col[(b + 1) * 32] : __ldg(col + (b + 1) * 32);
        uint32_t u[8], d[8];
        const uint32_t ids[8] = {i0.x & 0xffffu, i0.x >> 16, i0.y & 0xffffu, i0.y >> 16,
                                 i1.x & 0xffffu, i1.x >> 16, i1.y & 0xffffu, i1.y >> 16};
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const uint32_t xj = sP[ids[q]], t = mi & ~(xj ^ sX[ids[q]]), cut = xi ^ xj;
          u[q] = t & ~cut, d[q] = t & cut;
        }
        up.add8(u[0], u[1], u[2], u[3], u[4], u[5], u[6], u[7]);
        down.add8(d[0], d[1], d[2], d[3], d[4], d[5], d[6], d[7]);
      }
    }
    const int delta = up.flush_warp(lane) - down.flush_warp(lane);
    if (delta) atomicAdd(&sCnt[lane], delta);
  }
}

template <int P>
__device__ __forceinline__ void evaluate_delta_and_accept(const GraphDev& g, const LsArgs& a, const char* sSweep,
                                                          const uint16_t* sInv, const uint16_t* sF, int* sNF,
                                                          uint32_t* sP, const uint32_t* sX, int* sCnt,
                                                          uint32_t* sAccept, int valid, int64_t& my_vs) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nf = *sNF;
  if (a.stage_sweep)
    delta_partial<P, true>(sweep_view(g, sSweep), sInv, sF, nf, sP, sX, sCnt);
  else
    delta_partial<P, false>(sweep_view(g, g.sweep_blob), sInv, sF, nf, sP, sX, sCnt);
  ls_sync();
  if (warp == 0) {
    const int delta = sCnt[lane];
    const bool keep = lane < valid && delta >= 0;                  // vs' >= vs (util_read_data.py:199)
    if (keep) my_vs += delta;
    const unsigned acc = __ballot_sync(kFull, keep);
    if (lane == 0) *sAccept = acc, *sNF = 0;
    sCnt[lane] = 0;
  }
  ls_sync();
  const uint32_t acc = *sAccept;
  for (int k = threadIdx.x; k < nf; k += kLSThreads) {
    const uint32_t i = sF[k];
    sP[i] = (sX[i] & acc) | (sP[i] & ~acc);
  }
  ls_sync();
}

// Hand-shake with the streaming mask generator (noise_masks.cu): the draws of group g are complete when
// ctl[2g + 1] == units.  One thread polls (acquire at gpu scope), the CTA barrier that follows publishes it.
__device__ __forceinline__ bool gen_group_ready(const uint32_t* ctl, int group, uint32_t units) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctl + 2 * group + 1) : "memory");
  return v >= units;
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// bounded: a generator that never shows up (it would mean the two kernels were not scheduled together) must not hang
// the device; the stall is recorded and reported by rlsb_ls_fused_status
__device__ __forceinline__ void gen_group_wait(const uint32_t* ctl, int group, uint32_t units) {
  if (gen_group_ready(ctl, group, units)) return;
  if (*reinterpret_cast<const volatile uint32_t*>(ctl + kLsCtlStalled) != 0u) return;     // already given up
  const unsigned long long t0 = global_ns();
  while (!gen_group_ready(ctl, group, units)) {
    __nanosleep(200);
    if (global_ns() - t0 > kLsStallNs) {
      atomicAdd(const_cast<uint32_t*>(ctl) + kLsCtlStalled, 1u);
      return;
    }
  }
}

// The body of the bit-mask tile kernel, shared by its two entry points: ls_bits_kernel (alone on the SM: the
// compiler may use 128 registers) and ls_bits_fused_kernel (96 registers: a CTA of 512 threads leaves a quarter
// of the register file to the generator block that runs on the same SM, rlsb_ls_fused_search).
// FUSED: persistent over tiles (grid <= one CTA per SM, so that every CTA of this kernel and every block of the
// generator is resident at the same time whatever order the two grids are dispatched in -- a CTA that waits for a
// group of draws can never keep the generator's blocks from being scheduled), waits on the generator's counters,
// L2-coherent mask loads.  Not FUSED: one tile per CTA, masks complete before the launch.
template <int P, bool FUSED>
__device__ __forceinline__ void ls_bits_body(const GraphDev& g, const LsArgs& a, const uint32_t* __restrict__ masks,
                                             int64_t mask_words, int use_delta, const uint32_t* __restrict__ ctl,
                                             uint32_t units) {
  extern __shared__ __align__(1024) uint32_t smem[];
  uint32_t* sP = smem;
  uint32_t* sX = smem + g.np;
  uint16_t* sInv = reinterpret_cast<uint16_t*>(smem + 2 * g.np);     // node -> slot of the sweep structure
  uint16_t* sF = sInv + g.np;                                        // nodes flipped by the current candidate
  char* sSweep = reinterpret_cast<char*>(smem + 3 * g.np);
  __shared__ int sCnt[kTileEnvs];
  __shared__ uint32_t sAccept;
  __shared__ int sNF;
  __shared__ int sNextReady;
  __shared__ __align__(8) uint64_t sBar;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t tiles = (a.num_envs + kTileEnvs - 1) / kTileEnvs;
  if (threadIdx.x == 0) {
    mbar_init(&sBar, 1);
    sNF = 0;
    sNextReady = 0;
    if (a.stage_sweep) stage_sweep_blob(g, sSweep, &sBar);
  }
  if (use_delta && a.num_iters > 0) {
    const uint16_t* node = sweep_view(g, g.sweep_blob).sell.node;
    for (int slot = threadIdx.x; slot < g.num_sweep_slices * 32; slot += kLSThreads) {
      const uint32_t v = __ldg(node + slot);
      if (v != 0xFFFFu) sInv[v] = (uint16_t)slot;
    }
  }
  const int blocks = (g.n + 31) >> 5;
  uint16_t* flist = use_delta ? sF : nullptr;
  int tk = 0;
  auto process = [&](const int64_t tile) {
    const int64_t env0 = tile * kTileEnvs;
    const int valid = (int)min((int64_t)kTileEnvs, a.num_envs - env0);
    for (int i = threadIdx.x; i < g.np; i += kLSThreads) {
      const uint32_t w = a.packed[tile * g.np + i];
      sP[i] = w, sX[i] = w;
    }
    if (threadIdx.x < kTileEnvs) sCnt[threadIdx.x] = 0;
    int64_t my_vs = 0;
    if (warp == 0 && lane < valid) my_vs = a.vs[env0 + lane];
    if constexpr (FUSED)
      if (a.num_iters > 0 && threadIdx.x == 0) gen_group_wait(ctl, 0, units);     // the first group of draws
    ls_sync();
    const uint64_t row = (uint64_t)(env0 + lane) * (uint64_t)g.n;     // first bit of the lane's env row
    const bool live = lane < valid;
    stamp(a, tk);
    constexpr bool coh = FUSED;
    MaskChunk pre = mask_chunk_load(masks, row, warp, a.num_iters > 0 ? blocks : 0, live, coh);
    bool have_pre = true;
    if (use_delta && a.stage_sweep && a.num_iters > 0) mbar_wait(&sBar, 0);   // the neighbour lists have landed
    stamp(a, tk);
    for (int it = 0; it < a.num_iters; ++it) {
      const uint32_t* mask = masks + it * mask_words;
      if constexpr (FUSED) {
        if (!have_pre) {      // the draw was not complete when the previous iteration could have fetched ahead
          if (threadIdx.x == 0) gen_group_wait(ctl, it / kGenGroup, units);
          ls_sync();
          pre = mask_chunk_load(mask, row, warp, blocks, live, coh);
        }
      }
      mask_chunk_apply(pre, row, warp, g.n, sP, sX, flist, &sNF);
      for (int b0 = warp + 4 * kLSWarps; b0 < blocks; b0 += 4 * kLSWarps)
        mask_chunk_apply(mask_chunk_load(mask, row, b0, blocks, live, coh), row, b0, g.n, sP, sX, flist, &sNF);
      // fetch ahead only when the next draw's group is known to be complete (same group: it is; next group: one
      // poll, published by the barrier below)
      const bool more = it + 1 < a.num_iters;
      if constexpr (FUSED) {
        const bool same_group = (it + 1) / kGenGroup == it / kGenGroup;
        if (more && !same_group && threadIdx.x == 0)
          sNextReady = gen_group_ready(ctl, (it + 1) / kGenGroup, units) ? 1 : 0;
        ls_sync();
        have_pre = !more || same_group || sNextReady != 0;
        if (have_pre) pre = mask_chunk_load(mask + mask_words, row, warp, more ? blocks : 0, live, coh);
      } else {
        ls_sync();
        pre = mask_chunk_load(mask + mask_words, row, warp, more ? blocks : 0, live, coh);
      }
      stamp(a, tk);
      if (use_delta)
        evaluate_delta_and_accept<P>(g, a, sSweep, sInv, sF, &sNF, sP, sX, sCnt, &sAccept, valid, my_vs);
      else
        evaluate_and_accept(g, a, sP, sX, sCnt, &sAccept, valid, my_vs);
      stamp(a, tk);
    }
    const uint32_t vmask = valid == 32 ? kFull : ((1u << valid) - 1u);
    if (a.finish) {
      finish_tile<P>(g, a, sP, sSweep, &sBar, sCnt, tile, valid, &tk);
    } else {
      if (warp == 0 && lane < valid) a.vs[env0 + lane] = my_vs;
    }
    stamp(a, tk);
    for (int i = threadIdx.x; i < g.np; i += kLSThreads) a.packed[tile * g.np + i] = sP[i] & vmask;
  };
  if constexpr (FUSED) {
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
      process(tile);
      ls_sync();          // the tile copies are free for the next tile
    }
  } else {
    process(blockIdx.x);
  }
}

template <int P>
__global__ void __launch_bounds__(kLSThreads) ls_bits_kernel(GraphDev g, LsArgs a, const uint32_t* __restrict__ masks,
                                                             int64_t mask_words, int use_delta) {
  ls_bits_body<P, false>(g, a, masks, mask_words, use_delta, nullptr, 0u);
}
template <int P>
__global__ void __maxnreg__(96) ls_bits_fused_kernel(GraphDev g, LsArgs a, const uint32_t* __restrict__ masks,
                                                     int64_t mask_words, int use_delta, const uint32_t* __restrict__ ctl,
                                                     uint32_t units) {
  ls_bits_body<P, true>(g, a, masks, mask_words, use_delta, ctl, units);
}

// single-flip pass on packed tiles only (rlsb_flip_sweep)
template <int P>
__global__ void __launch_bounds__(kLSThreads) flip_sweep_kernel(GraphDev g, uint32_t* __restrict__ packed,
                                                                int64_t* __restrict__ vs, int64_t num_envs,
                                                                int cut_warps, int sweep_warps, int stage) {
  extern __shared__ __align__(1024) uint32_t sP[];
  __shared__ int sCnt[kTileEnvs];
  __shared__ __align__(8) uint64_t sBar;
  char* sSweep = reinterpret_cast<char*>(sP + g.np);       // stage: the sweep structure, copied once per CTA (TMA)
  const int lane = threadIdx.x & 31;
  const int64_t tiles = (num_envs + kTileEnvs - 1) / kTileEnvs;
  if (stage) {
    if (threadIdx.x == 0) {
      mbar_init(&sBar, 1);
      stage_sweep_blob(g, sSweep, &sBar);
    }
    __syncthreads();
  }
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t env0 = tile * kTileEnvs;
    const int valid = (int)min((int64_t)kTileEnvs, num_envs - env0);
    const uint32_t vmask = valid == 32 ? kFull : ((1u << valid) - 1u);
    for (int i = threadIdx.x; i < g.np; i += blockDim.x) sP[i] = packed[tile * g.np + i];
    if (threadIdx.x < kTileEnvs) sCnt[threadIdx.x] = 0;
    __syncthreads();
    if (stage) {
      mbar_wait(&sBar, 0);
      sweep_tile<P, true>(g, sweep_view(g, sSweep), sP, sweep_warps);
    } else {
      sweep_tile<P, false>(g, sweep_view(g, g.sweep_blob), sP, sweep_warps);
    }
    __syncthreads();
    const int cnt = tile_cut_partial(g, sP, cut_warps);
    if (cnt) atomicAdd(&sCnt[lane], cnt);
    __syncthreads();
    if (threadIdx.x < valid) vs[env0 + threadIdx.x] = sCnt[threadIdx.x];
    for (int i = threadIdx.x; i < g.np; i += blockDim.x) packed[tile * g.np + i] = sP[i] & vmask;
    __syncthreads();
  }
}

template <typename K>
static int allow_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024)
    RLSB_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return RLSB_OK;
}

static int sweep_warps_for(const GraphDev& g, int nwarps) {
  const int w = g.max_level_slices < 1 ? 1 : g.max_level_slices;
  return w < nwarps ? w : nwarps;
}

constexpr size_t kSmemBudget = 220 * 1024;   // dynamic shared memory a search CTA may use

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// float32 [E][N] row-major seen as {16 floats, E rows, N/16 column groups}; box = 16 x 32 x kChunkG16: ONE copy
// per 128-node chunk (the producer thread retires a TMA instruction only every ~100 cycles, so the number of
// copies per chunk, not their size, bounds the ring's fill rate).  Rows / groups past the tensor read as zeros.
static int make_noise_map(CUtensorMap* map, const float* base, int64_t num_envs, int n) {
  EncodeTiledFn fn = encode_tiled_fn();
  RLSB_REQUIRE(fn != nullptr, RLSB_ERR_CUDA, "ls_run: cuTensorMapEncodeTiled is not available from this driver");
  const cuuint64_t dims[3] = {16, (cuuint64_t)num_envs, (cuuint64_t)(n / 16)};
  const cuuint64_t strides[2] = {(cuuint64_t)n * 4, 64};
  const cuuint32_t box[3] = {16, (cuuint32_t)kTileEnvs, (cuuint32_t)kChunkG16};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  RLSB_REQUIRE(r == CUDA_SUCCESS, RLSB_ERR_CUDA, "ls_run: cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return RLSB_OK;
}

// cross counts [tiles * Np/4 groups][32 envs x 4 counts] -> rows of 128 (uint8) / 256 (uint16) bytes, box = 32 groups;
// rd_std / degree words [Np] -> 1-D, box = 128 elements
static int make_side_maps(TmapPack* maps, const GraphDev& g, const LsArgs& a, size_t cross_elt) {
  EncodeTiledFn fn = encode_tiled_fn();
  RLSB_REQUIRE(fn != nullptr, RLSB_ERR_CUDA, "ls_run: cuTensorMapEncodeTiled is not available from this driver");
  const int64_t tiles = (a.num_envs + kTileEnvs - 1) / kTileEnvs;
  const cuuint32_t row = (cuuint32_t)(kTileEnvs * 4 * cross_elt);
  const cuuint64_t cdims[2] = {row, (cuuint64_t)(tiles * (g.np / 4))};
  const cuuint64_t cstr[1] = {row};
  const cuuint32_t cbox[2] = {row, (cuuint32_t)kChunkGroups};
  const cuuint32_t one[2] = {1, 1};
  CUresult r = fn(&maps->cross, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(a.cross), cdims, cstr, cbox, one,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  RLSB_REQUIRE(r == CUDA_SUCCESS, RLSB_ERR_CUDA, "ls_run: tensor map for the cross counts failed (CUresult %d)", (int)r);
  const cuuint64_t ndims[1] = {(cuuint64_t)2 * g.np};
  const cuuint32_t nbox[1] = {(cuuint32_t)(kChunkGroups * 8)};
  r = fn(&maps->nd, CU_TENSOR_MAP_DATA_TYPE_UINT32, 1, const_cast<uint32_t*>(a.nd), ndims, cstr, nbox, one,
         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  RLSB_REQUIRE(r == CUDA_SUCCESS, RLSB_ERR_CUDA, "ls_run: tensor map for the node words failed (CUresult %d)", (int)r);
  return RLSB_OK;
}

template <int P, typename CrossT, bool THRESH>
static int launch_pipe(const GraphDev& g, LsArgs a, const TmapPack& maps, cudaStream_t st) {
  a.stage_bytes = kNoiseStageBytes + kChunkGroups * kTileEnvs * 4 * (int)sizeof(CrossT) + 1024;   // + rd_std, degm
  const int passes = (a.thresh_noise ? 1 : 0) + a.num_iters;
  const int chunks = passes * ((g.n / 4 + kChunkGroups - 1) / kChunkGroups);
  a.ring_off = (int)((2 * (size_t)g.np * sizeof(uint32_t) + 1023) / 1024 * 1024);
  const size_t tiles_bytes = a.ring_off + (THRESH ? kSelBytes : 0);
  int stages = (int)((kSmemBudget - tiles_bytes) / a.stage_bytes);
  stages = stages > kMaxStages ? kMaxStages : stages;
  stages = stages > chunks ? (chunks > 0 ? chunks : 1) : stages;
  a.stages = stages;
  size_t ring = (size_t)stages * a.stage_bytes;
  a.stage_sweep = (a.finish && tiles_bytes + (size_t)g.sweep_blob_bytes <= kSmemBudget) ? 1 : 0;
  if (a.stage_sweep && (size_t)g.sweep_blob_bytes > ring) ring = (size_t)g.sweep_blob_bytes;
  const size_t smem = tiles_bytes + ring;      // the selection buffers sit behind the ring stages
  const int64_t tiles = (a.num_envs + kTileEnvs - 1) / kTileEnvs;
  if (int rc = allow_smem(ls_pipe_kernel<P, CrossT, THRESH>, smem)) return rc;
  ls_pipe_kernel<P, CrossT, THRESH><<<(unsigned)tiles, kPipeThreads, smem, st>>>(g, a, maps);
  RLSB_LAUNCH_OK();
  return RLSB_OK;
}

template <int P, typename CrossT>
static int launch_pipe_k(const GraphDev& g, const LsArgs& a, const TmapPack& maps, cudaStream_t st) {
  if (!a.thresh_noise) return launch_pipe<P, CrossT, false>(g, a, maps, st);
  return launch_pipe<P, CrossT, true>(g, a, maps, st);
}

template <int P, typename CrossT>
static int launch_generic(const GraphDev& g, LsArgs a, cudaStream_t st) {
  const size_t tiles_bytes = 2 * (size_t)g.np * sizeof(uint32_t);
  a.stage_sweep = (a.finish && tiles_bytes + (size_t)g.sweep_blob_bytes <= kSmemBudget) ? 1 : 0;
  const size_t smem = tiles_bytes + (a.stage_sweep ? (size_t)g.sweep_blob_bytes : 0);
  const int64_t tiles = (a.num_envs + kTileEnvs - 1) / kTileEnvs;
  if (int rc = allow_smem(ls_generic_kernel<P, CrossT>, smem)) return rc;
  ls_generic_kernel<P, CrossT><<<(unsigned)tiles, kLSThreads, smem, st>>>(g, a);
  RLSB_LAUNCH_OK();
  return RLSB_OK;
}

// smem_budget: the fused search leaves room for the generator block on the same SM
template <int P>
static int launch_bits(const GraphDev& g, LsArgs a, const uint32_t* masks, cudaStream_t st, const uint32_t* ctl = nullptr,
                       uint32_t units = 0, size_t smem_budget = kSmemBudget) {
  const size_t tiles_bytes = 3 * (size_t)g.np * sizeof(uint32_t);     // two tile copies + node->slot map + flipped list
  a.stage_sweep = ((a.finish || a.num_iters > 0) && tiles_bytes + (size_t)g.sweep_blob_bytes <= smem_budget) ? 1 : 0;
  const size_t smem = tiles_bytes + (a.stage_sweep ? (size_t)g.sweep_blob_bytes : 0);
  RLSB_REQUIRE(smem <= smem_budget, RLSB_ERR_UNSUPPORTED, "ls_run_masks: %d nodes exceed the shared-memory tile", g.n);
  const int64_t tiles = (a.num_envs + kTileEnvs - 1) / kTileEnvs;
  // slots of the sweep structure are addressed with 16 bits; RLSB_DEBUG_FULL_CUT keeps the full re-count (cross-check)
  const int use_delta = (g.num_sweep_slices * 32 <= 65536 && !(debug_flags() & RLSB_DEBUG_FULL_CUT)) ? 1 : 0;
  const int64_t words = ls_mask_words(a.num_envs, g.n);
  if (ctl) {
    const unsigned grid = (unsigned)(tiles < kNumSMs ? tiles : kNumSMs);      // persistent: at most one CTA per SM
    if (int rc = allow_smem(ls_bits_fused_kernel<P>, smem)) return rc;
    if (!(debug_flags() & RLSB_DEBUG_CARVEOUT_DEFAULT))      // same carve-out as the generator (see noise_masks.cu)
      RLSB_CUDA_OK(cudaFuncSetAttribute(ls_bits_fused_kernel<P>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                        cudaSharedmemCarveoutMaxShared));
    ls_bits_fused_kernel<P><<<grid, kLSThreads, smem, st>>>(g, a, masks, words, use_delta, ctl, units);
  } else {
    if (int rc = allow_smem(ls_bits_kernel<P>, smem)) return rc;
    ls_bits_kernel<P><<<(unsigned)tiles, kLSThreads, smem, st>>>(g, a, masks, words, use_delta);
  }
  RLSB_LAUNCH_OK();
  return RLSB_OK;
}

template <typename CrossT>
static int launch_thresh(const GraphDev& g, const LsWorkspace& w, int mult, const float* noise, int kth_big,
                         int64_t num_envs, cudaStream_t st) {
  const unsigned grid = (unsigned)((num_envs + 7) / 8);
  const CrossT* cross = static_cast<const CrossT*>(w.cross);
  if (kth_big <= 10)
    ls_thresh_kernel<10, CrossT><<<grid, 256, 0, st>>>(g, cross, w.rd_std, w.degm, mult, noise, kth_big, num_envs,
                                                       w.thresh);
  else
    ls_thresh_kernel<32, CrossT><<<grid, 256, 0, st>>>(g, cross, w.rd_std, w.degm, mult, noise, kth_big, num_envs,
                                                       w.thresh);
  RLSB_LAUNCH_OK();
  return RLSB_OK;
}

// RLSB_LS_TIMES=1: CTA 0 of every pipe launch writes clock64 phase stamps into a device buffer
// that rlsb_ls_debug_times() copies out (profiling aid, tools/ls_phase_times.py)
static long long* ls_debug_times() {
  static long long* bufs[64] = {};      // one buffer per device (the kernels run on the current device)
  if (!(debug_flags() & RLSB_DEBUG_LS_TIMES)) return nullptr;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  if (!bufs[dev]) {
    if (cudaMalloc(&bufs[dev], 64 * sizeof(long long)) == cudaSuccess)
      cudaMemset(bufs[dev], 0, 64 * sizeof(long long));
    else
      bufs[dev] = nullptr;
  }
  return bufs[dev];
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// thresh (optional) + iterations + finish (optional), in launches of at most kLSMaxIters tensors
static int run_search(const GraphDev& g, int64_t num_envs, int64_t* vs, int ws_mult, const float* thresh_noise,
                      int num_spin, const float* const* h_noise_ptrs, int num_iters, int finish, uint8_t* xs_out,
                      void* workspace, cudaStream_t st) {
  const LsWorkspace w = carve(g, num_envs, workspace);
  const int dc = degree_class(g);
  const int kth_big = num_spin + 1;   // kth smallest with k = N - num_spin  ==  (num_spin+1)-th largest
  // the pipelined kernel needs rows made of whole 64-byte column groups (tensor-map dim 0) and room for
  // two ring stages next to the tile copies
  bool pipe = g.n % 16 == 0 &&
              2 * (size_t)g.np * 4 + 1024 + kSelBytes + 2 * (size_t)(kNoiseStageBytes + 8192 + 1024) <= kSmemBudget;
  if (thresh_noise) pipe = pipe && aligned16(thresh_noise);
  for (int k = 0; k < num_iters; ++k) {
    RLSB_REQUIRE(h_noise_ptrs[k] != nullptr, RLSB_ERR_INVALID, "ls_search: null noise tensor %d", k);
    pipe = pipe && aligned16(h_noise_ptrs[k]);
  }
  // the threshold on its own (what the fused-RNG path asks for): one warp per env over the row-major counts
  if (thresh_noise && num_iters == 0 && !finish && w.cross_rows && dc != 2 && g.n % 4 == 0 && aligned16(thresh_noise) &&
      kth_big <= 10 && !(debug_flags() & RLSB_DEBUG_THRESH_PIPE)) {
    ls_thresh_rows_kernel<10><<<(unsigned)((num_envs + 3) / 4), 128, 0, st>>>(g.n, g.np, w.cross_rows, w.rd_std, w.degm,
                                                                            ws_mult, thresh_noise, kth_big, num_envs,
                                                                            w.thresh);
    RLSB_LAUNCH_OK();
    return RLSB_OK;
  }
  if (thresh_noise && !pipe) {
    if (int rc = dc == 2 ? launch_thresh<uint16_t>(g, w, ws_mult, thresh_noise, kth_big, num_envs, st)
                         : launch_thresh<uint8_t>(g, w, ws_mult, thresh_noise, kth_big, num_envs, st))
      return rc;
    thresh_noise = nullptr;
    if (num_iters == 0 && !finish) return RLSB_OK;
  }
  int done = 0;
  do {
    const int now = num_iters - done < kLSMaxIters ? num_iters - done : kLSMaxIters;
    LsArgs a{};
    a.packed = w.packed, a.vs = vs, a.cross = w.cross, a.rd_std = w.rd_std, a.degm = w.degm, a.thresh = w.thresh;
    a.nd = w.nd;
    a.thresh_noise = done == 0 ? thresh_noise : nullptr;
    a.kth_big = kth_big;
    for (int k = 0; k < now; ++k) a.noise.p[k] = h_noise_ptrs[done + k];
    a.num_iters = now, a.mult = ws_mult, a.num_envs = num_envs;
    a.finish = (finish && done + now == num_iters) ? 1 : 0;
    a.xs_out = xs_out, a.unpack_vec4 = (xs_out && rows_vec4_ok(xs_out, g.n)) ? 1 : 0;
    a.cut_warps = cut_warps_for(g.m, kLSWarps);
    a.sweep_warps = sweep_warps_for(g, kLSWarps), a.negmult = -ws_mult;
    a.times = ls_debug_times();
    {
      a.skip = (debug_flags() & RLSB_DEBUG_LS_SKIP) ? 1 : 0;
    }
    int rc;
    if (pipe) {
      TmapPack maps;
      memset(&maps, 0, sizeof(maps));
      if (a.thresh_noise)
        if ((rc = make_noise_map(&maps.m[0], a.thresh_noise, num_envs, g.n))) return rc;
      for (int k = 0; k < now; ++k)
        if ((rc = make_noise_map(&maps.m[1 + k], a.noise.p[k], num_envs, g.n))) return rc;
      if ((rc = make_side_maps(&maps, g, a, dc == 2 ? 2 : 1))) return rc;
      rc = dc == 0   ? launch_pipe_k<6, uint8_t>(g, a, maps, st)
           : dc == 1 ? launch_pipe_k<8, uint8_t>(g, a, maps, st)
                     : launch_pipe_k<12, uint16_t>(g, a, maps, st);
    } else
      rc = dc == 0   ? launch_generic<6, uint8_t>(g, a, st)
           : dc == 1 ? launch_generic<8, uint8_t>(g, a, st)
                     : launch_generic<12, uint16_t>(g, a, st);
    if (rc) return rc;
    done += now;
  } while (done < num_iters);
  return RLSB_OK;
}

}  // namespace rlsb

extern "C" {

int64_t rlsb_ls_workspace_bytes(const rlsb_graph_t* gh, int64_t num_envs) {
  using namespace rlsb;
  const GraphDev* g;
  if (graph_check(gh, &g, "ls_workspace_bytes") || num_envs < 0) return -1;
  return (int64_t)carve(*g, num_envs, nullptr).bytes;
}

int64_t rlsb_ls_workspace_offset(const rlsb_graph_t* gh, int64_t num_envs, int32_t section) {
  using namespace rlsb;
  const GraphDev* g;
  if (graph_check(gh, &g, "ls_workspace_offset") || num_envs < 0) return -1;
  char* base = reinterpret_cast<char*>(uintptr_t(4096));
  const LsWorkspace w = carve(*g, num_envs, base);
  const void* at[9] = {w.packed, w.cross, w.col_min, w.col_max, w.degm, w.rd_std, w.thresh, w.cross_rows, w.ctl};
  if (section < 0 || section > 8 || !at[section]) return -1;
  return static_cast<const char*>(at[section]) - base;
}

int rlsb_ls_begin(const rlsb_graph_t* gh, const uint8_t* xs, int64_t num_envs, int64_t* vs, int32_t compute_vs,
                  int32_t ws_mult, float noise_std, void* workspace, void* stream) {
  using namespace rlsb;
  const GraphDev* g;
  if (int rc = graph_check(gh, &g, "ls_begin")) return rc;
  RLSB_REQUIRE(num_envs >= 0, RLSB_ERR_INVALID, "ls_begin: negative num_envs");
  RLSB_REQUIRE(ws_mult == 1 || ws_mult == 2, RLSB_ERR_INVALID, "ls_begin: ws_mult must be 1 or 2");
  RLSB_REQUIRE(2 * (size_t)g->np * 4 <= kSmemBudget, RLSB_ERR_UNSUPPORTED,
               "ls_begin: %d nodes exceed the two-copy shared-memory tile of the search kernel", g->n);
  if (num_envs == 0 || g->n == 0) return RLSB_OK;
  RLSB_REQUIRE(xs && workspace && (vs || !compute_vs), RLSB_ERR_INVALID, "ls_begin: null pointer");
  RLSB_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255u) == 0, RLSB_ERR_INVALID,
               "ls_begin: workspace must be 256-byte aligned");
  auto st = static_cast<cudaStream_t>(stream);
  const LsWorkspace w = carve(*g, num_envs, workspace);
  if (int rc = prepare_tiles(*g, xs, nullptr, num_envs, w.packed, w.cross, degree_class(*g) != 2 ? 1 : 2, w.cross_rows,
                             w.col_min, w.col_max, compute_vs ? vs : nullptr, st))
    return rc;
  ls_rdstd_kernel<<<(g->np + 255) / 256, 256, 0, st>>>(*g, w.col_min, w.col_max, ws_mult, noise_std, w.rd_std, w.degm,
                                                       w.nd);
  RLSB_LAUNCH_OK();
  return RLSB_OK;
}

static int ls_check(const rlsb_graph_t* gh, const rlsb::GraphDev** g, const char* what, int64_t num_envs,
                    int32_t ws_mult) {
  using namespace rlsb;
  if (int rc = graph_check(gh, g, what)) return rc;
  RLSB_REQUIRE(num_envs >= 0, RLSB_ERR_INVALID, "%s: negative num_envs", what);
  RLSB_REQUIRE(ws_mult == 1 || ws_mult == 2, RLSB_ERR_INVALID, "%s: ws_mult must be 1 or 2", what);
  RLSB_REQUIRE(2 * (size_t)(*g)->np * 4 <= kSmemBudget, RLSB_ERR_UNSUPPORTED,
               "%s: %d nodes exceed the two-copy shared-memory tile", what, (*g)->n);
  return RLSB_OK;
}

static int spin_check(const rlsb::GraphDev& g, int32_t num_spin, const char* what) {
  using namespace rlsb;
  // torch.kthvalue(k = N - num_spin) needs 1 <= k <= N
  RLSB_REQUIRE(num_spin >= 0 && num_spin < g.n, RLSB_ERR_INVALID,
               "%s: k = N - num_spin = %d out of range for N = %d (torch.kthvalue raises)", what, g.n - num_spin, g.n);
  RLSB_REQUIRE(num_spin + 1 <= 32, RLSB_ERR_UNSUPPORTED, "%s: num_spin %d above the in-register limit 31", what,
               num_spin);
  return RLSB_OK;
}

int rlsb_ls_thresh(const rlsb_graph_t* gh, int64_t num_envs, int32_t ws_mult, const float* noise, int32_t num_spin,
                   void* workspace, void* stream) {
  using namespace rlsb;
  const GraphDev* g;
  if (int rc = ls_check(gh, &g, "ls_thresh", num_envs, ws_mult)) return rc;
  if (int rc = spin_check(*g, num_spin, "ls_thresh")) return rc;
  if (num_envs == 0) return RLSB_OK;
  RLSB_REQUIRE(noise && workspace, RLSB_ERR_INVALID, "ls_thresh: null pointer");
  const LsWorkspace w = carve(*g, num_envs, workspace);
  auto st = static_cast<cudaStream_t>(stream);
  return degree_class(*g) == 2 ? launch_thresh<uint16_t>(*g, w, ws_mult, noise, num_spin + 1, num_envs, st)
                               : launch_thresh<uint8_t>(*g, w, ws_mult, noise, num_spin + 1, num_envs, st);
}

int rlsb_ls_run(const rlsb_graph_t* gh, int64_t num_envs, int64_t* vs, int32_t ws_mult, const float* thresh_noise,
                int32_t num_spin, const float* const* h_noise_ptrs, int32_t num_iters, int32_t finish,
                uint8_t* xs_out, void* workspace, void* stream) {
  using namespace rlsb;
  const GraphDev* g;
  if (int rc = ls_check(gh, &g, "ls_run", num_envs, ws_mult)) return rc;
  RLSB_REQUIRE(num_iters >= 0, RLSB_ERR_INVALID, "ls_run: negative num_iters");
  if (thresh_noise)
    if (int rc = spin_check(*g, num_spin, "ls_run")) return rc;
  if (num_envs == 0 || g->n == 0) return RLSB_OK;
  RLSB_REQUIRE(vs && workspace && (num_iters == 0 || h_noise_ptrs) && (!finish || xs_out), RLSB_ERR_INVALID,
               "ls_run: null pointer");
  return run_search(*g, num_envs, vs, ws_mult, thresh_noise, thresh_noise ? num_spin : 0, h_noise_ptrs, num_iters,
                    finish, xs_out, workspace, static_cast<cudaStream_t>(stream));
}

int rlsb_ls_run_masks(const rlsb_graph_t* gh, int64_t num_envs, int64_t* vs, const uint32_t* masks, int32_t num_iters,
                      int32_t finish, uint8_t* xs_out, void* workspace, void* stream) {
  using namespace rlsb;
  const GraphDev* g;
  if (int rc = ls_check(gh, &g, "ls_run_masks", num_envs, 1)) return rc;
  RLSB_REQUIRE(num_iters >= 0, RLSB_ERR_INVALID, "ls_run_masks: negative num_iters");
  if (num_envs == 0 || g->n == 0) return RLSB_OK;
  RLSB_REQUIRE(vs && workspace && (num_iters == 0 || masks), RLSB_ERR_INVALID, "ls_run_masks: null pointer");
  const LsWorkspace w = carve(*g, num_envs, workspace);
  LsArgs a{};
  a.packed = w.packed, a.vs = vs, a.num_iters = num_iters, a.num_envs = num_envs;
  a.finish = finish ? 1 : 0, a.xs_out = xs_out, a.unpack_vec4 = (xs_out && rows_vec4_ok(xs_out, g->n)) ? 1 : 0;
  a.cut_warps = cut_warps_for(g->m, kLSWarps), a.sweep_warps = sweep_warps_for(*g, kLSWarps);
  a.times = ls_debug_times();
  auto st = static_cast<cudaStream_t>(stream);
  const int dc = degree_class(*g);
  return dc == 0 ? launch_bits<6>(*g, a, masks, st) : dc == 1 ? launch_bits<8>(*g, a, masks, st)
                                                              : launch_bits<12>(*g, a, masks, st);
}

// Noisy iterations + single-flip pass with the generator running NEXT TO the tile kernel: the streaming mask
// generator (noise_masks.cu) on the graph's side stream, the tile kernel on the caller's stream, coupled through
// per-group counters in the workspace.  Both kernels fit on an SM together (96 + 64 registers per thread, the
// generator's 5 KB of shared memory), the generator never waits for a tile CTA, so any interleaving the
// hardware picks makes progress.  Same masks, same decisions, same results as rlsb_ls_noise_masks followed by
// rlsb_ls_run_masks.
int rlsb_ls_fused_search(const rlsb_graph_t* gh, int64_t num_envs, int64_t* vs, int32_t ws_mult, uint64_t seed,
                         uint64_t offset, const uint64_t* rng_dev, int32_t rng_threads, int32_t rng_iters,
                         int32_t num_iters, int32_t finish, uint8_t* xs_out, uint32_t* masks, void* workspace,
                         void* stream) {
  using namespace rlsb;
  const GraphDev* g;
  if (int rc = ls_check(gh, &g, "ls_fused_search", num_envs, ws_mult)) return rc;
  RLSB_REQUIRE(num_iters >= 0 && num_iters <= kLsMaxFusedDraws, RLSB_ERR_INVALID,
               "ls_fused_search: num_iters must be in [0, %d]", kLsMaxFusedDraws);
  if (num_envs == 0 || g->n == 0) return RLSB_OK;
  RLSB_REQUIRE(vs && workspace && (num_iters == 0 || masks), RLSB_ERR_INVALID, "ls_fused_search: null pointer");
  auto st = static_cast<cudaStream_t>(stream);
  const LsWorkspace w = carve(*g, num_envs, workspace);
  LsArgs a{};
  a.packed = w.packed, a.vs = vs, a.num_iters = num_iters, a.num_envs = num_envs;
  a.finish = finish ? 1 : 0, a.xs_out = xs_out, a.unpack_vec4 = (xs_out && rows_vec4_ok(xs_out, g->n)) ? 1 : 0;
  a.cut_warps = cut_warps_for(g->m, kLSWarps), a.sweep_warps = sweep_warps_for(*g, kLSWarps);
  a.times = ls_debug_times();
  const int dc = degree_class(*g);
  const uint32_t* ctl = nullptr;
  uint32_t units = 0;
  GraphSide side{};
  MaskPlan plan;
  if (num_iters > 0) {
    if (int rc = graph_side(gh, &side)) return rc;
    if (int rc = mask_plan(*g, "ls_fused_search", num_envs, ws_mult, seed, offset, rng_dev, rng_threads, rng_iters,
                           num_iters, masks, workspace, &plan))
      return rc;
    if (int rc = mask_stream_preload(plan, st)) return rc;
    if (int rc = mask_prepare(plan, true, true, st)) return rc;
    RLSB_CUDA_OK(cudaEventRecord(side.fork, st));
    ctl = plan.ctl, units = (uint32_t)rng_threads / 256u;
  }
  // the tile kernel first: its CTAs take their SMs, the generator blocks fill in beside them
  constexpr size_t kFusedSmem = 212 * 1024;
  if (int rc = dc == 0 ? launch_bits<6>(*g, a, masks, st, ctl, units, kFusedSmem)
               : dc == 1 ? launch_bits<8>(*g, a, masks, st, ctl, units, kFusedSmem)
                         : launch_bits<12>(*g, a, masks, st, ctl, units, kFusedSmem))
    return rc;
  if (num_iters > 0) {
    RLSB_CUDA_OK(cudaStreamWaitEvent(side.stream, side.fork, 0));
    if (int rc = mask_stream_launch(plan, side.stream)) return rc;
    RLSB_CUDA_OK(cudaEventRecord(side.join, side.stream));
    RLSB_CUDA_OK(cudaStreamWaitEvent(st, side.join, 0));
  }
  return RLSB_OK;
}

// diagnostics of the last fused search on this workspace: out3 = {generator blocks that started, tile CTAs whose wait
// for a group of draws ran out (non-zero: the results of that call are invalid), units finished of the first group}
int rlsb_ls_fused_status(const rlsb_graph_t* gh, int64_t num_envs, const void* workspace, uint32_t* h_out3, void* stream) {
  using namespace rlsb;
  const GraphDev* g;
  if (int rc = graph_check(gh, &g, "ls_fused_status")) return rc;
  RLSB_REQUIRE(workspace && h_out3, RLSB_ERR_INVALID, "ls_fused_status: null pointer");
  const LsWorkspace w = carve(*g, num_envs, const_cast<void*>(workspace));
  auto st = static_cast<cudaStream_t>(stream);
  uint32_t tmp[3];
  RLSB_CUDA_OK(cudaMemcpyAsync(&tmp[0], w.ctl + kLsCtlStarted, 8, cudaMemcpyDeviceToHost, st));
  RLSB_CUDA_OK(cudaMemcpyAsync(&tmp[2], w.ctl + 1, 4, cudaMemcpyDeviceToHost, st));
  RLSB_CUDA_OK(cudaStreamSynchronize(st));
  h_out3[0] = tmp[0], h_out3[1] = tmp[1], h_out3[2] = tmp[2];
  return RLSB_OK;
}

// rlsb_ls_begin with the state given as packed tiles (uint32 [ceil(E/32)][Np], bit b of word [t][i] = node i of
// env 32t + b): the layout a host that keeps its spins packed hands over (1 bit instead of 1 byte per spin on
// the wire).  packed_in may be the workspace's own packed section (rlsb_ls_workspace_offset(.., 0)): no copy.
int rlsb_ls_begin_packed(const rlsb_graph_t* gh, const uint32_t* packed_in, int64_t num_envs, int64_t* vs,
                         int32_t compute_vs, int32_t ws_mult, float noise_std, void* workspace, void* stream) {
  using namespace rlsb;
  const GraphDev* g;
  if (int rc = ls_check(gh, &g, "ls_begin_packed", num_envs, ws_mult)) return rc;
  if (num_envs == 0 || g->n == 0) return RLSB_OK;
  RLSB_REQUIRE(packed_in && workspace && (vs || !compute_vs), RLSB_ERR_INVALID, "ls_begin_packed: null pointer");
  RLSB_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255u) == 0, RLSB_ERR_INVALID,
               "ls_begin_packed: workspace must be 256-byte aligned");
  auto st = static_cast<cudaStream_t>(stream);
  const LsWorkspace w = carve(*g, num_envs, workspace);
  const int64_t tiles = (num_envs + kTileEnvs - 1) / kTileEnvs;
  if (packed_in != w.packed)
    RLSB_CUDA_OK(cudaMemcpyAsync(w.packed, packed_in, (size_t)tiles * g->np * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
  if (int rc = prepare_tiles(*g, nullptr, w.packed, num_envs, nullptr, w.cross, degree_class(*g) != 2 ? 1 : 2,
                             w.cross_rows, w.col_min, w.col_max, compute_vs ? vs : nullptr, st))
    return rc;
  ls_rdstd_kernel<<<(g->np + 255) / 256, 256, 0, st>>>(*g, w.col_min, w.col_max, ws_mult, noise_std, w.rd_std, w.degm,
                                                       w.nd);
  RLSB_LAUNCH_OK();
  return RLSB_OK;
}

int rlsb_ls_debug_times(int64_t* out64) {
  long long* buf = rlsb::ls_debug_times();
  if (!buf) return RLSB_ERR_INVALID;
  RLSB_CUDA_OK(cudaMemcpy(out64, buf, 64 * sizeof(long long), cudaMemcpyDeviceToHost));
  return RLSB_OK;
}

int rlsb_ls_search(const rlsb_graph_t* gh, int64_t num_envs, int64_t* vs, int32_t ws_mult,
                   const float* const* h_noise_ptrs, int32_t num_iters, int32_t finish, uint8_t* xs_out,
                   void* workspace, void* stream) {
  return rlsb_ls_run(gh, num_envs, vs, ws_mult, nullptr, 0, h_noise_ptrs, num_iters, finish, xs_out, workspace, stream);
}

int rlsb_flip_sweep(const rlsb_graph_t* gh, uint32_t* packed, int64_t* vs, int64_t num_envs, void* stream) {
  using namespace rlsb;
  const GraphDev* g;
  if (int rc = graph_check(gh, &g, "flip_sweep")) return rc;
  RLSB_REQUIRE(num_envs >= 0, RLSB_ERR_INVALID, "flip_sweep: negative num_envs");
  if (num_envs == 0 || g->n == 0) return RLSB_OK;
  RLSB_REQUIRE(packed && vs, RLSB_ERR_INVALID, "flip_sweep: null pointer");
  const int64_t tiles = (num_envs + kTileEnvs - 1) / kTileEnvs;
  // few tiles (a level is latency bound): the sweep structure is staged into shared memory, one CTA per SM; many
  // tiles: several CTAs per SM hide the L2 latency of the structure instead
  const size_t tile_bytes = (size_t)g->np * sizeof(uint32_t);
  const int stage = (tiles <= 2 * kNumSMs && tile_bytes + (size_t)g->sweep_blob_bytes <= kSmemBudget) ? 1 : 0;
  const size_t smem = tile_bytes + (stage ? (size_t)g->sweep_blob_bytes : 0);
  const int64_t max_ctas = stage ? kNumSMs : 4 * kNumSMs;
  const unsigned grid = (unsigned)(tiles < max_ctas ? tiles : max_ctas);
  auto st = static_cast<cudaStream_t>(stream);
  const int cw = cut_warps_for(g->m, kLSThreads / 32);
  const int dc = g->max_full_deg <= 63 ? 0 : g->max_full_deg <= 255 ? 1 : 2;
  int rc;
#define RLSB_SWEEP(P)                                                \
  if ((rc = allow_smem(flip_sweep_kernel<P>, smem))) return rc;      \
  flip_sweep_kernel<P><<<grid, kLSThreads, smem, st>>>(*g, packed, vs, num_envs, cw, sweep_warps_for(*g, kLSThreads / 32), stage)
  if (dc == 0) { RLSB_SWEEP(6); } else if (dc == 1) { RLSB_SWEEP(8); } else { RLSB_SWEEP(12); }
#undef RLSB_SWEEP
  RLSB_LAUNCH_OK();
  return RLSB_OK;
}

}  // extern "C"
