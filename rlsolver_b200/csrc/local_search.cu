// Local search: the three phases of EnvMaxcut.local_search_inplace
// (rlsolver/envs/env_L2A.py:87-116) and LocalSearch.random_search
// (rlsolver/methods/LocalSearch.py:53-86) as four launches:
//
//   ls_begin   (cross_counts.cu prepare_kernel) pack + objective + per-node cross counts + their
//              max/min over the env batch (the `ws_std` coupling, env_L2A.py:92-93)
//   ls_rdstd   rd_std[i] = float(mult * (max_i - min_i)) * noise_std                  (N floats)
//   ls_thresh  kth-value threshold of the noise-perturbed weights (env_L2A.py:94-96)
//   ls_search  ALL noisy multi-flip iterations (97-107) + the exhaustive single-flip pass
//              (110-115) + unpack, one CTA per tile of 32 envs, state in shared memory.  Only the
//              float32 noise (4 B per env-node-iteration) and the 1-byte cross counts stream in.
//
// Floating point: spin_rand = ws + noise * rd_std is evaluated exactly as the reference's torch
// kernels do -- one IEEE round-to-nearest multiply, one add, no FMA contraction; float(ws) is
// exact (small integer).
#include <math.h>

#include "tile_ops.cuh"

namespace rlsb {

int prepare_tiles(const GraphDev& g, const uint8_t* xs, const uint32_t* packed_in, int64_t num_envs,
                  uint32_t* packed_out, void* cross, bool cross_is_u8, int32_t* col_min, int32_t* col_max, int64_t* vs,
                  cudaStream_t st);

// float(k) for |k| < 2^22 without the conversion pipe: 0x4B400000 is 12582912.0f (1.5 * 2^23)
constexpr int kMagicI = 0x4B400000;
constexpr float kMagicF = 12582912.0f;

__device__ __forceinline__ float spin_rand(int degm, int mult, int cross, float noise, float rd_std) {
  const float wsf = __fadd_rn(__int_as_float(degm - mult * cross), -kMagicF);   // exact float(deg - mult*cross)
  return __fadd_rn(wsf, __fmul_rn(noise, rd_std));
}

// order-preserving float -> uint32 key (so REDUX max works on floats)
__device__ __forceinline__ uint32_t float_key(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_float(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// cross counts of VEC consecutive nodes, kept packed until they are used
template <typename CrossT, int VEC>
struct CrossVec {
  uint32_t lo, hi;
  __device__ __forceinline__ void load(const CrossT* p) {
    if constexpr (VEC == 4 && sizeof(CrossT) == 1) {
      lo = __ldg(reinterpret_cast<const uint32_t*>(p)), hi = 0;
    } else if constexpr (VEC == 4) {
      const uint2 c = __ldg(reinterpret_cast<const uint2*>(p));
      lo = c.x, hi = c.y;
    } else {
      lo = (uint32_t)__ldg(p), hi = 0;
    }
  }
  __device__ __forceinline__ int get(int b) const {
    if constexpr (sizeof(CrossT) == 1) return (int)__byte_perm(lo, 0, 0x4440 + b);   // byte b, zero-extended
    return (int)(((b < 2 ? lo : hi) >> (16 * (b & 1))) & 0xffffu);
  }
};

__global__ void ls_rdstd_kernel(GraphDev g, const int32_t* __restrict__ col_min, const int32_t* __restrict__ col_max,
                                int mult, float noise_std, float* __restrict__ rd_std, int32_t* __restrict__ degm) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= g.np) return;
  const bool real = i < g.n;
  rd_std[i] = real ? __fmul_rn((float)(mult * (col_max[i] - col_min[i])), noise_std) : 0.f;
  degm[i] = (real ? g.listed_deg[i] : 0) + kMagicI;
}

// ---------------------------------------------------------------- thresh (kthvalue)
// One warp per environment.  Each lane keeps the KMAX largest values of its share in a sorted
// register list; the lists are then merged by K rounds of warp-max.
template <int KMAX, typename CrossT, int VEC>
__global__ void __launch_bounds__(256) ls_thresh_kernel(GraphDev g, const CrossT* __restrict__ cross,
                                                        const float* __restrict__ rd_std,
                                                        const int32_t* __restrict__ degm, int mult,
                                                        const float* __restrict__ noise, int kth_big,
                                                        int64_t num_envs, float* __restrict__ thresh) {
  const int lane = threadIdx.x & 31;
  const int64_t env = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (env >= num_envs) return;
  float top[KMAX];
#pragma unroll
  for (int t = 0; t < KMAX; ++t) top[t] = -INFINITY;
  const CrossT* crow = cross + env * (int64_t)g.np;
  const float* nrow = noise + env * (int64_t)g.n;
  auto push = [&](float s) {
    if (s > top[KMAX - 1]) {
#pragma unroll
      for (int t = 0; t < KMAX; ++t) {
        const float hi = fmaxf(top[t], s);
        s = fminf(top[t], s);
        top[t] = hi;
      }
    }
  };
  if (VEC == 4) {
    for (int i = lane * 4; i < g.n; i += 128) {
      const float4 nz = ldg_stream4(nrow + i);
      const float4 rd = __ldg(reinterpret_cast<const float4*>(rd_std + i));
      const int4 dm = __ldg(reinterpret_cast<const int4*>(degm + i));
      CrossVec<CrossT, 4> cv;
      cv.load(crow + i);
      const int c0 = cv.get(0), c1 = cv.get(1), c2 = cv.get(2), c3 = cv.get(3);
      push(spin_rand(dm.x, mult, c0, nz.x, rd.x));
      push(spin_rand(dm.y, mult, c1, nz.y, rd.y));
      push(spin_rand(dm.z, mult, c2, nz.z, rd.z));
      push(spin_rand(dm.w, mult, c3, nz.w, rd.w));
    }
  } else {
    for (int i = lane; i < g.n; i += 32)
      push(spin_rand(__ldg(degm + i), mult, (int)__ldg(crow + i), ldg_stream(nrow + i), __ldg(rd_std + i)));
  }
  uint32_t best = 0;
  for (int r = 0; r < kth_big; ++r) {
    const uint32_t head = float_key(top[0]);
    best = __reduce_max_sync(kFull, head);
    const unsigned who = __ballot_sync(kFull, head == best);
    if (lane == __ffs(who) - 1) {
#pragma unroll
      for (int t = 0; t + 1 < KMAX; ++t) top[t] = top[t + 1];
      top[KMAX - 1] = -INFINITY;
    }
  }
  if (lane == 0) thresh[env] = key_float(best);
}

// ---------------------------------------------------------------- fused search kernel
constexpr int kLSThreads = 512;
constexpr int kLSMaxIters = 16;   // noise tensors per launch (pointers travel by value)
struct NoisePtrs {
  const float* p[kLSMaxIters];
};

struct LsArgs {
  uint32_t* packed;        // [W][Np] in/out (workspace)
  int64_t* vs;             // [E] in/out
  const void* cross;       // [E][Np] uint8 / uint16
  const float* rd_std;     // [Np]
  const int32_t* degm;     // [Np] listed degree + kMagicI
  const float* thresh;     // [E]
  NoisePtrs noise;
  int num_iters, mult;
  int64_t num_envs;
  int finish;              // 1: run the single-flip pass and write the bool rows
  uint8_t* xs_out;         // [E][N] bool rows (finish)
  int cut_warps;
  int stage_sweep;         // 1: the sweep structure fits in shared memory next to the two tile copies
  int sweep_warps;         // warps that take part in the single-flip pass
  int negmult;             // -mult
};

// Flip-mask of one noisy iteration for the whole tile, written as candidate = accepted ^ mask.
// Work item = (4 consecutive nodes) x (8 consecutive envs): 8 coalesced 16-byte noise loads in
// flight per thread, one result byte per node (byte g of a word = envs 8g..8g+7).
template <typename CrossT, int VEC, bool FULL>
__device__ __forceinline__ void noisy_candidate(const GraphDev& g, const LsArgs& a, const float* __restrict__ noise,
                                                int64_t env0, int valid, const float* sThresh, const uint32_t* sP,
                                                uint32_t* sX) {
  // tile bases once (64-bit); everything inside the tile is a 32-bit offset (32 * N < 2^31)
  const float* __restrict__ nbase = noise + env0 * (int64_t)g.n;
  const CrossT* __restrict__ cbase = static_cast<const CrossT*>(a.cross) + env0 * (int64_t)g.np;
  const uint32_t n = (uint32_t)g.n, np = (uint32_t)g.np;
  const uint32_t groups = (n + VEC - 1) / VEC;        // node groups
  const uint8_t* sPb = reinterpret_cast<const uint8_t*>(sP);
  uint8_t* sXb = reinterpret_cast<uint8_t*>(sX);
  const int negmult = a.negmult;
  for (uint32_t task = threadIdx.x; task < groups * 4; task += blockDim.x) {
    const uint32_t eg = task / groups, i0 = (task - eg * groups) * VEC;
    const uint32_t e0 = eg * 8;
    float rd[VEC];
    int dm[VEC];
    if constexpr (VEC == 4) {
      const float4 r = __ldg(reinterpret_cast<const float4*>(a.rd_std + i0));
      const int4 d = __ldg(reinterpret_cast<const int4*>(a.degm + i0));
      rd[0] = r.x, rd[1] = r.y, rd[2] = r.z, rd[3] = r.w;
      dm[0] = d.x, dm[1] = d.y, dm[2] = d.z, dm[3] = d.w;
    } else {
      rd[0] = __ldg(a.rd_std + i0), dm[0] = __ldg(a.degm + i0);
    }
    float nz[8][VEC];
    CrossVec<CrossT, VEC> cr[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (FULL || (int)(e0 + j) < valid) {
        const uint32_t no = (e0 + j) * n + i0, co = (e0 + j) * np + i0;
        if constexpr (VEC == 4) {
          const float4 v = ldg_stream4(nbase + no);
          nz[j][0] = v.x, nz[j][1] = v.y, nz[j][2] = v.z, nz[j][3] = v.w;
        } else {
          nz[j][0] = ldg_stream(nbase + no);
        }
        cr[j].load(cbase + co);
      } else {
#pragma unroll
        for (int b = 0; b < VEC; ++b) nz[j][b] = 0.f;
        cr[j].lo = cr[j].hi = 0;
      }
    }
    uint32_t bits[VEC];
#pragma unroll
    for (int b = 0; b < VEC; ++b) bits[b] = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float th = sThresh[e0 + j];      // +inf for envs past the batch: never flips
#pragma unroll
      for (int b = 0; b < VEC; ++b) {
        // exact float(deg - mult*cross) via the magic-number add, then one multiply and one add
        const float wsf = __fadd_rn(__int_as_float(cr[j].get(b) * negmult + dm[b]), -kMagicF);
        const float sr = __fadd_rn(wsf, __fmul_rn(nz[j][b], rd[b]));
        if (sr > th) bits[b] |= (1u << j);
      }
    }
#pragma unroll
    for (int b = 0; b < VEC; ++b) {
      const uint32_t at = (i0 + b) * 4 + eg;
      sXb[at] = sPb[at] ^ (uint8_t)bits[b];
    }
  }
}

// Exhaustive single-flip pass, Gauss-Seidel over nodes 0..N-1 with acceptance gain >= 0.
// Nodes of one dependency level (graph_store.cu) are pairwise non-adjacent: one lane per
// node decides them concurrently and a barrier separates levels, which reproduces the
// sequential order exactly.  gain = deg - 2*cross >= 0  <=>  cross <= floor(deg/2).
template <int P, bool SMEM>
__device__ __forceinline__ void sweep_tile(const GraphDev& g, const SweepView& sv, uint32_t* sP, int sweep_warps) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp >= sweep_warps) return;      // a level never has more slices than sweep_warps (or all warps take part)
  for (int l = 0; l < g.levels; ++l) {
    const int sb = SMEM ? sv.level_slice[l] : __ldg(sv.level_slice + l);
    const int se = SMEM ? sv.level_slice[l + 1] : __ldg(sv.level_slice + l + 1);
    for (int s = sb + warp; s < se; s += sweep_warps) {
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
  }
}

template <int P, typename CrossT, int VEC>
__global__ void __launch_bounds__(kLSThreads) ls_search_kernel(GraphDev g, LsArgs a) {
  extern __shared__ __align__(16) uint32_t smem[];
  uint32_t* sP = smem;           // accepted state of the tile
  uint32_t* sX = smem + g.np;    // candidate state
  char* sSweep = reinterpret_cast<char*>(smem + 2 * g.np);   // staged sweep structure (a.stage_sweep)
  __shared__ float sThresh[kTileEnvs];
  __shared__ int sCnt[kTileEnvs];
  __shared__ uint32_t sAccept;
  __shared__ __align__(8) uint64_t sBar;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t tiles = (a.num_envs + kTileEnvs - 1) / kTileEnvs;
  // The single-flip pass walks the graph level by level, one dependent load after another: have
  // the TMA engine copy the whole sweep structure into shared memory now; it lands while the
  // noisy iterations run.
  const bool staged = a.finish && a.stage_sweep;
  if (staged && threadIdx.x == 0) {
    mbar_init(&sBar, 1);
    mbar_expect_tx(&sBar, (uint32_t)g.sweep_blob_bytes);
    for (int off = 0; off < g.sweep_blob_bytes; off += 32768) {
      const int len = g.sweep_blob_bytes - off < 32768 ? g.sweep_blob_bytes - off : 32768;
      bulk_g2s(sSweep + off, g.sweep_blob + off, (uint32_t)len, &sBar);
    }
  }
  bool sweep_landed = false;
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t env0 = tile * kTileEnvs;
    const int valid = (int)min((int64_t)kTileEnvs, a.num_envs - env0);
    const uint32_t vmask = valid == 32 ? kFull : ((1u << valid) - 1u);
    for (int i = threadIdx.x; i < g.np; i += blockDim.x) {
      const uint32_t w = a.packed[tile * g.np + i];
      sP[i] = w, sX[i] = w;        // padding nodes stay equal in both copies
    }
    if (threadIdx.x < kTileEnvs)
      sThresh[threadIdx.x] = (a.num_iters > 0 && threadIdx.x < valid) ? a.thresh[env0 + threadIdx.x] : INFINITY;
    int64_t my_vs = 0;   // warp 0: lane e owns env e's value
    if (warp == 0 && lane < valid) my_vs = a.vs[env0 + lane];
    __syncthreads();
    for (int it = 0; it < a.num_iters; ++it) {
      if (threadIdx.x < kTileEnvs) sCnt[threadIdx.x] = 0;
      if (valid == kTileEnvs)
        noisy_candidate<CrossT, VEC, true>(g, a, a.noise.p[it], env0, valid, sThresh, sP, sX);
      else
        noisy_candidate<CrossT, VEC, false>(g, a, a.noise.p[it], env0, valid, sThresh, sP, sX);
      __syncthreads();
      const int cnt = tile_cut_partial(g, sX, a.cut_warps);
      if (cnt) atomicAdd(&sCnt[lane], cnt);
      __syncthreads();
      // keep rows that are not worse (vs1 >= vs0, util_read_data.py:199)
      if (warp == 0) {
        const int64_t cand = sCnt[lane];
        const bool keep = lane < valid && cand >= my_vs;
        if (keep) my_vs = cand;
        const unsigned acc = __ballot_sync(kFull, keep);
        if (lane == 0) sAccept = acc;
      }
      __syncthreads();
      const uint32_t acc = sAccept;
      for (int i = threadIdx.x; i < g.n; i += blockDim.x) sP[i] = (sX[i] & acc) | (sP[i] & ~acc);
      __syncthreads();
    }
    if (a.finish) {
      if (staged) {
        if (!sweep_landed) mbar_wait(&sBar, 0), sweep_landed = true;
        sweep_tile<P, true>(g, sweep_view(g, sSweep), sP, a.sweep_warps);
      } else {
        sweep_tile<P, false>(g, sweep_view(g, g.sweep_blob), sP, a.sweep_warps);
      }
      if (threadIdx.x < kTileEnvs) sCnt[threadIdx.x] = 0;
      __syncthreads();
      const int cnt = tile_cut_partial(g, sP, a.cut_warps);
      if (cnt) atomicAdd(&sCnt[lane], cnt);
      __syncthreads();
      if (threadIdx.x < valid) a.vs[env0 + threadIdx.x] = sCnt[threadIdx.x];
      unpack_tile_from_smem<VEC>(sP, a.xs_out, a.num_envs, g.n, g.np, tile);
    } else {
      if (warp == 0 && lane < valid) a.vs[env0 + lane] = my_vs;
    }
    for (int i = threadIdx.x; i < g.np; i += blockDim.x) a.packed[tile * g.np + i] = sP[i] & vmask;
    __syncthreads();
  }
}

// single-flip pass on packed tiles only (rlsb_flip_sweep)
template <int P>
__global__ void __launch_bounds__(kLSThreads) flip_sweep_kernel(GraphDev g, uint32_t* __restrict__ packed,
                                                                int64_t* __restrict__ vs, int64_t num_envs,
                                                                int cut_warps, int sweep_warps) {
  extern __shared__ uint32_t sP[];
  __shared__ int sCnt[kTileEnvs];
  const int lane = threadIdx.x & 31;
  const int64_t tiles = (num_envs + kTileEnvs - 1) / kTileEnvs;
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t env0 = tile * kTileEnvs;
    const int valid = (int)min((int64_t)kTileEnvs, num_envs - env0);
    const uint32_t vmask = valid == 32 ? kFull : ((1u << valid) - 1u);
    for (int i = threadIdx.x; i < g.np; i += blockDim.x) sP[i] = packed[tile * g.np + i];
    if (threadIdx.x < kTileEnvs) sCnt[threadIdx.x] = 0;
    __syncthreads();
    sweep_tile<P, false>(g, sweep_view(g, g.sweep_blob), sP, sweep_warps);
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

static int degree_class(const GraphDev& g) {
  const int d = g.max_listed_deg > g.max_full_deg ? g.max_listed_deg : g.max_full_deg;
  return d <= 63 ? 0 : d <= 255 ? 1 : 2;
}

// workspace carving (all sections 256-byte aligned)
struct LsWorkspace {
  uint32_t* packed;
  void* cross;
  int32_t *col_min, *col_max, *degm;
  float *rd_std, *thresh;
  size_t bytes;
};

static LsWorkspace carve(const GraphDev& g, int64_t num_envs, void* base) {
  const int64_t tiles = (num_envs + kTileEnvs - 1) / kTileEnvs;
  const size_t cross_elt = degree_class(g) == 2 ? 2 : 1;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* p = base ? static_cast<char*>(base) + off : nullptr;
    off += (bytes + 255) / 256 * 256;
    return p;
  };
  LsWorkspace w;
  w.packed = reinterpret_cast<uint32_t*>(take((size_t)tiles * g.np * 4));
  w.cross = take((size_t)num_envs * g.np * cross_elt);
  w.col_min = reinterpret_cast<int32_t*>(take((size_t)g.np * 4));
  w.col_max = reinterpret_cast<int32_t*>(take((size_t)g.np * 4));
  w.degm = reinterpret_cast<int32_t*>(take((size_t)g.np * 4));
  w.rd_std = reinterpret_cast<float*>(take((size_t)g.np * 4));
  w.thresh = reinterpret_cast<float*>(take((size_t)num_envs * 4));
  w.bytes = off + 256;
  return w;
}

template <int P, typename CrossT>
static int launch_search(const GraphDev& g, const LsArgs& a, bool vec4, cudaStream_t st) {
  const size_t smem = 2 * (size_t)g.np * sizeof(uint32_t) + (a.stage_sweep ? (size_t)g.sweep_blob_bytes : 0);
  const int64_t tiles = (a.num_envs + kTileEnvs - 1) / kTileEnvs;
  const unsigned grid = (unsigned)(tiles < 4 * kNumSMs ? tiles : 4 * kNumSMs);
  int rc;
  if (vec4) {
    if ((rc = allow_smem(ls_search_kernel<P, CrossT, 4>, smem))) return rc;
    ls_search_kernel<P, CrossT, 4><<<grid, kLSThreads, smem, st>>>(g, a);
  } else {
    if ((rc = allow_smem(ls_search_kernel<P, CrossT, 1>, smem))) return rc;
    ls_search_kernel<P, CrossT, 1><<<grid, kLSThreads, smem, st>>>(g, a);
  }
  RLSB_LAUNCH_OK();
  return RLSB_OK;
}

template <typename CrossT>
static int launch_thresh(const GraphDev& g, const LsWorkspace& w, int mult, const float* noise, int kth_big,
                         int64_t num_envs, bool vec4, cudaStream_t st) {
  const unsigned grid = (unsigned)((num_envs + 7) / 8);
  const CrossT* cross = static_cast<const CrossT*>(w.cross);
#define RLSB_THRESH(KMAX, V) \
  ls_thresh_kernel<KMAX, CrossT, V><<<grid, 256, 0, st>>>(g, cross, w.rd_std, w.degm, mult, noise, kth_big, num_envs, w.thresh)
  if (kth_big <= 10) {
    if (vec4) RLSB_THRESH(10, 4); else RLSB_THRESH(10, 1);
  } else {
    if (vec4) RLSB_THRESH(32, 4); else RLSB_THRESH(32, 1);
  }
#undef RLSB_THRESH
  RLSB_LAUNCH_OK();
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
  const void* at[7] = {w.packed, w.cross, w.col_min, w.col_max, w.degm, w.rd_std, w.thresh};
  if (section < 0 || section > 6) return -1;
  return static_cast<const char*>(at[section]) - base;
}

int rlsb_ls_begin(const rlsb_graph_t* gh, const uint8_t* xs, int64_t num_envs, int64_t* vs, int32_t compute_vs,
                  int32_t ws_mult, float noise_std, void* workspace, void* stream) {
  using namespace rlsb;
  const GraphDev* g;
  if (int rc = graph_check(gh, &g, "ls_begin")) return rc;
  RLSB_REQUIRE(num_envs >= 0, RLSB_ERR_INVALID, "ls_begin: negative num_envs");
  RLSB_REQUIRE(ws_mult == 1 || ws_mult == 2, RLSB_ERR_INVALID, "ls_begin: ws_mult must be 1 or 2");
  RLSB_REQUIRE(2 * (size_t)g->np * 4 <= 220 * 1024, RLSB_ERR_UNSUPPORTED,
               "ls_begin: %d nodes exceed the two-copy shared-memory tile of the search kernel", g->n);
  if (num_envs == 0 || g->n == 0) return RLSB_OK;
  RLSB_REQUIRE(xs && workspace && (vs || !compute_vs), RLSB_ERR_INVALID, "ls_begin: null pointer");
  RLSB_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255u) == 0, RLSB_ERR_INVALID,
               "ls_begin: workspace must be 256-byte aligned");
  auto st = static_cast<cudaStream_t>(stream);
  const LsWorkspace w = carve(*g, num_envs, workspace);
  if (int rc = prepare_tiles(*g, xs, nullptr, num_envs, w.packed, w.cross, degree_class(*g) != 2, w.col_min, w.col_max,
                             compute_vs ? vs : nullptr, st))
    return rc;
  ls_rdstd_kernel<<<(g->np + 255) / 256, 256, 0, st>>>(*g, w.col_min, w.col_max, ws_mult, noise_std, w.rd_std, w.degm);
  RLSB_LAUNCH_OK();
  return RLSB_OK;
}

int rlsb_ls_thresh(const rlsb_graph_t* gh, int64_t num_envs, int32_t ws_mult, const float* noise, int32_t num_spin,
                   void* workspace, void* stream) {
  using namespace rlsb;
  const GraphDev* g;
  if (int rc = graph_check(gh, &g, "ls_thresh")) return rc;
  RLSB_REQUIRE(num_envs >= 0, RLSB_ERR_INVALID, "ls_thresh: negative num_envs");
  // torch.kthvalue(k = N - num_spin) needs 1 <= k <= N
  RLSB_REQUIRE(num_spin >= 0 && num_spin < g->n, RLSB_ERR_INVALID,
               "ls_thresh: k = N - num_spin = %d out of range for N = %d (torch.kthvalue raises)", g->n - num_spin,
               g->n);
  RLSB_REQUIRE(ws_mult == 1 || ws_mult == 2, RLSB_ERR_INVALID, "ls_thresh: ws_mult must be 1 or 2");
  if (num_envs == 0) return RLSB_OK;
  RLSB_REQUIRE(noise && workspace, RLSB_ERR_INVALID, "ls_thresh: null pointer");
  const int kth_big = num_spin + 1;   // kth smallest with k = N - num_spin  ==  (num_spin+1)-th largest
  RLSB_REQUIRE(kth_big <= 32, RLSB_ERR_UNSUPPORTED, "ls_thresh: num_spin %d above the in-register limit 31", num_spin);
  auto st = static_cast<cudaStream_t>(stream);
  const LsWorkspace w = carve(*g, num_envs, workspace);
  const bool vec4 = rows_vec4_ok(noise, g->n) && (reinterpret_cast<uintptr_t>(noise) & 15u) == 0;
  return degree_class(*g) == 2 ? launch_thresh<uint16_t>(*g, w, ws_mult, noise, kth_big, num_envs, vec4, st)
                               : launch_thresh<uint8_t>(*g, w, ws_mult, noise, kth_big, num_envs, vec4, st);
}

int rlsb_ls_search(const rlsb_graph_t* gh, int64_t num_envs, int64_t* vs, int32_t ws_mult,
                   const float* const* h_noise_ptrs, int32_t num_iters, int32_t finish, uint8_t* xs_out,
                   void* workspace, void* stream) {
  using namespace rlsb;
  const GraphDev* g;
  if (int rc = graph_check(gh, &g, "ls_search")) return rc;
  RLSB_REQUIRE(num_envs >= 0 && num_iters >= 0, RLSB_ERR_INVALID, "ls_search: negative size");
  RLSB_REQUIRE(ws_mult == 1 || ws_mult == 2, RLSB_ERR_INVALID, "ls_search: ws_mult must be 1 or 2");
  if (num_envs == 0 || g->n == 0) return RLSB_OK;
  RLSB_REQUIRE(vs && workspace && (num_iters == 0 || h_noise_ptrs) && (!finish || xs_out), RLSB_ERR_INVALID,
               "ls_search: null pointer");
  RLSB_REQUIRE(2 * (size_t)g->np * 4 <= 220 * 1024, RLSB_ERR_UNSUPPORTED,
               "ls_search: %d nodes exceed the two-copy shared-memory tile", g->n);
  auto st = static_cast<cudaStream_t>(stream);
  const LsWorkspace w = carve(*g, num_envs, workspace);
  const int dc = degree_class(*g);
  int done = 0;
  do {
    const int now = num_iters - done < kLSMaxIters ? num_iters - done : kLSMaxIters;
    LsArgs a{};
    a.packed = w.packed, a.vs = vs, a.cross = w.cross, a.rd_std = w.rd_std, a.degm = w.degm, a.thresh = w.thresh;
    bool vec4 = g->n % 4 == 0 && (!xs_out || rows_vec4_ok(xs_out, g->n));
    for (int k = 0; k < now; ++k) {
      RLSB_REQUIRE(h_noise_ptrs[done + k] != nullptr, RLSB_ERR_INVALID, "ls_search: null noise tensor %d", done + k);
      a.noise.p[k] = h_noise_ptrs[done + k];
      vec4 = vec4 && (reinterpret_cast<uintptr_t>(a.noise.p[k]) & 15u) == 0;
    }
    a.num_iters = now, a.mult = ws_mult, a.num_envs = num_envs;
    a.finish = (finish && done + now == num_iters) ? 1 : 0;
    a.xs_out = xs_out, a.cut_warps = cut_warps_for(g->m, kLSThreads / 32);
    a.sweep_warps = sweep_warps_for(*g, kLSThreads / 32), a.negmult = -ws_mult;
    a.stage_sweep = (a.finish && 2 * (size_t)g->np * 4 + (size_t)g->sweep_blob_bytes <= 200 * 1024) ? 1 : 0;
    int rc = dc == 0   ? launch_search<6, uint8_t>(*g, a, vec4, st)
             : dc == 1 ? launch_search<8, uint8_t>(*g, a, vec4, st)
                       : launch_search<12, uint16_t>(*g, a, vec4, st);
    if (rc) return rc;
    done += now;
  } while (done < num_iters);
  return RLSB_OK;
}

int rlsb_flip_sweep(const rlsb_graph_t* gh, uint32_t* packed, int64_t* vs, int64_t num_envs, void* stream) {
  using namespace rlsb;
  const GraphDev* g;
  if (int rc = graph_check(gh, &g, "flip_sweep")) return rc;
  RLSB_REQUIRE(num_envs >= 0, RLSB_ERR_INVALID, "flip_sweep: negative num_envs");
  if (num_envs == 0 || g->n == 0) return RLSB_OK;
  RLSB_REQUIRE(packed && vs, RLSB_ERR_INVALID, "flip_sweep: null pointer");
  const size_t smem = (size_t)g->np * sizeof(uint32_t);
  const int64_t tiles = (num_envs + kTileEnvs - 1) / kTileEnvs;
  const unsigned grid = (unsigned)(tiles < 4 * kNumSMs ? tiles : 4 * kNumSMs);
  auto st = static_cast<cudaStream_t>(stream);
  const int cw = cut_warps_for(g->m, kLSThreads / 32);
  const int dc = g->max_full_deg <= 63 ? 0 : g->max_full_deg <= 255 ? 1 : 2;
  int rc;
#define RLSB_SWEEP(P)                                                \
  if ((rc = allow_smem(flip_sweep_kernel<P>, smem))) return rc;      \
  flip_sweep_kernel<P><<<grid, kLSThreads, smem, st>>>(*g, packed, vs, num_envs, cw, sweep_warps_for(*g, kLSThreads / 32))
  if (dc == 0) { RLSB_SWEEP(6); } else if (dc == 1) { RLSB_SWEEP(8); } else { RLSB_SWEEP(12); }
#undef RLSB_SWEEP
  RLSB_LAUNCH_OK();
  return RLSB_OK;
}

}  // extern "C"
