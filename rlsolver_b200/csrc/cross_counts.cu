// Tile "prepare" kernel: everything the reference computes from a batch of states before it
// starts flipping -- EnvMaxcut.calculate_obj_values_for_loop (rlsolver/envs/env_L2A.py:68-80,
// N Python iterations of three tiny kernels there), its cross-env max/min (`ws_std`,
// env_L2A.py:93 / LocalSearch.py:65) and the objective (env_L2A.py:54-66).
//
// One CTA per tile of 32 envs, spins bit-packed in shared memory.  One LANE per node: the lane
// walks its node's listed neighbours (SELL-32 slice, coalesced uint16 ids), XORs the
// neighbour word with its own and adds the result into bit-sliced counters, so a LOP3 serves
// 32 environments and there are no shuffles in the counting loop.  The planes are expanded to
// one byte (or halfword) per env only when the counts leave the chip.
#include <limits.h>

#include "tile_ops.cuh"

namespace rlsb {

constexpr int kPrepThreads = 512;

// Row-major layout [E][Np] (the public rlsb_node_cross_counts result).
template <int P, typename CrossT>
__device__ __forceinline__ void store_cross(const VCount<P>& vc, CrossT* __restrict__ cross, int64_t env0, int valid,
                                            int np, int i) {
  CrossT* base = cross + env0 * (int64_t)np + i;
  if (sizeof(CrossT) == 1) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const uint32_t b = vc.bytes4(q);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (4 * q + j < valid) base[(int64_t)(4 * q + j) * np] = (CrossT)((b >> (8 * j)) & 0xffu);
    }
  } else {
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const uint32_t h = vc.halves2(q);
#pragma unroll
      for (int j = 0; j < 2; ++j)
        if (2 * q + j < valid) base[(int64_t)(2 * q + j) * np] = (CrossT)((h >> (16 * j)) & 0xffffu);
    }
  }
}

// Tiled layout of the local-search workspace: per tile [node/4][env][node%4], so that the search
// kernel's lane (= env) finds the 4 counts of a node group in one word and a warp reads 128
// contiguous bytes.  Lane = node holds, per q, the counts of envs 4q..4q+3 as 4 bytes; a 4x4 byte
// transpose among the 4 lanes of a node group (2 shuffles) turns that into "4 nodes of one env",
// which is one aligned 4-byte store.  All 32 env slots are written (idle envs count 0).
// `rows` (uint8 only, optional): the same words also go to a row-major [E][Np] copy -- "4 nodes of one
// env" is an aligned 4-byte store there too -- which the mask generator (noise_masks.cu) reads with
// lane = node.
template <int P, typename CrossT>
__device__ __forceinline__ void store_cross_tiled(const VCount<P>& vc, CrossT* __restrict__ cross, int64_t tile,
                                                  int np, int i, int lane, uint8_t* __restrict__ rows) {
  CrossT* tbase = cross + tile * (int64_t)np * kTileEnvs;
  if (sizeof(CrossT) == 1) {
    const int j = lane & 3;
    uint32_t* gbase = reinterpret_cast<uint32_t*>(tbase) + (i >> 2) * kTileEnvs;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      uint32_t x = vc.bytes4(q);
      uint32_t y = __shfl_xor_sync(kFull, x, 2);
      x = (j & 2) ? __byte_perm(x, y, 0x3276) : __byte_perm(x, y, 0x5410);
      y = __shfl_xor_sync(kFull, x, 1);
      x = (j & 1) ? __byte_perm(x, y, 0x3715) : __byte_perm(x, y, 0x6240);
      gbase[4 * q + j] = x;      // env 4q+j, nodes 4*(i/4) .. +3
      if (rows)
        *reinterpret_cast<uint32_t*>(rows + (tile * kTileEnvs + 4 * q + j) * (int64_t)np + (i & ~3)) = x;
    }
  } else {
    CrossT* gbase = tbase + ((int64_t)(i >> 2) * kTileEnvs << 2) + (i & 3);
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const uint32_t h = vc.halves2(q);
      gbase[(2 * q) << 2] = (CrossT)(h & 0xffffu);
      gbase[(2 * q + 1) << 2] = (CrossT)(h >> 16);
    }
  }
}

// VEC: 0 = packed input, 1 / 4 = bool rows read bytewise / as 4-byte words
template <int P, typename CrossT, int VEC>
__global__ void __launch_bounds__(kPrepThreads) prepare_kernel(GraphDev g, const uint8_t* __restrict__ xs,
                                                               const uint32_t* __restrict__ packed_in,
                                                               int64_t num_envs, uint32_t* __restrict__ packed_out,
                                                               CrossT* __restrict__ cross, int tiled,
                                                               uint8_t* __restrict__ cross_rows,
                                                               int32_t* col_min, int32_t* col_max,
                                                               int64_t* __restrict__ vs, int cut_warps) {
  extern __shared__ uint32_t sP[];
  __shared__ int sCnt[kTileEnvs];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int64_t tiles = (num_envs + kTileEnvs - 1) / kTileEnvs;
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t env0 = tile * kTileEnvs;
    const int valid = (int)min((int64_t)kTileEnvs, num_envs - env0);
    const uint32_t vmask = valid == 32 ? kFull : ((1u << valid) - 1u);
    if (threadIdx.x < kTileEnvs) sCnt[threadIdx.x] = 0;
    if (VEC == 0) {
      for (int i = threadIdx.x; i < g.np; i += blockDim.x) sP[i] = __ldg(packed_in + tile * g.np + i);
    } else {
      pack_tile_to_smem<(VEC == 0 ? 1 : VEC)>(xs, num_envs, g.n, g.np, tile, sP);
    }
    __syncthreads();
    if (VEC != 0 && packed_out)
      for (int i = threadIdx.x; i < g.np; i += blockDim.x) packed_out[tile * g.np + i] = sP[i];
    if (vs) {
      const int cnt = tile_cut_partial(g, sP, cut_warps);
      if (cnt) atomicAdd(&sCnt[lane], cnt);
    }
    if (cross || col_min) {
      for (int slice = warp; slice * 32 < g.np; slice += nwarps) {
        const int i = slice * 32 + lane;
        VCount<P> vc;
        sell_cross<P, false>(g.listed, slice, lane, sP, sP[i], vc);
        if (cross) {
          if (tiled) store_cross_tiled<P, CrossT>(vc, cross, tile, g.np, i, lane, cross_rows);
          else store_cross<P, CrossT>(vc, cross, env0, valid, g.np, i);
        }
        if (col_min && i < g.n) {
          atomicMin(col_min + i, (int)vc.min_over(vmask));
          atomicMax(col_max + i, (int)vc.max_over(vmask));
        }
      }
    }
    __syncthreads();
    if (vs && threadIdx.x < valid) vs[env0 + threadIdx.x] = sCnt[threadIdx.x];
    __syncthreads();
  }
}

__global__ void fill_minmax_kernel(int32_t* mn, int32_t* mx, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) mn[i] = INT_MAX, mx[i] = 0;
}

template <int P, typename CrossT, int VEC>
static int launch_prepare(const GraphDev& g, const uint8_t* xs, const uint32_t* packed_in, int64_t num_envs,
                          uint32_t* packed_out, CrossT* cross, int tiled, uint8_t* cross_rows, int32_t* col_min,
                          int32_t* col_max, int64_t* vs, cudaStream_t st) {
  const size_t smem = (size_t)g.np * sizeof(uint32_t);
  auto kernel = prepare_kernel<P, CrossT, VEC>;
  RLSB_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  if (smem > 48 * 1024)
    RLSB_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t tiles = (num_envs + kTileEnvs - 1) / kTileEnvs;
  const unsigned grid = (unsigned)(tiles < 8 * kNumSMs ? tiles : 8 * kNumSMs);
  kernel<<<grid, kPrepThreads, smem, st>>>(g, xs, packed_in, num_envs, packed_out, cross, tiled, cross_rows, col_min, col_max, vs,
                                           cut_warps_for(g.m, kPrepThreads / 32));
  RLSB_LAUNCH_OK();
  return RLSB_OK;
}

template <typename CrossT, int VEC>
static int dispatch_planes(const GraphDev& g, const uint8_t* xs, const uint32_t* packed_in, int64_t num_envs,
                           uint32_t* packed_out, CrossT* cross, int tiled, uint8_t* cross_rows, int32_t* col_min,
                           int32_t* col_max, int64_t* vs, cudaStream_t st) {
  if (g.max_listed_deg <= 63)
    return launch_prepare<6, CrossT, VEC>(g, xs, packed_in, num_envs, packed_out, cross, tiled, cross_rows, col_min, col_max, vs, st);
  if (g.max_listed_deg <= 255)
    return launch_prepare<8, CrossT, VEC>(g, xs, packed_in, num_envs, packed_out, cross, tiled, cross_rows, col_min, col_max, vs, st);
  if (sizeof(CrossT) == 1) {
    set_error("prepare: listed degree %d needs uint16 counts", g.max_listed_deg);
    return RLSB_ERR_INVALID;
  }
  return launch_prepare<12, CrossT, VEC>(g, xs, packed_in, num_envs, packed_out, cross, tiled, cross_rows, col_min, col_max, vs, st);
}

// Used by rlsb_ls_begin (local_search.cu).  cross_layout: 0 = uint16 [E][Np] row-major,
// 1 = uint8 tiled, 2 = uint16 tiled (store_cross_tiled).  cross_rows: optional row-major uint8 copy (layout 1).
int prepare_tiles(const GraphDev& g, const uint8_t* xs, const uint32_t* packed_in, int64_t num_envs,
                  uint32_t* packed_out, void* cross, int cross_layout, uint8_t* cross_rows, int32_t* col_min,
                  int32_t* col_max, int64_t* vs, cudaStream_t st) {
  if (cross_layout != 1) cross_rows = nullptr;
  const bool cross_is_u8 = cross_layout == 1;
  const int tiled = cross_layout != 0;
  if (col_min && g.n > 0) {
    fill_minmax_kernel<<<(g.n + 255) / 256, 256, 0, st>>>(col_min, col_max, g.n);
    RLSB_LAUNCH_OK();
  }
  if (num_envs == 0 || g.n == 0) return RLSB_OK;
#define RLSB_PREP(T, V) dispatch_planes<T, V>(g, xs, packed_in, num_envs, packed_out, (T*)cross, tiled, cross_rows, col_min, col_max, vs, st)
  if (packed_in) return cross_is_u8 ? RLSB_PREP(uint8_t, 0) : RLSB_PREP(uint16_t, 0);
  if (rows_vec4_ok(xs, g.n)) return cross_is_u8 ? RLSB_PREP(uint8_t, 4) : RLSB_PREP(uint16_t, 4);
  return cross_is_u8 ? RLSB_PREP(uint8_t, 1) : RLSB_PREP(uint16_t, 1);
#undef RLSB_PREP
}

}  // namespace rlsb

extern "C" int rlsb_node_cross_counts(const rlsb_graph_t* gh, const uint32_t* packed, int64_t num_envs,
                                      uint16_t* cross, int32_t* col_min, int32_t* col_max, void* stream) {
  using namespace rlsb;
  const GraphDev* g;
  if (int rc = graph_check(gh, &g, "node_cross_counts")) return rc;
  RLSB_REQUIRE(num_envs >= 0, RLSB_ERR_INVALID, "node_cross_counts: negative num_envs");
  RLSB_REQUIRE((col_min == nullptr) == (col_max == nullptr), RLSB_ERR_INVALID,
               "node_cross_counts: col_min and col_max must both be given or both be null");
  RLSB_REQUIRE(num_envs == 0 || g->n == 0 || (packed && cross), RLSB_ERR_INVALID, "node_cross_counts: null pointer");
  return prepare_tiles(*g, nullptr, packed, num_envs, nullptr, cross, 0, nullptr, col_min, col_max, nullptr,
                       static_cast<cudaStream_t>(stream));
}
