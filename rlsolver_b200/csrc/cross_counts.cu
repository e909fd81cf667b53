// Per-node cross counts: integer core of EnvMaxcut.calculate_obj_values_for_loop
// (rlsolver/envs/env_L2A.py:68-80), which the reference runs as N Python iterations of
// three tiny kernels.  One CTA per tile of 32 envs; one warp per node: lanes take the
// node's listed neighbours, XOR their packed words with the node's word, and a 32x32 bit
// transpose turns "bit = env" into "lane = env" so a POPC yields the count for env == lane.
// Rows are staged through shared memory so the uint16 [E][Np] output is written coalesced.
// Also produces the cross-env min / max per node that `ws_std` needs (env_L2A.py:93).
#include <limits.h>

#include "tile_ops.cuh"

namespace rlsb {

constexpr int kCCThreads = 512;
constexpr int kCCChunk = 256;            // nodes per staging pass
constexpr int kCCRow = kCCChunk + 2;     // halfwords per staged row: odd word stride -> conflict-free

__global__ void __launch_bounds__(kCCThreads) cross_counts_kernel(GraphDev g, const uint32_t* __restrict__ packed,
                                                                  int64_t num_envs, uint16_t* __restrict__ cross,
                                                                  int32_t* col_min, int32_t* col_max) {
  extern __shared__ uint32_t smem[];
  uint32_t* sP = smem;
  uint16_t* sOut = reinterpret_cast<uint16_t*>(smem + g.np);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int64_t tiles = (num_envs + kTileEnvs - 1) / kTileEnvs;
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int valid = (int)min((int64_t)kTileEnvs, num_envs - tile * kTileEnvs);
    for (int i = threadIdx.x; i < g.np; i += blockDim.x) sP[i] = __ldg(packed + tile * g.np + i);
    __syncthreads();
    for (int base = 0; base < g.n; base += kCCChunk) {
      const int stop = min(base + kCCChunk, g.n);
      for (int i = base + warp; i < stop; i += nwarps) {
        const int rb = __ldg(g.listed_ptr + i), re = __ldg(g.listed_ptr + i + 1);
        const uint32_t pi = sP[i];
        int cnt = 0;
        for (int c = rb; c < re; c += 32) {
          const int k = c + lane;
          uint32_t x = 0;
          if (k < re) x = sP[__ldg(g.listed_col + k)] ^ pi;
          cnt += __popc(transpose32(x, lane));
        }
        sOut[lane * kCCRow + (i - base)] = (uint16_t)cnt;
        if (col_min) {
          const unsigned mn = __reduce_min_sync(kFull, lane < valid ? (unsigned)cnt : 0xffffffffu);
          const unsigned mx = __reduce_max_sync(kFull, lane < valid ? (unsigned)cnt : 0u);
          if (lane == 0) {
            atomicMin(col_min + i, (int)mn);
            atomicMax(col_max + i, (int)mx);
          }
        }
      }
      __syncthreads();
      const int width = stop - base;              // base is a multiple of 256 -> even
      for (int e = warp; e < valid; e += nwarps) {
        uint16_t* dst = cross + (tile * kTileEnvs + e) * (int64_t)g.np + base;
        const uint16_t* src = sOut + e * kCCRow;
        for (int j = 2 * lane; j < width; j += 64) {
          if (j + 1 < width) *reinterpret_cast<uint32_t*>(dst + j) = *reinterpret_cast<const uint32_t*>(src + j);
          else dst[j] = src[j];
        }
      }
      __syncthreads();
    }
  }
}

__global__ void fill_i32_kernel(int32_t* p, int32_t v, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

}  // namespace rlsb

extern "C" int rlsb_node_cross_counts(const rlsb_graph_t* gh, const uint32_t* packed, int64_t num_envs,
                                      uint16_t* cross, int32_t* col_min, int32_t* col_max, void* stream) {
  using namespace rlsb;
  const GraphDev* g = graph_dev(gh);
  RLSB_REQUIRE(gh != nullptr, RLSB_ERR_INVALID, "node_cross_counts: null graph");
  RLSB_REQUIRE(g != nullptr, RLSB_ERR_NODEVICE, "node_cross_counts: graph has no device image");
  RLSB_REQUIRE(num_envs >= 0, RLSB_ERR_INVALID, "node_cross_counts: negative num_envs");
  RLSB_REQUIRE((col_min == nullptr) == (col_max == nullptr), RLSB_ERR_INVALID,
               "node_cross_counts: col_min and col_max must both be given or both be null");
  RLSB_REQUIRE(rlsb_graph_max_listed_degree(gh) <= 65535, RLSB_ERR_UNSUPPORTED,
               "node_cross_counts: listed degree above 65535 does not fit the uint16 counts");
  auto st = static_cast<cudaStream_t>(stream);
  if (col_min && g->n > 0) {
    fill_i32_kernel<<<(g->n + 255) / 256, 256, 0, st>>>(col_min, INT_MAX, g->n);
    fill_i32_kernel<<<(g->n + 255) / 256, 256, 0, st>>>(col_max, 0, g->n);
  }
  if (num_envs == 0 || g->n == 0) return RLSB_OK;
  RLSB_REQUIRE(packed && cross, RLSB_ERR_INVALID, "node_cross_counts: null pointer");
  const size_t smem = (size_t)g->np * sizeof(uint32_t) + (size_t)kTileEnvs * kCCRow * sizeof(uint16_t);
  RLSB_REQUIRE(smem <= 220 * 1024, RLSB_ERR_UNSUPPORTED, "node_cross_counts: %d nodes exceed the shared-memory tile",
               g->n);
  if (smem > 48 * 1024)
    RLSB_CUDA_OK(cudaFuncSetAttribute(cross_counts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t tiles = (num_envs + kTileEnvs - 1) / kTileEnvs;
  const unsigned grid = (unsigned)(tiles < 8 * kNumSMs ? tiles : 8 * kNumSMs);
  cross_counts_kernel<<<grid, kCCThreads, smem, st>>>(*g, packed, num_envs, cross, col_min, col_max);
  RLSB_LAUNCH_OK();
  return RLSB_OK;
}
