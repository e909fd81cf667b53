// Objective evaluation: replaces EnvMaxcut.calculate_obj_values (rlsolver/envs/env_L2A.py:54-66).
// The reference gathers xs through three int64 [E][Md] index tensors (~27 B per env-edge);
// here one CTA owns a tile of 32 envs: it bit-packs the 32 bool rows into shared memory
// (coalesced row reads, the only HBM traffic: N bytes per env) and streams the edge list
// once for all 32 envs, each lane adding its edges' XOR words into bit-sliced counters
// (vcount.cuh).  Algorithmic bytes per env-eval: N + 8 + 4M/E (bool API) or N/8 + 8 + 4M/E (packed).
#include "tile_ops.cuh"

namespace rlsb {

constexpr int kCutThreads = 256;

template <int VEC, bool PACKED_IN>
__global__ void __launch_bounds__(kCutThreads) cut_eval_kernel(GraphDev g, const uint8_t* __restrict__ xs,
                                                               const uint32_t* __restrict__ packed,
                                                               int64_t num_envs, int64_t* __restrict__ vs,
                                                               int cut_warps) {
  extern __shared__ uint32_t sP[];
  __shared__ int sCnt[kTileEnvs];
  const int64_t tiles = (num_envs + kTileEnvs - 1) / kTileEnvs;
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    if (threadIdx.x < kTileEnvs) sCnt[threadIdx.x] = 0;
    if (PACKED_IN) {
      for (int i = threadIdx.x; i < g.np; i += blockDim.x) sP[i] = __ldg(packed + tile * g.np + i);
    } else {
      pack_tile_to_smem<VEC>(xs, num_envs, g.n, g.np, tile, sP);
    }
    __syncthreads();
    const int cnt = tile_cut_partial(g, sP, cut_warps);
    if (cnt) atomicAdd(&sCnt[threadIdx.x & 31], cnt);
    __syncthreads();
    const int64_t env = tile * kTileEnvs + threadIdx.x;
    if (threadIdx.x < kTileEnvs && env < num_envs) vs[env] = sCnt[threadIdx.x];
    __syncthreads();
  }
}

// Weighted objective: sum_e w_e [x_u != x_v] (PISCO's -1/4 s^T A s with a weighted adjacency,
// rlsolver/envs/env_ISCO.py:436-444; the expected cut of MCPG's weighted sampler, MCPG/sampling.py:121).
template <int VEC, bool PACKED_IN>
__global__ void __launch_bounds__(kCutThreads) cut_eval_weighted_kernel(GraphDev g, const uint8_t* __restrict__ xs,
                                                                        const uint32_t* __restrict__ packed,
                                                                        int64_t num_envs, int64_t* __restrict__ vs,
                                                                        int cut_warps) {
  extern __shared__ uint32_t sP[];
  __shared__ unsigned long long sCnt[kTileEnvs];
  const int64_t tiles = (num_envs + kTileEnvs - 1) / kTileEnvs;
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    if (threadIdx.x < kTileEnvs) sCnt[threadIdx.x] = 0;
    if (PACKED_IN) {
      for (int i = threadIdx.x; i < g.np; i += blockDim.x) sP[i] = __ldg(packed + tile * g.np + i);
    } else {
      pack_tile_to_smem<VEC>(xs, num_envs, g.n, g.np, tile, sP);
    }
    __syncthreads();
    const long long cnt = tile_cut_weighted_partial(g, sP, cut_warps);
    if (cnt) atomicAdd(&sCnt[threadIdx.x & 31], (unsigned long long)cnt);       // two's complement: signed sums work
    __syncthreads();
    const int64_t env = tile * kTileEnvs + threadIdx.x;
    if (threadIdx.x < kTileEnvs && env < num_envs) vs[env] = (int64_t)sCnt[threadIdx.x];
    __syncthreads();
  }
}

// Weighted local fields: out[e][i] = sum_j w_ij [x_i != x_j] over the full neighbourhood of node i (every
// undirected edge seen from both ends, self loops dropped).  The flip gain of node i is wdeg_i - 2 out[e][i].
// Lane = env, a warp walks nodes: the neighbour word is one broadcast shared-memory load, the weight one
// broadcast load from L2 -- 4 instructions per (env, neighbour), exact for any int32 weights.  (The unit-weight
// path counts 32 envs per LOP3 with bit-sliced counters; chains of weighted samplers come in hundreds, not
// hundreds of thousands.)
__global__ void __launch_bounds__(256) node_fields_weighted_kernel(GraphDev g, const uint32_t* __restrict__ packed,
                                                                   int64_t num_envs, int32_t* __restrict__ out) {
  extern __shared__ uint32_t sP[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int64_t tiles = (num_envs + kTileEnvs - 1) / kTileEnvs;
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    for (int i = threadIdx.x; i < g.np; i += blockDim.x) sP[i] = __ldg(packed + tile * g.np + i);
    __syncthreads();
    const int64_t env = tile * kTileEnvs + lane;
    for (int i = warp; i < g.n; i += nwarps) {
      const uint32_t self = sP[i];
      const int kb = __ldg(g.full_ptr + i), ke = __ldg(g.full_ptr + i + 1);
      int acc = 0;
      for (int k = kb; k < ke; ++k) {
        const uint32_t diff = self ^ sP[__ldg(g.full_col + k)];
        acc += ((diff >> lane) & 1u) ? __ldg(g.full_w + k) : 0;
      }
      if (env < num_envs) out[env * (int64_t)g.np + i] = acc;
    }
    __syncthreads();
  }
}

// if_sum=False: one indicator byte per (env, listed edge), in the reference's n0/n1 order.
__global__ void __launch_bounds__(256) cut_edges_kernel(GraphDev g, const uint8_t* __restrict__ xs, int64_t num_envs,
                                                        uint8_t* __restrict__ out) {
  const int64_t env = blockIdx.y;
  const uint8_t* row = xs + env * (int64_t)g.n;
  uint8_t* orow = out + env * (int64_t)g.md;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < g.md; k += gridDim.x * blockDim.x)
    orow[k] = (uint8_t)((row[__ldg(g.listed_row + k)] != 0) ^ (row[__ldg(g.listed_col + k)] != 0));
}

template <typename K>
static int set_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) RLSB_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return RLSB_OK;
}

}  // namespace rlsb

extern "C" {

static int cut_eval_common(const rlsb_graph_t* gh, const uint8_t* xs, const uint32_t* packed, int64_t num_envs,
                           int64_t* vs, void* stream) {
  using namespace rlsb;
  const GraphDev* g;
  if (int rc0 = graph_check(gh, &g, "cut_eval")) return rc0;
  RLSB_REQUIRE(num_envs >= 0, RLSB_ERR_INVALID, "cut_eval: negative num_envs");
  if (num_envs == 0) return RLSB_OK;
  RLSB_REQUIRE((xs || packed) && vs, RLSB_ERR_INVALID, "cut_eval: null pointer");
  const size_t smem = (size_t)g->np * sizeof(uint32_t);
    const int64_t tiles = (num_envs + kTileEnvs - 1) / kTileEnvs;
  const unsigned grid = (unsigned)(tiles < 8 * kNumSMs ? tiles : 8 * kNumSMs);
  auto st = static_cast<cudaStream_t>(stream);
  const int cw = cut_warps_for(g->m, kCutThreads / 32);
  int rc;
  if (packed) {
    if ((rc = set_smem(cut_eval_kernel<1, true>, smem))) return rc;
    cut_eval_kernel<1, true><<<grid, kCutThreads, smem, st>>>(*g, nullptr, packed, num_envs, vs, cw);
  } else if (rows_vec4_ok(xs, g->n)) {
    if ((rc = set_smem(cut_eval_kernel<4, false>, smem))) return rc;
    cut_eval_kernel<4, false><<<grid, kCutThreads, smem, st>>>(*g, xs, nullptr, num_envs, vs, cw);
  } else {
    if ((rc = set_smem(cut_eval_kernel<1, false>, smem))) return rc;
    cut_eval_kernel<1, false><<<grid, kCutThreads, smem, st>>>(*g, xs, nullptr, num_envs, vs, cw);
  }
  RLSB_LAUNCH_OK();
  return RLSB_OK;
}

int rlsb_cut_eval(const rlsb_graph_t* g, const uint8_t* xs, int64_t num_envs, int64_t* vs, void* stream) {
  return cut_eval_common(g, xs, nullptr, num_envs, vs, stream);
}

int rlsb_cut_eval_packed(const rlsb_graph_t* g, const uint32_t* packed, int64_t num_envs, int64_t* vs, void* stream) {
  return cut_eval_common(g, nullptr, packed, num_envs, vs, stream);
}

static int weighted_check(const rlsb_graph_t* gh, const rlsb::GraphDev** g, const char* what) {
  using namespace rlsb;
  if (int rc0 = graph_check(gh, g, what)) return rc0;
  RLSB_REQUIRE((*g)->wbuckets > 0 || (*g)->m == 0, RLSB_ERR_INVALID,
               "%s: the graph carries unit weights only (use the unweighted entry point)", what);
  return RLSB_OK;
}

int rlsb_graph_is_weighted(const rlsb_graph_t* gh) {
  const rlsb::GraphDev* g = rlsb::graph_dev(gh);
  return g && g->wbuckets > 0;
}

int rlsb_cut_eval_weighted(const rlsb_graph_t* gh, const uint8_t* xs, const uint32_t* packed, int64_t num_envs,
                           int64_t* vs, void* stream) {
  using namespace rlsb;
  const GraphDev* g;
  if (int rc0 = weighted_check(gh, &g, "cut_eval_weighted")) return rc0;
  RLSB_REQUIRE(num_envs >= 0, RLSB_ERR_INVALID, "cut_eval_weighted: negative num_envs");
  if (num_envs == 0) return RLSB_OK;
  RLSB_REQUIRE((xs || packed) && vs, RLSB_ERR_INVALID, "cut_eval_weighted: null pointer");
  const size_t smem = (size_t)g->np * sizeof(uint32_t);
  const int64_t tiles = (num_envs + kTileEnvs - 1) / kTileEnvs;
  const unsigned grid = (unsigned)(tiles < 8 * kNumSMs ? tiles : 8 * kNumSMs);
  auto st = static_cast<cudaStream_t>(stream);
  const int cw = cut_warps_for(g->m, kCutThreads / 32);
  int rc;
  if (packed) {
    if ((rc = set_smem(cut_eval_weighted_kernel<1, true>, smem))) return rc;
    cut_eval_weighted_kernel<1, true><<<grid, kCutThreads, smem, st>>>(*g, nullptr, packed, num_envs, vs, cw);
  } else if (rows_vec4_ok(xs, g->n)) {
    if ((rc = set_smem(cut_eval_weighted_kernel<4, false>, smem))) return rc;
    cut_eval_weighted_kernel<4, false><<<grid, kCutThreads, smem, st>>>(*g, xs, nullptr, num_envs, vs, cw);
  } else {
    if ((rc = set_smem(cut_eval_weighted_kernel<1, false>, smem))) return rc;
    cut_eval_weighted_kernel<1, false><<<grid, kCutThreads, smem, st>>>(*g, xs, nullptr, num_envs, vs, cw);
  }
  RLSB_LAUNCH_OK();
  return RLSB_OK;
}

int rlsb_node_fields_weighted(const rlsb_graph_t* gh, const uint32_t* packed, int64_t num_envs, int32_t* out,
                              void* stream) {
  using namespace rlsb;
  const GraphDev* g;
  if (int rc0 = weighted_check(gh, &g, "node_fields_weighted")) return rc0;
  RLSB_REQUIRE(num_envs >= 0, RLSB_ERR_INVALID, "node_fields_weighted: negative num_envs");
  if (num_envs == 0 || g->n == 0) return RLSB_OK;
  RLSB_REQUIRE(packed && out, RLSB_ERR_INVALID, "node_fields_weighted: null pointer");
  const size_t smem = (size_t)g->np * sizeof(uint32_t);
  const int64_t tiles = (num_envs + kTileEnvs - 1) / kTileEnvs;
  const unsigned grid = (unsigned)(tiles < 8 * kNumSMs ? tiles : 8 * kNumSMs);
  if (int rc = set_smem(node_fields_weighted_kernel, smem)) return rc;
  node_fields_weighted_kernel<<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(*g, packed, num_envs, out);
  RLSB_LAUNCH_OK();
  return RLSB_OK;
}

int rlsb_cut_edges(const rlsb_graph_t* gh, const uint8_t* xs, int64_t num_envs, uint8_t* out, void* stream) {
  using namespace rlsb;
  const GraphDev* g;
  if (int rc0 = graph_check(gh, &g, "cut_edges")) return rc0;
  if (num_envs <= 0 || g->md == 0) return RLSB_OK;
  RLSB_REQUIRE(xs && out, RLSB_ERR_INVALID, "cut_edges: null pointer");
  RLSB_REQUIRE(num_envs <= 65535, RLSB_ERR_UNSUPPORTED, "cut_edges: more than 65535 envs per call");
  dim3 grid((unsigned)((g->md + 1023) / 1024), (unsigned)num_envs);
  cut_edges_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(*g, xs, num_envs, out);
  RLSB_LAUNCH_OK();
  return RLSB_OK;
}

}  // extern "C"
