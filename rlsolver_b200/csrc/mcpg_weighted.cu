// Weighted max-cut sampler of MCPG: the local-search sweeps and the expected cut of mcpg_sampling_maxcut
// (rlsolver/methods/MCPG/sampling.py:101-121) for float `edge_attr` (rlsolver/methods/MCPG/dataloader.py:53-103).
//
// Reference, per sweep and node (in descending |weighted degree| order): a [1, deg] x [deg, C] torch.mm, a
// torch.rand(C), a compare and a row assignment -- num_ls * N rounds of four tiny launches.  Here a LANE is a
// chain and a warp walks the nodes of its 32 chains in order (Gauss-Seidel is sequential per chain, the chains
// are independent): the state of a tile is two words per node in shared memory (value bit, "not visited yet"
// bit: the reference starts from {-0.5, 1.5} and writes {0, 1}), the neighbour list and weights are broadcast
// loads, the sum is float32.  A decision needs its random number only when the sum is within 1/4 of the
// threshold (u / 4 < 1/4); then the lane evaluates torch's Philox stream for that (visit, chain) in place and
// applies the reference's expression -- one add, one compare -- so the generator ends where the reference leaves it.
//
// Sum order: four interleaved float32 accumulators (torch.mm's own order is unspecified).  For weights whose
// partial sums are exact in float32 (integers, dyadic fractions) every decision equals the reference's bit for bit;
// for arbitrary floats a decision within rounding of its threshold may differ -- the reference's CPU and CUDA
// builds disagree with each other in the same way (tests state the tolerance).
#include "philox.cuh"
#include "common.cuh"

namespace rlsb {

constexpr int kWmWarps = 4;

struct WmArgs {
  int n, np;
  int64_t num_chains;
  const int32_t* order;     // [n] visiting order
  const int32_t* nbr_ptr;   // [n + 1]
  const int32_t* nbr_col;   // neighbours in edge order, both directions
  const float* nbr_w;
  const float* thr;         // [n] float32(weighted_degree / 2 + 0.125)
  int num_edges;
  const int32_t *edge_u, *edge_v;
  const float* edge_w;
  float* xs;                // [n][C] in/out
  int sweeps;
  const float* explicit_u;  // [sweeps * n][C] or null
  TorchRng rng;
  float* expected;          // [C]
};

__device__ __forceinline__ float wm_value(uint32_t b, uint32_t u, int lane) {
  const float bit = (float)((b >> lane) & 1u), unv = (float)((u >> lane) & 1u);
  return bit * (1.f + unv) - 0.5f * unv;             // {0, 1} once visited, {-0.5, 1.5} before
}

__global__ void __launch_bounds__(kWmWarps * 32) mcpg_weighted_kernel(WmArgs a) {
  extern __shared__ uint32_t smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t* sB = smem + (size_t)warp * 2 * a.np;
  uint32_t* sU = sB + a.np;
  const int64_t tile = (int64_t)blockIdx.x * kWmWarps + warp;
  const int64_t chain = tile * kTileEnvs + lane;
  if (tile * kTileEnvs >= a.num_chains) return;       // whole warp idle (no block-wide barrier below)
  const bool live = chain < a.num_chains;
  const bool direct = a.num_chains <= (int64_t)a.rng.threads;
  const PhiloxKey key = philox_key(a.rng);
  const uint2 pkey = make_uint2((uint32_t)key.seed, (uint32_t)(key.seed >> 32));
  for (int i = 0; i < a.n; ++i) {
    const float v = live ? __ldg(a.xs + (int64_t)i * a.num_chains + chain) : 0.f;
    const uint32_t w = __ballot_sync(kFull, v != 0.f);
    if (lane == 0) sB[i] = w, sU[i] = kFull;
  }
  __syncwarp();
  {   // graph_probs += graph_probs[top]; % 2  (sampling.py:102-104)
    const uint32_t top = sB[__ldg(a.order)];
    __syncwarp();
    for (int i = lane; i < a.n; i += 32) sB[i] ^= top;
    __syncwarp();
  }
  for (int sweep = 0; sweep < a.sweeps; ++sweep) {
    for (int pos = 0; pos < a.n; ++pos) {
      const int node = __ldg(a.order + pos);
      const int kb = __ldg(a.nbr_ptr + node), ke = __ldg(a.nbr_ptr + node + 1);
      float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
      int k = kb;
      for (; k + 4 <= ke; k += 4) {
        const int j0 = __ldg(a.nbr_col + k), j1 = __ldg(a.nbr_col + k + 1), j2 = __ldg(a.nbr_col + k + 2),
                  j3 = __ldg(a.nbr_col + k + 3);
        acc0 = fmaf(__ldg(a.nbr_w + k), wm_value(sB[j0], sU[j0], lane), acc0);
        acc1 = fmaf(__ldg(a.nbr_w + k + 1), wm_value(sB[j1], sU[j1], lane), acc1);
        acc2 = fmaf(__ldg(a.nbr_w + k + 2), wm_value(sB[j2], sU[j2], lane), acc2);
        acc3 = fmaf(__ldg(a.nbr_w + k + 3), wm_value(sB[j3], sU[j3], lane), acc3);
      }
      for (; k < ke; ++k) {
        const int j = __ldg(a.nbr_col + k);
        acc0 = fmaf(__ldg(a.nbr_w + k), wm_value(sB[j], sU[j], lane), acc0);
      }
      const float s = (acc0 + acc1) + (acc2 + acc3);
      const float thr = __ldg(a.thr + node);
      bool bit;
      if (s >= thr) {
        bit = false;                                   // fl(s + u/4) >= s >= thr
      } else if ((double)s + 0.25 <= (double)thr - 1e-3) {
        bit = true;                                    // s + u/4 < s + 1/4 stays clear of thr for every u in [0, 1)
      } else {
        const uint64_t call = (uint64_t)sweep * a.n + pos;
        float u;
        if (a.explicit_u) {
          u = live ? __ldg(a.explicit_u + call * a.num_chains + chain) : 0.f;
        } else if (direct) {
          const uint64_t ctr = key.offset4 + call * a.rng.iters_per_call;
          u = torch_uniform_from_u32(
              curand_Philox4x32_10(make_uint4((uint32_t)ctr, (uint32_t)(ctr >> 32), (uint32_t)chain, 0u), pkey).x);
        } else {
          TorchRng r = a.rng;
          r.seed = key.seed, r.offset4 = key.offset4;
          u = torch_uniform_from_u32(torch_philox_u32(r, call, (uint32_t)chain));
        }
        bit = __fadd_rn(s, __fmul_rn(u, 0.25f)) < thr;   // node_temp_v += rand / 4; (node_temp_v < wdeg / 2 + 0.125)
      }
      const uint32_t word = __ballot_sync(kFull, bit);
      if (lane == 0) sB[node] = word, sU[node] = 0u;
      __syncwarp();
    }
  }
  // expected cut: sum_e (2 x_u - 1)(2 x_v - 1) w_e  (sampling.py:121); every node has been visited: x in {0, 1}
  float e0 = 0.f, e1 = 0.f;
  int k = 0;
  for (; k + 2 <= a.num_edges; k += 2) {
    const uint32_t d0 = sB[__ldg(a.edge_u + k)] ^ sB[__ldg(a.edge_v + k)];
    const uint32_t d1 = sB[__ldg(a.edge_u + k + 1)] ^ sB[__ldg(a.edge_v + k + 1)];
    const float w0 = __ldg(a.edge_w + k), w1 = __ldg(a.edge_w + k + 1);
    e0 += ((d0 >> lane) & 1u) ? -w0 : w0;
    e1 += ((d1 >> lane) & 1u) ? -w1 : w1;
  }
  if (k < a.num_edges) {
    const uint32_t d0 = sB[__ldg(a.edge_u + k)] ^ sB[__ldg(a.edge_v + k)];
    const float w0 = __ldg(a.edge_w + k);
    e0 += ((d0 >> lane) & 1u) ? -w0 : w0;
  }
  if (live) a.expected[chain] = e0 + e1;
  for (int i = 0; i < a.n; ++i)
    if (live) a.xs[(int64_t)i * a.num_chains + chain] = (float)((sB[i] >> lane) & 1u);
}

}  // namespace rlsb

extern "C" int rlsb_mcpg_weighted_sweeps(int32_t num_nodes, int64_t num_chains, const int32_t* order,
                                         const int32_t* nbr_ptr, const int32_t* nbr_col, const float* nbr_w,
                                         const float* thr, int32_t num_edges, const int32_t* edge_u,
                                         const int32_t* edge_v, const float* edge_w, float* xs, int32_t num_sweeps,
                                         const float* explicit_u, uint64_t seed, uint64_t offset, uint32_t rng_threads,
                                         uint32_t rng_iters, float* expected, void* stream) {
  using namespace rlsb;
  RLSB_REQUIRE(num_nodes > 0 && num_chains >= 0 && num_edges >= 0 && num_sweeps >= 1, RLSB_ERR_INVALID,
               "mcpg_weighted_sweeps: bad size (at least one sweep: the reference's loop runs once for num_ls <= 1)");
  RLSB_REQUIRE(num_chains < (int64_t(1) << 31), RLSB_ERR_UNSUPPORTED, "mcpg_weighted_sweeps: more than 2^31 chains");
  if (num_chains == 0) return RLSB_OK;
  RLSB_REQUIRE(order && nbr_ptr && thr && xs && expected && (num_edges == 0 || (nbr_col && nbr_w && edge_u && edge_v && edge_w)),
               RLSB_ERR_INVALID, "mcpg_weighted_sweeps: null pointer");
  RLSB_REQUIRE(explicit_u || (rng_threads > 0 && rng_iters > 0 && offset % 4 == 0), RLSB_ERR_INVALID,
               "mcpg_weighted_sweeps: no random source");
  WmArgs a{};
  a.n = num_nodes, a.np = (num_nodes + 31) / 32 * 32, a.num_chains = num_chains;
  a.order = order, a.nbr_ptr = nbr_ptr, a.nbr_col = nbr_col, a.nbr_w = nbr_w, a.thr = thr;
  a.num_edges = num_edges, a.edge_u = edge_u, a.edge_v = edge_v, a.edge_w = edge_w;
  a.xs = xs, a.sweeps = num_sweeps, a.explicit_u = explicit_u, a.expected = expected;
  a.rng = TorchRng{seed, offset / 4, rng_threads, rng_iters, nullptr};
  const size_t smem = (size_t)kWmWarps * 2 * a.np * sizeof(uint32_t);
  RLSB_REQUIRE(smem <= 220 * 1024, RLSB_ERR_UNSUPPORTED, "mcpg_weighted_sweeps: %d nodes exceed the shared-memory tiles",
               num_nodes);
  if (smem > 48 * 1024)
    RLSB_CUDA_OK(cudaFuncSetAttribute(mcpg_weighted_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t tiles = (num_chains + kTileEnvs - 1) / kTileEnvs;
  mcpg_weighted_kernel<<<(unsigned)((tiles + kWmWarps - 1) / kWmWarps), kWmWarps * 32, smem,
                         static_cast<cudaStream_t>(stream)>>>(a);
  RLSB_LAUNCH_OK();
  return RLSB_OK;
}
