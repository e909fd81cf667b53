// Single-flip moves chosen per environment (pattern I and greedy):
//
//   step_flip    env_PPO.EnvMaxcut.step (rlsolver/envs/env_PPO.py:92-106): every env flips the node
//                its action names; reward = new cut - old cut.  The reference runs a Python loop over
//                the envs and then re-evaluates all M edges of every env; here the reward is the
//                single-flip gain deg_a - 2*cross_a read off the acted node's neighbours (O(degree),
//                the batched form of S2V_PPO/env.py:197-206).  State = the reference's float32 [E][N].
//   greedy       best-single-flip ascent with the contract of greedy_maxcut
//                (rlsolver/methods/greedy.py:33-78): all N single-flip gains, lowest index among the
//                best, accept only a strictly better cut, stop otherwise.  One CTA per tile of 32
//                envs; the per-(node, env) gains ("local fields") stay resident in shared memory as
//                int8/int16 [node][env] and are updated in O(degree) per accepted flip; the argmax
//                is a lane-per-env scan split over the warps and merged through shared memory.
#include <limits.h>

#include "tile_ops.cuh"

namespace rlsb {

// ---------------------------------------------------------------- pattern-I step
__global__ void __launch_bounds__(256) step_flip_kernel(GraphDev g, float* __restrict__ xs,
                                                        const int64_t* __restrict__ action, int64_t num_envs,
                                                        float* __restrict__ reward, float* __restrict__ cut,
                                                        int32_t* __restrict__ bad_actions) {
  const int lane = threadIdx.x & 31;
  const int64_t env = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (env >= num_envs) return;
  const int64_t a = action[env];
  if (a < 0 || a >= g.n) {               // the reference raises IndexError; here: no move, flagged
    if (lane == 0) {
      reward[env] = 0.f;
      atomicAdd(bad_actions, 1);
    }
    return;
  }
  float* row = xs + env * (int64_t)g.n;
  const bool xa = row[a] > 0.f;
  const int rb = __ldg(g.full_ptr + a), re = __ldg(g.full_ptr + a + 1);
  int cross = 0;
  for (int k = rb + lane; k < re; k += 32) cross += (int)((row[__ldg(g.full_col + k)] > 0.f) != xa);
  cross = warp_sum(cross);
  if (lane == 0) {
    const float gain = (float)((re - rb) - 2 * cross);
    row[a] = xa ? 0.f : 1.f;             // logical_not of a {0,1} float
    const float cur = cut[env] + gain;
    reward[env] = gain;
    cut[env] = cur;
  }
}

// ---------------------------------------------------------------- greedy best flip
constexpr int kGreedyThreads = 512;

template <typename F, int P>
__global__ void __launch_bounds__(kGreedyThreads) greedy_kernel(GraphDev g, uint8_t* __restrict__ xs,
                                                                int64_t num_envs, int64_t* __restrict__ vs,
                                                                int32_t* __restrict__ flips, int max_flips,
                                                                int strict, int cut_warps, int vec4) {
  extern __shared__ __align__(16) uint32_t smem[];
  uint32_t* sP = smem;                                        // packed tile
  F* sF = reinterpret_cast<F*>(smem + g.np);                  // gain[node][env]
  __shared__ int sVal[kGreedyThreads / 32][kTileEnvs];
  __shared__ int sIdx[kGreedyThreads / 32][kTileEnvs];
  __shared__ int sCnt[kTileEnvs];
  __shared__ int sAny;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int64_t tiles = (num_envs + kTileEnvs - 1) / kTileEnvs;
  const SweepView sv = sweep_view(g, g.sweep_blob);
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t env0 = tile * kTileEnvs;
    const int valid = (int)min((int64_t)kTileEnvs, num_envs - env0);
    if (threadIdx.x < kTileEnvs) sCnt[threadIdx.x] = 0;
    if (vec4) pack_tile_to_smem<4>(xs, num_envs, g.n, g.np, tile, sP);
    else pack_tile_to_smem<1>(xs, num_envs, g.n, g.np, tile, sP);
    __syncthreads();
    // gains of every (node, env): one lane per node over the full-neighbour SELL slices
    for (int s = warp; s < sv.sell.num_slices; s += nwarps) {
      const uint32_t node = __ldg(sv.sell.node + s * 32 + lane);
      const bool active = node != 0xFFFFu;
      const uint32_t self = active ? sP[node] : 0u;
      VCount<P> vc;
      sell_cross<P, false>(sv.sell, s, lane, sP, self, vc);
      if (active) {
        const int deg = __ldg(g.full_ptr + node + 1) - __ldg(g.full_ptr + node);
        F* dst = sF + (size_t)node * kTileEnvs;
        if (sizeof(F) == 1) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const uint32_t c = vc.bytes4(q);
            uint32_t out = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) out |= (uint32_t)(uint8_t)(int8_t)(deg - 2 * (int)((c >> (8 * j)) & 0xffu)) << (8 * j);
            reinterpret_cast<uint32_t*>(dst)[q] = out;
          }
        } else {
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            const uint32_t c = vc.halves2(q);
            const uint32_t lo = (uint32_t)(uint16_t)(int16_t)(deg - 2 * (int)(c & 0xffffu));
            const uint32_t hi = (uint32_t)(uint16_t)(int16_t)(deg - 2 * (int)(c >> 16));
            reinterpret_cast<uint32_t*>(dst)[q] = lo | (hi << 16);
          }
        }
      }
    }
    const int cnt = tile_cut_partial(g, sP, cut_warps);
    if (cnt) atomicAdd(&sCnt[lane], cnt);
    __syncthreads();
    int64_t my_vs = sCnt[lane];          // lane = env (every warp holds a copy; warp 0's is written back)
    int my_flips = 0;
    bool alive = lane < valid;
    for (int step = 0; step < max_flips; ++step) {
      // lane = env: best gain over this warp's share of the nodes, lowest index on ties
      int best = INT_MIN, arg = 0;
      for (int i = warp; i < g.n; i += nwarps) {
        const int v = (int)sF[(size_t)i * kTileEnvs + lane];
        if (v > best) best = v, arg = i;
      }
      sVal[warp][lane] = best, sIdx[warp][lane] = arg;
      if (threadIdx.x == 0) sAny = 0;
      __syncthreads();
      best = INT_MIN, arg = 0;
#pragma unroll 4
      for (int w = 0; w < nwarps; ++w) {
        const int v = sVal[w][lane], ix = sIdx[w][lane];
        if (v > best || (v == best && ix < arg)) best = v, arg = ix;
      }
      const bool go = alive && (strict ? best > 0 : best >= 0);
      if (!go) alive = false;            // greedy.py: stop at the first step without a strictly better cut
      if (warp == 0 && go) sAny = 1;
      __syncthreads();
      if (!sAny) break;
      if (warp == 0 && go) {             // O(degree) update of the resident gains, one lane per env
        const uint32_t bit = 1u << lane;
        const uint32_t xa = sP[arg] & bit;
        const int rb = __ldg(g.full_ptr + arg), re = __ldg(g.full_ptr + arg + 1);
        for (int k = rb; k < re; ++k) {
          const int j = __ldg(g.full_col + k);
          F* cell = sF + (size_t)j * kTileEnvs + lane;
          *cell = (F)((int)*cell + (((sP[j] & bit) == xa) ? -2 : 2));   // same side before the flip: now cut
        }
        F* own = sF + (size_t)arg * kTileEnvs + lane;
        *own = (F)(-(int)*own);
        atomicXor(&sP[arg], bit);
        my_vs += best;
        ++my_flips;
      }
      __syncthreads();
    }
    __syncthreads();
    if (warp == 0 && lane < valid) {
      vs[env0 + lane] = my_vs;
      if (flips) flips[env0 + lane] = my_flips;
    }
    if (vec4) unpack_tile_from_smem<4>(sP, xs, num_envs, g.n, g.np, tile);
    else unpack_tile_from_smem<1>(sP, xs, num_envs, g.n, g.np, tile);
    __syncthreads();
  }
}

}  // namespace rlsb

extern "C" {

int rlsb_step_flip(const rlsb_graph_t* gh, float* xs, const int64_t* action, int64_t num_envs, float* reward,
                   float* cut, int32_t* bad_actions, void* stream) {
  using namespace rlsb;
  const GraphDev* g;
  if (int rc = graph_check(gh, &g, "step_flip")) return rc;
  RLSB_REQUIRE(num_envs >= 0, RLSB_ERR_INVALID, "step_flip: negative num_envs");
  if (num_envs == 0) return RLSB_OK;
  RLSB_REQUIRE(xs && action && reward && cut && bad_actions, RLSB_ERR_INVALID, "step_flip: null pointer");
  step_flip_kernel<<<(unsigned)((num_envs + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      *g, xs, action, num_envs, reward, cut, bad_actions);
  RLSB_LAUNCH_OK();
  return RLSB_OK;
}

int rlsb_greedy_best_flip(const rlsb_graph_t* gh, uint8_t* xs, int64_t num_envs, int64_t* vs, int32_t* flips,
                          int32_t max_flips, int32_t strict, void* stream) {
  using namespace rlsb;
  const GraphDev* g;
  if (int rc = graph_check(gh, &g, "greedy_best_flip")) return rc;
  RLSB_REQUIRE(num_envs >= 0 && max_flips >= 0, RLSB_ERR_INVALID, "greedy_best_flip: negative size");
  if (num_envs == 0 || g->n == 0) return RLSB_OK;
  RLSB_REQUIRE(xs && vs, RLSB_ERR_INVALID, "greedy_best_flip: null pointer");
  const bool small = g->max_full_deg <= 127;
  const size_t smem = (size_t)g->np * 4 + (size_t)g->np * kTileEnvs * (small ? 1 : 2);
  RLSB_REQUIRE(smem <= 220 * 1024, RLSB_ERR_UNSUPPORTED,
               "greedy_best_flip: %d nodes x 32 envs of resident gains (%zu bytes) exceed shared memory", g->n, smem);
  const int64_t tiles = (num_envs + kTileEnvs - 1) / kTileEnvs;
  const unsigned grid = (unsigned)(tiles < 4 * kNumSMs ? tiles : 4 * kNumSMs);
  auto st = static_cast<cudaStream_t>(stream);
  const int cw = cut_warps_for(g->m, kGreedyThreads / 32);
  const int vec4 = rows_vec4_ok(xs, g->n) ? 1 : 0;
#define RLSB_GREEDY(F, P)                                                                                   \
  do {                                                                                                      \
    if (smem > 48 * 1024)                                                                                   \
      RLSB_CUDA_OK(cudaFuncSetAttribute(greedy_kernel<F, P>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                        (int)smem));                                                        \
    greedy_kernel<F, P><<<grid, kGreedyThreads, smem, st>>>(*g, xs, num_envs, vs, flips, max_flips, strict, \
                                                            cw, vec4);                                      \
  } while (0)
  if (g->max_full_deg <= 63) RLSB_GREEDY(int8_t, 6);
  else if (small) RLSB_GREEDY(int8_t, 8);
  else RLSB_GREEDY(int16_t, 12);
#undef RLSB_GREEDY
  RLSB_LAUNCH_OK();
  return RLSB_OK;
}

}  // extern "C"
