// Local search kernels: the three phases of EnvMaxcut.local_search_inplace
// (rlsolver/envs/env_L2A.py:87-116) and LocalSearch.random_search
// (rlsolver/methods/LocalSearch.py:53-86).
//
//   ls_thresh       kth-value threshold of the noise-perturbed weights (env_L2A.py:94-96)
//   ls_noisy_iters  noisy multi-flip + full re-evaluation + accept-if-not-worse (97-107),
//                   all iterations of one call fused: the tile lives in shared memory and
//                   only the float32 noise (4 B per env-node-iteration) streams from HBM
//   flip_sweep      the exhaustive single-flip pass (110-115): the reference does N full
//                   evaluations on N clones; here each node's gain is recomputed from the
//                   packed tile in O(degree) and nodes are scheduled by dependency level.
//
// Floating point: spin_rand = ws + noise * rd_std is evaluated exactly as the reference's
// two torch kernels do -- one IEEE round-to-nearest multiply, one add, no FMA contraction.
#include <math.h>

#include "tile_ops.cuh"

namespace rlsb {

__device__ __forceinline__ float spin_rand(int deg, int mult, int cross, float noise, float rd_std) {
  return __fadd_rn((float)(deg - mult * cross), __fmul_rn(noise, rd_std));
}

__device__ __forceinline__ float rd_std_of(int mult, int cmin, int cmax, float noise_std) {
  return __fmul_rn((float)(mult * (cmax - cmin)), noise_std);
}

// order-preserving float -> uint32 key (so REDUX max works on floats)
__device__ __forceinline__ uint32_t float_key(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_float(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// ---------------------------------------------------------------- thresh (kthvalue)
// One warp per environment.  Each lane keeps the KMAX largest values of its strided
// share in a sorted register list; the lists are then merged by K rounds of warp-max.
template <int KMAX>
__global__ void __launch_bounds__(256) ls_thresh_kernel(GraphDev g, const uint16_t* __restrict__ cross,
                                                        const int32_t* __restrict__ col_min,
                                                        const int32_t* __restrict__ col_max, int mult,
                                                        float noise_std, const float* __restrict__ noise, int kth_big,
                                                        int64_t num_envs, float* __restrict__ thresh) {
  const int lane = threadIdx.x & 31;
  const int64_t env = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (env >= num_envs) return;
  float top[KMAX];
#pragma unroll
  for (int t = 0; t < KMAX; ++t) top[t] = -INFINITY;
  const uint16_t* crow = cross + env * (int64_t)g.np;
  const float* nrow = noise + env * (int64_t)g.n;
  for (int i = lane; i < g.n; i += 32) {
    const int deg = __ldg(g.listed_ptr + i + 1) - __ldg(g.listed_ptr + i);
    const float rd = rd_std_of(mult, __ldg(col_min + i), __ldg(col_max + i), noise_std);
    float s = spin_rand(deg, mult, __ldg(crow + i), __ldg(nrow + i), rd);
    if (s > top[KMAX - 1]) {
#pragma unroll
      for (int t = 0; t < KMAX; ++t) {
        const float hi = fmaxf(top[t], s);
        s = fminf(top[t], s);
        top[t] = hi;
      }
    }
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

// ---------------------------------------------------------------- noisy multi-flip iterations
constexpr int kNIThreads = 1024;
constexpr int kNIMaxIters = 16;   // noise tensors per launch (pointers travel by value)
struct NoisePtrs {
  const float* p[kNIMaxIters];
};

__global__ void __launch_bounds__(kNIThreads) ls_noisy_iters_kernel(
    GraphDev g, uint32_t* __restrict__ packed, int64_t* __restrict__ vs, const uint16_t* __restrict__ cross,
    const int32_t* __restrict__ col_min, const int32_t* __restrict__ col_max, int mult, float noise_std,
    NoisePtrs noise_ptrs, int num_iters, const float* __restrict__ thresh, int64_t num_envs) {
  extern __shared__ uint32_t smem[];
  uint32_t* sP = smem;           // accepted state of the tile
  uint32_t* sX = smem + g.np;    // candidate state
  __shared__ float sThresh[kTileEnvs];
  __shared__ int sCnt[kTileEnvs];
  __shared__ uint32_t sAccept;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int64_t tiles = (num_envs + kTileEnvs - 1) / kTileEnvs;
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t env0 = tile * kTileEnvs;
    const int valid = (int)min((int64_t)kTileEnvs, num_envs - env0);
    for (int i = threadIdx.x; i < g.np; i += blockDim.x) sP[i] = packed[tile * g.np + i];
    if (threadIdx.x < kTileEnvs) sThresh[threadIdx.x] = threadIdx.x < valid ? thresh[env0 + threadIdx.x] : INFINITY;
    int64_t my_vs = 0;   // warp 0: lane e owns env e's value
    if (warp == 0 && lane < valid) my_vs = vs[env0 + lane];
    __syncthreads();
    for (int it = 0; it < num_iters; ++it) {
      const float* __restrict__ noise = noise_ptrs.p[it];
      if (threadIdx.x < kTileEnvs) sCnt[threadIdx.x] = 0;
      // phase A: flip mask, one strip of 32 nodes per warp pass; every load is a coalesced row segment
      for (int strip = warp; strip * 32 < g.np; strip += nwarps) {
        const int i = strip * 32 + lane;
        uint32_t word = 0;
        if (i < g.n) {
          const int deg = __ldg(g.listed_ptr + i + 1) - __ldg(g.listed_ptr + i);
          const float rd = rd_std_of(mult, __ldg(col_min + i), __ldg(col_max + i), noise_std);
          const float* np_ = noise + env0 * (int64_t)g.n + i;
          const uint16_t* cp_ = cross + env0 * (int64_t)g.np + i;
#pragma unroll 8
          for (int e = 0; e < kTileEnvs; ++e) {
            if (e < valid) {
              const float s = spin_rand(deg, mult, __ldg(cp_ + (int64_t)e * g.np), __ldg(np_ + (int64_t)e * g.n), rd);
              word |= (uint32_t)(s > sThresh[e]) << e;
            }
          }
        }
        sX[i] = sP[i] ^ word;
      }
      __syncthreads();
      // phase B: objective of the candidate
      const int cnt = tile_cut_partial(g, sX);
      if (cnt) atomicAdd(&sCnt[lane], cnt);
      __syncthreads();
      // phase C: keep rows that are not worse (vs1 >= vs0, util_read_data.py:199)
      if (warp == 0) {
        const int64_t cand = sCnt[lane];
        const bool keep = lane < valid && cand >= my_vs;
        if (keep) my_vs = cand;
        const unsigned a = __ballot_sync(kFull, keep);
        if (lane == 0) sAccept = a;
      }
      __syncthreads();
      const uint32_t a = sAccept;
      for (int i = threadIdx.x; i < g.np; i += blockDim.x) sP[i] = (sX[i] & a) | (sP[i] & ~a);
      __syncthreads();
    }
    for (int i = threadIdx.x; i < g.np; i += blockDim.x) packed[tile * g.np + i] = sP[i];
    if (warp == 0 && lane < valid) vs[env0 + lane] = my_vs;
    __syncthreads();
  }
}

// ---------------------------------------------------------------- exhaustive single-flip sweep
// Gauss-Seidel over nodes 0..N-1 with acceptance gain >= 0.  Nodes of one dependency level
// (graph_store.cu) are pairwise non-adjacent, so the warps of the CTA decide them
// concurrently and a barrier separates levels: the result equals the sequential order.
constexpr int kSweepThreads = 1024;

__global__ void __launch_bounds__(kSweepThreads) flip_sweep_kernel(GraphDev g, uint32_t* __restrict__ packed,
                                                                   int64_t* __restrict__ vs, int64_t num_envs) {
  extern __shared__ uint32_t sP[];
  __shared__ int sGain[kTileEnvs];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int64_t tiles = (num_envs + kTileEnvs - 1) / kTileEnvs;
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t env0 = tile * kTileEnvs;
    for (int i = threadIdx.x; i < g.np; i += blockDim.x) sP[i] = packed[tile * g.np + i];
    if (threadIdx.x < kTileEnvs) sGain[threadIdx.x] = 0;
    __syncthreads();
    int gained = 0;   // lane e: total gain this warp accepted for env e
    for (int l = 0; l < g.levels; ++l) {
      const int lb = __ldg(g.level_ptr + l), le = __ldg(g.level_ptr + l + 1);
      for (int idx = lb + warp; idx < le; idx += nwarps) {
        const int i = __ldg(g.level_nodes + idx);
        const int rb = __ldg(g.full_ptr + i), re = __ldg(g.full_ptr + i + 1);
        const uint32_t pi = sP[i];
        int cross = 0;
        for (int c = rb; c < re; c += 32) {
          const int k = c + lane;
          uint32_t x = 0;
          if (k < re) x = sP[__ldg(g.full_col + k)] ^ pi;
          cross += __popc(transpose32(x, lane));
        }
        const int gain = (re - rb) - 2 * cross;      // same-side minus other-side neighbours
        const bool keep = gain >= 0;
        const unsigned mask = __ballot_sync(kFull, keep);
        if (lane == 0) sP[i] = pi ^ mask;
        if (keep) gained += gain;
      }
      __syncthreads();
    }
    if (gained) atomicAdd(&sGain[lane], gained);
    __syncthreads();
    for (int i = threadIdx.x; i < g.np; i += blockDim.x) packed[tile * g.np + i] = sP[i];
    if (threadIdx.x < kTileEnvs && env0 + threadIdx.x < num_envs) vs[env0 + threadIdx.x] += sGain[threadIdx.x];
    __syncthreads();
  }
}

template <typename K>
static int allow_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024)
    RLSB_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return RLSB_OK;
}

}  // namespace rlsb

extern "C" {

int rlsb_ls_thresh(const rlsb_graph_t* gh, const uint16_t* cross, const int32_t* col_min, const int32_t* col_max,
                   int32_t ws_mult, float noise_std, const float* noise, int32_t num_spin, int64_t num_envs,
                   float* thresh, void* stream) {
  using namespace rlsb;
  const GraphDev* g = graph_dev(gh);
  RLSB_REQUIRE(gh != nullptr, RLSB_ERR_INVALID, "ls_thresh: null graph");
  RLSB_REQUIRE(g != nullptr, RLSB_ERR_NODEVICE, "ls_thresh: graph has no device image");
  RLSB_REQUIRE(num_envs >= 0, RLSB_ERR_INVALID, "ls_thresh: negative num_envs");
  // torch.kthvalue(k = N - num_spin) needs 1 <= k <= N
  RLSB_REQUIRE(num_spin >= 0 && num_spin < g->n, RLSB_ERR_INVALID,
               "ls_thresh: k = N - num_spin = %d out of range for N = %d (torch.kthvalue raises)", g->n - num_spin,
               g->n);
  RLSB_REQUIRE(ws_mult == 1 || ws_mult == 2, RLSB_ERR_INVALID, "ls_thresh: ws_mult must be 1 or 2");
  if (num_envs == 0) return RLSB_OK;
  RLSB_REQUIRE(cross && col_min && col_max && noise && thresh, RLSB_ERR_INVALID, "ls_thresh: null pointer");
  const int kth_big = num_spin + 1;   // kth smallest with k = N - num_spin  ==  (num_spin+1)-th largest
  RLSB_REQUIRE(kth_big <= 32, RLSB_ERR_UNSUPPORTED, "ls_thresh: num_spin %d above the in-register limit 31", num_spin);
  auto st = static_cast<cudaStream_t>(stream);
  const unsigned grid = (unsigned)((num_envs + 7) / 8);
#define RLSB_THRESH(KMAX)                                                                                         \
  ls_thresh_kernel<KMAX><<<grid, 256, 0, st>>>(*g, cross, col_min, col_max, ws_mult, noise_std, noise, kth_big, \
                                               num_envs, thresh)
  if (kth_big <= 6) RLSB_THRESH(6);
  else if (kth_big <= 10) RLSB_THRESH(10);
  else if (kth_big <= 18) RLSB_THRESH(18);
  else RLSB_THRESH(32);
#undef RLSB_THRESH
  RLSB_LAUNCH_OK();
  return RLSB_OK;
}

int rlsb_ls_noisy_iters(const rlsb_graph_t* gh, uint32_t* packed, int64_t* vs, const uint16_t* cross,
                        const int32_t* col_min, const int32_t* col_max, int32_t ws_mult, float noise_std,
                        const float* const* h_noise_ptrs, int32_t num_iters, const float* thresh,
                        int64_t num_envs, void* stream) {
  using namespace rlsb;
  const GraphDev* g = graph_dev(gh);
  RLSB_REQUIRE(gh != nullptr, RLSB_ERR_INVALID, "ls_noisy_iters: null graph");
  RLSB_REQUIRE(g != nullptr, RLSB_ERR_NODEVICE, "ls_noisy_iters: graph has no device image");
  RLSB_REQUIRE(num_envs >= 0 && num_iters >= 0, RLSB_ERR_INVALID, "ls_noisy_iters: negative size");
  RLSB_REQUIRE(ws_mult == 1 || ws_mult == 2, RLSB_ERR_INVALID, "ls_noisy_iters: ws_mult must be 1 or 2");
  if (num_envs == 0 || num_iters == 0 || g->n == 0) return RLSB_OK;
  RLSB_REQUIRE(packed && vs && cross && col_min && col_max && h_noise_ptrs && thresh, RLSB_ERR_INVALID,
               "ls_noisy_iters: null pointer");
  const size_t smem = 2 * (size_t)g->np * sizeof(uint32_t);
  RLSB_REQUIRE(smem <= 220 * 1024, RLSB_ERR_UNSUPPORTED, "ls_noisy_iters: %d nodes exceed the shared-memory tile",
               g->n);
  int rc;
  if ((rc = allow_smem(ls_noisy_iters_kernel, smem))) return rc;
  const int64_t tiles = (num_envs + kTileEnvs - 1) / kTileEnvs;
  const unsigned grid = (unsigned)(tiles < 8 * kNumSMs ? tiles : 8 * kNumSMs);
  for (int done = 0; done < num_iters; done += kNIMaxIters) {
    const int now = num_iters - done < kNIMaxIters ? num_iters - done : kNIMaxIters;
    NoisePtrs np{};
    for (int k = 0; k < now; ++k) {
      RLSB_REQUIRE(h_noise_ptrs[done + k] != nullptr, RLSB_ERR_INVALID, "ls_noisy_iters: null noise tensor %d",
                   done + k);
      np.p[k] = h_noise_ptrs[done + k];
    }
    ls_noisy_iters_kernel<<<grid, kNIThreads, smem, static_cast<cudaStream_t>(stream)>>>(
        *g, packed, vs, cross, col_min, col_max, ws_mult, noise_std, np, now, thresh, num_envs);
    RLSB_LAUNCH_OK();
  }
  return RLSB_OK;
}

int rlsb_flip_sweep(const rlsb_graph_t* gh, uint32_t* packed, int64_t* vs, int64_t num_envs, void* stream) {
  using namespace rlsb;
  const GraphDev* g = graph_dev(gh);
  RLSB_REQUIRE(gh != nullptr, RLSB_ERR_INVALID, "flip_sweep: null graph");
  RLSB_REQUIRE(g != nullptr, RLSB_ERR_NODEVICE, "flip_sweep: graph has no device image");
  RLSB_REQUIRE(num_envs >= 0, RLSB_ERR_INVALID, "flip_sweep: negative num_envs");
  if (num_envs == 0 || g->n == 0) return RLSB_OK;
  RLSB_REQUIRE(packed && vs, RLSB_ERR_INVALID, "flip_sweep: null pointer");
  const size_t smem = (size_t)g->np * sizeof(uint32_t);
  RLSB_REQUIRE(smem <= 220 * 1024, RLSB_ERR_UNSUPPORTED, "flip_sweep: %d nodes exceed the shared-memory tile", g->n);
  int rc;
  if ((rc = allow_smem(flip_sweep_kernel, smem))) return rc;
  const int64_t tiles = (num_envs + kTileEnvs - 1) / kTileEnvs;
  const unsigned grid = (unsigned)(tiles < 8 * kNumSMs ? tiles : 8 * kNumSMs);
  flip_sweep_kernel<<<grid, kSweepThreads, smem, static_cast<cudaStream_t>(stream)>>>(*g, packed, vs, num_envs);
  RLSB_LAUNCH_OK();
  return RLSB_OK;
}

}  // extern "C"
