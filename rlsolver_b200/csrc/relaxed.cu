// Relaxed (probabilistic) max-cut objective and its gradient.
//
//   SimulatorMaxcut.get_objectives        rlsolver/envs/env_k_spin.py:191-193   -(p0 + p1 - 2 p0 p1).sum(1)
//   SimulatorMaxcut.get_objectives_using_for_loop            :164-189          the same sum, node by node
//   PIGNN hamiltonian_maxcut              rlsolver/methods/PIGNN/util.py:4-8    sum(2 p_i p_j - p_i - p_j)
//
// The reference materialises two [E, M] gathers (env_k_spin.py:199-208: 8 bytes of index + 8 bytes of
// gathered floats per env-edge) and lets autograd scatter the gradient back.  Here a CTA keeps the
// probability rows of ENVS environments in shared memory (interleaved [node][env]: one LDS gives a node's
// value for all ENVS envs), streams the original edge list (u | v << 16, 4 bytes per edge, L2 resident)
// once for all of them and reduces; HBM traffic is the 4*E*N bytes of the rows.  The backward kernel walks the
// same list and accumulates d/dp_u = sum over incident edges of (1 - 2 p_other) with shared-memory atomics
// (self loops and duplicate edges come out right by construction).  float32, summation order differs from
// torch's: parity within 1e-5 relative (tests/test_gpu_relaxed.py).
#include "common.cuh"

namespace rlsb {

constexpr int kRelaxThreads = 256;

template <int ENVS>
__device__ __forceinline__ void stage_rows(const float* __restrict__ probs, int64_t num_envs, int64_t env0, int n,
                                           float* sp) {
#pragma unroll
  for (int e = 0; e < ENVS; ++e) {
    const bool live = env0 + e < num_envs;
    const float* row = probs + (env0 + e) * (int64_t)n;
    for (int i = threadIdx.x; i < n; i += kRelaxThreads) sp[i * ENVS + e] = live ? __ldg(row + i) : 0.f;
  }
}

template <int ENVS>
__global__ void __launch_bounds__(kRelaxThreads) relaxed_cut_kernel(GraphDev g, const float* __restrict__ probs,
                                                                    int64_t num_envs, float* __restrict__ out) {
  extern __shared__ float sp[];
  __shared__ float sRed[kRelaxThreads / 32][ENVS];
  const int64_t env0 = (int64_t)blockIdx.x * ENVS;
  stage_rows<ENVS>(probs, num_envs, env0, g.n, sp);
  __syncthreads();
  float acc[ENVS];
#pragma unroll
  for (int e = 0; e < ENVS; ++e) acc[e] = 0.f;
  for (int k = threadIdx.x; k < g.m; k += kRelaxThreads) {
    const uint32_t pr = __ldg(g.edge_pair + k);
    const float* a = sp + (pr & 0xffffu) * ENVS;
    const float* b = sp + (pr >> 16) * ENVS;
#pragma unroll
    for (int e = 0; e < ENVS; ++e) acc[e] += fmaf(-2.f * a[e], b[e], a[e] + b[e]);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int e = 0; e < ENVS; ++e) {
    float v = acc[e];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    if (lane == 0) sRed[warp][e] = v;
  }
  __syncthreads();
  if (threadIdx.x < ENVS && env0 + threadIdx.x < num_envs) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < kRelaxThreads / 32; ++w) v += sRed[w][threadIdx.x];
    out[env0 + threadIdx.x] = -v;
  }
}

template <int ENVS>
__global__ void __launch_bounds__(kRelaxThreads) relaxed_cut_grad_kernel(GraphDev g, const float* __restrict__ probs,
                                                                         const float* __restrict__ grad_out,
                                                                         int64_t num_envs, float* __restrict__ grad_probs) {
  extern __shared__ float sp[];
  float* sg = sp + (size_t)g.np * ENVS;
  const int64_t env0 = (int64_t)blockIdx.x * ENVS;
  stage_rows<ENVS>(probs, num_envs, env0, g.n, sp);
  for (int i = threadIdx.x; i < g.n * ENVS; i += kRelaxThreads) sg[i] = 0.f;
  __syncthreads();
  for (int k = threadIdx.x; k < g.m; k += kRelaxThreads) {
    const uint32_t pr = __ldg(g.edge_pair + k);
    const uint32_t u = pr & 0xffffu, v = pr >> 16;
#pragma unroll
    for (int e = 0; e < ENVS; ++e) {
      const float a = sp[u * ENVS + e], b = sp[v * ENVS + e];
      atomicAdd(&sg[u * ENVS + e], 1.f - 2.f * b);
      atomicAdd(&sg[v * ENVS + e], 1.f - 2.f * a);
    }
  }
  __syncthreads();
#pragma unroll
  for (int e = 0; e < ENVS; ++e) {
    if (env0 + e >= num_envs) break;
    const float go = -__ldg(grad_out + env0 + e);
    float* row = grad_probs + (env0 + e) * (int64_t)g.n;
    for (int i = threadIdx.x; i < g.n; i += kRelaxThreads) row[i] = go * sg[i * ENVS + e];
  }
}

constexpr size_t kRelaxSmem = 200 * 1024;

template <typename K>
static int relax_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024)
    RLSB_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return RLSB_OK;
}

}  // namespace rlsb

extern "C" {

int rlsb_relaxed_cut(const rlsb_graph_t* gh, const float* probs, int64_t num_envs, float* out, void* stream) {
  using namespace rlsb;
  const GraphDev* g;
  if (int rc = graph_check(gh, &g, "relaxed_cut")) return rc;
  RLSB_REQUIRE(num_envs >= 0, RLSB_ERR_INVALID, "relaxed_cut: negative num_envs");
  if (num_envs == 0) return RLSB_OK;
  RLSB_REQUIRE(probs && out, RLSB_ERR_INVALID, "relaxed_cut: null pointer");
  auto st = static_cast<cudaStream_t>(stream);
  const size_t row = (size_t)g->np * sizeof(float);
  RLSB_REQUIRE(row <= kRelaxSmem, RLSB_ERR_UNSUPPORTED, "relaxed_cut: %d nodes exceed the shared-memory row", g->n);
#define RLSB_RELAX(E)                                                                                   \
  {                                                                                                     \
    if (int rc = relax_smem(relaxed_cut_kernel<E>, row * E)) return rc;                                 \
    relaxed_cut_kernel<E><<<(unsigned)((num_envs + E - 1) / E), kRelaxThreads, row * E, st>>>(*g, probs, num_envs, out); \
  }
  if (row * 4 <= 100 * 1024 && num_envs >= 4 * 2 * kNumSMs) RLSB_RELAX(4)
  else if (row * 2 <= kRelaxSmem && num_envs >= 2 * 2 * kNumSMs) RLSB_RELAX(2)
  else RLSB_RELAX(1)
#undef RLSB_RELAX
  RLSB_LAUNCH_OK();
  return RLSB_OK;
}

int rlsb_relaxed_cut_grad(const rlsb_graph_t* gh, const float* probs, const float* grad_out, int64_t num_envs,
                          float* grad_probs, void* stream) {
  using namespace rlsb;
  const GraphDev* g;
  if (int rc = graph_check(gh, &g, "relaxed_cut_grad")) return rc;
  RLSB_REQUIRE(num_envs >= 0, RLSB_ERR_INVALID, "relaxed_cut_grad: negative num_envs");
  if (num_envs == 0) return RLSB_OK;
  RLSB_REQUIRE(probs && grad_out && grad_probs, RLSB_ERR_INVALID, "relaxed_cut_grad: null pointer");
  auto st = static_cast<cudaStream_t>(stream);
  const size_t row = 2 * (size_t)g->np * sizeof(float);       // probabilities + gradient accumulators
  RLSB_REQUIRE(row <= kRelaxSmem, RLSB_ERR_UNSUPPORTED, "relaxed_cut_grad: %d nodes exceed the shared-memory rows", g->n);
#define RLSB_RELAX(E)                                                                                   \
  {                                                                                                     \
    if (int rc = relax_smem(relaxed_cut_grad_kernel<E>, row * E)) return rc;                            \
    relaxed_cut_grad_kernel<E><<<(unsigned)((num_envs + E - 1) / E), kRelaxThreads, row * E, st>>>(*g, probs, grad_out, \
                                                                                                 num_envs, grad_probs); \
  }
  if (row * 4 <= 100 * 1024 && num_envs >= 4 * 2 * kNumSMs) RLSB_RELAX(4)
  else if (row * 2 <= kRelaxSmem && num_envs >= 2 * 2 * kNumSMs) RLSB_RELAX(2)
  else RLSB_RELAX(1)
#undef RLSB_RELAX
  RLSB_LAUNCH_OK();
  return RLSB_OK;
}

}  // extern "C"
