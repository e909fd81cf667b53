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
  // four edges per 16-byte load (the list is zero-padded to whole quads), the next quad in flight while this
  // one is evaluated; pad edges (k >= m) are skipped -- (0, 0) would add p0 + p0 - 2 p0^2
  const int quads = (g.m + 3) >> 2;
  const uint4* pairs = reinterpret_cast<const uint4*>(g.edge_pair);
  uint4 cur = threadIdx.x < quads ? __ldg(pairs + threadIdx.x) : make_uint4(0, 0, 0, 0);
  for (int q = threadIdx.x; q < quads; q += kRelaxThreads) {
    const int qn = q + kRelaxThreads;
    const uint4 nxt = qn < quads ? __ldg(pairs + qn) : make_uint4(0, 0, 0, 0);
    const uint32_t pr[4] = {cur.x, cur.y, cur.z, cur.w};
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (4 * q + c < g.m) {
        const float* a = sp + (pr[c] & 0xffffu) * ENVS;
        const float* b = sp + (pr[c] >> 16) * ENVS;
#pragma unroll
        for (int e = 0; e < ENVS; ++e) acc[e] += fmaf(-2.f * a[e], b[e], a[e] + b[e]);
      }
    }
    cur = nxt;
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

// Gradient by gathering (graphs without self loops): d/dp_u = deg_u - 2 * sum over the neighbours v of p_v.  One
// lane per row of the full-neighbour SELL structure (the sweep's; short rows are padded with the node's own id,
// subtracted again through the true degree), one LDS per neighbour for all ENVS envs, no atomics, deterministic.
template <int ENVS>
__global__ void __launch_bounds__(kRelaxThreads) relaxed_cut_grad_gather_kernel(GraphDev g, const float* __restrict__ probs,
                                                                                const float* __restrict__ grad_out,
                                                                                int64_t num_envs,
                                                                                float* __restrict__ grad_probs) {
  extern __shared__ float sp[];
  const int64_t env0 = (int64_t)blockIdx.x * ENVS;
  stage_rows<ENVS>(probs, num_envs, env0, g.n, sp);
  __syncthreads();
  const SweepView sv = sweep_view(g, g.sweep_blob);
  float go[ENVS];
#pragma unroll
  for (int e = 0; e < ENVS; ++e) go[e] = env0 + e < num_envs ? -__ldg(grad_out + env0 + e) : 0.f;
  for (int slot = threadIdx.x; slot < g.num_sweep_slices * 32; slot += kRelaxThreads) {
    const uint32_t u = __ldg(sv.sell.node + slot);
    if (u == 0xFFFFu) continue;
    const int slice = slot >> 5;
    const int gb = __ldg(sv.sell.off + slice), nb = __ldg(sv.sell.off + slice + 1) - gb;
    const uint2* col = reinterpret_cast<const uint2*>(sv.sell.col) + (int64_t)gb * 32 + (slot & 31);
    float sum[ENVS];
#pragma unroll
    for (int e = 0; e < ENVS; ++e) sum[e] = 0.f;
    for (int b = 0; b < nb; ++b) {
      const uint2 id = __ldg(col + b * 32);
      const uint32_t ids[4] = {id.x & 0xffffu, id.x >> 16, id.y & 0xffffu, id.y >> 16};
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int e = 0; e < ENVS; ++e) sum[e] += sp[ids[c] * ENVS + e];
    }
    const int deg = __ldg(g.full_ptr + u + 1) - __ldg(g.full_ptr + u);
    const float pad = (float)(4 * nb - deg), fdeg = (float)deg;
#pragma unroll
    for (int e = 0; e < ENVS; ++e)
      if (env0 + e < num_envs)
        grad_probs[(env0 + e) * (int64_t)g.n + u] = go[e] * (fdeg - 2.f * (sum[e] - pad * sp[u * ENVS + e]));
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
  if (g->m == g->mf / 2) {          // no self loops: every edge sits twice in the full-neighbour structure
    const size_t row1 = (size_t)g->np * sizeof(float);
    RLSB_REQUIRE(row1 <= kRelaxSmem, RLSB_ERR_UNSUPPORTED, "relaxed_cut_grad: %d nodes exceed the shared-memory row", g->n);
#define RLSB_RELAX(E)                                                                                   \
  {                                                                                                     \
    if (int rc = relax_smem(relaxed_cut_grad_gather_kernel<E>, row1 * E)) return rc;                    \
    relaxed_cut_grad_gather_kernel<E><<<(unsigned)((num_envs + E - 1) / E), kRelaxThreads, row1 * E, st>>>(             \
        *g, probs, grad_out, num_envs, grad_probs);                                                     \
  }
    if (row1 * 4 <= 100 * 1024 && num_envs >= 4 * 2 * kNumSMs) RLSB_RELAX(4)
    else if (row1 * 2 <= kRelaxSmem && num_envs >= 2 * 2 * kNumSMs) RLSB_RELAX(2)
    else RLSB_RELAX(1)
#undef RLSB_RELAX
    RLSB_LAUNCH_OK();
    return RLSB_OK;
  }
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
