// Float tail of one ISCO / PISCO Metropolis-Hastings step as two kernels (one CTA per chain):
//
//   isco_propose : local distribution of x (log-softmax of the flip gains / 2T), Gumbel top-k choice of
//                  path_length[b] sites (rlsolver/methods/ISCO/util.py:19-60 `multinomial`), the log-probability of
//                  drawing them in that order without replacement (`noreplacement_sampling_renormalize`, :12-17), and
//                  the proposal y = x with those sites flipped (rlsolver/envs/env_ISCO.py:37-49 / 394-405).
//   isco_accept  : local distribution of y, the reverse-order log-probability (env_ISCO.py:65-77 / 420-431), the MH
//                  ratio and the accept (util.py:62-75 `mh_step`), the next state.
//
// The reference spends 2 full sorts (`torch.sort`, `argsort`), an `argsort` for the reverse order, gather / scatter /
// cumsum passes over [B, N] and ~15 elementwise launches per step on this.  Here a chain lives in shared memory: one
// bitonic sort of (perturbed value, site) keys gives both the top-k set and its order; only the k chosen sites
// enter the renormalisation (every other term of the reference's cumsum is masked out or exactly zero).  The integer
// part -- cut values and per-site cross counts -- comes from the packed-spin kernels (cross_counts.cu, cut_eval.cu).
//
// Numerics: same float32 expressions as the reference's torch ops (division by T, exp / log / log1p / expm1 of
// libdevice, the -0.693 switch of log1mexp); sums (softmax normaliser, cumulative probabilities) run in a fixed
// order that is not torch's, so values agree to a few ulp and decisions agree unless two candidates are within that
// distance (tests: states equal on the reference's trajectories, floats to 1e-5).
#include <cuda_fp16.h>
#include <float.h>

#include "common.cuh"

namespace rlsb {

constexpr int kIscoThreads = 256;

struct IscoArgs {
  const void* x;            // [B][ld] float32 or float16, entries {0, 1}
  void* y;                  // [B][ld] proposal (propose) / next state (accept)
  int half;                 // 1: float16 state (PISCO)
  const void* cross;        // [B][cross_ld]: uint16 cross counts (weighted == 0) or int32 weighted cut sums
  int weighted, cross_ld;
  const int32_t* deg;       // [n] degree / weighted degree
  const int64_t* cut;       // [B]
  int pisco;                // energy formula: 0 ISCO (#cut / T), 1 PISCO (-1/4 fp16(s^T A s) / T)
  const float* temperature; // scalar on the device
  const int64_t* path_length;   // [B]
  const float* u;           // propose: [B][ld] uniforms of gumbel(); accept: [B] uniforms of bernoulli_logp()
  int32_t* sel;             // [B][kmax] chosen sites in forward order
  int kmax;
  float* ll_x;              // [B] energy of x            (propose out / accept in)
  float* ll_x2y;            // [B]                        (propose out / accept in)
  float* ll_y_t;            // [B] ll_y * T               (accept out)
  float* acc;               // [B] exp(log_acc)           (accept out)
  int n, ld, npow2;
  int64_t num_chains;
};

__device__ __forceinline__ float block_reduce_max(float v, float* red) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, off));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = red[0];
  for (int w = 1; w < kIscoThreads / 32; ++w) r = fmaxf(r, red[w]);
  return r;
}
__device__ __forceinline__ float block_reduce_sum(float v, float* red) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(kFull, v, off);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
  for (int w = 0; w < kIscoThreads / 32; ++w) r += red[w];
  return r;
}
__device__ __forceinline__ long long block_reduce_sum_ll(long long v, long long* red) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(kFull, v, off);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  long long r = 0;
  for (int w = 0; w < kIscoThreads / 32; ++w) r += red[w];
  return r;
}

__device__ __forceinline__ float isco_x(const IscoArgs& a, int64_t b, int i) {
  return a.half ? __half2float(reinterpret_cast<const __half*>(a.x)[b * a.ld + i])
                : reinterpret_cast<const float*>(a.x)[b * a.ld + i];
}
// d_i * sum_j A_ij d_j for site i of chain b (0 for padding sites i >= n)
__device__ __forceinline__ int isco_gain2(const IscoArgs& a, int64_t b, int i) {
  if (i >= a.n) return 0;
  const int c = a.weighted ? reinterpret_cast<const int32_t*>(a.cross)[b * a.cross_ld + i]
                           : (int)reinterpret_cast<const uint16_t*>(a.cross)[b * a.cross_ld + i];
  return __ldg(a.deg + i) - 2 * c;
}
// log(1 - exp(-|v|)) with the reference's switch (util.py:7-10)
__device__ __forceinline__ float log1mexp_ref(float v) {
  const float t = -fabsf(v);
  return t > -0.693f ? logf(-expm1f(t)) : log1pf(-expf(t));
}
// orderable 32-bit image of a float (ascending)
__device__ __forceinline__ uint32_t fkey(float f) {
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// log-softmax of score_i = (gain2_i / T) / 2 over the ld sites of chain b into sLp; returns the energy of the state
__device__ __forceinline__ float local_dist(const IscoArgs& a, int64_t b, float T, float* sLp, float* redf, long long* redl) {
  float mx = -FLT_MAX;
  long long quad = 0;
  for (int i = threadIdx.x; i < a.ld; i += kIscoThreads) {
    const int g2 = isco_gain2(a, b, i);
    const float s = __fmul_rn(__fdiv_rn((float)g2, T), 0.5f);
    sLp[i] = s;
    mx = fmaxf(mx, s);
    quad += g2;
  }
  mx = block_reduce_max(mx, redf);
  float sum = 0.f;
  for (int i = threadIdx.x; i < a.ld; i += kIscoThreads) sum += expf(sLp[i] - mx);
  sum = block_reduce_sum(sum, redf);
  const float lse = logf(sum);
  for (int i = threadIdx.x; i < a.ld; i += kIscoThreads) sLp[i] = (sLp[i] - mx) - lse;
  float energy;
  if (a.pisco) {      // -1/4 * fp16(s^T A s), s^T A s = sum_i gain2_i, in fp16 as the reference's matmul leaves it
    quad = block_reduce_sum_ll(quad, redl);
    const __half q = __float2half_rn((float)quad);
    const __half e16 = __hmul(__float2half_rn(-0.25f), q);
    energy = __fdiv_rn(__half2float(e16), T);
  } else {
    energy = __fdiv_rn((float)a.cut[b], T);
  }
  __syncthreads();
  return energy;
}

__global__ void __launch_bounds__(kIscoThreads) isco_propose_kernel(IscoArgs a) {
  extern __shared__ unsigned long long sKey[];          // [npow2] (value key << 32 | site), then float sLp[ld]
  float* sLp = reinterpret_cast<float*>(sKey + a.npow2);
  __shared__ float redf[kIscoThreads / 32];
  __shared__ long long redl[kIscoThreads / 32];
  const int64_t b = blockIdx.x;
  const float T = __ldg(a.temperature);
  const float energy = local_dist(a, b, T, sLp, redf, redl);
  // perturbed_ll = log_prob - log(-log(u))  (util.py:3-5); keys sort DESCENDING: the largest perturbed value first
  for (int i = threadIdx.x; i < a.npow2; i += kIscoThreads) {
    unsigned long long key = 0ull;                       // padding keys sort last
    if (i < a.ld) {
      const float u = __ldg(a.u + b * a.ld + i);
      const float pert = sLp[i] - logf(-logf(u));
      key = ((unsigned long long)fkey(pert) << 32) | (unsigned)(0xFFFFFFFFu - (unsigned)i);
    }
    sKey[i] = key;
  }
  __syncthreads();
  for (int size = 2; size <= a.npow2; size <<= 1) {      // bitonic sort, descending
    for (int stride = size >> 1; stride >= 1; stride >>= 1) {
      for (int t = threadIdx.x; t < a.npow2 / 2; t += kIscoThreads) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool desc = (lo & size) == 0;
        const unsigned long long x0 = sKey[lo], x1 = sKey[hi];
        if ((x0 < x1) == desc) sKey[lo] = x1, sKey[hi] = x0;
      }
      __syncthreads();
    }
  }
  // the k = path_length[b] largest, in order; ties of the k-th value are all taken by the reference's `>= threshold`
  // (probability zero for continuous noise) -- here exactly k
  int k = (int)min((int64_t)a.ld, max((int64_t)0, a.path_length[b]));
  k = min(k, a.kmax);
  if (threadIdx.x == 0) {
    // ll of drawing the chosen sites in this order without replacement (util.py:12-17): the reference's base is the
    // max of log_prob over ALL sites in sorted order = max over all sites
    float base = -FLT_MAX;
    for (int i = 0; i < a.ld; ++i) base = fmaxf(base, sLp[i]);
    float csum = 0.f, total = 0.f;
    for (int r = 0; r < k; ++r) {
      const int site = (int)(0xFFFFFFFFu - (unsigned)(sKey[r] & 0xFFFFFFFFull));
      const float ll = sLp[site];
      const float prob = expf(ll - base);
      const float incl = csum + prob;                    // cumsum, then `- prob_idx`
      const float taken = logf(incl - prob) + base;
      total += fminf(ll - log1mexp_ref(taken), 0.f);
      csum = incl;
      a.sel[b * a.kmax + r] = site;
    }
    for (int r = k; r < a.kmax; ++r) a.sel[b * a.kmax + r] = -1;
    a.ll_x[b] = energy;
    a.ll_x2y[b] = total;
  }
  __syncthreads();
  // y = x with the chosen sites flipped
  for (int i = threadIdx.x; i < a.ld; i += kIscoThreads) {
    const float v = isco_x(a, b, i);
    if (a.half) reinterpret_cast<__half*>(a.y)[b * a.ld + i] = __float2half_rn(v);
    else reinterpret_cast<float*>(a.y)[b * a.ld + i] = v;
  }
  __syncthreads();
  for (int r = threadIdx.x; r < k; r += kIscoThreads) {
    const int site = (int)(0xFFFFFFFFu - (unsigned)(sKey[r] & 0xFFFFFFFFull));
    const float v = 1.f - isco_x(a, b, site);
    if (a.half) reinterpret_cast<__half*>(a.y)[b * a.ld + site] = __float2half_rn(v);
    else reinterpret_cast<float*>(a.y)[b * a.ld + site] = v;
  }
}

// a.x = the CURRENT state x, a.y = the proposal (in) / the next state (out); cross / cut are those of the proposal
__global__ void __launch_bounds__(kIscoThreads) isco_accept_kernel(IscoArgs a) {
  extern __shared__ unsigned long long sKey[];
  float* sLp = reinterpret_cast<float*>(sKey);
  __shared__ float redf[kIscoThreads / 32];
  __shared__ long long redl[kIscoThreads / 32];
  __shared__ int sAccept;
  const int64_t b = blockIdx.x;
  const float T = __ldg(a.temperature);
  const float ll_y = local_dist(a, b, T, sLp, redf, redl);
  if (threadIdx.x == 0) {
    int k = 0;
    while (k < a.kmax && a.sel[b * a.kmax + k] >= 0) ++k;
    // reverse order: ascending perturbed value = the forward order backwards; unchosen sites carry -1e18 in the
    // reference (exp -> 0: they add nothing to the cumulative sum and come first)
    float base = -FLT_MAX;
    for (int r = 0; r < k; ++r) base = fmaxf(base, sLp[a.sel[b * a.kmax + r]]);
    float csum = 0.f, ll_y2x = 0.f;
    for (int r = k - 1; r >= 0; --r) {
      const float ll = sLp[a.sel[b * a.kmax + r]];
      const float prob = expf(ll - base);
      const float incl = csum + prob;
      const float taken = logf(incl - prob) + base;
      ll_y2x += fminf(ll - log1mexp_ref(taken), 0.f);
      csum = incl;
    }
    const float log_acc = fminf(((ll_y + ll_y2x) - a.ll_x[b]) - a.ll_x2y[b], 0.f);
    const float noise = __ldg(a.u + b);
    sAccept = logf(noise + 1e-24f) < log_acc ? 1 : 0;
    a.ll_y_t[b] = ll_y * T;
    a.acc[b] = expf(log_acc);
  }
  __syncthreads();
  if (!sAccept) {       // keep x: overwrite the proposal
    for (int i = threadIdx.x; i < a.ld; i += kIscoThreads) {
      if (a.half) reinterpret_cast<__half*>(a.y)[b * a.ld + i] = reinterpret_cast<const __half*>(a.x)[b * a.ld + i];
      else reinterpret_cast<float*>(a.y)[b * a.ld + i] = reinterpret_cast<const float*>(a.x)[b * a.ld + i];
    }
  }
}

static int isco_check(const IscoArgs& a, const char* what) {
  RLSB_REQUIRE(a.num_chains >= 0 && a.n > 0 && a.ld >= a.n && a.ld <= 16384 && a.kmax >= 1 && a.cross_ld >= a.n, RLSB_ERR_INVALID,
               "%s: bad shape (n <= ld <= 16384 sites per chain)", what);
  RLSB_REQUIRE(a.x && a.y && a.cross && a.deg && a.temperature && a.u && a.sel && a.ll_x && a.ll_x2y, RLSB_ERR_INVALID,
               "%s: null pointer", what);
  return RLSB_OK;
}

}  // namespace rlsb

extern "C" {

int rlsb_isco_propose(const void* x, void* y, int32_t is_half, const void* cross, int32_t weighted, int32_t cross_ld,
                      const int32_t* deg, const int64_t* cut, int32_t pisco, const float* temperature,
                      const int64_t* path_length, const float* u, int32_t* sel, int32_t kmax, float* ll_x, float* ll_x2y,
                      int32_t num_nodes, int32_t ld, int64_t num_chains, void* stream) {
  using namespace rlsb;
  IscoArgs a{};
  a.x = x, a.y = y, a.half = is_half, a.cross = cross, a.weighted = weighted, a.cross_ld = cross_ld, a.deg = deg, a.cut = cut;
  a.pisco = pisco, a.temperature = temperature, a.path_length = path_length, a.u = u, a.sel = sel, a.kmax = kmax;
  a.ll_x = ll_x, a.ll_x2y = ll_x2y, a.n = num_nodes, a.ld = ld, a.num_chains = num_chains;
  if (int rc = isco_check(a, "isco_propose")) return rc;
  RLSB_REQUIRE(path_length && (pisco || cut), RLSB_ERR_INVALID, "isco_propose: null pointer");
  if (num_chains == 0) return RLSB_OK;
  a.npow2 = 2;
  while (a.npow2 < ld) a.npow2 <<= 1;
  const size_t smem = (size_t)a.npow2 * 8 + (size_t)ld * 4;
  if (smem > 48 * 1024)
    RLSB_CUDA_OK(cudaFuncSetAttribute(isco_propose_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  isco_propose_kernel<<<(unsigned)num_chains, kIscoThreads, smem, static_cast<cudaStream_t>(stream)>>>(a);
  RLSB_LAUNCH_OK();
  return RLSB_OK;
}

int rlsb_isco_accept(const void* x, void* y, int32_t is_half, const void* cross_y, int32_t weighted, int32_t cross_ld,
                     const int32_t* deg, const int64_t* cut_y, int32_t pisco, const float* temperature, const float* u,
                     const int32_t* sel, int32_t kmax, const float* ll_x, const float* ll_x2y, float* ll_y_times_t,
                     float* acc, int32_t num_nodes, int32_t ld, int64_t num_chains, void* stream) {
  using namespace rlsb;
  IscoArgs a{};
  a.x = x, a.y = y, a.half = is_half, a.cross = cross_y, a.weighted = weighted, a.cross_ld = cross_ld, a.deg = deg, a.cut = cut_y;
  a.pisco = pisco, a.temperature = temperature, a.u = u, a.sel = const_cast<int32_t*>(sel), a.kmax = kmax;
  a.ll_x = const_cast<float*>(ll_x), a.ll_x2y = const_cast<float*>(ll_x2y), a.ll_y_t = ll_y_times_t, a.acc = acc;
  a.n = num_nodes, a.ld = ld, a.num_chains = num_chains;
  if (int rc = isco_check(a, "isco_accept")) return rc;
  RLSB_REQUIRE(ll_y_times_t && acc && (pisco || cut_y), RLSB_ERR_INVALID, "isco_accept: null pointer");
  if (num_chains == 0) return RLSB_OK;
  const size_t smem = (size_t)ld * 4 + 16;
  if (smem > 48 * 1024)
    RLSB_CUDA_OK(cudaFuncSetAttribute(isco_accept_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  isco_accept_kernel<<<(unsigned)num_chains, kIscoThreads, smem, static_cast<cudaStream_t>(stream)>>>(a);
  RLSB_LAUNCH_OK();
  return RLSB_OK;
}

}  // extern "C"
