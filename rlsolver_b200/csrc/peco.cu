// Pattern-I environment with one graph PER environment (distribution-wise training):
// SpinSystemUnbiased of rlsolver/methods/ECO_S2V/src/envs/spinsystem_PECO.py -- reset (151-193),
// step (306-486), calculate_cut (601-607), _get_immeditate_cuts_avaialable (660-662) -- and the
// visited-state test of HistoryBuffer (util_envs_PECO.py:228-288).
//
// The reference recomputes every local field with a batched matmul (E*N^2 flops) and clones the
// [E, 7, N] state each step.  Here the local fields (A s) stay resident: flipping node a changes
// (A s)_j by -2 A[a][j] s_a, one row of the env's matrix (O(N), or O(degree) of information), and one
// warp per environment rewrites the observables in place.  Every float32 value is produced with
// the same single operations as the reference's torch code (exact for integer-valued weights).
#include "common.cuh"

namespace rlsb {

// (A s)_j, fields_j = s_j (A s)_j and the cut for a given spin configuration
__global__ void __launch_bounds__(256) peco_fields_kernel(const float* __restrict__ matrix, const float* __restrict__ spins,
                                                          int64_t num_envs, int n, float* __restrict__ as,
                                                          float* __restrict__ fields, float* __restrict__ cut) {
  const int lane = threadIdx.x & 31;
  const int64_t env = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (env >= num_envs) return;
  const float* a = matrix + env * (int64_t)n * n;
  const float* s = spins + env * (int64_t)n;
  float sas = 0.f, suma = 0.f;
  for (int j = 0; j < n; ++j) {
    float acc = 0.f, rs = 0.f;
    for (int k = lane; k < n; k += 32) {
      const float v = __ldg(a + (int64_t)j * n + k);
      acc += v * s[k];
      rs += v;
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
      acc += __shfl_xor_sync(kFull, acc, off);
      rs += __shfl_xor_sync(kFull, rs, off);
    }
    if (lane == 0) {
      if (as) as[env * (int64_t)n + j] = acc;
      if (fields) fields[env * (int64_t)n + j] = acc * s[j];
    }
    sas += acc * s[j];
    suma += rs;
  }
  // calculate_cut: (1/4) * sum(-s * A s) + (1/4) * sum(A)
  if (lane == 0 && cut) cut[env] = __fadd_rn(__fmul_rn(0.25f, -sas), __fmul_rn(0.25f, suma));
}

struct PecoStep {
  const float* matrix;      // [E][N][N] symmetric
  float* state;             // [E][num_obs][N]
  float* as;                // [E][N] resident local fields (A s)
  const int64_t* action;    // [E]
  float *score, *best_score, *best_spins;   // [E], [E], [E][N]
  const float* max_local;   // [E]
  float* reward;            // [E]
  uint32_t* history;        // [cap][E][W] packed visited states (nullable)
  int32_t* bad_actions;
  int64_t num_envs;
  int n, num_obs, words, hist_len;
  int idx_imm, idx_tsf, idx_ept, idx_term, idx_greedy, idx_dscore, idx_dstate;   // observable rows, -1 = absent
  int reward_signal;        // 1 DENSE, 2 BLS, 4 CUSTOM_BLS
  int norm_rewards, use_stag, use_basin;
  float inv_steps, termination, stag, basin;
  // x / python_scalar: torch CUDA multiplies by the float32 reciprocal (div_true_kernel_cuda), torch CPU divides
  int recip_div;
  float inv_n;
};

__global__ void __launch_bounds__(256) peco_step_kernel(PecoStep p) {
  __shared__ uint32_t sMine[8][32];                    // packed spins of the warp's env (up to 1024 spins)
  uint32_t* mine = sMine[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  const int64_t env = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (env >= p.num_envs) return;
  const int n = p.n;
  float* st = p.state + env * (int64_t)p.num_obs * n;
  float* spins = st;                                   // observable 0 is always the spin state
  float* as = p.as + env * (int64_t)n;
  const int64_t a = p.action[env];
  if (a < 0 || a >= n) {                               // IndexError in the reference
    if (lane == 0) p.reward[env] = 0.f, atomicAdd(p.bad_actions, 1);
    return;
  }
  const float s_old = spins[a];
  __syncwarp();
  if (lane == 0) spins[a] = -s_old;
  __syncwarp();
  // (A s)_j -= 2 A[a][j] s_old;  fields_j = s_j (A s)_j   (the matrix is symmetric: row a == column a)
  const float* arow = p.matrix + (env * (int64_t)n + a) * n;
  const float maxl = p.max_local[env];
  float* best_spins = p.best_spins + env * (int64_t)n;
  int nonpos = 0;
  float delta = 0.f;
  for (int j = lane; j < n; j += 32) {
    const float v = as[j] - 2.f * __ldg(arow + j) * s_old;     // exact for integer-valued weights
    as[j] = v;
    const float f = v * spins[j];
    nonpos += (int)(f <= 0.f);
    if (j == (int)a) delta = -f;
    if (p.idx_imm >= 0) st[p.idx_imm * n + j] = __fdiv_rn(f, maxl);
    if (p.idx_tsf >= 0) st[p.idx_tsf * n + j] = (j == (int)a) ? 0.f : __fadd_rn(st[p.idx_tsf * n + j], p.inv_steps);
  }
  nonpos = warp_sum(nonpos);
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) delta += __shfl_xor_sync(kFull, delta, off);
  const float score = __fadd_rn(p.score[env], delta);
  const float best_obs = p.best_score[env];            // infinite memory: best observable == best ever
  // reward (always w.r.t. the best observable score BEFORE this step)
  const float improvement = __fsub_rn(score, best_obs);
  float rew = 0.f;
  if (p.reward_signal == 2) rew = improvement > 0.f ? improvement : 0.f;
  else if (p.reward_signal == 4) rew = improvement > 0.f ? __fdiv_rn(improvement, __fadd_rn(improvement, 0.1f)) : 0.f;
  else if (p.reward_signal == 1) rew = delta;
  if (p.norm_rewards) rew = p.recip_div ? __fmul_rn(rew, p.inv_n) : __fdiv_rn(rew, (float)n);
  // visited-state test against every state seen since reset (HistoryBuffer.update)
  if (p.history) {
    bool fresh = true;
    for (int w = 0; w < p.words; ++w) {
      const int j = w * 32 + lane;
      const uint32_t bits = __ballot_sync(kFull, j < n && spins[j] > 0.f);
      if (lane == 0) mine[w] = bits;
    }
    __syncwarp();
    const int64_t slab = p.num_envs * (int64_t)p.words;
    for (int t0 = 0; t0 < p.hist_len; t0 += 32) {
      const int t = t0 + lane;
      bool same = t < p.hist_len;
      if (same) {
        const uint32_t* h = p.history + t * slab + env * (int64_t)p.words;
        for (int w = 0; w < p.words; ++w) same = same && (h[w] == mine[w]);
      }
      if (__any_sync(kFull, same)) fresh = false;
    }
    if (lane == 0) {
      uint32_t* h = p.history + p.hist_len * slab + env * (int64_t)p.words;
      for (int w = 0; w < p.words; ++w) h[w] = mine[w];
    }
    if (p.use_stag && !fresh) rew = __fsub_rn(rew, p.stag);
    if (p.use_basin && nonpos == n && fresh) rew = __fadd_rn(rew, p.basin);
  }
  // best score / spins
  const bool better = score > best_obs;
  const float new_best = better ? score : best_obs;
  int dist = 0;
  for (int j = lane; j < n; j += 32) {
    const float sj = spins[j];
    if (better) best_spins[j] = sj;
    else dist += (int)(best_spins[j] != sj);
  }
  dist = warp_sum(dist);
  // global observables (one value broadcast over the row)
  const float g_greedy = __fsub_rn(1.f, p.recip_div ? __fmul_rn((float)nonpos, p.inv_n) : __fdiv_rn((float)nonpos, (float)n));
  const float g_dscore = __fdiv_rn(fabsf(__fsub_rn(score, new_best)), maxl);
  const float g_dstate = (float)dist;
  for (int j = lane; j < n; j += 32) {
    if (p.idx_ept >= 0) st[p.idx_ept * n + j] = __fadd_rn(st[p.idx_ept * n + j], p.inv_steps);
    if (p.idx_term >= 0) st[p.idx_term * n + j] = p.termination;
    if (p.idx_greedy >= 0) st[p.idx_greedy * n + j] = g_greedy;
    if (p.idx_dscore >= 0) st[p.idx_dscore * n + j] = g_dscore;
    if (p.idx_dstate >= 0) st[p.idx_dstate * n + j] = g_dstate;
  }
  if (lane == 0) {
    p.score[env] = score;
    p.best_score[env] = new_best;
    p.reward[env] = rew;
  }
}

}  // namespace rlsb

extern "C" {

int rlsb_peco_fields(const float* matrix, const float* spins, int64_t num_envs, int32_t num_spins, float* as,
                     float* fields, float* cut, void* stream) {
  using namespace rlsb;
  RLSB_REQUIRE(num_envs >= 0 && num_spins > 0, RLSB_ERR_INVALID, "peco_fields: bad shape");
  if (num_envs == 0) return RLSB_OK;
  RLSB_REQUIRE(matrix && spins, RLSB_ERR_INVALID, "peco_fields: null pointer");
  peco_fields_kernel<<<(unsigned)((num_envs + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      matrix, spins, num_envs, num_spins, as, fields, cut);
  RLSB_LAUNCH_OK();
  return RLSB_OK;
}

int rlsb_peco_step(const float* matrix, float* state, float* as, const int64_t* action, float* score,
                   float* best_score, float* best_spins, const float* max_local, float* reward, uint32_t* history,
                   int32_t hist_len, int32_t* bad_actions, int64_t num_envs, int32_t num_spins, int32_t num_obs,
                   const int32_t* h_obs_rows, int32_t reward_signal, int32_t norm_rewards, float inv_steps,
                   float termination, int32_t use_stag, float stag, int32_t use_basin, float basin, int32_t scalar_div_as_cuda,
                   void* stream) {
  using namespace rlsb;
  RLSB_REQUIRE(num_envs >= 0 && num_spins > 0 && num_spins <= 1024 && num_obs >= 1, RLSB_ERR_INVALID,
               "peco_step: bad shape (1 <= n_spins <= 1024)");
  if (num_envs == 0) return RLSB_OK;
  RLSB_REQUIRE(matrix && state && as && action && score && best_score && best_spins && max_local && reward &&
                   bad_actions && h_obs_rows,
               RLSB_ERR_INVALID, "peco_step: null pointer");
  RLSB_REQUIRE(reward_signal == 1 || reward_signal == 2 || reward_signal == 4, RLSB_ERR_UNSUPPORTED,
               "peco_step: reward signal %d (DENSE = 1, BLS = 2, CUSTOM_BLS = 4 are implemented)", reward_signal);
  RLSB_REQUIRE(!(use_stag || use_basin) || history, RLSB_ERR_INVALID, "peco_step: stag/basin rewards need the history buffer");
  PecoStep p{};
  p.matrix = matrix, p.state = state, p.as = as, p.action = action, p.score = score, p.best_score = best_score;
  p.best_spins = best_spins, p.max_local = max_local, p.reward = reward, p.history = history, p.bad_actions = bad_actions;
  p.num_envs = num_envs, p.n = num_spins, p.num_obs = num_obs, p.words = (num_spins + 31) / 32, p.hist_len = hist_len;
  for (int k = 0; k < 7; ++k)
    RLSB_REQUIRE(h_obs_rows[k] < num_obs, RLSB_ERR_INVALID, "peco_step: observable row %d out of range", h_obs_rows[k]);
  p.idx_imm = h_obs_rows[0], p.idx_tsf = h_obs_rows[1], p.idx_ept = h_obs_rows[2], p.idx_term = h_obs_rows[3];
  p.idx_greedy = h_obs_rows[4], p.idx_dscore = h_obs_rows[5], p.idx_dstate = h_obs_rows[6];
  p.reward_signal = reward_signal, p.norm_rewards = norm_rewards, p.use_stag = use_stag, p.use_basin = use_basin;
  p.inv_steps = inv_steps, p.termination = termination, p.stag = stag, p.basin = basin;
  p.recip_div = scalar_div_as_cuda, p.inv_n = 1.0f / (float)num_spins;
  peco_step_kernel<<<(unsigned)((num_envs + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  RLSB_LAUNCH_OK();
  return RLSB_OK;
}

}  // extern "C"
