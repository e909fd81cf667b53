/*
 * rlsolver_b200 -- C ABI of the B200-native max-cut / QUBO environment hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types.  The
 * reference (Open-Finance-Lab/RLSolver) is 100% Python with no FFI of its own;
 * each entry point below names the reference function (file:line, relative to
 * the reference checkout) whose work it replaces.  INTEGRATION.md shows the
 * ctypes binding a reference maintainer would add.
 *
 * Conventions
 *   - every function returns 0 (RLSB_OK) or a positive error code; the message
 *     for the calling thread's last error is rlsb_last_error().
 *   - pointers are DEVICE pointers unless the parameter name starts with `h_`.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *   - spins ("xs") at the API surface are the reference's layout: row-major
 *     bool/uint8 [E][N], one byte per (environment, node).
 *   - packed spins are uint32 [W][Np], W = ceil(E/32) env tiles, Np =
 *     rlsb_graph_padded_nodes(): word [t][i] holds node i of envs 32t..32t+31
 *     (bit b <-> env 32t+b).  Padding bits/words are zero on pack.
 *   - objective values ("vs") are int64 [E] as in the reference (th.long).
 *   - nothing here allocates device memory except rlsb_graph_create.
 */
#ifndef RLSOLVER_B200_H
#define RLSOLVER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RLSB_OK 0
#define RLSB_ERR_INVALID 1     /* bad argument (reference would raise IndexError/ValueError/assert) */
#define RLSB_ERR_CUDA 2        /* CUDA runtime error, see rlsb_last_error() */
#define RLSB_ERR_UNSUPPORTED 3 /* shape outside what the kernels are built for */
#define RLSB_ERR_NODEVICE 4    /* graph was built host-only (device < 0) */

typedef struct rlsb_graph rlsb_graph_t;
typedef struct rlsb_mcpg_plan rlsb_mcpg_plan_t;
typedef struct rlsb_qubo rlsb_qubo_t;

int rlsb_version(void);
const char* rlsb_last_error(void);

/* Diagnostic switches (cross-check paths of the tests and the profiling tools; none changes a result).
 * The word is initialised ONCE, when the library is loaded, from the environment variables named below;
 * afterwards only this call changes it: new = (old & ~clear_mask) | set_mask.  Returns the new word.
 * Nothing on the data path reads the environment. */
#define RLSB_DEBUG_PLAIN_MASKS 1 /* RLSB_LS_PLAIN_MASKS=1: rlsb_ls_noise_masks evaluates every normal (no early-out bytes) */
#define RLSB_DEBUG_FULL_CUT 2    /* RLSB_LS_FULL_CUT=1: rlsb_ls_run_masks re-counts all edges for every candidate */
#define RLSB_DEBUG_LS_SKIP 4     /* RLSB_LS_SKIP=1: rlsb_ls_run consumers skip the arithmetic (streaming-rate probe) */
#define RLSB_DEBUG_LS_TIMES 8    /* RLSB_LS_TIMES=1: CTA 0 of the local-search kernels stamps clock64 per phase */
#define RLSB_DEBUG_CARVEOUT_DEFAULT 16 /* RLSB_CARVEOUT_DEFAULT=1: the fused search leaves the shared-memory carve-out of its two kernels to the driver (they then do not share SMs) */
#define RLSB_DEBUG_GEN_PER_DRAW 32 /* RLSB_GEN_PER_DRAW=1: rlsb_ls_noise_masks uses the one-draw-per-thread generator (round-1 form) */
#define RLSB_DEBUG_PECO_WARP_PER_ENV 64 /* RLSB_PECO_WARP_PER_ENV=1: rlsb_peco_compact_step uses one warp per env also for N <= 128 */
#define RLSB_DEBUG_QUBO_NO_SPLITK 128 /* RLSB_QUBO_NO_SPLITK=1: rlsb_qubo_sweeps never splits the K range over several CTAs per chain group */
#define RLSB_DEBUG_THRESH_PIPE 256 /* RLSB_LS_THRESH_PIPE=1: a threshold-only rlsb_ls_run keeps the pipelined tile kernel (round-2a form) instead of the warp-per-env kernel */
int32_t rlsb_debug_flags(int32_t set_mask, int32_t clear_mask);

/* ---- graph store: replaces EnvMaxcut.__init__ (rlsolver/envs/env_L2A.py:25-52),
 * build_adjacency_indies (rlsolver/methods/util_read_data.py:144-187) and
 * calc_num_nodes_in_mygraph (rlsolver/methods/util.py:35-40).
 * h_n0/h_n1 are HOST arrays of 0-based endpoints in file order; h_w (nullable)
 * integer weights.  num_nodes <= 0 -> number of distinct endpoints (reference
 * quirk); ids must then be < that number.  device < 0 builds host-side only. */
int rlsb_graph_create(int32_t num_nodes, int64_t num_edges, const int32_t* h_n0, const int32_t* h_n1,
                      const int32_t* h_w, int32_t bidirectional, int32_t device, rlsb_graph_t** out);
int rlsb_graph_destroy(rlsb_graph_t* g);
int32_t rlsb_graph_num_nodes(const rlsb_graph_t* g);
int32_t rlsb_graph_padded_nodes(const rlsb_graph_t* g);
int64_t rlsb_graph_num_edges(const rlsb_graph_t* g);      /* len(mygraph) */
int64_t rlsb_graph_num_listed(const rlsb_graph_t* g);     /* Md: M or 2M */
int64_t rlsb_graph_num_full(const rlsb_graph_t* g);       /* undirected neighbour slots, self loops dropped */
int32_t rlsb_graph_num_levels(const rlsb_graph_t* g);     /* Gauss-Seidel dependency levels of the sweep */
int32_t rlsb_graph_max_listed_degree(const rlsb_graph_t* g);
int32_t rlsb_graph_max_full_degree(const rlsb_graph_t* g);
/* copy host-side arrays out (each pointer nullable).  listed_*: sorted listed-neighbour
 * CSR (== n0_ids/n1_ids, n0_num_n1 of the reference); full_*: undirected CSR;
 * level_*: nodes grouped by dependency level. */
int rlsb_graph_export(const rlsb_graph_t* g, int32_t* h_listed_ptr, int32_t* h_listed_col, int32_t* h_full_ptr,
                      int32_t* h_full_col, int32_t* h_level_ptr, int32_t* h_level_nodes);

/* The tile kernels read neighbour lists as SELL-32 slices (32 node slots per slice, column ids
 * round-major, short rows padded with the node's own id, unused slots = 0xFFFF).  which = 0:
 * listed neighbours, slot == node over the padded node range; which = 1: full neighbours in
 * sweep order (dependency level, then degree-descending), `half` = floor(degree/2), and
 * level_slice[l] .. level_slice[l+1] = the slices of level l.  sizes: out3 = {num_slices,
 * column entries, levels + 1}.  Host-side copies for tests and tools. */
int rlsb_graph_sell_sizes(const rlsb_graph_t* g, int32_t which, int64_t* out3);
int rlsb_graph_sell_export(const rlsb_graph_t* g, int32_t which, int32_t* h_off, uint16_t* h_node, uint16_t* h_half,
                           uint16_t* h_col, int32_t* h_level_slice);

/* ---- spin (de)packing: layout change only, no reference counterpart */
int rlsb_pack_spins(const uint8_t* xs, int64_t num_envs, int32_t num_nodes, int32_t padded_nodes, uint32_t* packed,
                    void* stream);
int rlsb_unpack_spins(const uint32_t* packed, int64_t num_envs, int32_t num_nodes, int32_t padded_nodes, uint8_t* xs,
                      void* stream);

/* ---- objective: replaces EnvMaxcut.calculate_obj_values (env_L2A.py:54-66)
 * cut_eval: bool [E][N] in, int64 [E] out (if_sum=True; '//2' of the bidirectional
 * listing is folded in).  cut_eval_packed: same on packed spins.
 * cut_edges: if_sum=False, uint8 [E][Md] per-listed-edge indicators. */
int rlsb_cut_eval(const rlsb_graph_t* g, const uint8_t* xs, int64_t num_envs, int64_t* vs, void* stream);
int rlsb_cut_eval_packed(const rlsb_graph_t* g, const uint32_t* packed, int64_t num_envs, int64_t* vs, void* stream);
int rlsb_cut_edges(const rlsb_graph_t* g, const uint8_t* xs, int64_t num_envs, uint8_t* out, void* stream);

/* ---- integer-weighted objective (row W of the scope table).  The reference's EnvMaxcut ignores edge weights
 * (env_L2A.py:54-66 counts XORs); its callers that do not are PISCO's adjacency energy -1/4 s^T A s
 * (rlsolver/envs/env_ISCO.py:436-444) and MCPG's weighted sampler (rlsolver/methods/MCPG/sampling.py:89-127).
 * A graph created with weights other than 1 additionally holds its edges bucketed by (bit of |w|, sign):
 * the weighted cut is sum_k scale_k * popcount-per-env(bucket k), +-1 weights (Gset) being two buckets.
 *   cut_eval_weighted   : vs[e] = sum over edges of w * [x_u != x_v]; xs bool [E][N] OR packed tiles (the other NULL).
 *   node_fields_weighted: out[e][i] = sum_j w_ij [x_i != x_j] over the undirected neighbourhood (int32 [E][Np]);
 *                         the gain of flipping node i is wdeg_i - 2 out[e][i].
 * |w| < 2^24; self loops never count.  RLSB_ERR_INVALID on a graph whose weights are all 1. */
int rlsb_graph_is_weighted(const rlsb_graph_t* g);
int rlsb_cut_eval_weighted(const rlsb_graph_t* g, const uint8_t* xs, const uint32_t* packed, int64_t num_envs,
                           int64_t* vs, void* stream);
int rlsb_node_fields_weighted(const rlsb_graph_t* g, const uint32_t* packed, int64_t num_envs, int32_t* out,
                              void* stream);

/* ---- per-node cross counts: integer core of calculate_obj_values_for_loop
 * (env_L2A.py:68-80).  cross is uint16 [E][Np] (row-major, rows padded to Np):
 * number of LISTED neighbours of node i on the other side in env e.
 * col_min/col_max (nullable, int32 [N]) receive min/max over envs per node --
 * the cross-env coupling `ws_std` of env_L2A.py:93 / LocalSearch.py:65.  Listed degree <= 4095. */
int rlsb_node_cross_counts(const rlsb_graph_t* g, const uint32_t* packed, int64_t num_envs, uint16_t* cross,
                           int32_t* col_min, int32_t* col_max, void* stream);

/* ---- local search: EnvMaxcut.local_search_inplace (env_L2A.py:87-116) and
 * LocalSearch.random_search (LocalSearch.py:53-86) in two calls on one stream (ls_begin, ls_run;
 * ls_thresh / ls_search are the two halves of ls_run for callers that want them apart).  State between
 * the calls lives in a caller-provided, 256-byte aligned device workspace of
 * rlsb_ls_workspace_bytes(g, E) bytes (returns -1 on a bad handle).
 *
 *   ws[e][i]  = listed_degree[i] - ws_mult * cross[e][i]   (ws_mult: 1 = local_search_inplace,
 *               2 = LocalSearch.random_search), rd_std[i] = float(ws_mult*(max_e cross - min_e cross)) * noise_std,
 *   spin_rand = ws + noise * rd_std in float32, one rounding per op (as the reference's torch kernels).
 *
 * ls_begin : bool xs [E][N] -> packed tiles, per-node cross counts, their max/min over the env
 *            batch (the `ws_std` coupling, env_L2A.py:92-93), rd_std; if compute_vs != 0 also
 *            vs[e] = cut (the `()` sentinel of env_L2A.py:91), else vs is left as given.
 * ls_thresh: thresh[e] = kth smallest of spin_rand[e][:], k = N - num_spin (torch.kthvalue,
 *            env_L2A.py:94-96) from one noise tensor float32 [E][N].
 * ls_search: for each of num_iters noise tensors (h_noise_ptrs: HOST array of device pointers,
 *            float32 [E][N]): flip where spin_rand > thresh, evaluate, keep rows with vs' >= vs
 *            (env_L2A.py:97-107).  If finish != 0 it then runs the exhaustive single-flip pass
 *            (env_L2A.py:110-115: node 0..N-1 in order, flip where the cut does not get worse,
 *            Gauss-Seidel; O(2M) word operations per 32 envs instead of N full evaluations; the
 *            gain rule is the batched form of S2V_PPO/env.py:197-206), writes the bool rows to
 *            xs_out [E][N] and the final cut values to vs.  vs must be consistent with the state
 *            (vs[e] == cut of row e) when ls_begin was called with compute_vs == 0.
 * ls_run   : ls_thresh on thresh_noise (skipped when NULL; num_spin is then ignored) followed by
 *            ls_search, fused into one launch per 16 noise tensors: a TMA-fed shared-memory ring
 *            streams noise and cross counts, the next iteration's noise lands while the current
 *            candidate is evaluated.  thresh_noise may alias h_noise_ptrs[0] (LocalSearch.py:68). */
int64_t rlsb_ls_workspace_bytes(const rlsb_graph_t* g, int64_t num_envs);
/* byte offset of a workspace section (tests / debugging): 0 packed u32 [W][Np], 1 cross counts
 * (uint8, or uint16 when a degree exceeds 255) tiled [W][Np/4][32 envs][4 nodes], 2 col_min i32 [Np], 3 col_max i32 [Np],
 * 4 listed degree + 0x4B400000 i32 [Np], 5 rd_std f32 [Np], 6 thresh f32 [E], 7 cross counts uint8 row-major
 * [32W][Np] (absent when a degree exceeds 255), 8 unit counters of the streaming generator (u32 pairs); -1 on error. */
int64_t rlsb_ls_workspace_offset(const rlsb_graph_t* g, int64_t num_envs, int32_t section);
int rlsb_ls_begin(const rlsb_graph_t* g, const uint8_t* xs, int64_t num_envs, int64_t* vs, int32_t compute_vs,
                  int32_t ws_mult, float noise_std, void* workspace, void* stream);
int rlsb_ls_thresh(const rlsb_graph_t* g, int64_t num_envs, int32_t ws_mult, const float* noise, int32_t num_spin,
                   void* workspace, void* stream);
int rlsb_ls_search(const rlsb_graph_t* g, int64_t num_envs, int64_t* vs, int32_t ws_mult,
                   const float* const* h_noise_ptrs, int32_t num_iters, int32_t finish, uint8_t* xs_out,
                   void* workspace, void* stream);
int rlsb_ls_run(const rlsb_graph_t* g, int64_t num_envs, int64_t* vs, int32_t ws_mult, const float* thresh_noise,
                int32_t num_spin, const float* const* h_noise_ptrs, int32_t num_iters, int32_t finish,
                uint8_t* xs_out, void* workspace, void* stream);

/* ---- the noisy iterations without noise tensors.  The reference draws `randn [E,N] float32` once per
 * iteration (env_L2A.py:99, LocalSearch.py:66) and uses one bit of each number (`spin_rand > thresh`,
 * env_L2A.py:100-101).  torch's CUDA generator is counter based (Philox4x32-10 keyed by seed and offset), so
 * ls_noise_masks recomputes exactly the numbers `num_draws` consecutive torch.randn([E,N]) calls starting
 * at generator state (seed, offset) would have produced (curand_normal4, as ATen's normal_ kernel), evaluates
 * spin_rand on them in registers against the thresholds in the workspace (ls_run / ls_thresh put them there)
 * and writes, per draw, the flip mask as a flat bit array: bit (e*N + n) of masks[k*mask_words ...], where
 * mask_words = rlsb_ls_mask_words(g, E) uint32 words per draw (-1: not available -- counts wider than
 * 8 bits or 2^31 and more elements per draw -- use rlsb_ls_run with explicit noise).  rng_threads /
 * rng_iters are the call geometry of ONE such torch call (ATen calc_execution_policy: T = 256 * min(SMs *
 * blocks per SM, ceil(numel / 256)), iters = ceil(numel / 4T)); the caller advances the generator offset by
 * 4 * rng_iters per draw.  reuse_bound != 0: the early-out bytes written by the previous call on this workspace
 * are still valid (same state, same thresholds: a later group of draws of the same search).  ls_run_masks then runs the iterations (one per mask array) and, if finish != 0,
 * the single-flip pass, as rlsb_ls_search does.  rlsb_torch_randn writes the same draws as float32
 * [num_draws][numel] (equal to torch.randn bit for bit; tests pin it). */
int64_t rlsb_ls_mask_words(const rlsb_graph_t* g, int64_t num_envs);
int rlsb_ls_noise_masks(const rlsb_graph_t* g, int64_t num_envs, int32_t ws_mult, uint64_t seed, uint64_t offset,
                        const uint64_t* rng_dev, int32_t rng_threads, int32_t rng_iters, int32_t num_draws,
                        int32_t reuse_bound, uint32_t* masks, void* workspace, void* stream);
int rlsb_ls_run_masks(const rlsb_graph_t* g, int64_t num_envs, int64_t* vs, const uint32_t* masks, int32_t num_iters,
                      int32_t finish, uint8_t* xs_out, void* workspace, void* stream);
/* Noisy iterations + single-flip pass with the generator running NEXT TO the tile kernel (round 2): same
 * arguments and results as rlsb_ls_noise_masks(num_draws = num_iters) followed by rlsb_ls_run_masks, issued as
 * early-out bytes + zeroing on `stream`, then the tile kernel on `stream` and the streaming generator on a side
 * stream of the graph handle (joined back into `stream` before the call returns; capturable in a CUDA graph).
 * The generator finishes the draws in groups of two; a tile CTA starts iteration k when the group of draw k is
 * complete (counters in the workspace).  masks: scratch, uint32 [num_iters][rlsb_ls_mask_words].  xs_out may be
 * NULL when only the packed tiles (workspace section 0) and vs are wanted; also for rlsb_ls_run_masks. */
int rlsb_ls_fused_search(const rlsb_graph_t* g, int64_t num_envs, int64_t* vs, int32_t ws_mult, uint64_t seed,
                         uint64_t offset, const uint64_t* rng_dev, int32_t rng_threads, int32_t rng_iters,
                         int32_t num_iters, int32_t finish, uint8_t* xs_out, uint32_t* masks, void* workspace,
                         void* stream);
/* Diagnostics of the last rlsb_ls_fused_search on this workspace (synchronises `stream`): h_out3 = {generator blocks
 * that started, tile CTAs whose bounded wait for a group of draws ran out, units finished of the first group}.
 * A non-zero second word means the two kernels were not scheduled together and the call's results are invalid. */
int rlsb_ls_fused_status(const rlsb_graph_t* g, int64_t num_envs, const void* workspace, uint32_t* h_out3, void* stream);
/* rlsb_ls_begin for a state handed over as packed tiles (uint32 [ceil(E/32)][Np], bit b of word [t][i] = node i
 * of env 32t + b; bits of envs >= E must be 0): what a host that keeps spins packed sends, 1 bit instead of
 * 1 byte per spin.  packed_in may be the workspace's own packed section (no copy then). */
int rlsb_ls_begin_packed(const rlsb_graph_t* g, const uint32_t* packed_in, int64_t num_envs, int64_t* vs,
                         int32_t compute_vs, int32_t ws_mult, float noise_std, void* workspace, void* stream);
int rlsb_torch_randn(float* out, int64_t numel, uint64_t seed, uint64_t offset, const uint64_t* rng_dev,
                     int32_t rng_threads, int32_t rng_iters, int32_t num_draws, void* stream);
/* Device-resident generator state for CUDA-graph replays: rng_dev -> {seed, offset} (2 x uint64 in device
 * memory).  When rng_dev is non-NULL the two calls above take the seed from it and count `offset` from its
 * offset (pass the offset of the draw relative to the cursor); rlsb_rng_cursor_advance moves the cursor on
 * the device, so every replay of a captured sequence continues the stream where the last one stopped. */
int rlsb_rng_cursor_advance(uint64_t* rng_dev, uint64_t delta, void* stream);

/* profiling aid: with RLSB_LS_TIMES=1 in the environment CTA 0 of every pipelined ls_run launch
 * records clock64 stamps (start, then per pass: candidate done / accepted, then finish done);
 * this copies the 64 slots out (host pointer).  RLSB_ERR_INVALID when the variable is unset. */
int rlsb_ls_debug_times(int64_t* out64);

/* ---- relaxed (probabilistic) max-cut objective: SimulatorMaxcut.get_objectives (rlsolver/envs/env_k_spin.py:191-193)
 * and PIGNN's hamiltonian_maxcut (rlsolver/methods/PIGNN/util.py:4-8):
 *   out[e] = -sum over the ORIGINAL edge list (u, v) of  p[e][u] + p[e][v] - 2 p[e][u] p[e][v]      (weights ignored, as there)
 * probs float32 [E][N], out float32 [E].  relaxed_cut_grad is its vector-Jacobian product: grad_probs[e][u] =
 * -grad_out[e] * sum over edges incident to u of (1 - 2 p[e][other end]) (what autograd computes through the
 * reference's gathers).  float32; results agree with the torch expression within 1e-5 relative. */
int rlsb_relaxed_cut(const rlsb_graph_t* g, const float* probs, int64_t num_envs, float* out, void* stream);
int rlsb_relaxed_cut_grad(const rlsb_graph_t* g, const float* probs, const float* grad_out, int64_t num_envs,
                          float* grad_probs, void* stream);

/* ---- the exhaustive single-flip pass alone, on packed tiles (vs is recomputed) */
int rlsb_flip_sweep(const rlsb_graph_t* g, uint32_t* packed, int64_t* vs, int64_t num_envs, void* stream);

/* ---- pattern-I step: env_PPO.EnvMaxcut.step (rlsolver/envs/env_PPO.py:92-106).
 * xs is the reference's observation tensor, float32 [E][N] with values {0,1}; action int64 [E].
 * Every env flips xs[e][action[e]]; reward[e] = new cut - old cut (the single-flip gain
 * deg - 2*cross over the acted node's neighbours, O(degree) instead of a full re-evaluation;
 * update rule of S2V_PPO/env.py:197-206); cut[e] (the env's `last_reward`) += reward[e].
 * An action outside [0, N) (IndexError in the reference) leaves the env untouched, gives
 * reward 0 and increments *bad_actions (device int32, caller-zeroed). */
int rlsb_step_flip(const rlsb_graph_t* g, float* xs, const int64_t* action, int64_t num_envs, float* reward,
                   float* cut, int32_t* bad_actions, void* stream);

/* ---- greedy best-single-flip ascent, contract of greedy_maxcut (rlsolver/methods/greedy.py:33-78)
 * batched over envs: per step all N single-flip gains, the LOWEST index among the best, accept
 * only if the cut gets strictly better (strict != 0; strict == 0 also takes zero-gain moves, the
 * stop rule of ECO_S2V/src/agents/solver.py:95-121), else that env stops; at most max_flips steps
 * (greedy.py caps them at N).  max_flips = 1, strict = 1 is one call of PECO's local_search
 * (ECO_S2V/util.py:66-75) without its spin-zeroing bug.  xs: bool [E][N] in/out, vs: int64 [E]
 * out (final cut), flips: int32 [E] out (nullable).  The per-(node, env) gains stay resident in
 * shared memory (int8 when every degree <= 127, else int16) and are updated in O(degree) per flip. */
int rlsb_greedy_best_flip(const rlsb_graph_t* g, uint8_t* xs, int64_t num_envs, int64_t* vs, int32_t* flips,
                          int32_t max_flips, int32_t strict, void* stream);

/* ---- samplers.  Random numbers: every sampler either reads EXPLICIT arrays (replay of recorded
 * draws) or, when the explicit pointers are null, computes torch's own Philox4x32-10 stream in
 * place from (seed, offset) of the torch CUDA generator: element li of the k-th consecutive torch
 * distribution call (torch.rand / rand_like / randint over `numel` elements) is
 *   Philox(counter = offset/4 + k*rng_iters + (li / rng_threads)/4, subsequence = li % rng_threads)[(li / rng_threads) % 4]
 * with rng_threads = 256 * min(#SM * (maxThreadsPerSM / 256), ceil(numel / 256)) and rng_iters =
 * ceil(numel / (4 * rng_threads)) (ATen/native/cuda/DistributionTemplates.h).  The caller advances
 * the generator offset by 4 * rng_iters per call the reference would have made.
 * rlsb_torch_rand / rlsb_torch_randint regenerate `calls` consecutive torch.rand(numel) /
 * torch.randint(0, range, [numel]) results (tests pin the stream identity with them). */
/* One round of metropolis_hastings_sampling_TNCO (rlsolver/envs/env_L2A.py:233-276, the row-major variant of
 * metro_sampling): xs bool [R*S][dim] in place, probs float32 [S][dim] (row r uses probs[r % S]), perm int64 [dim] =
 * the round's torch.randperm(dim).  Position i proposes a flip of column perm[i] in every row, accepted when
 * torch.rand(R*S) call number i (generator state (seed, offset), geometry of numel = R*S) is below (1 - q) / q;
 * state[0] (int64, in/out) = accepts so far, state[1] (out) = positions visited = index of the first position at
 * which the count reached `target`, plus one (dim when it never did).  counts: int32 [dim] scratch.  The caller
 * advances the generator by state[1] calls. */
int rlsb_mh_rows_round(const float* probs, uint8_t* xs, const int64_t* perm, int64_t num_rows, int64_t num_sims,
                       int32_t dim, int64_t target, uint64_t seed, uint64_t offset, uint32_t rng_threads, uint32_t rng_iters,
                       int64_t* state, int32_t* counts, void* stream);
int rlsb_torch_rand(uint64_t seed, uint64_t offset, uint32_t rng_threads, uint32_t rng_iters, int64_t calls,
                    int64_t numel, float* out, void* stream);
int rlsb_torch_randint(uint64_t seed, uint64_t offset, uint32_t rng_threads, uint32_t rng_iters, int64_t calls,
                       int64_t numel, uint32_t range, int64_t* out, void* stream);

/* sampler_func (rlsolver/methods/MCPG.py:120-166): num_ls Gauss-Seidel sweeps over the nodes in
 * `order` (data.sorted_degree_nodes), x_i <- [sum_{j in N(i)} x_j + rand/4 < (deg_i + 1/4)/2] with
 * one torch.rand(C) per node visit, then expected_cut[c] = sum_e (2x_u-1)(2x_v-1).
 * xs: float32 [N][C] node-major {0,1}, updated in place to the swept state ({0,1} floats; call with
 * num_ls >= 1); expected: float32 [C].  The plan holds the order-dependent structures (neighbours
 * split into visited-before / visited-after, dependency levels of the visiting order).
 * explicit_u: float32 [num_ls*N][C] (draw of node visit k of sweep s at row s*N + k) or null.
 * Graphs with self loops are rejected (the reference double-lists the node as its own neighbour). */
int rlsb_mcpg_plan_create(const rlsb_graph_t* g, const int32_t* h_order, rlsb_mcpg_plan_t** out);
int rlsb_mcpg_plan_destroy(rlsb_mcpg_plan_t* plan);
int32_t rlsb_mcpg_plan_num_levels(const rlsb_mcpg_plan_t* plan);
/* Note on rlsb_mcpg_sweeps: the plan owns device scratch for the tie-break words of a call (allocated on first use
 * and when a larger chain count arrives -- cudaMalloc, so not inside a CUDA-graph capture; freed by
 * rlsb_mcpg_plan_destroy).  Calls that share a plan must be ordered on one stream. */
int rlsb_mcpg_sweeps(const rlsb_graph_t* g, const rlsb_mcpg_plan_t* plan, float* xs, int64_t num_chains,
                     int32_t num_ls, const float* explicit_u, uint64_t seed, uint64_t offset, uint32_t rng_threads,
                     uint32_t rng_iters, float* expected, void* stream);

/* metro_sampling (rlsolver/methods/MCPG.py:88-117 == MCPG/sampling.py:67-86): per iteration every
 * chain picks a node r = randint(0, N), accepts a flip with probability (1-q)/q, q = p_r if the
 * bit is set else 1-p_r (rand < rate).  The reference stops once the accepted moves of ALL chains
 * reach C*max_transfer_time, checked before every iteration: call with count_only = 1 to get the
 * accepted count of each of max_iters iterations in acc (int32 [max_iters], caller-zeroed), derive
 * the number of iterations the reference executes, then call with count_only = 0 and that number
 * in *num_iters_dev (device int32) to produce out.  probs: float32 [N]; start / out: float32
 * [N][C] node-major; explicit_idx int64 / explicit_u float32: [max_iters][C] or both null. */
int rlsb_metro_sampling(int32_t num_nodes, const float* probs, const float* start, float* out, int64_t num_chains,
                        int32_t max_iters, const int32_t* num_iters_dev, const int64_t* explicit_idx,
                        const float* explicit_u, uint64_t seed, uint64_t offset, uint32_t rng_threads,
                        uint32_t rng_iters, int32_t* acc, int32_t count_only, void* stream);
/* The same sampler in split form (the default of the Python mirror): the draws of all iterations are computed in
 * parallel first, the chain runs once over all max_iters iterations logging its accepted moves, the stop rule
 * `accepted moves before iteration t >= stop_count` (MCPG.py:101-103, stop_count = C * max_transfer_time) gives the
 * number of executed iterations *num_iters_dev (device int32, output) and the moves of the iterations that were not
 * executed are taken back (flips commute).  Same result as rlsb_metro_sampling's count + apply passes; the caller
 * advances the generator by 2 calls per executed iteration.  workspace: rlsb_metro_workspace_bytes(N, C, max_iters)
 * bytes of device memory, 256-byte aligned. */
int64_t rlsb_metro_workspace_bytes(int32_t num_nodes, int64_t num_chains, int32_t max_iters);
int rlsb_metro_sampling_split(int32_t num_nodes, const float* probs, const float* start, float* out, int64_t num_chains,
                              int32_t max_iters, int64_t stop_count, uint64_t seed, uint64_t offset, uint32_t rng_threads,
                              uint32_t rng_iters, int32_t* num_iters_dev, void* workspace, void* stream);

/* sub_set_sampling, the resampling loop (rlsolver/methods/L2A/transformer.py:346-352):
 * xs[row][ids[row % S][k]] = rand_k[row] < vals[row % S][k] for k < top_k, one rand_like draw of
 * `rows` floats per k.  xs: bool [rows][N] (rows = num_repeats*S) in place; ids int64 / vals
 * float32: [S][top_k] (the topk of the determinism, smallest first); explicit_u [top_k][rows] or null. */
int rlsb_subset_sampling(uint8_t* xs, int64_t rows, int32_t num_nodes, int64_t num_sims, int32_t top_k,
                         const int64_t* ids, const float* vals, const float* explicit_u, uint64_t seed,
                         uint64_t offset, uint32_t rng_threads, uint32_t rng_iters, void* stream);

/* ---- dense QUBO Hamiltonian on the tensor cores: E[c] = x_c^T Q x_c, the "compute value" step of
 * mcpg_sampling_qubo / _qubo_bin (rlsolver/methods/MCPG/sampling.py:339-340, 364-365) and the
 * dense energy of PISCO (rlsolver/envs/env_ISCO.py:436-444).
 * qubo_create: q = float32 [N][N] DEVICE pointer; splits Q once into three bf16 limbs (exact) and
 *   builds the TMA descriptor.  N is padded to a multiple of 128 internally.
 * qubo_energy: x = float32 [N][C] node-major DEVICE pointer whose entries are exactly representable
 *   in bf16 (the reference's {-1,+1}, {0,1}, or 0 for a masked variable); energy = float32 [C].
 *   tcgen05.mma (bf16 x bf16 -> fp32 in tensor memory) with the x^T (Q x) reduction fused into the
 *   epilogue; deterministic.  Accuracy: fp32 accumulation like the reference's SGEMM; the tests hold
 *   it to 1e-5 relative against float64.  workspace: 256-byte aligned, rlsb_qubo_workspace_bytes(). */
int rlsb_qubo_create(const float* q, int32_t num_vars, int32_t device, rlsb_qubo_t** out, void* stream);
int rlsb_qubo_destroy(rlsb_qubo_t* h);
int32_t rlsb_qubo_padded_vars(const rlsb_qubo_t* h);
int64_t rlsb_qubo_workspace_bytes(const rlsb_qubo_t* h, int64_t num_chains);
int rlsb_qubo_energy(const rlsb_qubo_t* h, const float* x, int64_t num_chains, float* energy, void* workspace,
                     void* stream);
/* qubo_sweeps: the local-search sweeps of mcpg_sampling_qubo (binary == 0: x in {-1,+1},
 *   x_i <- sign(Q_i . x with x_i zeroed), ties -> -1; sampling.py:331-337) and mcpg_sampling_qubo_bin
 *   (binary != 0: x in {0,1}, x_i <- [Q_i . x > -Q_ii / 2]; sampling.py:356-362), num_sweeps passes
 *   over index 0..N-1 in order, in place on x = float32 [N][C].  q = the float32 [N][N] matrix the
 *   handle was created from (diagonal blocks are read in full precision).  Blocked Gauss-Seidel:
 *   per 128 rows one tensor-core GEMM tile + the in-order corrections; decisions are exact for
 *   integer-valued Q, and differ from an fp32 GEMV only where |Q_i . x| is below its rounding error. */
int rlsb_qubo_sweeps(const rlsb_qubo_t* h, const float* q, float* x, int64_t num_chains, int32_t num_sweeps,
                     int32_t binary, void* workspace, void* stream);

/* ---- pattern-I environment with one graph per environment: SpinSystemUnbiased of
 * rlsolver/methods/ECO_S2V/src/envs/spinsystem_PECO.py (reset 151-193, step 306-486).
 * matrix: float32 [E][N][N], symmetric (the reference's generators produce symmetric matrices);
 * spins are +-1 floats; state is the reference's observation block float32 [E][num_obs][N] whose
 * row 0 is the spin state.
 * peco_fields: as[e][j] = (A s)_j, fields[e][j] = s_j (A s)_j (_get_immeditate_cuts_avaialable,
 *   660-662), cut[e] = calculate_cut (601-607); each output nullable.
 * peco_step: flips state[e][0][action[e]], updates the resident (A s) by one matrix row, score +=
 *   -fields_new[action], reward (reward_signal 1 DENSE, 2 BLS, 4 CUSTOM_BLS; / N if norm_rewards;
 *   stag punishment / basin reward against the visited-state history), best_score / best_spins, and
 *   rewrites the observables in place.  h_obs_rows: HOST int32[7] = state row of
 *   {IMMEDIATE_REWARD_AVAILABLE, TIME_SINCE_FLIP, EPISODE_TIME, TERMINATION_IMMANENCY,
 *   NUMBER_OF_GREEDY_ACTIONS_AVAILABLE, DISTANCE_FROM_BEST_SCORE, DISTANCE_FROM_BEST_STATE}, -1 =
 *   absent.  inv_steps = float32(1/max_steps); termination = the TERMINATION_IMMANENCY value of
 *   this step (a host scalar).  history: uint32 [capacity][E][ceil(N/32)] (nullable), hist_len =
 *   states recorded since reset; the kernel appends the new state at index hist_len.
 *   scalar_div_as_cuda: the reference divides by the Python scalar n_spins in two places; torch's CUDA
 *   kernel multiplies by the float32 reciprocal there, torch's CPU kernel divides (1 ulp apart): 1
 *   reproduces the reference running on a GPU, 0 the reference running on the CPU (the goldens).
 *   Infinite memory (memory_length None), reversible spins, ExtraAction.NONE only. */
int rlsb_peco_fields(const float* matrix, const float* spins, int64_t num_envs, int32_t num_spins, float* as,
                     float* fields, float* cut, void* stream);
int rlsb_peco_step(const float* matrix, float* state, float* as, const int64_t* action, float* score,
                   float* best_score, float* best_spins, const float* max_local, float* reward, uint32_t* history,
                   int32_t hist_len, int32_t* bad_actions, int64_t num_envs, int32_t num_spins, int32_t num_obs,
                   const int32_t* h_obs_rows, int32_t reward_signal, int32_t norm_rewards, float inv_steps,
                   float termination, int32_t use_stag, float stag, int32_t use_basin, float basin,
                   int32_t scalar_div_as_cuda, void* stream);

/* ---- select ops on the reference's bool layout
 * select_rows: update_xs_by_vs (util_read_data.py:190-202): rows of (xs1,vs1) replace
 *   rows of (xs0,vs0) where vs1 >= vs0 (<= when maximize == 0).
 * pick_best: pick_xs_by_vs (util_read_data.py:204-216): input viewed [R][S][N];
 *   per sim the repeat with the best value, lowest repeat index on ties. */
int rlsb_select_rows(uint8_t* xs0, int64_t* vs0, const uint8_t* xs1, const int64_t* vs1, int64_t num_envs,
                     int32_t num_nodes, int32_t maximize, void* stream);
int rlsb_pick_best(const uint8_t* xs, const int64_t* vs, int32_t num_repeats, int64_t num_sims, int32_t num_nodes,
                   int32_t maximize, uint8_t* out_xs, int64_t* out_vs, void* stream);
/* rows dst_ids[i] <- rows src_ids[i] (xs bool [E][N], vs int64 [E]) for disjoint index sets: the row copy of
 * evolutionary_replacement (rlsolver/methods/util.py:87-94; its argsort and randperm stay torch calls so that ties and
 * the random stream are the reference's).  *bad_ids counts pairs with an index outside [0, E). */
int rlsb_copy_rows(uint8_t* xs, int64_t* vs, const int64_t* dst_ids, const int64_t* src_ids, int64_t count, int64_t num_envs,
                   int32_t num_nodes, int32_t* bad_ids, void* stream);

/* ---- multi-GPU best-cut exchange (new; the reference is single-process): the record a rank
 * contributes to the all-gather -- 64-bit key ((cut + 2^31) << 32) | (0xFFFFFFFF - global_env_id) of its best
 * row, compared unsigned (negative values order below positive ones, values saturate to the int32 range; ties
 * go to the lowest global env id), then the row's N bytes.
 * vs int64 [E], xs bool [E][N], record uint8 [8 + N] (8-byte aligned); env_offset = rank * E. */
int rlsb_best_record(const int64_t* vs, const uint8_t* xs, int64_t num_envs, int32_t num_nodes, int64_t env_offset,
                     uint8_t* record, void* stream);
/* the same record taken from packed tiles (uint32 [ceil(E/32)][Np]); the row is expanded to one byte per node */
int rlsb_best_record_packed(const int64_t* vs, const uint32_t* packed, int64_t num_envs, int32_t num_nodes,
                            int32_t padded_nodes, int64_t env_offset, uint8_t* record, void* stream);
/* winner of `world` gathered records (uint8 [world][8 + N]): out2[0] = best cut, out2[1] = its global
 * env id, row = its N spin bytes. */
int rlsb_best_pick(const uint8_t* gathered, int32_t world, int32_t num_nodes, int64_t* out2, uint8_t* row, void* stream);
/* the same for records `stride` bytes apart (stride >= 8 + N; padded so that every record is 8-byte aligned) */
int rlsb_best_pick_strided(const uint8_t* gathered, int32_t world, int32_t num_nodes, int64_t stride, int64_t* out2,
                           uint8_t* row, void* stream);

/* ---- the same exchange as ONE kernel over NVLink peer memory (csrc/peer_exchange.cu), one process per GPU of one box.
 * Replaces record kernel + ncclAllGather + pick kernel of the sharded best-cut exchange (the reference has no
 * multi-GPU form of this path: rlsolver/envs/env_L2A.py:118-165 runs one device).  Every rank owns a
 * mailbox in its HBM which the other ranks map through CUDA IPC; a call stores the rank's record into every mailbox,
 * raises an arrival word, polls its own arrival words and picks the winner.  Collective: every rank must make the same
 * sequence of calls.  The poll is bounded (20 s, RLSB_PEER_TIMEOUT_MS overrides): time-outs are counted, never hang.
 * The launch has no per-call arguments (the call counter lives in the mailbox), so it may be captured in a CUDA graph.
 *   create : allocates the mailbox on the current device, handle_out receives rlsb_peer_exchange_handle_bytes() bytes
 *   connect: handles = the world handles, rank-major (all-gathered by the host side) */
typedef struct rlsb_peer_exchange rlsb_peer_exchange_t;
int64_t rlsb_peer_exchange_handle_bytes(void);
int rlsb_peer_exchange_create(int32_t rank, int32_t world, int32_t num_nodes, rlsb_peer_exchange_t** out, uint8_t* handle_out);
int rlsb_peer_exchange_connect(rlsb_peer_exchange_t* ex, const uint8_t* handles);
/* vs int64 [E], xs bool [E][N]; env_offset = global id of local env 0; out2[0] = best cut, out2[1] = its global env id,
 * row = the winner's N spin bytes -- identical on every rank */
int rlsb_peer_exchange_best(rlsb_peer_exchange_t* ex, const int64_t* vs, const uint8_t* xs, int64_t num_envs, int64_t env_offset,
                            int64_t* out2, uint8_t* row, void* stream);
/* the same for a batch kept as packed tiles (uint32 [ceil(E/32)][Np]) */
int rlsb_peer_exchange_best_packed(rlsb_peer_exchange_t* ex, const int64_t* vs, const uint32_t* packed, int64_t num_envs,
                                   int32_t padded_nodes, int64_t env_offset, int64_t* out2, uint8_t* row, void* stream);
/* out[0] = calls completed, out[1] = polls that ran into the time-out; synchronises `stream` */
int rlsb_peer_exchange_status(rlsb_peer_exchange_t* ex, uint32_t* out, void* stream);
int rlsb_peer_exchange_destroy(rlsb_peer_exchange_t* ex);

/* ---- weighted max-cut sampler of MCPG: the local-search sweeps + expected cut of mcpg_sampling_maxcut
 * (rlsolver/methods/MCPG/sampling.py:101-121) for float edge weights (MCPG/dataloader.py:53-103).
 * xs float32 [N][C] node-major, in: the 0/1 output of metro_sampling, out: the states after the sweeps (0/1).
 * order int32 [N] (data.sorted_degree_nodes); nbr_* = data.neighbors / data.neighbor_edges as CSR (edge order, both
 * directions); thr[i] = float32(data.weighted_degree[i] / 2 + 0.125); edge_* = data.edge_index / edge_attr.
 * The kernel applies the symmetry-breaking XOR with row order[0] (:102-104), runs num_sweeps >= 1 Gauss-Seidel
 * sweeps `x_i <- [sum_j w_ij x_j + rand / 4 < thr_i]` and writes expected[c] = sum_e (2x_u - 1)(2x_v - 1) w_e.
 * Random numbers: explicit_u float32 [num_sweeps * N][C] (replayed draws), or torch's CUDA Philox stream from
 * (seed, offset) with the call geometry of torch.rand(C) -- one call per node visit; the caller advances the
 * generator by 4 * rng_iters * num_sweeps * N. */
int rlsb_mcpg_weighted_sweeps(int32_t num_nodes, int64_t num_chains, const int32_t* order, const int32_t* nbr_ptr,
                              const int32_t* nbr_col, const float* nbr_w, const float* thr, int32_t num_edges,
                              const int32_t* edge_u, const int32_t* edge_v, const float* edge_w, float* xs,
                              int32_t num_sweeps, const float* explicit_u, uint64_t seed, uint64_t offset,
                              uint32_t rng_threads, uint32_t rng_iters, float* expected, void* stream);

/* ---- pattern-I environment, COMPACT resident state (round 2; csrc/peco_compact.cu).  Same semantics as
 * rlsb_peco_step for matrices with entries in {-1, 0, +1} (what util_envs_PECO.py:15-113 generates), with the
 * state that lives in HBM reduced to bits and small integers:
 *   adj uint32 [E][N][W] adjacency bit rows (W = ceil(N/32), diagonal bit = self loop); sgn uint32 bit = weight -1,
 *   laid out [E][N][W] (sgn_stride = N*W), one shared [N][W] matrix (sgn_stride = 0) or NULL (all +1);
 *   spins / best_spins uint32 [E][W] (bit = spin +1); fields [E][Np] = (A s)_j, int8 when N <= 128 (|field| <= N - 1) and int16 otherwise; last_flip uint16 [E][Np];
 *   score / best_score / max_local / reward float32 [E].
 * compact_step    : SpinSystemUnbiased.step (spinsystem_PECO.py:306-486) for ExtraAction.NONE, reversible spins,
 *                   infinite memory; `step` = 1-based index of this step.  Visited-state test (HistoryBuffer,
 *                   util_envs_PECO.py:228-288) as a per-env open-addressing set of 64-bit Zobrist keys: hset uint64
 *                   [E][hcap] (hcap a power of two >= 2 * (max_steps + 1), zeroed at reset), hkey uint64 [E] (0 at
 *                   reset), zobrist uint64 [N] random constants; NULL when no stag / basin reward is used.
 *                   done_out uint8 [E] (nullable) = last_step, or -- irreversible spins, the S2V-DQN pattern of
 *                   ECO_S2V/src/envs/spinsystem.py:476-480 -- no +1 spin left in the env after this flip.
 * compact_fields  : fields, cut (:601-607) and max_local (:163-171, from the all-ones state) for given spins;
 *                   *empty_graphs counts envs whose all-ones fields are all zero or whose largest is zero (:166-171).
 * compact_from_dense / expand_matrix : dense float32 [E][N][N] <-> bit rows (*bad_entries counts entries not in
 *                   {-1, 0, 1}).  expand_state writes the reference's float32 state rows (obs row indices as in
 *                   rlsb_peco_step; table[k] = k-fold float32 accumulation of 1/max_steps, k <= max_steps);
 *                   *_env_stride = floats between consecutive envs of the destination (so both can write straight
 *                   into an observation tensor [E][num_obs + N][N]).
 * gen_er / gen_ba : RandomERGraphGenerator.get / RandomBAGraphGenerator.get (util_envs_PECO.py:40-52, 87-107)
 *                   written as bit rows from torch's CUDA Philox stream (ER: the one rand(E, N, N) call; BA: the
 *                   exponential_ call inside each torch.multinomial): same graphs as the reference's torch ops for
 *                   the same seed.  The caller advances the generator (ER: 4 * rng_iters; BA: (N - m - 1) calls of
 *                   E * N elements).  E * N * N (ER) and E * N (BA) must stay below 2^31 per call. */
int rlsb_peco_compact_step(const uint32_t* adj, const uint32_t* sgn, int64_t sgn_stride, uint32_t* spins,
                           void* fields, uint16_t* last_flip, uint32_t* best_spins, float* score, float* best_score,
                           const float* max_local, float* reward, const int64_t* action, uint64_t* hset,
                           int32_t hcap, uint64_t* hkey, const uint64_t* zobrist, int32_t* bad_actions,
                           int64_t num_envs, int32_t num_spins, int32_t step, int32_t reward_signal,
                           int32_t norm_rewards, int32_t use_stag, float stag, int32_t use_basin, float basin,
                           int32_t scalar_div_as_cuda, uint8_t* done_out, int32_t last_step, int32_t irreversible,
                           void* stream);
int rlsb_peco_compact_fields(const uint32_t* adj, const uint32_t* sgn, int64_t sgn_stride, const uint32_t* spins,
                             int64_t num_envs, int32_t num_spins, void* fields, float* cut, float* max_local,
                             int32_t* empty_graphs, void* stream);
int rlsb_peco_compact_from_dense(const float* matrix, int64_t num_envs, int32_t num_spins, uint32_t* adj, uint32_t* sgn,
                                 int32_t* bad_entries, void* stream);
int rlsb_peco_compact_expand_matrix(const uint32_t* adj, const uint32_t* sgn, int64_t sgn_stride, int64_t num_envs,
                                    int32_t num_spins, float* out, int64_t out_env_stride, void* stream);
int rlsb_peco_compact_expand_state(const uint32_t* spins, const uint32_t* best_spins, const void* fields,
                                   const uint16_t* last_flip, const float* score, const float* best_score,
                                   const float* max_local, const float* table, float* state, int64_t state_env_stride,
                                   int64_t num_envs, int32_t num_spins, int32_t num_obs, const int32_t* h_obs_rows,
                                   int32_t step, int32_t binary_spins, float termination, int32_t scalar_div_as_cuda,
                                   int32_t at_reset, void* stream);
/* expand_state with float16 rows: the inference env's use_tensor_core mode (inference_network_env.py:143-145, 212-236).
 * state: __half [E][...] with state_env_stride counted in halves; table[k] = the k-fold accumulation of 1/max_steps as
 * torch performs it on a half tensor (float32 add, rounded to half each time), stored as float32. */
int rlsb_peco_compact_expand_state_half(const uint32_t* spins, const uint32_t* best_spins, const void* fields,
                                        const uint16_t* last_flip, const float* score, const float* best_score,
                                        const float* max_local, const float* table, void* state, int64_t state_env_stride,
                                        int64_t num_envs, int32_t num_spins, int32_t num_obs, const int32_t* h_obs_rows,
                                        int32_t step, int32_t binary_spins, float termination, int32_t scalar_div_as_cuda,
                                        int32_t at_reset, void* stream);
int rlsb_peco_gen_er(uint32_t* adj, int64_t num_envs, int32_t num_spins, float p_connection, uint64_t seed,
                     uint64_t offset, uint32_t rng_threads, uint32_t rng_iters, void* stream);
int rlsb_peco_gen_ba(uint32_t* adj, int64_t num_envs, int32_t num_spins, int32_t m_insertion_edges, uint64_t seed,
                     uint64_t offset, uint32_t rng_threads, uint32_t rng_iters, void* stream);
/* Power-law cluster graphs (Holme-Kim growth = networkx.powerlaw_cluster_graph, the reference's third graph family:
 * rlsolver/methods/util_generate.py:75-93 with m = 4, p = 0.05) as bit rows, one graph per env, from a Philox stream
 * (seed, subsequence = env, offset).  The reference grows single graphs on the host from Python's `random`, so the
 * agreement is statistical (edge count m (n - m) up to repeated picks, degree tail, clustering), not bitwise.
 * num_spins <= 128. */
int rlsb_peco_gen_pl(uint32_t* adj, int64_t num_envs, int32_t num_spins, int32_t m_insertion_edges, float p_triangle,
                     uint64_t seed, uint64_t offset, void* stream);

/* ---- float tail of an ISCO / PISCO Metropolis-Hastings step (csrc/isco.cu), one CTA per chain:
 * propose = get_local_dist + multinomial + the flip (rlsolver/envs/env_ISCO.py:37-63 / 394-418,
 *           rlsolver/methods/ISCO/util.py:3-60); accept = get_local_dist(y) + ll_y2x + mh_step (env_ISCO.py:27-35, 65-77,
 *           util.py:62-75).  x / y: [B][ld] float32 (is_half = 0) or float16 (1), entries {0, 1}; ld >= num_nodes, sites
 *           >= num_nodes are padding with gain 0 (PISCO pads to a multiple of 8 and its softmax includes them).
 * cross: per (chain, site) uint16 cross counts [B][cross_ld] of rlsb_node_cross_counts (weighted = 0) or the int32
 *           sums of rlsb_node_fields_weighted (weighted = 1); deg int32 [N] = (weighted) degree; the flip gain is
 *           (deg - 2 cross) / (2 T).  cut int64 [B] (ISCO energy = cut / T; pisco = 1: -1/4 fp16(sum of gains) / T).
 * u: the uniform draws the reference makes (gumbel: [B][ld], then bernoulli_logp: [B]).  sel int32 [B][kmax]: the
 *           chosen sites in order (-1 padded), kmax >= max path_length.  accept overwrites y with the next state. */
int rlsb_isco_propose(const void* x, void* y, int32_t is_half, const void* cross, int32_t weighted, int32_t cross_ld,
                      const int32_t* deg, const int64_t* cut, int32_t pisco, const float* temperature,
                      const int64_t* path_length, const float* u, int32_t* sel, int32_t kmax, float* ll_x, float* ll_x2y,
                      int32_t num_nodes, int32_t ld, int64_t num_chains, void* stream);
int rlsb_isco_accept(const void* x, void* y, int32_t is_half, const void* cross_y, int32_t weighted, int32_t cross_ld,
                     const int32_t* deg, const int64_t* cut_y, int32_t pisco, const float* temperature, const float* u,
                     const int32_t* sel, int32_t kmax, const float* ll_x, const float* ll_x2y, float* ll_y_times_t,
                     float* acc, int32_t num_nodes, int32_t ld, int64_t num_chains, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RLSOLVER_B200_H */
