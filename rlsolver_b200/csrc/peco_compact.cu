// Pattern-I environment with one graph per environment, COMPACT resident state (round 2).
//
// The reference keeps, per env, a dense float32 matrix [N, N] and a float32 state [7, N] and recomputes every
// local field with a batched matmul each step (rlsolver/methods/ECO_S2V/src/envs/spinsystem_PECO.py:306-486);
// round 1 kept those tensors and updated them in place (peco.cu, 4.4 KB of traffic per env-step).  For the graphs
// the generators produce (util_envs_PECO.py:15-113: weights in {-1, 0, +1}) everything is a small integer or a
// bit, so the state that lives in HBM here is
//
//   adj    uint32 [E][N][W]   adjacency rows as bit masks (W = ceil(N / 32); the diagonal bit is a self loop)
//   sgn    uint32 [S][N][W]   bit = 1: weight -1.  S = E (per-env signs), 1 (one sign matrix shared by every env:
//                             EdgeType.DISCRETE draws ONE [N, N] mask, util_envs_PECO.py:27-29) or absent (all +1)
//   spins  uint32 [E][W]      bit = 1: s = +1
//   fields int8 / int16 [E][Np]  (A s)_j -- one byte when N <= 128 (|(A s)_j| <= N - 1), two otherwise
//   last_flip uint16 [E][Np]  step at which node j was flipped last (TIME_SINCE_FLIP is a function of the gap)
//   best_spins uint32 [E][W], score / best_score / max_local float32 [E], the visited set (below)
//
// and a step reads one 2 * 4W-byte matrix row instead of 4N bytes, touches the fields of the acted node's
// neighbours only and rewrites no observable: 3.3 KB per env resident instead of 43 KB, ~0.5 KB of DRAM sectors
// touched per env-step instead of 4.4 KB.  The reference's tensors -- `state [E, obs, N]`, `matrix [E, N, N]`,
// the observation -- are MATERIALISED on request by the expand kernels, with exactly the float32 values the
// reference's step-by-step updates produce (every float is either an integer, one IEEE division of integers, or
// a k-fold accumulation of 1 / max_steps looked up in a table built with the same float32 additions).
//
// Visited-state test (HistoryBuffer, util_envs_PECO.py:228-288: an ever-growing [T, E, N/8] byte buffer XOR-scanned
// every step): a 64-bit Zobrist key per env (key ^= Z[a] per flip) in a per-env open-addressing table -- O(1) per
// step.  Two different states collide with probability ~2^-64 per pair; peco.cu keeps the exact scan.
//
// Generators (util_envs_PECO.py:40-52 ER, 87-107 BA) write the bit rows directly from torch's Philox stream: the
// same graphs as the reference's torch ops on the same device and seed, no [E, N, N] float tensor on the way.
#include <cuda_fp16.h>

#include "common.cuh"
#include "philox.cuh"

namespace rlsb {

constexpr int kPcWarps = 8;
constexpr int kPcByteFields = 128;     // up to this many spins the fields are int8 (peco_field_bytes)

struct PecoC {
  const uint32_t* adj;
  const uint32_t* sgn;        // null: every weight +1
  int64_t sgn_stride;         // words between the sign matrices of consecutive envs (0: shared)
  uint32_t* spins;
  void* fields;               // int8 [E][Np] when n <= kPcByteFields, int16 [E][Np] otherwise
  uint16_t* last_flip;
  uint32_t* best_spins;
  float *score, *best_score;
  const float* max_local;
  float* reward;
  const int64_t* action;
  unsigned long long* hset;   // [E][hcap] visited keys (0 = empty), nullable
  unsigned long long* hkey;   // [E] Zobrist key of the current state
  const unsigned long long* zobrist;   // [N]
  int32_t* bad_actions;
  uint8_t* done;              // [E] nullable: last_step, or (irreversible) no +1 spin left after this flip
  int last_step, irreversible;
  int64_t num_envs;
  int n, np, words, hcap, step;
  int reward_signal;          // 1 DENSE, 2 BLS, 4 CUSTOM_BLS
  int norm_rewards, use_stag, use_basin, recip_div;
  float stag, basin, inv_n;
};

__device__ __forceinline__ int warp_sum_i(int v) { return __reduce_add_sync(kFull, v); }

// One warp per env.  Everything that does not depend on the action -- the spin words, the env's fields, its best
// spins and scalars -- is requested BEFORE the action is looked at, so that a step is two dependent trips to
// HBM (action -> matrix row) instead of four: the kernel is bound by that latency chain, not by bandwidth.
// KW: words per spin vector held in registers (N <= 32 KW); 0 = generic loop for larger N.
template <int KW, typename FT>
__global__ void __launch_bounds__(kPcWarps * 32) peco_compact_step_kernel(PecoC p) {
  const int lane = threadIdx.x & 31;
  const int64_t env = (int64_t)blockIdx.x * kPcWarps + (threadIdx.x >> 5);
  if (env >= p.num_envs) return;
  const int n = p.n, W = p.words;
  FT* fl = static_cast<FT*>(p.fields) + env * (int64_t)p.np;
  const int64_t a = p.action[env];
  uint32_t sp = lane < W ? p.spins[env * W + lane] : 0u;           // lane w holds word w
  int pre[KW > 0 ? KW : 1];
  if (KW > 0) {
#pragma unroll
    for (int k = 0; k < KW; ++k) pre[k] = (k < W && 32 * k + lane < n) ? (int)fl[32 * k + lane] : 0;
  }
  const uint32_t bs = lane < W ? p.best_spins[env * W + lane] : 0u;
  const float score0 = p.score[env], best_obs = p.best_score[env];
  unsigned long long key0 = 0ull;
  if (p.hset) key0 = p.hkey[env];
  (void)bs;
  if (a < 0 || a >= n) {                               // IndexError in the reference
    if (lane == 0) {
      p.reward[env] = 0.f, atomicAdd(p.bad_actions, 1);
      if (p.done) p.done[env] = (uint8_t)p.last_step;
    }
    return;
  }
  const int wa = (int)a >> 5, ba = (int)a & 31;
  const uint32_t arow = lane < W ? __ldg(p.adj + (env * n + a) * W + lane) : 0u;
  const uint32_t srow = (p.sgn && lane < W) ? __ldg(p.sgn + env * p.sgn_stride + a * W + lane) : 0u;
  const int s_old = ((__shfl_sync(kFull, sp, wa) >> ba) & 1u) ? 1 : -1;
  if (lane == wa) {
    sp ^= 1u << ba;
    p.spins[env * W + lane] = sp;
  }
  int nonpos = 0, delta = 0;
  auto node = [&](int k, int v0) {
    const uint32_t aw = __shfl_sync(kFull, arow, k), sw = __shfl_sync(kFull, srow, k), spw = __shfl_sync(kFull, sp, k);
    const int j = 32 * k + lane;
    if (j < n) {
      int v = v0;
      if ((aw >> lane) & 1u) {                         // (A s)_j -= 2 A[a][j] s_old
        v -= ((sw >> lane) & 1u) ? -2 * s_old : 2 * s_old;
        fl[j] = (FT)v;
      }
      const int f = ((spw >> lane) & 1u) ? v : -v;     // fields_j = s_j (A s)_j with the flipped spin
      nonpos += (int)(f <= 0);
      if (j == (int)a) delta = -f;
    }
  };
  if (KW > 0) {
#pragma unroll
    for (int k = 0; k < KW; ++k)
      if (k < W) node(k, pre[k]);
  } else {
    for (int k = 0; k < W; ++k) node(k, 32 * k + lane < n ? (int)fl[32 * k + lane] : 0);
  }
  nonpos = warp_sum_i(nonpos);
  delta = warp_sum_i(delta);
  if (lane == 0) p.last_flip[env * (int64_t)p.np + a] = (uint16_t)p.step;
  const float score = __fadd_rn(score0, (float)delta);
  const float improvement = __fsub_rn(score, best_obs);
  float rew = 0.f;
  if (p.reward_signal == 2) rew = improvement > 0.f ? improvement : 0.f;
  else if (p.reward_signal == 4) rew = improvement > 0.f ? __fdiv_rn(improvement, __fadd_rn(improvement, 0.1f)) : 0.f;
  else if (p.reward_signal == 1) rew = (float)delta;
  if (p.norm_rewards) rew = p.recip_div ? __fmul_rn(rew, p.inv_n) : __fdiv_rn(rew, (float)n);
  if (p.hset) {
    // the state's key moves by one table entry; look it up / insert it in the env's table (linear probing, a
    // window of 32 slots per round trip)
    unsigned long long key = key0 ^ __ldg(p.zobrist + a);
    if (lane == 0) p.hkey[env] = key;
    if (key == 0ull) key = 1ull;                       // 0 marks an empty slot
    unsigned long long* tab = p.hset + env * (int64_t)p.hcap;
    const uint32_t mask = (uint32_t)p.hcap - 1u;
    uint32_t h = (uint32_t)(key >> 17) & mask;
    bool fresh = true;
    for (int probe = 0; probe < p.hcap; probe += 32) {
      const uint32_t slot = (h + probe + lane) & mask;
      const bool in_range = probe + lane < p.hcap;
      const unsigned long long v = in_range ? tab[slot] : ~0ull;
      const uint32_t hit = __ballot_sync(kFull, in_range && v == key);
      const uint32_t empty = __ballot_sync(kFull, in_range && v == 0ull);
      const int first_empty = empty ? __ffs(empty) - 1 : 32;
      if (hit && (__ffs(hit) - 1) < first_empty) {     // found before the first hole of the probe sequence
        fresh = false;
        break;
      }
      if (empty) {
        if (lane == first_empty) tab[slot] = key;
        break;
      }
    }
    if (p.use_stag && !fresh) rew = __fsub_rn(rew, p.stag);
    if (p.use_basin && nonpos == n && fresh) rew = __fadd_rn(rew, p.basin);
  }
  const bool better = score > best_obs;
  if (better && lane < W) p.best_spins[env * W + lane] = sp;
  const bool none_left = __ballot_sync(kFull, sp != 0u) == 0u;        // padding bits are never set
  if (lane == 0) {
    p.score[env] = score;
    p.best_score[env] = better ? score : best_obs;
    p.reward[env] = rew;
    if (p.done) p.done[env] = (uint8_t)(p.last_step || (p.irreversible && none_left));
  }
}

// A few lanes per env (N <= 128): several envs share a warp, so several times as many dependent load chains are in
// flight per resident warp as with a warp per env.  Lane `sub` of a group owns kPcNpl consecutive nodes: they sit in ONE
// word of the spin / adjacency / sign rows, which the lane loads itself, and their byte-sized fields are 16-byte loads --
// no shuffles in the node loop.  History of this kernel (2^20 ER-100 envs): warp per env 0.35 ms; 8 lanes x 16 nodes
// with shuffles 0.233; no shuffles 0.186; byte fields 0.178; four fields per word 0.1435 (ncu: 94 % issue-active before
// that step, so instructions per env are what counts); 4 lanes x 32 nodes halves the per-env share of everything
// outside the node loop (addresses, scalars, reward arithmetic are computed by every lane of a group).
constexpr int kPcLpe = 4, kPcNpl = 32;       // lanes per env, nodes per lane (kPcLpe * kPcNpl = 128 >= N)
constexpr int kPcLpw = 32 / kPcNpl;          // lanes per 32-bit word of the bit rows
constexpr int kPcFw = kPcNpl / 4;            // 32-bit words of byte fields per lane
constexpr uint32_t kPcGroupMask = (1u << kPcLpe) - 1u;
static_assert(kPcLpe * kPcNpl == kPcByteFields && kPcNpl % 16 == 0 && 32 % kPcNpl == 0, "lane geometry");
// bits 0..3 -> 0x01 in bytes 0..3 (the partial products of the multiplier land on distinct bits: no carries)
__device__ __forceinline__ uint32_t spread4(uint32_t nib) { return ((nib & 0xFu) * 0x00204081u) & 0x01010101u; }
__global__ void __launch_bounds__(kPcWarps * 32) peco_compact_step8_kernel(PecoC p) {
  const int lane = threadIdx.x & 31, sub = lane & (kPcLpe - 1), grp = lane / kPcLpe, base = grp * kPcLpe;
  const int64_t env = ((int64_t)blockIdx.x * kPcWarps + (threadIdx.x >> 5)) * (32 / kPcLpe) + grp;
  const bool live = env < p.num_envs;
  const int64_t ev = live ? env : 0;             // dead groups shadow env 0 read-only (all lanes stay in the shuffles)
  const int n = p.n, W = p.words;
  const int myw = sub / kPcLpw, j0 = kPcNpl * sub;   // my word, my first node
  const bool has = myw < W;                      // the lane owns nodes of this graph (np >= 32 W covers them)
  int8_t* fl = static_cast<int8_t*>(p.fields) + ev * (int64_t)p.np;        // N <= 128: byte fields
  const int64_t a_raw = p.action[ev];
  const uint32_t sp0 = has ? p.spins[ev * W + myw] : 0u;
  uint32_t fw[kPcFw];
#pragma unroll
  for (int v4 = 0; v4 < kPcFw / 4; ++v4) {
    uint4 f = make_uint4(0, 0, 0, 0);
    if (has) f = *reinterpret_cast<const uint4*>(fl + j0 + 16 * v4);
    fw[4 * v4] = f.x, fw[4 * v4 + 1] = f.y, fw[4 * v4 + 2] = f.z, fw[4 * v4 + 3] = f.w;
  }
  const float score0 = p.score[ev], best_obs = p.best_score[ev];
  unsigned long long key0 = 0ull;
  if (p.hset) key0 = p.hkey[ev];
  const bool valid = live && a_raw >= 0 && a_raw < n;
  if (live && !valid && sub == 0) p.reward[env] = 0.f, atomicAdd(p.bad_actions, 1);     // IndexError in the reference
  const int a = valid ? (int)a_raw : 0;
  const int wa = a >> 5, ba = a & 31;
  const uint32_t arow = has ? __ldg(p.adj + (ev * n + a) * W + myw) : 0u;
  const uint32_t srow = (p.sgn && has) ? __ldg(p.sgn + ev * p.sgn_stride + (int64_t)a * W + myw) : 0u;
  const int s_old = ((__shfl_sync(kFull, sp0, base + kPcLpw * wa) >> ba) & 1u) ? 1 : -1;
  const uint32_t sp = (myw == wa) ? sp0 ^ (1u << ba) : sp0;      // every lane of the word sees the flipped spin
  if (valid && sub == kPcLpw * wa) p.spins[env * W + wa] = sp;
  const int sh = (sub % kPcLpw) * kPcNpl;        // my bits inside the word
  constexpr uint32_t kMine = kPcNpl == 32 ? 0xFFFFFFFFu : ((1u << (kPcNpl & 31)) - 1u);
  const uint32_t abits = (arow >> sh) & kMine, sbits = (srow >> sh) & kMine, pbits = (sp >> sh) & kMine;
  // Four byte-sized fields per 32-bit word, handled together.  spread4 turns 4 bits into 0x01 flags of 4 bytes; the
  // update adds +2 / -2 per flagged byte with a carry-free byte-wise add; "s_j (A s)_j <= 0" per byte is
  // zero | (negative == spin up).
  const uint32_t plus = s_old > 0 ? sbits : ~sbits;            // neighbours whose field moves by +2 (the others: -2)
  const int rest = n - j0;                                      // my nodes below n
  const uint32_t vbits = rest >= kPcNpl ? kMine : (rest > 0 ? (1u << rest) - 1u : 0u);
  int nonpos = 0;
#pragma unroll
  for (int q = 0; q < kPcFw; ++q) {
    const uint32_t a4 = (abits >> (4 * q)) & 0xFu;
    const uint32_t y = spread4(a4 & (plus >> (4 * q))) * 2u + spread4(a4 & ~(plus >> (4 * q))) * 0xFEu;
    const uint32_t w = ((fw[q] & 0x7F7F7F7Fu) + (y & 0x7F7F7F7Fu)) ^ ((fw[q] ^ y) & 0x80808080u);
    fw[q] = w;
    const uint32_t neg = (w >> 7) & 0x01010101u;
    const uint32_t zero = (~(((w & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | w) >> 7) & 0x01010101u;
    const uint32_t up = spread4((pbits >> (4 * q)) & 0xFu);
    nonpos += __popc((zero | (~(neg ^ up) & 0x01010101u)) & spread4((vbits >> (4 * q)) & 0xFu));
  }
  int delta = 0;
  const int ta = a - j0;                                         // the acted node, if it is one of mine
  if (ta >= 0 && ta < kPcNpl) {
    uint32_t wsel = fw[0];
#pragma unroll
    for (int q = 1; q < kPcFw; ++q) wsel = (ta >> 2) == q ? fw[q] : wsel;
    const int v = (int)(int8_t)((wsel >> (8 * (ta & 3))) & 0xFFu);
    delta = ((pbits >> ta) & 1u) ? -v : v;                       // -(s_a (A s)_a) with the flipped spin
  }
  // the lane's fields go back as 16-byte stores (a group rewrites at most the env's 128 bytes, whole sectors)
  if (valid && has && abits) {
#pragma unroll
    for (int v4 = 0; v4 < kPcFw / 4; ++v4)
      *reinterpret_cast<uint4*>(fl + j0 + 16 * v4) = make_uint4(fw[4 * v4], fw[4 * v4 + 1], fw[4 * v4 + 2], fw[4 * v4 + 3]);
  }
#pragma unroll
  for (int off = kPcLpe / 2; off >= 1; off >>= 1) {
    nonpos += __shfl_xor_sync(kFull, nonpos, off);
    delta += __shfl_xor_sync(kFull, delta, off);
  }
  if (valid && sub == 0) p.last_flip[env * (int64_t)p.np + a] = (uint16_t)p.step;
  const float score = __fadd_rn(score0, (float)delta);
  const float improvement = __fsub_rn(score, best_obs);
  float rew = 0.f;
  if (p.reward_signal == 2) rew = improvement > 0.f ? improvement : 0.f;
  else if (p.reward_signal == 4) rew = improvement > 0.f ? __fdiv_rn(improvement, __fadd_rn(improvement, 0.1f)) : 0.f;
  else if (p.reward_signal == 1) rew = (float)delta;
  if (p.norm_rewards) rew = p.recip_div ? __fmul_rn(rew, p.inv_n) : __fdiv_rn(rew, (float)n);
  if (p.hset) {
    unsigned long long key = key0 ^ __ldg(p.zobrist + a);
    if (valid && sub == 0) p.hkey[env] = key;
    if (key == 0ull) key = 1ull;
    unsigned long long* tab = p.hset + ev * (int64_t)p.hcap;
    const uint32_t mask = (uint32_t)p.hcap - 1u;
    const uint32_t h = (uint32_t)(key >> 17) & mask;
    bool fresh = true, done = !valid;
    for (int probe = 0; probe < p.hcap; probe += kPcLpe) {      // windows of kPcLpe slots per group; groups finish on their own
      const uint32_t slot = (h + probe + sub) & mask;
      const unsigned long long v = done ? ~0ull : tab[slot];
      const uint32_t hit = (__ballot_sync(kFull, !done && v == key) >> base) & kPcGroupMask;
      const uint32_t empty = (__ballot_sync(kFull, !done && v == 0ull) >> base) & kPcGroupMask;
      const int first_empty = empty ? __ffs(empty) - 1 : kPcLpe;
      if (!done) {
        if (hit && (__ffs(hit) - 1) < first_empty) fresh = false, done = true;
        else if (empty) {
          if (sub == first_empty) tab[slot] = key;
          done = true;
        }
      }
      if (__all_sync(kFull, done)) break;
    }
    if (p.use_stag && !fresh) rew = __fsub_rn(rew, p.stag);
    if (p.use_basin && nonpos == n && fresh) rew = __fadd_rn(rew, p.basin);
  }
  const bool none_left = ((__ballot_sync(kFull, has && sp != 0u) >> base) & kPcGroupMask) == 0u;
  if (live && sub == 0 && p.done) p.done[env] = (uint8_t)(p.last_step || (valid && p.irreversible && none_left));
  if (!valid) return;
  const bool better = score > best_obs;
  if (better && has && sub % kPcLpw == 0) p.best_spins[env * W + myw] = sp;
  if (sub == 0) {
    p.score[env] = score;
    p.best_score[env] = better ? score : best_obs;
    p.reward[env] = rew;
  }
}

// ---- reset pieces ----------------------------------------------------------------------------------------------
// fields = A s for the given spins, the cut (spinsystem_PECO.py:601-607: 1/4 sum(-s A s) + 1/4 sum(A)), the largest
// field of the all-ones state (max_local_reward_available, :163-171) and a flag for envs whose graph is empty
__global__ void __launch_bounds__(kPcWarps * 32) peco_compact_fields_kernel(const uint32_t* __restrict__ adj,
                                                                            const uint32_t* __restrict__ sgn,
                                                                            int64_t sgn_stride,
                                                                            const uint32_t* __restrict__ spins,
                                                                            int64_t num_envs, int n, int np, int W,
                                                                            void* __restrict__ fields,
                                                                            float* __restrict__ cut,
                                                                            float* __restrict__ max_local,
                                                                            int32_t* __restrict__ empty_graphs) {
  const int lane = threadIdx.x & 31;
  const int64_t env = (int64_t)blockIdx.x * kPcWarps + (threadIdx.x >> 5);
  if (env >= num_envs) return;
  const uint32_t sp = lane < W ? spins[env * W + lane] : 0u;
  int sas = 0, suma = 0, best_one = INT_MIN, abs_one = 0;
  for (int j = 0; j < n; ++j) {                       // lane w handles word w of row j
    uint32_t aw = 0, sw = 0;
    if (lane < W) {
      aw = __ldg(adj + (env * n + j) * W + lane);
      if (sgn) sw = __ldg(sgn + env * sgn_stride + (int64_t)j * W + lane);
    }
    // sum_k A[j][k] s_k = (+1 edges to +spins) + (-1 edges to -spins) - (+1 edges to -spins) - (-1 edges to +spins)
    const uint32_t pos = aw & ~sw, neg = aw & sw;
    int as = __popc(pos & sp) + __popc(neg & ~sp) - __popc(pos & ~sp) - __popc(neg & sp);
    int rs = __popc(pos) - __popc(neg);                // row sum = (A 1)_j
    as = warp_sum_i(as), rs = warp_sum_i(rs);
    const int sj = ((__shfl_sync(kFull, sp, j >> 5) >> (j & 31)) & 1u) ? 1 : -1;
    if (lane == 0 && fields) {
      if (n <= kPcByteFields) static_cast<int8_t*>(fields)[env * (int64_t)np + j] = (int8_t)as;
      else static_cast<int16_t*>(fields)[env * (int64_t)np + j] = (int16_t)as;
    }
    sas += as * sj, suma += rs;
    best_one = max(best_one, rs), abs_one += abs(rs);
  }
  if (lane == 0) {
    if (cut) cut[env] = __fadd_rn(__fmul_rn(0.25f, (float)(-sas)), __fmul_rn(0.25f, (float)suma));
    if (max_local) max_local[env] = (float)best_one;
    if (empty_graphs && (abs_one == 0 || best_one == 0)) atomicAdd(empty_graphs, 1);
  }
}

// dense float32 [E][N][N] with entries in {-1, 0, +1} -> bit rows (+ per-env sign rows); *bad counts other entries
__global__ void __launch_bounds__(256) peco_compact_from_dense_kernel(const float* __restrict__ matrix,
                                                                      int64_t num_envs, int n, int W,
                                                                      uint32_t* __restrict__ adj,
                                                                      uint32_t* __restrict__ sgn,
                                                                      int32_t* __restrict__ bad) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);      // (env, node) pair
  if (row >= num_envs * n) return;
  const float* src = matrix + row * n;
  for (int k = 0; k < W; ++k) {
    const int j = 32 * k + lane;
    const float v = j < n ? __ldg(src + j) : 0.f;
    if (v != 0.f && v != 1.f && v != -1.f) atomicAdd(bad, 1);
    const uint32_t aw = __ballot_sync(kFull, v != 0.f), sw = __ballot_sync(kFull, v < 0.f);
    if (lane == 0) adj[row * W + k] = aw, sgn[row * W + k] = sw;
  }
}

// bit rows -> dense float32 [E][N][N] (the reference's `matrix` attribute / the lower block of the observation)
__global__ void __launch_bounds__(256) peco_compact_expand_matrix_kernel(const uint32_t* __restrict__ adj,
                                                                         const uint32_t* __restrict__ sgn,
                                                                         int64_t sgn_stride, int64_t num_envs, int n,
                                                                         int W, float* __restrict__ out,
                                                                         int64_t out_env_stride) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= num_envs * n) return;
  const int64_t env = row / n;
  const int i = (int)(row - env * n);
  float* dst = out + env * out_env_stride + (int64_t)i * n;
  for (int k = 0; k < W; ++k) {
    const int j = 32 * k + lane;
    if (j >= n) break;
    const uint32_t aw = __ldg(adj + row * W + k);
    const uint32_t sw = sgn ? __ldg(sgn + env * sgn_stride + (int64_t)i * W + k) : 0u;
    dst[j] = ((aw >> lane) & 1u) ? (((sw >> lane) & 1u) ? -1.f : 1.f) : 0.f;
  }
}

// compact state -> the reference's float32 state [E][num_obs][N] (rows by observable index, -1 = absent)
struct PecoExpand {
  const uint32_t *spins, *best_spins;
  const void* fields;
  const uint16_t* last_flip;
  const float *score, *best_score, *max_local;
  const float* table;        // table[k] = k-fold float32 accumulation of 1 / max_steps
  float* state;
  int64_t state_env_stride;  // floats between consecutive envs (num_obs * N, or (num_obs + N) * N inside an observation)
  int64_t num_envs;
  int n, np, words, step, binary_spins;
  int idx_imm, idx_tsf, idx_ept, idx_term, idx_greedy, idx_dscore, idx_dstate;
  float termination, inv_n;
  int recip_div, at_reset;
};

__global__ void __launch_bounds__(kPcWarps * 32) peco_compact_expand_state_kernel(PecoExpand p) {
  const int lane = threadIdx.x & 31;
  const int64_t env = (int64_t)blockIdx.x * kPcWarps + (threadIdx.x >> 5);
  if (env >= p.num_envs) return;
  const int n = p.n, W = p.words;
  const uint32_t sp = lane < W ? p.spins[env * W + lane] : 0u;
  const uint32_t bs = lane < W ? p.best_spins[env * W + lane] : 0u;
  const int dist = warp_sum_i(__popc(sp ^ bs));
  const float maxl = p.max_local[env];
  const int8_t* fl8 = static_cast<const int8_t*>(p.fields) + env * (int64_t)p.np;
  const int16_t* fl16 = static_cast<const int16_t*>(p.fields) + env * (int64_t)p.np;
  const uint16_t* lf = p.last_flip + env * (int64_t)p.np;
  float* st = p.state + env * p.state_env_stride;
  int nonpos = 0;
  for (int k = 0; k < W; ++k) {
    const uint32_t spw = __shfl_sync(kFull, sp, k);
    const int j = 32 * k + lane;
    if (j < n) {
      const bool up = (spw >> lane) & 1u;
      const int as = n <= kPcByteFields ? (int)fl8[j] : (int)fl16[j];
      const int f = up ? as : -as;
      nonpos += (int)(f <= 0);
      st[j] = p.binary_spins ? (up ? 0.f : 1.f) : (up ? 1.f : -1.f);       // BINARY basis: (1 - s) / 2
      if (p.idx_imm >= 0) st[p.idx_imm * n + j] = __fdiv_rn((float)f, maxl);
      if (p.idx_tsf >= 0) st[p.idx_tsf * n + j] = __ldg(p.table + (p.step - (int)lf[j]));
    }
  }
  nonpos = warp_sum_i(nonpos);
  const float score = p.score[env], best = p.best_score[env];
  const float g_greedy = __fsub_rn(1.f, p.recip_div ? __fmul_rn((float)nonpos, p.inv_n) : __fdiv_rn((float)nonpos, (float)n));
  // rows the reference's reset leaves at zero stay zero until the first step writes them
  const float g_dscore = p.at_reset ? 0.f : __fdiv_rn(fabsf(__fsub_rn(score, best)), maxl);
  const float g_dstate = p.at_reset ? 0.f : (float)dist;
  const float g_ept = __ldg(p.table + p.step);
  for (int j = lane; j < n; j += 32) {
    if (p.idx_ept >= 0) st[p.idx_ept * n + j] = g_ept;
    if (p.idx_term >= 0) st[p.idx_term * n + j] = p.termination;
    if (p.idx_greedy >= 0) st[p.idx_greedy * n + j] = g_greedy;
    if (p.idx_dscore >= 0) st[p.idx_dscore * n + j] = g_dscore;
    if (p.idx_dstate >= 0) st[p.idx_dstate * n + j] = g_dstate;
  }
}

// The same state as float16 rows: the inference env's `use_tensor_core=True` mode (inference_network_env.py:143-145,
// 212-236) keeps `state` in half precision.  torch evaluates every half operation in float32 and rounds the result to
// half, so each observable is the float32 value of the kernel above rounded once -- except the greedy-actions row, whose
// two operations (count / n, then 1 - x) round twice -- and the time observables, whose k-fold accumulation of
// 1 / max_steps happens in half (the table handed in holds those values).  Integers below 2048 are exact in half; the
// Python side admits this mode only for graphs whose sums stay below that.
__global__ void __launch_bounds__(kPcWarps * 32) peco_compact_expand_state_half_kernel(PecoExpand p, __half* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t env = (int64_t)blockIdx.x * kPcWarps + (threadIdx.x >> 5);
  if (env >= p.num_envs) return;
  const int n = p.n, W = p.words;
  const uint32_t sp = lane < W ? p.spins[env * W + lane] : 0u;
  const uint32_t bs = lane < W ? p.best_spins[env * W + lane] : 0u;
  const int dist = warp_sum_i(__popc(sp ^ bs));
  const float maxl = p.max_local[env];
  const int8_t* fl8 = static_cast<const int8_t*>(p.fields) + env * (int64_t)p.np;
  const int16_t* fl16 = static_cast<const int16_t*>(p.fields) + env * (int64_t)p.np;
  const uint16_t* lf = p.last_flip + env * (int64_t)p.np;
  __half* st = out + env * p.state_env_stride;
  int nonpos = 0;
  for (int k = 0; k < W; ++k) {
    const uint32_t spw = __shfl_sync(kFull, sp, k);
    const int j = 32 * k + lane;
    if (j < n) {
      const bool up = (spw >> lane) & 1u;
      const int as = n <= kPcByteFields ? (int)fl8[j] : (int)fl16[j];
      const int f = up ? as : -as;
      nonpos += (int)(f <= 0);
      st[j] = __float2half_rn(p.binary_spins ? (up ? 0.f : 1.f) : (up ? 1.f : -1.f));
      if (p.idx_imm >= 0) st[p.idx_imm * n + j] = __float2half_rn(__fdiv_rn((float)f, maxl));
      if (p.idx_tsf >= 0) st[p.idx_tsf * n + j] = __float2half_rn(__ldg(p.table + (p.step - (int)lf[j])));
    }
  }
  nonpos = warp_sum_i(nonpos);
  const float score = p.score[env], best = p.best_score[env];
  const __half frac = __float2half_rn(p.recip_div ? __fmul_rn((float)nonpos, p.inv_n) : __fdiv_rn((float)nonpos, (float)n));
  const __half g_greedy = __float2half_rn(__fsub_rn(1.f, __half2float(frac)));
  const __half g_dscore = __float2half_rn(p.at_reset ? 0.f : __fdiv_rn(fabsf(__fsub_rn(score, best)), maxl));
  const __half g_dstate = __float2half_rn(p.at_reset ? 0.f : (float)dist);
  const __half g_ept = __float2half_rn(__ldg(p.table + p.step));
  const __half g_term = __float2half_rn(p.termination);
  for (int j = lane; j < n; j += 32) {
    if (p.idx_ept >= 0) st[p.idx_ept * n + j] = g_ept;
    if (p.idx_term >= 0) st[p.idx_term * n + j] = g_term;
    if (p.idx_greedy >= 0) st[p.idx_greedy * n + j] = g_greedy;
    if (p.idx_dscore >= 0) st[p.idx_dscore * n + j] = g_dscore;
    if (p.idx_dstate >= 0) st[p.idx_dstate * n + j] = g_dstate;
  }
}

// ---- generators ----------------------------------------------------------------------------------------------
// ER (util_envs_PECO.py:40-52): adj[e][i][j] = adj[e][j][i] = [rand(E, n, n)[e][i][j] < p] for i < j.  The kernel
// keeps torch's decomposition of that ONE rand call: thread idx, round r owns elements idx + T (4 r + c), all four
// outputs of its Philox block are used; a set bit is OR-ed into both rows.
__global__ void __launch_bounds__(256) peco_gen_er_kernel(uint32_t* __restrict__ adj, int64_t numel, int n, int W, float p,
                                                          TorchRng r) {
  const uint32_t idx = blockIdx.x * 256 + threadIdx.x;
  const uint32_t round = blockIdx.y;
  const PhiloxKey key = philox_key(r);
  const uint64_t ctr = key.offset4 + round;
  const uint4 o = curand_Philox4x32_10(make_uint4((uint32_t)ctr, (uint32_t)(ctr >> 32), idx, 0u),
                                       make_uint2((uint32_t)key.seed, (uint32_t)(key.seed >> 32)));
  const uint32_t x[4] = {o.x, o.y, o.z, o.w};
  const int64_t nn = (int64_t)n * n;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int64_t li = (int64_t)idx + (int64_t)r.threads * (4 * (int64_t)round + c);
    if (li >= numel) break;
    if (!(torch_uniform_from_u32(x[c]) < p)) continue;
    const int64_t e = li / nn;
    const int rem = (int)(li - e * nn), i = rem / n, j = rem - i * n;
    if (i >= j) continue;                                  // strict upper triangle, mirrored
    atomicOr(adj + (e * n + i) * W + (j >> 5), 1u << (j & 31));
    atomicOr(adj + (e * n + j) * W + (i >> 5), 1u << (i & 31));
  }
}

// BA (util_envs_PECO.py:87-107), one warp per env.  Nodes 0..m form a clique INCLUDING their self loops (the
// reference's loop sets adj[:, i, :i+1] = 1: kept for parity); every later node t draws m distinct targets with
// torch.multinomial(prob = degree / degree.sum(), m, replacement=False), which torch evaluates as
// topk(prob / q, m) with q ~ Exp(1) drawn by ONE exponential_ call over [E, n] (ATen Distributions.cpp,
// multinomial fast path): element (e, j) of call number t - m - 1.  exponential_ on CUDA = -log(curand_uniform)
// with log replaced by -eps/2 at 1 (TransformationHelper.h); the division is IEEE.
// kMaxSlots: candidate nodes per lane held in registers (N <= 32 kMaxSlots)
template <int kMaxSlots>
__global__ void __launch_bounds__(kPcWarps * 32) peco_gen_ba_kernel(uint32_t* __restrict__ adj, int64_t num_envs, int n,
                                                                    int W, int m, TorchRng r) {
  const int lane = threadIdx.x & 31;
  const int64_t env = (int64_t)blockIdx.x * kPcWarps + (threadIdx.x >> 5);
  if (env >= num_envs) return;
  uint32_t* rows = adj + env * n * W;
  int deg[kMaxSlots];
  const int slots = (n + 31) >> 5;
#pragma unroll
  for (int k = 0; k < kMaxSlots; ++k) deg[k] = 0;
  for (int i = lane; i <= m && i < n; i += 32) {           // clique with self loops on nodes 0..m
    for (int j = 0; j <= m && j < n; ++j) atomicOr(rows + (int64_t)i * W + (j >> 5), 1u << (j & 31));
  }
  if (lane <= m && lane < n) deg[0] = min(m + 1, n);       // m + 1 <= 32 is checked by the host
  __syncwarp();
  const PhiloxKey key = philox_key(r);
  const uint2 pkey = make_uint2((uint32_t)key.seed, (uint32_t)(key.seed >> 32));
  for (int t = m + 1; t < n; ++t) {
    int total = 0;
#pragma unroll
    for (int k = 0; k < kMaxSlots; ++k)
      if (k < slots) total += deg[k];
    total = warp_sum_i(total);
    const uint64_t call = (uint64_t)(t - m - 1);
    float val[kMaxSlots];
#pragma unroll
    for (int k = 0; k < kMaxSlots; ++k) {
      val[k] = -1.f;
      if (k < slots) {
        const int j = 32 * k + lane;
        if (j < n) {
          const float prob = __fdiv_rn((float)deg[k], (float)total);
          TorchRng rr = r;
          rr.seed = key.seed, rr.offset4 = key.offset4;
          const uint64_t li = (uint64_t)env * n + j;       // element of the [E, n] exponential_ call
          const uint32_t tid = (uint32_t)(li % r.threads);
          const uint64_t kk = li / r.threads;
          const uint64_t ctr = key.offset4 + call * r.iters_per_call + (kk >> 2);
          const uint4 o = curand_Philox4x32_10(make_uint4((uint32_t)ctr, (uint32_t)(ctr >> 32), tid, 0u), pkey);
          const uint32_t c = (uint32_t)(kk & 3u);
          const float u = _curand_uniform(c == 0 ? o.x : c == 1 ? o.y : c == 2 ? o.z : o.w);
          const float lg = u >= 1.f - 1.1920928955078125e-07f / 2 ? -1.1920928955078125e-07f / 2 : logf(u);
          const float q = -lg;                             // (-1 / lambda) * log, lambda = 1
          val[k] = __fdiv_rn(prob, q);
        }
      }
    }
    for (int pick = 0; pick < m; ++pick) {                 // topk(m): largest first
      float best = -1.f;
      int arg = -1;
#pragma unroll
      for (int k = 0; k < kMaxSlots; ++k)
        if (k < slots && val[k] > best) best = val[k], arg = 32 * k + lane;
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) {
        const float ob = __shfl_xor_sync(kFull, best, off);
        const int oa = __shfl_xor_sync(kFull, arg, off);
        if (ob > best || (ob == best && oa >= 0 && (arg < 0 || oa < arg))) best = ob, arg = oa;
      }
      // arg: the chosen target (the same in every lane)
#pragma unroll
      for (int k = 0; k < kMaxSlots; ++k)
        if (k < slots && arg == 32 * k + lane) val[k] = -2.f, deg[k] += 1;
      if (lane == 0) {
        atomicOr(rows + (int64_t)t * W + (arg >> 5), 1u << (arg & 31));
        atomicOr(rows + (int64_t)arg * W + (t >> 5), 1u << (t & 31));
      }
    }
#pragma unroll
    for (int k = 0; k < kMaxSlots; ++k)
      if (k < slots && t == 32 * k + lane) deg[k] += m;
    __syncwarp();
  }
}

// Power-law cluster graphs (Holme-Kim), the third graph family of the reference: `nx.powerlaw_cluster_graph(n, m, p)`
// (rlsolver/methods/util_generate.py:75-93 with m = 4, p = 0.05).  networkx grows ONE graph on the host from Python's
// `random`; there is no device-side stream to reproduce, so this is the same growth process per env on a Philox stream
// of its own (subsequence = env): node t >= m draws m distinct targets with probability proportional to networkx's
// `repeated_nodes` multiplicities (one entry per initial node, one per edge end at a target, m per finished source),
// connects to the first, and for each further edge either -- with probability p, if the previous target has a
// neighbour that is not yet adjacent to t -- closes a triangle with a uniformly chosen such neighbour, or takes the
// next drawn target.  One thread per env, weights in local memory, the adjacency bit rows are the output (zeroed by
// the caller).  Statistical, not bitwise, agreement with networkx (tests: edge counts, degree tail, clustering).
constexpr int kPlMaxNodes = 128;
__global__ void __launch_bounds__(128) peco_gen_pl_kernel(uint32_t* __restrict__ adj, int64_t num_envs, int n, int W, int m,
                                                          float p_tri, uint64_t seed, uint64_t offset) {
  const int64_t env = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= num_envs) return;
  curandStatePhilox4_32_10_t rs;
  curand_init(seed, (unsigned long long)env, offset, &rs);
  uint32_t* rows = adj + env * (int64_t)n * W;
  uint16_t wt[kPlMaxNodes];
  for (int v = 0; v < n; ++v) wt[v] = v < m ? 1 : 0;
  int total = m;
  auto connect = [&](int u, int v) {               // networkx add_edge + repeated_nodes.append(v)
    rows[u * W + (v >> 5)] |= 1u << (v & 31);
    rows[v * W + (u >> 5)] |= 1u << (u & 31);
    ++wt[v], ++total;
  };
  for (int src = m; src < n; ++src) {
    int targets[32];
    int nt = 0;
    while (nt < m) {                                // _random_subset: draw until m distinct
      int r = min((int)(curand_uniform(&rs) * (float)total), total - 1);
      int v = 0;
      while (r >= (int)wt[v]) r -= wt[v], ++v;
      bool seen = false;
      for (int k = 0; k < nt; ++k) seen |= targets[k] == v;
      if (!seen) targets[nt++] = v;
    }
    // networkx pops its targets from a Python set: ascending slot order of an open-addressing table (8 slots up to
    // four small ints, linear probing) filled in draw order.  The order matters statistically -- the first target's
    // neighbourhood feeds the triangle steps, and a later target that a triangle step already connected is a repeated
    // edge (draw order gives 0.16 % fewer edges at p = 0.6) -- so it is reproduced (exactly for m <= 4, the reference's
    // setting; with the table size of the final set beyond that).
    {
      const int slots = m <= 4 ? 8 : (m <= 18 ? 32 : 128);
      int table[128];
      for (int k = 0; k < slots; ++k) table[k] = -1;
      for (int k = 0; k < m; ++k) {
        int i = targets[k] & (slots - 1);
        while (table[i] >= 0) i = (i + 1) & (slots - 1);
        table[i] = targets[k];
      }
      int k = 0;
      for (int i = 0; i < slots; ++i)
        if (table[i] >= 0) targets[k++] = table[i];
    }
    int next = 0;
    int target = targets[next++];
    connect(src, target);
    for (int count = 1; count < m; ++count) {
      if (curand_uniform(&rs) < p_tri) {
        int cand = 0;                               // neighbours of `target` not adjacent to src (and not src)
        for (int w = 0; w < W; ++w) {
          uint32_t c = rows[target * W + w] & ~rows[src * W + w];
          if ((src >> 5) == w) c &= ~(1u << (src & 31));
          cand += __popc(c);
        }
        if (cand > 0) {
          int pick = min((int)(curand_uniform(&rs) * (float)cand), cand - 1);
          int nbr = -1;
          for (int w = 0; w < W && nbr < 0; ++w) {
            uint32_t c = rows[target * W + w] & ~rows[src * W + w];
            if ((src >> 5) == w) c &= ~(1u << (src & 31));
            const int here = __popc(c);
            if (pick < here) {
              for (int k = 0; k < pick; ++k) c &= c - 1;
              nbr = 32 * w + __ffs(c) - 1;
            } else {
              pick -= here;
            }
          }
          connect(src, nbr);
          continue;
        }
      }
      target = targets[next++];
      connect(src, target);
    }
    wt[src] += (uint16_t)m, total += m;             // repeated_nodes.extend([source] * m)
  }
}

}  // namespace rlsb

extern "C" {

static int pc_shape_ok(int64_t num_envs, int32_t n, const char* what) {
  using namespace rlsb;
  RLSB_REQUIRE(num_envs >= 0 && n > 0 && n <= 1024, RLSB_ERR_INVALID, "%s: bad shape (1 <= n_spins <= 1024)", what);
  return RLSB_OK;
}

int rlsb_peco_compact_step(const uint32_t* adj, const uint32_t* sgn, int64_t sgn_stride, uint32_t* spins,
                           void* fields, uint16_t* last_flip, uint32_t* best_spins, float* score, float* best_score,
                           const float* max_local, float* reward, const int64_t* action, uint64_t* hset,
                           int32_t hcap, uint64_t* hkey, const uint64_t* zobrist, int32_t* bad_actions,
                           int64_t num_envs, int32_t num_spins, int32_t step, int32_t reward_signal,
                           int32_t norm_rewards, int32_t use_stag, float stag, int32_t use_basin, float basin,
                           int32_t scalar_div_as_cuda, uint8_t* done_out, int32_t last_step, int32_t irreversible,
                           void* stream) {
  using namespace rlsb;
  if (int rc = pc_shape_ok(num_envs, num_spins, "peco_compact_step")) return rc;
  if (num_envs == 0) return RLSB_OK;
  RLSB_REQUIRE(adj && spins && fields && last_flip && best_spins && score && best_score && max_local && reward &&
                   action && bad_actions,
               RLSB_ERR_INVALID, "peco_compact_step: null pointer");
  RLSB_REQUIRE(reward_signal == 1 || reward_signal == 2 || reward_signal == 4, RLSB_ERR_UNSUPPORTED,
               "peco_compact_step: reward signal %d (DENSE = 1, BLS = 2, CUSTOM_BLS = 4 are implemented)", reward_signal);
  RLSB_REQUIRE(!(use_stag || use_basin) || (hset && hkey && zobrist && hcap >= 32 && (hcap & (hcap - 1)) == 0),
               RLSB_ERR_INVALID, "peco_compact_step: stag / basin rewards need the visited set (power-of-two capacity >= 32)");
  RLSB_REQUIRE(step >= 1 && step < 65536, RLSB_ERR_INVALID, "peco_compact_step: step out of range");
  PecoC p{};
  p.adj = adj, p.sgn = sgn, p.sgn_stride = sgn_stride, p.spins = spins, p.fields = fields, p.last_flip = last_flip;
  p.best_spins = best_spins, p.score = score, p.best_score = best_score, p.max_local = max_local, p.reward = reward;
  p.action = action, p.hset = reinterpret_cast<unsigned long long*>(hset), p.hcap = hcap;
  p.hkey = reinterpret_cast<unsigned long long*>(hkey), p.zobrist = reinterpret_cast<const unsigned long long*>(zobrist);
  p.bad_actions = bad_actions, p.num_envs = num_envs, p.n = num_spins, p.np = (num_spins + 31) / 32 * 32;
  p.words = (num_spins + 31) / 32, p.step = step, p.reward_signal = reward_signal, p.norm_rewards = norm_rewards;
  p.use_stag = use_stag, p.use_basin = use_basin, p.stag = stag, p.basin = basin;
  p.recip_div = scalar_div_as_cuda, p.inv_n = 1.0f / (float)num_spins;
  p.done = done_out, p.last_step = last_step != 0, p.irreversible = irreversible != 0;
  const unsigned grid = (unsigned)((num_envs + kPcWarps - 1) / kPcWarps);
  auto st = static_cast<cudaStream_t>(stream);
  if (p.n <= kPcLpe * kPcNpl && p.hcap % kPcLpe == 0 && !(debug_flags() & RLSB_DEBUG_PECO_WARP_PER_ENV)) {
    const int64_t per_block = (int64_t)kPcWarps * (32 / kPcLpe);
    peco_compact_step8_kernel<<<(unsigned)((num_envs + per_block - 1) / per_block), kPcWarps * 32, 0, st>>>(p);
  } else if (p.n <= kPcByteFields) peco_compact_step_kernel<4, int8_t><<<grid, kPcWarps * 32, 0, st>>>(p);
  else if (p.words <= 8) peco_compact_step_kernel<8, int16_t><<<grid, kPcWarps * 32, 0, st>>>(p);
  else peco_compact_step_kernel<0, int16_t><<<grid, kPcWarps * 32, 0, st>>>(p);
  RLSB_LAUNCH_OK();
  return RLSB_OK;
}

int rlsb_peco_compact_fields(const uint32_t* adj, const uint32_t* sgn, int64_t sgn_stride, const uint32_t* spins,
                             int64_t num_envs, int32_t num_spins, void* fields, float* cut, float* max_local,
                             int32_t* empty_graphs, void* stream) {
  using namespace rlsb;
  if (int rc = pc_shape_ok(num_envs, num_spins, "peco_compact_fields")) return rc;
  if (num_envs == 0) return RLSB_OK;
  RLSB_REQUIRE(adj && spins, RLSB_ERR_INVALID, "peco_compact_fields: null pointer");
  peco_compact_fields_kernel<<<(unsigned)((num_envs + kPcWarps - 1) / kPcWarps), kPcWarps * 32, 0,
                               static_cast<cudaStream_t>(stream)>>>(adj, sgn, sgn_stride, spins, num_envs, num_spins,
                                                                    (num_spins + 31) / 32 * 32, (num_spins + 31) / 32,
                                                                    fields, cut, max_local, empty_graphs);
  RLSB_LAUNCH_OK();
  return RLSB_OK;
}

int rlsb_peco_compact_from_dense(const float* matrix, int64_t num_envs, int32_t num_spins, uint32_t* adj, uint32_t* sgn,
                                 int32_t* bad_entries, void* stream) {
  using namespace rlsb;
  if (int rc = pc_shape_ok(num_envs, num_spins, "peco_compact_from_dense")) return rc;
  if (num_envs == 0) return RLSB_OK;
  RLSB_REQUIRE(matrix && adj && sgn && bad_entries, RLSB_ERR_INVALID, "peco_compact_from_dense: null pointer");
  const int64_t rows = num_envs * num_spins;
  peco_compact_from_dense_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      matrix, num_envs, num_spins, (num_spins + 31) / 32, adj, sgn, bad_entries);
  RLSB_LAUNCH_OK();
  return RLSB_OK;
}

int rlsb_peco_compact_expand_matrix(const uint32_t* adj, const uint32_t* sgn, int64_t sgn_stride, int64_t num_envs,
                                    int32_t num_spins, float* out, int64_t out_env_stride, void* stream) {
  using namespace rlsb;
  if (int rc = pc_shape_ok(num_envs, num_spins, "peco_compact_expand_matrix")) return rc;
  if (num_envs == 0) return RLSB_OK;
  RLSB_REQUIRE(adj && out && out_env_stride >= (int64_t)num_spins * num_spins, RLSB_ERR_INVALID,
               "peco_compact_expand_matrix: bad argument");
  const int64_t rows = num_envs * num_spins;
  peco_compact_expand_matrix_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      adj, sgn, sgn_stride, num_envs, num_spins, (num_spins + 31) / 32, out, out_env_stride);
  RLSB_LAUNCH_OK();
  return RLSB_OK;
}

int rlsb_peco_compact_expand_state(const uint32_t* spins, const uint32_t* best_spins, const void* fields,
                                   const uint16_t* last_flip, const float* score, const float* best_score,
                                   const float* max_local, const float* table, float* state, int64_t state_env_stride,
                                   int64_t num_envs, int32_t num_spins, int32_t num_obs, const int32_t* h_obs_rows,
                                   int32_t step, int32_t binary_spins, float termination, int32_t scalar_div_as_cuda,
                                   int32_t at_reset, void* stream) {
  using namespace rlsb;
  if (int rc = pc_shape_ok(num_envs, num_spins, "peco_compact_expand_state")) return rc;
  if (num_envs == 0) return RLSB_OK;
  RLSB_REQUIRE(spins && best_spins && fields && last_flip && score && best_score && max_local && table && state &&
                   h_obs_rows && num_obs >= 1 && state_env_stride >= (int64_t)num_obs * num_spins,
               RLSB_ERR_INVALID, "peco_compact_expand_state: bad argument");
  PecoExpand p{};
  p.spins = spins, p.best_spins = best_spins, p.fields = fields, p.last_flip = last_flip, p.score = score;
  p.best_score = best_score, p.max_local = max_local, p.table = table, p.state = state;
  p.state_env_stride = state_env_stride, p.num_envs = num_envs, p.n = num_spins, p.np = (num_spins + 31) / 32 * 32;
  p.words = (num_spins + 31) / 32, p.step = step, p.binary_spins = binary_spins;
  for (int k = 0; k < 7; ++k)
    RLSB_REQUIRE(h_obs_rows[k] < num_obs, RLSB_ERR_INVALID, "peco_compact_expand_state: observable row %d out of range",
                 h_obs_rows[k]);
  p.idx_imm = h_obs_rows[0], p.idx_tsf = h_obs_rows[1], p.idx_ept = h_obs_rows[2], p.idx_term = h_obs_rows[3];
  p.idx_greedy = h_obs_rows[4], p.idx_dscore = h_obs_rows[5], p.idx_dstate = h_obs_rows[6];
  p.termination = termination, p.inv_n = 1.0f / (float)num_spins, p.recip_div = scalar_div_as_cuda, p.at_reset = at_reset;
  peco_compact_expand_state_kernel<<<(unsigned)((num_envs + kPcWarps - 1) / kPcWarps), kPcWarps * 32, 0,
                                     static_cast<cudaStream_t>(stream)>>>(p);
  RLSB_LAUNCH_OK();
  return RLSB_OK;
}

int rlsb_peco_compact_expand_state_half(const uint32_t* spins, const uint32_t* best_spins, const void* fields,
                                        const uint16_t* last_flip, const float* score, const float* best_score,
                                        const float* max_local, const float* table, void* state, int64_t state_env_stride,
                                        int64_t num_envs, int32_t num_spins, int32_t num_obs, const int32_t* h_obs_rows,
                                        int32_t step, int32_t binary_spins, float termination, int32_t scalar_div_as_cuda,
                                        int32_t at_reset, void* stream) {
  using namespace rlsb;
  if (int rc = pc_shape_ok(num_envs, num_spins, "peco_compact_expand_state_half")) return rc;
  if (num_envs == 0) return RLSB_OK;
  RLSB_REQUIRE(spins && best_spins && fields && last_flip && score && best_score && max_local && table && state &&
                   h_obs_rows && num_obs >= 1 && state_env_stride >= (int64_t)num_obs * num_spins,
               RLSB_ERR_INVALID, "peco_compact_expand_state_half: bad argument");
  PecoExpand p{};
  p.spins = spins, p.best_spins = best_spins, p.fields = fields, p.last_flip = last_flip, p.score = score;
  p.best_score = best_score, p.max_local = max_local, p.table = table, p.state = nullptr;
  p.state_env_stride = state_env_stride, p.num_envs = num_envs, p.n = num_spins, p.np = (num_spins + 31) / 32 * 32;
  p.words = (num_spins + 31) / 32, p.step = step, p.binary_spins = binary_spins;
  for (int k = 0; k < 7; ++k)
    RLSB_REQUIRE(h_obs_rows[k] < num_obs, RLSB_ERR_INVALID,
                 "peco_compact_expand_state_half: observable row %d out of range", h_obs_rows[k]);
  p.idx_imm = h_obs_rows[0], p.idx_tsf = h_obs_rows[1], p.idx_ept = h_obs_rows[2], p.idx_term = h_obs_rows[3];
  p.idx_greedy = h_obs_rows[4], p.idx_dscore = h_obs_rows[5], p.idx_dstate = h_obs_rows[6];
  p.termination = termination, p.inv_n = 1.0f / (float)num_spins, p.recip_div = scalar_div_as_cuda, p.at_reset = at_reset;
  peco_compact_expand_state_half_kernel<<<(unsigned)((num_envs + kPcWarps - 1) / kPcWarps), kPcWarps * 32, 0,
                                          static_cast<cudaStream_t>(stream)>>>(p, static_cast<__half*>(state));
  RLSB_LAUNCH_OK();
  return RLSB_OK;
}

int rlsb_peco_gen_er(uint32_t* adj, int64_t num_envs, int32_t num_spins, float p_connection, uint64_t seed,
                     uint64_t offset, uint32_t rng_threads, uint32_t rng_iters, void* stream) {
  using namespace rlsb;
  if (int rc = pc_shape_ok(num_envs, num_spins, "peco_gen_er")) return rc;
  if (num_envs == 0) return RLSB_OK;
  const int64_t numel = num_envs * (int64_t)num_spins * num_spins;
  RLSB_REQUIRE(numel < (int64_t(1) << 31), RLSB_ERR_UNSUPPORTED,
               "peco_gen_er: %lld elements in one rand call (torch splits calls of 2^31 and more: generate in chunks)",
               (long long)numel);
  RLSB_REQUIRE(adj && rng_threads > 0 && rng_threads % 256 == 0 && rng_iters > 0 && rng_iters <= 65535 && offset % 4 == 0 &&
                   (int64_t)rng_threads * 4 * rng_iters >= numel,
               RLSB_ERR_INVALID, "peco_gen_er: bad argument");
  const int W = (num_spins + 31) / 32;
  auto st = static_cast<cudaStream_t>(stream);
  RLSB_CUDA_OK(cudaMemsetAsync(adj, 0, (size_t)num_envs * num_spins * W * sizeof(uint32_t), st));
  TorchRng r{seed, offset / 4, rng_threads, rng_iters, nullptr};
  peco_gen_er_kernel<<<dim3(rng_threads / 256, rng_iters), 256, 0, st>>>(adj, numel, num_spins, W, p_connection, r);
  RLSB_LAUNCH_OK();
  return RLSB_OK;
}

int rlsb_peco_gen_ba(uint32_t* adj, int64_t num_envs, int32_t num_spins, int32_t m_insertion_edges, uint64_t seed,
                     uint64_t offset, uint32_t rng_threads, uint32_t rng_iters, void* stream) {
  using namespace rlsb;
  if (int rc = pc_shape_ok(num_envs, num_spins, "peco_gen_ba")) return rc;
  if (num_envs == 0) return RLSB_OK;
  RLSB_REQUIRE(m_insertion_edges >= 1 && m_insertion_edges < 32 && m_insertion_edges + 1 <= num_spins, RLSB_ERR_INVALID,
               "peco_gen_ba: 1 <= m < 32 and m + 1 <= n_spins");
  RLSB_REQUIRE(num_envs * (int64_t)num_spins < (int64_t(1) << 31), RLSB_ERR_UNSUPPORTED,
               "peco_gen_ba: more than 2^31 elements per exponential_ call (generate in chunks)");
  RLSB_REQUIRE(adj && rng_threads > 0 && rng_threads % 256 == 0 && rng_iters > 0 && offset % 4 == 0, RLSB_ERR_INVALID,
               "peco_gen_ba: bad argument");
  const int W = (num_spins + 31) / 32;
  auto st = static_cast<cudaStream_t>(stream);
  RLSB_CUDA_OK(cudaMemsetAsync(adj, 0, (size_t)num_envs * num_spins * W * sizeof(uint32_t), st));
  TorchRng r{seed, offset / 4, rng_threads, rng_iters, nullptr};
  const unsigned grid = (unsigned)((num_envs + kPcWarps - 1) / kPcWarps);
  if (num_spins <= 128) peco_gen_ba_kernel<4><<<grid, kPcWarps * 32, 0, st>>>(adj, num_envs, num_spins, W, m_insertion_edges, r);
  else if (num_spins <= 256) peco_gen_ba_kernel<8><<<grid, kPcWarps * 32, 0, st>>>(adj, num_envs, num_spins, W, m_insertion_edges, r);
  else peco_gen_ba_kernel<32><<<grid, kPcWarps * 32, 0, st>>>(adj, num_envs, num_spins, W, m_insertion_edges, r);
  RLSB_LAUNCH_OK();
  return RLSB_OK;
}

int rlsb_peco_gen_pl(uint32_t* adj, int64_t num_envs, int32_t num_spins, int32_t m_insertion_edges, float p_triangle,
                     uint64_t seed, uint64_t offset, void* stream) {
  using namespace rlsb;
  if (int rc = pc_shape_ok(num_envs, num_spins, "peco_gen_pl")) return rc;
  if (num_envs == 0) return RLSB_OK;
  RLSB_REQUIRE(num_spins <= kPlMaxNodes, RLSB_ERR_UNSUPPORTED, "peco_gen_pl: at most %d spins", kPlMaxNodes);
  RLSB_REQUIRE(m_insertion_edges >= 1 && m_insertion_edges < 32 && m_insertion_edges < num_spins, RLSB_ERR_INVALID,
               "peco_gen_pl: 1 <= m < 32 and m < n_spins");
  RLSB_REQUIRE(adj && p_triangle >= 0.f && p_triangle <= 1.f, RLSB_ERR_INVALID, "peco_gen_pl: bad argument");
  const int W = (num_spins + 31) / 32;
  auto st = static_cast<cudaStream_t>(stream);
  RLSB_CUDA_OK(cudaMemsetAsync(adj, 0, (size_t)num_envs * num_spins * W * sizeof(uint32_t), st));
  peco_gen_pl_kernel<<<(unsigned)((num_envs + 127) / 128), 128, 0, st>>>(adj, num_envs, num_spins, W, m_insertion_edges,
                                                                       p_triangle, seed, offset);
  RLSB_LAUNCH_OK();
  return RLSB_OK;
}

}  // extern "C"
