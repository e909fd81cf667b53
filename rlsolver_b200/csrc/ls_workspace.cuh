// Pieces shared by the local-search translation units (local_search.cu, noise_masks.cu): the
// caller-owned workspace layout of rlsb_ls_begin / ls_run / ls_noise_masks and the one
// floating-point expression of the path, spin_rand (env_L2A.py:94-95, LocalSearch.py:66-67).
#pragma once
#include "common.cuh"
#include "philox.cuh"

namespace rlsb {

// float(k) for |k| < 2^22 without the conversion pipe: 0x4B400000 is 12582912.0f (1.5 * 2^23)
constexpr int kMagicI = 0x4B400000;
constexpr float kMagicF = 12582912.0f;

#ifdef __CUDACC__
// degm = listed degree + kMagicI, negmult = -mult.  Evaluated like the reference's two torch
// kernels: one IEEE multiply, one add, no FMA contraction.
__device__ __forceinline__ float spin_rand(int degm, int negmult, int cross, float noise, float rd_std) {
  const float wsf = __fadd_rn(__int_as_float(cross * negmult + degm), -kMagicF);   // exact float(deg - mult*cross)
  return __fadd_rn(wsf, __fmul_rn(noise, rd_std));
}
#endif

// 0: counters of 6 bit planes, 1: 8 planes (uint8 cross counts), 2: 12 planes (uint16 cross counts)
inline int degree_class(const GraphDev& g) {
  const int d = g.max_listed_deg > g.max_full_deg ? g.max_listed_deg : g.max_full_deg;
  return d <= 63 ? 0 : d <= 255 ? 1 : 2;
}

// Early-out section: 2 bytes (one per Box-Muller pair) per generator thread and round.  T * 4 * iters < numel + 4 T and T <= 2048 threads
// on each of the 148 SMs (ATen calc_execution_policy on a B200).
inline int64_t ls_bound_bytes(int64_t numel) { return (numel + 4 * (int64_t)kNumSMs * 2048) / 2 + 256; }

// Streaming mask generator <-> tile kernel hand-shake (rlsb_ls_fused_search): per group of draws two counters,
// ctl[2g] = units handed out, ctl[2g + 1] = units finished.
constexpr int kLsMaxFusedDraws = 1024;
constexpr int kLsCtlWords = (kLsMaxFusedDraws + 2) * 2 + 4;
constexpr size_t kLsCtlBytes = (size_t)kLsCtlWords * sizeof(uint32_t);
// diagnostics at the end of the counter block: generator blocks that started, tile CTAs whose wait ran out
constexpr int kLsCtlStarted = kLsCtlWords - 2, kLsCtlStalled = kLsCtlWords - 1;
// a tile CTA gives up waiting for a group of draws after this many nanoseconds (the call then reports
// RLSB_ERR_CUDA through rlsb_ls_fused_status instead of hanging the device)
constexpr unsigned long long kLsStallNs = 4000000000ull;

// workspace carving (all sections 256-byte aligned)
struct LsWorkspace {
  uint32_t* packed;
  void* cross;
  int32_t *col_min, *col_max, *degm;
  float *rd_std, *thresh;
  uint32_t* nd;
  uint8_t* cross_rows;   // [E][Np] uint8 row-major copy of the cross counts (mask generator); null for uint16 counts
  uint8_t* bound;        // early-out bytes of the mask generator (noise_masks.cu: two planes of one byte per thread-round)
  uint32_t* ctl;         // unit counters of the streaming generator (kLsCtlBytes)
  size_t bytes;
};

inline LsWorkspace carve(const GraphDev& g, int64_t num_envs, void* base) {
  const int64_t tiles = (num_envs + kTileEnvs - 1) / kTileEnvs;
  const size_t cross_elt = degree_class(g) == 2 ? 2 : 1;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* p = base ? static_cast<char*>(base) + off : nullptr;
    off += (bytes + 255) / 256 * 256;
    return p;
  };
  LsWorkspace w;
  w.packed = reinterpret_cast<uint32_t*>(take((size_t)tiles * g.np * 4));
  w.cross = take((size_t)tiles * kTileEnvs * g.np * cross_elt);
  w.col_min = reinterpret_cast<int32_t*>(take((size_t)g.np * 4));
  w.col_max = reinterpret_cast<int32_t*>(take((size_t)g.np * 4));
  w.degm = reinterpret_cast<int32_t*>(take((size_t)g.np * 4));
  w.rd_std = reinterpret_cast<float*>(take((size_t)g.np * 4));
  w.thresh = reinterpret_cast<float*>(take((size_t)num_envs * 4));
  w.nd = reinterpret_cast<uint32_t*>(take((size_t)g.np * 8));
  w.ctl = reinterpret_cast<uint32_t*>(take(kLsCtlBytes));
  w.cross_rows = nullptr;
  w.bound = nullptr;
  if (cross_elt == 1) {
    w.cross_rows = reinterpret_cast<uint8_t*>(take((size_t)tiles * kTileEnvs * g.np));
    w.bound = reinterpret_cast<uint8_t*>(take((size_t)ls_bound_bytes(num_envs * (int64_t)g.n)));
  }
  w.bytes = off + 256;
  return w;
}

// words of one draw's flip-mask bit array (bit e*N + n), padded so that a reader may always fetch
// the word after the one holding its first bit; multiple of 4 words
inline int64_t ls_mask_words(int64_t num_envs, int n) {
  return (((num_envs * n + 31) / 32 + 2) + 3) / 4 * 4;
}

// ---- mask generator (noise_masks.cu), shared with the fused search entry point (local_search.cu)
struct MaskArgs {
  const uint8_t* cross_rows;   // [E][Np] cross counts, row-major (ls_begin)
  const float* rd_std;         // [Np]
  const int32_t* degm;         // [Np] listed degree + kMagicI
  const float* thresh;         // [E]
  uint32_t* masks;             // [draws][mask_words]
  int64_t mask_words;
  uint32_t numel, n, np;
  uint32_t step_e, step_n;     // T / N, T % N
  uint32_t div_m, div_s;       // floor(l / N) = (l * div_m) >> (31 + div_s) for every l < 2^31
  int negmult;
};
struct MaskPlan {
  MaskArgs a;
  TorchRng r;
  uint8_t* bound;
  uint32_t* ctl;
  int num_draws;
};
int mask_plan(const GraphDev& g, const char* what, int64_t num_envs, int ws_mult, uint64_t seed, uint64_t offset,
              const uint64_t* rng_dev, int rng_threads, int rng_iters, int num_draws, uint32_t* masks, void* workspace,
              MaskPlan* plan);
int mask_prepare(const MaskPlan& p, bool write_bound, bool zero_ctl, cudaStream_t st);
int mask_stream_preload(const MaskPlan& p, cudaStream_t st);
int mask_stream_launch(const MaskPlan& p, cudaStream_t st);
constexpr int kGenGroup = 2;   // draws the streaming generator finishes together (= what a tile iteration waits for)

// side stream + fork / join events of a graph handle (graph_store.cu): the generator runs there
struct GraphSide {
  cudaStream_t stream;
  cudaEvent_t fork, join;
};
int graph_side(const rlsb_graph_t* g, GraphSide* out);

}  // namespace rlsb
