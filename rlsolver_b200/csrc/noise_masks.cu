// Flip masks of the noisy multi-flip iterations without the noise tensors.
//
// The reference draws `randn [E, N] float32` once per iteration (env_L2A.py:99, LocalSearch.py:66)
// and only ever looks at one bit of each number: `spin_rand > thresh` (env_L2A.py:100-101).  With
// explicit noise tensors (rlsb_ls_run) every draw costs torch's generator kernel, a 4-byte write
// and a 4-byte read per (env, node).  torch's CUDA generator is counter based (philox.cuh), so
// this kernel recomputes exactly the numbers those calls would have produced -- same Philox
// counters, curand's own Box-Muller (`curand_normal4`, the function torch's normal_ kernel calls,
// ATen/native/cuda/DistributionTemplates.h normal_and_transform) -- evaluates the reference's
// expression on them in registers and emits the flip bit.  Per draw the output is a flat bit
// array (bit e*N + n), E*N/8 bytes instead of 4*E*N; rlsb_ls_run_masks consumes it.
//
// Work decomposition = torch's own: thread `idx` of call `k`, round `j` owns the Philox block
// (counter offset/4 + k*iters + j, subsequence idx) whose four normals belong to the elements
// idx + T*(4j + {0,1,2,3}); consecutive lanes hold consecutive elements, so one BALLOT is one
// 32-bit word of the bit array (noise_mask_kernel, the plain form: every normal is evaluated).
//
// Early-out form (the default).  Only ~num_spin of N numbers per row pass the threshold, and
// whether one CAN pass is decided by the first uint32 of its Box-Muller pair alone: the two
// normals of a pair are s*sin(v), s*cos(v) with s = sqrt(-2 ln u), u = x * 2^-32 + 2^-33, so
// |normal| <= s and s is decreasing in x.  mask_bound_kernel turns the per-element requirement
// "noise > (thresh - ws) / rd_std" into one byte c per element, once per call (ws, rd_std and
// thresh are the same for all draws): the element can flip only if (x >> 24) <= c.  The
// generator then costs one Philox block + four byte compares per four elements; the few pairs
// that pass go to a shared-memory queue and are evaluated exactly -- curand's Box-Muller, the
// reference's expression, the exact compare -- by the first threads of the block, which OR the
// flip bits into the (zeroed) arrays.  The bound is conservative (margins below), the decision
// is always made by the exact expression, so the masks are identical to the plain form's.
#include <stdlib.h>

#include "ls_workspace.cuh"
#include "philox.cuh"

namespace rlsb {


// exact floor(l / n) for l < 2^31 (Granlund-Montgomery, 31-bit dividends: the magic number fits 32 bits)
__device__ __forceinline__ uint32_t div_n(const MaskArgs& a, uint32_t l) {
  return (uint32_t)(((uint64_t)l * a.div_m) >> (31u + a.div_s));
}

__global__ void __launch_bounds__(256) noise_mask_kernel(MaskArgs a, TorchRng r) {
  const uint32_t idx = blockIdx.x * 256 + threadIdx.x;
  const uint32_t j = blockIdx.y, k = blockIdx.z;
  const PhiloxKey key = philox_key(r);
  const uint64_t ctr = key.offset4 + (uint64_t)k * r.iters_per_call + j;
  const uint4 o = curand_Philox4x32_10(make_uint4((uint32_t)ctr, (uint32_t)(ctr >> 32), idx, 0u),
                                       make_uint2((uint32_t)key.seed, (uint32_t)(key.seed >> 32)));
  // curand_normal4: (x, y) -> the first two normals, (z, w) -> the other two.  torch's transform
  // `rand * std + mean` with std = 1, mean = 0 leaves the value as it is.
  const float2 n01 = _curand_box_muller(o.x, o.y), n23 = _curand_box_muller(o.z, o.w);
  const float z[4] = {n01.x, n01.y, n23.x, n23.y};
  uint32_t li = idx + r.threads * 4u * j;
  uint32_t e = div_n(a, li), nn = li - e * a.n;
  uint32_t* out = a.masks + (int64_t)k * a.mask_words;
#pragma unroll
  for (int ii = 0; ii < 4; ++ii) {
    bool bit = false;
    if (li < a.numel) {
      const int cross = (int)__ldg(a.cross_rows + (uint64_t)e * a.np + nn);
      const float sr = spin_rand(__ldg(a.degm + nn), a.negmult, cross, z[ii], __ldg(a.rd_std + nn));
      bit = sr > __ldg(a.thresh + e);
    }
    const uint32_t word = __ballot_sync(kFull, bit);
    if ((threadIdx.x & 31) == 0 && li < a.numel) out[li >> 5] = word;
    li += r.threads, e += a.step_e, nn += a.step_n;
    if (nn >= a.n) nn -= a.n, ++e;
  }
}


// ---- early-out bound.  flip  <=>  fl(ws + fl(z * rd)) > th.  Rounding is monotone and th is a float, so a
// flip implies ws + fl(z*rd) > th in real arithmetic, hence z * rd > t' for any float t' <= th - ws, hence
// (t' > 0) z > t' / rd.  With |z| <= s = sqrt(-2 ln u): a flip needs s > zmin = t' / rd, i.e.
// u < exp(-zmin^2 / 2).  (x >> 24) > c gives u >= (c + 1) / 256, so c + 1 >= 256 * exp(-zmin^2 / 2) makes
// "(x >> 24) > c  =>  no flip" true.  Margins: t' = fl(th - ws) * (1 - 2^-20) absorbs the rounding of the
// difference, zmin * 0.999 the division, logf / sqrtf (<= 2 ulp) and the intrinsic's error, * 1.001 __expf.
// t' <= 0 (or rd = 0 with ws > th, or NaN): c = 255 = always evaluated exactly.
__device__ __forceinline__ uint32_t bound_byte(const MaskArgs& a, float th, int cross, int degm, float rd) {
  const float wsf = __fadd_rn(__int_as_float(cross * a.negmult + degm), -kMagicF);
  const float t = __fadd_rn(th, -wsf);
  if (!(t > 0.f)) return 255u;
  const float zmin = __fdividef(t * 0.999999f, rd) * 0.999f;   // rd = 0 -> +inf -> c = 0; the approximate division's 2 ulp sit inside the 0.999
  const float p = 256.f * 1.001f * __expf(-0.5f * zmin * zmin);
  return p >= 255.f ? 255u : (uint32_t)p;
}

// One byte per Box-Muller PAIR (the larger of its two elements' bytes; 0 for slots past the tensor -- the
// exact path re-checks the range), in two planes [pair of the block][round j][thread idx] so that the
// generator's thread (idx, j) finds them with two coalesced byte loads.  blockIdx = (idx group, j, pair).
// VEC4 (N % 4 == 0): a thread covers four consecutive idx = four consecutive nodes of one env row for each
// of the pair's two elements (they are T apart), all loads and the store are words.
__device__ __forceinline__ uint32_t bound_of(const MaskArgs& a, uint32_t l) {
  if (l >= a.numel) return 0u;
  const uint32_t e = div_n(a, l), nn = l - e * a.n;
  return bound_byte(a, __ldg(a.thresh + e), (int)__ldg(a.cross_rows + (uint64_t)e * a.np + nn), __ldg(a.degm + nn),
                    __ldg(a.rd_std + nn));
}

__device__ __forceinline__ uint32_t bound_of4(const MaskArgs& a, uint32_t l) {
  if (l >= a.numel) return 0u;                              // numel % 4 == 0: a group is inside or outside
  const uint32_t e = div_n(a, l), nn = l - e * a.n;         // nn % 4 == 0
  const float th = __ldg(a.thresh + e);
  const uint32_t cr = __ldg(reinterpret_cast<const uint32_t*>(a.cross_rows + (uint64_t)e * a.np + nn));
  const int4 dm = __ldg(reinterpret_cast<const int4*>(a.degm + nn));
  const float4 rd = __ldg(reinterpret_cast<const float4*>(a.rd_std + nn));
  return bound_byte(a, th, (int)(cr & 0xffu), dm.x, rd.x) | (bound_byte(a, th, (int)((cr >> 8) & 0xffu), dm.y, rd.y) << 8) |
         (bound_byte(a, th, (int)((cr >> 16) & 0xffu), dm.z, rd.z) << 16) |
         (bound_byte(a, th, (int)(cr >> 24), dm.w, rd.w) << 24);
}

template <bool VEC4>
__global__ void __launch_bounds__(256) mask_bound_kernel(MaskArgs a, uint8_t* __restrict__ planes, uint32_t T,
                                                         uint32_t iters) {
  const uint32_t idx = (blockIdx.x * 256 + threadIdx.x) * (VEC4 ? 4u : 1u), j = blockIdx.y, pair = blockIdx.z;
  if (idx >= T) return;
  const uint32_t l = idx + T * (4u * j + 2u * pair);
  uint8_t* dst = planes + ((uint64_t)pair * iters + j) * T + idx;
  if (VEC4)
    *reinterpret_cast<uint32_t*>(dst) = __vmaxu4(bound_of4(a, l), bound_of4(a, l + T));
  else
    *dst = (uint8_t)max(bound_of(a, l), bound_of(a, l + T));
}

// Generator, early-out form.  grid = (T / 256, rounds split, draws); a thread walks the rounds j = blockIdx.y,
// blockIdx.y + gridDim.y, ... of its Philox subsequence: one 2-byte load, one Philox block, two byte
// compares.  A passing pair is queued PER WARP as one word (round, lane, which pair of the block)
// -- ballot-compacted, no atomics, no block barrier -- and whenever a warp has 32 of them all its lanes take
// one each: recompute that Philox block, curand's Box-Muller, the reference's expression, the exact
// compare.  The exact path so runs with full warps although only a few percent of the pairs need it.
constexpr int kWarpQueue = 96;   // < 32 left over + at most 64 new per round

__device__ __forceinline__ void mask_exact_pair(const MaskArgs& a, uint32_t* __restrict__ out, uint32_t x, uint32_t y,
                                                uint32_t l0, uint32_t T) {
  const float2 z = _curand_box_muller(x, y);
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const uint32_t l = l0 + h * T;
    if (l < a.numel) {
      const uint32_t e = div_n(a, l), nn = l - e * a.n;
      const int cross = (int)__ldg(a.cross_rows + (uint64_t)e * a.np + nn);
      const float sr = spin_rand(__ldg(a.degm + nn), a.negmult, cross, h ? z.y : z.x, __ldg(a.rd_std + nn));
      if (sr > __ldg(a.thresh + e)) atomicOr(out + (l >> 5), 1u << (l & 31u));
    }
  }
}

__global__ void __launch_bounds__(256) noise_mask_fast_kernel(MaskArgs a, const uint8_t* __restrict__ planes,
                                                              TorchRng r) {
  __shared__ uint32_t queue[8][kWarpQueue];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t lt = (1u << lane) - 1u;
  const uint32_t idx = blockIdx.x * 256 + threadIdx.x;
  const uint32_t k = blockIdx.z;
  const PhiloxKey key = philox_key(r);
  const uint64_t ctr0 = key.offset4 + (uint64_t)k * r.iters_per_call;
  const PhiloxKeys pkey = philox_keys(key.seed);
  const uint32_t T = r.threads;
  uint32_t* out = a.masks + (int64_t)k * a.mask_words;
  uint32_t* q = queue[warp];
  const uint32_t jstep = gridDim.y;
  const uint8_t* bp = planes + (uint64_t)blockIdx.y * T + idx;           // pair A plane; pair B one plane further
  const uint64_t bstride = (uint64_t)jstep * T, plane = (uint64_t)r.iters_per_call * T;

  auto exact = [&](uint32_t ent) {      // one queued pair: bit 0 = second pair of the block
    const uint32_t jj = ent >> 8, src = idx - lane + ((ent >> 2) & 31u);
    const uint64_t ctr = ctr0 + jj;
    const uint4 o = philox10(make_uint4((uint32_t)ctr, (uint32_t)(ctr >> 32), src, 0u), pkey);
    const bool second = (ent & 1u) != 0u;
    mask_exact_pair(a, out, second ? o.z : o.x, second ? o.w : o.y, src + T * (4u * jj + (second ? 2u : 0u)), T);
  };

  uint32_t ca = __ldg(bp), cb = __ldg(bp + plane);
  int cnt = 0;
  for (uint32_t j = blockIdx.y; j < r.iters_per_call; j += jstep) {
    bp += bstride;
    uint32_t can = 0, cbn = 0;
    if (j + jstep < r.iters_per_call) can = __ldg(bp), cbn = __ldg(bp + plane);   // next round's bytes travel during this Philox block
    const uint64_t ctr = ctr0 + j;
    const uint4 o = philox10(make_uint4((uint32_t)ctr, (uint32_t)(ctr >> 32), idx, 0u), pkey);
    const bool pa = (o.x >> 24) <= ca, pb = (o.z >> 24) <= cb;
    const uint32_t ba = __ballot_sync(kFull, pa), bb = __ballot_sync(kFull, pb);
    const uint32_t ent = (j << 8) | ((uint32_t)lane << 2);
    if (pa) q[cnt + __popc(ba & lt)] = ent;
    if (pb) q[cnt + __popc(ba) + __popc(bb & lt)] = ent | 1u;
    cnt += __popc(ba) + __popc(bb);
    __syncwarp();
    while (cnt >= 32) {
      cnt -= 32;
      const uint32_t mine = q[cnt + lane];
      __syncwarp();
      exact(mine);
    }
    ca = can, cb = cbn;
  }
  if (lane < cnt) exact(q[lane]);
}

// ---- streaming generator: the form that runs NEXT TO the tile kernel (rlsb_ls_fused_search).
//
// The masks do not depend on what the tile kernel accepts, so the generator can run a group of draws ahead of
// the tile CTAs on the same SMs.  One persistent block of 8 warps per SM (64 registers: it fits beside a tile CTA of
// 512 threads x 96 registers and its 170 KB of shared memory).  Work is handed out dynamically: the draws are
// processed in GROUPS of kGenGroup consecutive draws; a unit is one 256-thread slice of torch's call geometry
// for every round j and every draw of the group; `ctl[2g]` hands out the units of group g, `ctl[2g + 1]` counts
// the finished ones.  A tile CTA starts iteration it only when ctl[2 (it / kGenGroup) + 1] == units (release /
// acquire through a gpu-scope fence around the counter).  No block ever waits for another block, so the kernel
// makes progress whatever part of the grid is resident.
//
// Inside a unit a thread handles the kGenGroup draws of round j together: the two early-out bytes are loaded
// once per round, the Philox blocks of the draws are independent instruction streams (the latency of one
// hides behind the other: two warps per scheduler are enough), and passing pairs go to the warp's queue
// exactly as in noise_mask_fast_kernel.
constexpr int kGenQueue = 32 + 64 * kGenGroup;

// one unit: the 256-thread slice `unit` of the call geometry, every round, the draws k0 .. k0 + nd - 1
__device__ __forceinline__ void gen_unit(const MaskArgs& a, const uint8_t* __restrict__ planes, const PhiloxKey& key,
                                         const PhiloxKeys& pkey, uint32_t T, uint32_t iters, uint32_t unit, int k0, int nd,
                                         uint32_t* q) {
  const int lane = threadIdx.x & 31;
  const uint32_t lt = (1u << lane) - 1u;
  const uint64_t plane = (uint64_t)iters * T;
  const uint32_t idx = unit * 256 + threadIdx.x;
  const uint8_t* bp = planes + idx;
  int cnt = 0;

  auto exact = [&](uint32_t ent) {                // one queued pair: bit 0 = second pair, bit 1 = draw in the group
    const uint32_t jj = ent >> 8, src = idx - lane + ((ent >> 2) & 31u), d = (ent >> 1) & 1u;
    const uint64_t ctr = key.offset4 + (uint64_t)(k0 + d) * iters + jj;
    const uint4 o = philox10(make_uint4((uint32_t)ctr, (uint32_t)(ctr >> 32), src, 0u), pkey);
    const bool second = (ent & 1u) != 0u;
    mask_exact_pair(a, a.masks + (int64_t)(k0 + d) * a.mask_words, second ? o.z : o.x, second ? o.w : o.y,
                    src + T * (4u * jj + (second ? 2u : 0u)), T);
  };

  uint32_t ca = __ldg(bp), cb = __ldg(bp + plane);
  for (uint32_t j = 0; j < iters; ++j) {
    bp += T;
    uint32_t can = 0, cbn = 0;
    if (j + 1 < iters) can = __ldg(bp), cbn = __ldg(bp + plane);
    uint4 o[kGenGroup];
#pragma unroll
    for (int d = 0; d < kGenGroup; ++d) {
      const uint64_t ctr = key.offset4 + (uint64_t)(k0 + d) * iters + j;
      o[d] = philox10(make_uint4((uint32_t)ctr, (uint32_t)(ctr >> 32), idx, 0u), pkey);
    }
#pragma unroll
    for (int d = 0; d < kGenGroup; ++d) {
      const bool pa = d < nd && (o[d].x >> 24) <= ca, pb = d < nd && (o[d].z >> 24) <= cb;
      const uint32_t ba = __ballot_sync(kFull, pa), bb = __ballot_sync(kFull, pb);
      const uint32_t ent = (j << 8) | ((uint32_t)lane << 2) | ((uint32_t)d << 1);
      if (pa) q[cnt + __popc(ba & lt)] = ent;
      if (pb) q[cnt + __popc(ba) + __popc(bb & lt)] = ent | 1u;
      cnt += __popc(ba) + __popc(bb);
    }
    __syncwarp();
    while (cnt >= 32) {
      cnt -= 32;
      const uint32_t mine = q[cnt + lane];
      __syncwarp();
      exact(mine);
    }
    ca = can, cb = cbn;
  }
  if (lane < cnt) exact(q[lane]);
}

__global__ void __maxnreg__(64) noise_mask_stream_kernel(MaskArgs a, const uint8_t* __restrict__ planes, TorchRng r,
                                                         uint32_t* __restrict__ ctl, int num_draws, uint32_t units) {
  __shared__ uint32_t queue[8][kGenQueue];
  __shared__ uint32_t sUnit[2];
  const PhiloxKey key = philox_key(r);
  const PhiloxKeys pkey = philox_keys(key.seed);
  uint32_t* q = queue[threadIdx.x >> 5];
  const int groups = (num_draws + kGenGroup - 1) / kGenGroup;
  if (threadIdx.x == 0 && groups > 0) atomicAdd(ctl + kLsCtlStarted, 1u);
  for (int g = 0; g < groups; ++g) {
    const int k0 = g * kGenGroup;
    const int nd = min(kGenGroup, num_draws - k0);
    // units are claimed one ahead: the atomic of the next claim is in flight while the current unit is processed
    if (threadIdx.x == 0) sUnit[0] = atomicAdd(ctl + 2 * g, 1u);
    __syncthreads();
    for (int cur = 0;; cur ^= 1) {
      const uint32_t unit = sUnit[cur];
      if (unit >= units) break;                       // uniform: every thread read the same value
      uint32_t next_unit = 0;
      if (threadIdx.x == 0) next_unit = atomicAdd(ctl + 2 * g, 1u);
      gen_unit(a, planes, key, pkey, r.threads, r.iters_per_call, unit, k0, nd, q);
      if (threadIdx.x == 0) sUnit[cur ^ 1] = next_unit;
      __syncthreads();                                // every flip bit of the unit has been issued; the next claim is published
      if (threadIdx.x == 0) {
        __threadfence();                              // ... and is visible device-wide before the count moves
        atomicAdd(ctl + 2 * g + 1, 1u);
      }
    }
    __syncthreads();                                  // sUnit is rewritten for the next group
  }
}

// The same unit loop as an ordinary grid (one block per unit and group of draws): the generator of the SEQUENTIAL
// path.  Two draws per thread share the early-out bytes and give the scheduler two independent Philox blocks; the
// ten round keys sit in registers.
__global__ void __maxnreg__(64) noise_mask_group_kernel(MaskArgs a, const uint8_t* __restrict__ planes, TorchRng r,
                                                        int num_draws) {
  __shared__ uint32_t queue[8][kGenQueue];
  const PhiloxKey key = philox_key(r);
  const PhiloxKeys pkey = philox_keys(key.seed);
  const int k0 = blockIdx.y * kGenGroup;
  gen_unit(a, planes, key, pkey, r.threads, r.iters_per_call, blockIdx.x, k0, min(kGenGroup, num_draws - k0),
           queue[threadIdx.x >> 5]);
}

// the same draws as explicit float32 tensors (tests: must equal torch.randn bit for bit)
__global__ void __launch_bounds__(256) noise_values_kernel(float* __restrict__ out, uint32_t numel, TorchRng r) {
  const uint32_t idx = blockIdx.x * 256 + threadIdx.x;
  const uint32_t j = blockIdx.y, k = blockIdx.z;
  const PhiloxKey key = philox_key(r);
  const uint64_t ctr = key.offset4 + (uint64_t)k * r.iters_per_call + j;
  const uint4 o = curand_Philox4x32_10(make_uint4((uint32_t)ctr, (uint32_t)(ctr >> 32), idx, 0u),
                                       make_uint2((uint32_t)key.seed, (uint32_t)(key.seed >> 32)));
  const float2 n01 = _curand_box_muller(o.x, o.y), n23 = _curand_box_muller(o.z, o.w);
  const float z[4] = {n01.x, n01.y, n23.x, n23.y};
  uint32_t li = idx + r.threads * 4u * j;
#pragma unroll
  for (int ii = 0; ii < 4; ++ii) {
    if (li < numel) out[(uint64_t)k * numel + li] = z[ii];
    li += r.threads;
  }
}

__global__ void cursor_advance_kernel(uint64_t* cur, uint64_t delta) { cur[1] += delta; }

static int rng_check(const char* what, int64_t numel, uint64_t offset, int32_t rng_threads, int32_t rng_iters,
                     int32_t num_draws) {
  RLSB_REQUIRE(numel > 0 && numel < (int64_t(1) << 31), RLSB_ERR_UNSUPPORTED,
               "%s: %lld elements per draw (torch splits calls of 2^31 elements and more)", what, (long long)numel);
  RLSB_REQUIRE(rng_threads > 0 && rng_threads % 256 == 0 && rng_iters > 0 && rng_iters <= 65535, RLSB_ERR_INVALID,
               "%s: bad call geometry (threads %d, iters %d)", what, rng_threads, rng_iters);
  RLSB_REQUIRE((int64_t)rng_threads * 4 * rng_iters >= numel, RLSB_ERR_INVALID,
               "%s: call geometry (threads %d, iters %d) does not cover %lld elements", what, rng_threads, rng_iters,
               (long long)numel);
  RLSB_REQUIRE(offset % 4 == 0, RLSB_ERR_INVALID, "%s: generator offset must be a multiple of 4", what);
  RLSB_REQUIRE(num_draws >= 0 && num_draws <= 65535, RLSB_ERR_INVALID, "%s: num_draws out of range", what);
  return RLSB_OK;
}

// ---- host pieces shared by rlsb_ls_noise_masks and rlsb_ls_fused_search (local_search.cu)
int mask_plan(const GraphDev& g, const char* what, int64_t num_envs, int ws_mult, uint64_t seed, uint64_t offset,
              const uint64_t* rng_dev, int rng_threads, int rng_iters, int num_draws, uint32_t* masks, void* workspace,
              MaskPlan* plan) {
  RLSB_REQUIRE(num_envs > 0, RLSB_ERR_INVALID, "%s: num_envs must be positive", what);
  RLSB_REQUIRE(ws_mult == 1 || ws_mult == 2, RLSB_ERR_INVALID, "%s: ws_mult must be 1 or 2", what);
  RLSB_REQUIRE(degree_class(g) != 2, RLSB_ERR_UNSUPPORTED,
               "%s: degrees above 255 keep uint16 counts (use rlsb_ls_run with noise tensors)", what);
  const int64_t numel = num_envs * (int64_t)g.n;
  if (int rc = rng_check(what, numel, offset, rng_threads, rng_iters, num_draws)) return rc;
  RLSB_REQUIRE(masks && workspace, RLSB_ERR_INVALID, "%s: null pointer", what);
  const LsWorkspace w = carve(g, num_envs, workspace);
  RLSB_REQUIRE(w.bound != nullptr, RLSB_ERR_INVALID, "%s: workspace without the bound section", what);
  RLSB_REQUIRE((int64_t)rng_threads * rng_iters * 2 <= ls_bound_bytes(numel), RLSB_ERR_INVALID,
               "%s: call geometry (threads %d, iters %d) larger than a B200's", what, rng_threads, rng_iters);
  MaskArgs& a = plan->a;
  a.cross_rows = w.cross_rows, a.rd_std = w.rd_std, a.degm = w.degm, a.thresh = w.thresh;
  a.masks = masks, a.mask_words = ls_mask_words(num_envs, g.n);
  a.numel = (uint32_t)numel, a.n = (uint32_t)g.n, a.np = (uint32_t)g.np;
  a.step_e = (uint32_t)rng_threads / a.n, a.step_n = (uint32_t)rng_threads % a.n;
  a.negmult = -ws_mult;
  a.div_s = 0;
  while ((1u << a.div_s) < a.n) ++a.div_s;
  a.div_m = (uint32_t)(((uint64_t(1) << (31 + a.div_s)) + a.n - 1) / a.n);
  plan->r = TorchRng{seed, offset / 4, (uint32_t)rng_threads, (uint32_t)rng_iters, rng_dev};
  plan->bound = w.bound, plan->ctl = w.ctl, plan->num_draws = num_draws;
  return RLSB_OK;
}

// early-out bytes (unless still valid) + zeroed mask arrays (+ zeroed unit counters of the streaming form)
int mask_prepare(const MaskPlan& p, bool write_bound, bool zero_ctl, cudaStream_t st) {
  const uint32_t T = p.r.threads, iters = p.r.iters_per_call;
  if (write_bound) {
    if (p.a.n % 4 == 0)
      mask_bound_kernel<true><<<dim3((T + 1023) / 1024, iters, 2), 256, 0, st>>>(p.a, p.bound, T, iters);
    else
      mask_bound_kernel<false><<<dim3(T / 256, iters, 2), 256, 0, st>>>(p.a, p.bound, T, iters);
    RLSB_LAUNCH_OK();
  }
  RLSB_CUDA_OK(cudaMemsetAsync(p.a.masks, 0, (size_t)p.num_draws * p.a.mask_words * sizeof(uint32_t), st));
  if (zero_ctl) RLSB_CUDA_OK(cudaMemsetAsync(p.ctl, 0, kLsCtlBytes, st));
  return RLSB_OK;
}

// CUDA loads a kernel's code lazily, at its first launch, and that load may need the context to go idle: a FIRST
// launch of the generator while tile CTAs are already spinning on its counters would never start (CUDA Programming
// Guide, "Lazy Loading": preload kernels that must run concurrently).  Called on the caller's stream before the
// tile kernel of a fused search is launched: once per device it runs the generator with no work.
int mask_stream_preload(const MaskPlan& p, cudaStream_t st) {
  static bool loaded[64] = {};
  int dev = 0;
  RLSB_CUDA_OK(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && loaded[dev]) return RLSB_OK;
  cudaFuncAttributes attr;
  RLSB_CUDA_OK(cudaFuncGetAttributes(&attr, noise_mask_stream_kernel));
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  RLSB_CUDA_OK(cudaStreamIsCapturing(st, &cap));
  RLSB_REQUIRE(cap == cudaStreamCaptureStatusNone, RLSB_ERR_INVALID,
               "fused search: the first call on a device must not be captured into a CUDA graph (run it once eagerly: "
               "the generator kernel has to be loaded before it can run next to the tile kernel)");
  noise_mask_stream_kernel<<<1, 256, 0, st>>>(p.a, p.bound, p.r, p.ctl, 0, 0u);     // no draws: returns at once
  RLSB_LAUNCH_OK();
  if (dev >= 0 && dev < 64) loaded[dev] = true;
  return RLSB_OK;
}

// the streaming generator: one persistent block per SM on `st` (the tile kernel polls ctl)
int mask_stream_launch(const MaskPlan& p, cudaStream_t st) {
  RLSB_REQUIRE(p.num_draws <= kLsMaxFusedDraws, RLSB_ERR_INVALID, "fused search: at most %d draws per call", kLsMaxFusedDraws);
  // Two kernels only share an SM when they ask for the same shared-memory carve-out (measured: with the driver's
  // per-kernel default the generator's blocks wait for the tile CTAs to leave, profiles/r02_fused_debug.log), so both
  // kernels of the fused search ask for the largest one.
  if (!(debug_flags() & RLSB_DEBUG_CARVEOUT_DEFAULT))
    RLSB_CUDA_OK(cudaFuncSetAttribute(noise_mask_stream_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                      cudaSharedmemCarveoutMaxShared));
  noise_mask_stream_kernel<<<kNumSMs, 256, 0, st>>>(p.a, p.bound, p.r, p.ctl, p.num_draws, p.r.threads / 256);
  RLSB_LAUNCH_OK();
  return RLSB_OK;
}

}  // namespace rlsb

extern "C" {

int64_t rlsb_ls_mask_words(const rlsb_graph_t* gh, int64_t num_envs) {
  using namespace rlsb;
  const GraphDev* g;
  if (graph_check(gh, &g, "ls_mask_words") || num_envs < 0) return -1;
  if (degree_class(*g) == 2 || num_envs * (int64_t)g->n >= (int64_t(1) << 31)) return -1;
  return ls_mask_words(num_envs, g->n);
}

int rlsb_ls_noise_masks(const rlsb_graph_t* gh, int64_t num_envs, int32_t ws_mult, uint64_t seed, uint64_t offset,
                        const uint64_t* rng_dev, int32_t rng_threads, int32_t rng_iters, int32_t num_draws,
                        int32_t reuse_bound, uint32_t* masks, void* workspace, void* stream) {
  using namespace rlsb;
  const GraphDev* g;
  if (int rc = graph_check(gh, &g, "ls_noise_masks")) return rc;
  if (num_envs == 0 || g->n == 0 || num_draws == 0) return RLSB_OK;
  MaskPlan plan;
  if (int rc = mask_plan(*g, "ls_noise_masks", num_envs, ws_mult, seed, offset, rng_dev, rng_threads, rng_iters, num_draws,
                         masks, workspace, &plan))
    return rc;
  auto st = static_cast<cudaStream_t>(stream);
  if (debug_flags() & RLSB_DEBUG_PLAIN_MASKS) {   // evaluate every normal (cross-check of the early-out form)
    const dim3 grid((unsigned)(rng_threads / 256), (unsigned)rng_iters, (unsigned)num_draws);
    noise_mask_kernel<<<grid, 256, 0, st>>>(plan.a, plan.r);
    RLSB_LAUNCH_OK();
    return RLSB_OK;
  }
  if (int rc = mask_prepare(plan, !reuse_bound, false, st)) return rc;
  if (!(debug_flags() & RLSB_DEBUG_GEN_PER_DRAW) && (int64_t)(rng_threads / 256) * ((num_draws + kGenGroup - 1) / kGenGroup) >= 2 * kNumSMs) {
    noise_mask_group_kernel<<<dim3((unsigned)(rng_threads / 256), (unsigned)((num_draws + kGenGroup - 1) / kGenGroup)), 256, 0, st>>>(
        plan.a, plan.bound, plan.r, num_draws);
    RLSB_LAUNCH_OK();
    return RLSB_OK;
  }
  // rounds are split over blockIdx.y only when there would be too few blocks to fill the GPU otherwise
  dim3 grid((unsigned)(rng_threads / 256), 1, (unsigned)num_draws);
  int jsplit = 1;
  while ((int64_t)grid.x * jsplit * num_draws < 4 * 8 * kNumSMs && jsplit < rng_iters) jsplit *= 2;
  if (jsplit > rng_iters) jsplit = rng_iters;
  grid.y = (unsigned)jsplit;
  noise_mask_fast_kernel<<<grid, 256, 0, st>>>(plan.a, plan.bound, plan.r);
  RLSB_LAUNCH_OK();
  return RLSB_OK;
}

int rlsb_torch_randn(float* out, int64_t numel, uint64_t seed, uint64_t offset, const uint64_t* rng_dev,
                     int32_t rng_threads, int32_t rng_iters, int32_t num_draws, void* stream) {
  using namespace rlsb;
  if (numel == 0 || num_draws == 0) return RLSB_OK;
  if (int rc = rng_check("torch_randn", numel, offset, rng_threads, rng_iters, num_draws)) return rc;
  RLSB_REQUIRE(out, RLSB_ERR_INVALID, "torch_randn: null pointer");
  TorchRng r{seed, offset / 4, (uint32_t)rng_threads, (uint32_t)rng_iters, rng_dev};
  const dim3 grid((unsigned)(rng_threads / 256), (unsigned)rng_iters, (unsigned)num_draws);
  // same carve-out as the tile "prepare" kernel, so that the two can share SMs when a caller runs the threshold
  // draw on a second stream next to rlsb_ls_begin (GraphStore.ls_fused under graph capture)
  static bool carve_set = false;
  if (!carve_set) {
    RLSB_CUDA_OK(cudaFuncSetAttribute(noise_values_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                      cudaSharedmemCarveoutMaxShared));
    carve_set = true;
  }
  noise_values_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(out, (uint32_t)numel, r);
  RLSB_LAUNCH_OK();
  return RLSB_OK;
}

int rlsb_rng_cursor_advance(uint64_t* rng_dev, uint64_t delta, void* stream) {
  using namespace rlsb;
  RLSB_REQUIRE(rng_dev != nullptr && delta % 4 == 0, RLSB_ERR_INVALID, "rng_cursor_advance: bad argument");
  cursor_advance_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(rng_dev, delta);
  RLSB_LAUNCH_OK();
  return RLSB_OK;
}

}  // extern "C"
