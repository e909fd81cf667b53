// Dense QUBO Hamiltonian  E[c] = x_c^T Q x_c  for a batch of chains -- the "compute value" step of
// mcpg_sampling_qubo / mcpg_sampling_qubo_bin (rlsolver/methods/MCPG/sampling.py:339-340, 364-365:
// res = matmul(Q, X); sum(X * res, dim=0)) and the energy of PISCO's dense model
// (rlsolver/envs/env_ISCO.py:436-444).  This is the one place on the path where the work really is
// a batched contraction, so it runs on the 5th-generation tensor cores:
//
//   * Q (float32) is split once into three bf16 limbs hi + mid + lo (3 x 8 = 24 significand bits:
//     the split is exact), X in {-1, 0, +1} is exact in bf16, products are exact and the
//     accumulation is fp32 in tensor memory -- the same arithmetic class as the reference's SGEMM.
//   * One CTA computes a 128 (rows of Q) x 256 (chains) tile of Y = Q X over the whole K = N range:
//     TMA (cp.async.bulk.tensor, 128B swizzle) feeds a 2-stage shared-memory ring, one elected
//     thread issues tcgen05.mma (M128 N256 K16, kind::f16, bf16 in / f32 out) for the three limbs
//     into one TMEM accumulator, completion is signalled through tcgen05.commit -> mbarrier.
//   * The epilogue never writes Y: four warps read the accumulator with tcgen05.ld, multiply by
//     x[row][chain] (kept as two bit masks per thread), reduce over the 128 rows (transposing
//     butterfly, 31 shuffles per 32 columns) and store one partial energy per (row block, chain); a
//     tiny kernel adds the row blocks in a fixed order (deterministic, no atomics).
//   * K is accumulated in tensor memory in chunks of 512 and the two 256-column accumulators
//     alternate (MMA fills one while the epilogue drains the other): the tensor core's truncating
//     fp32 adds would otherwise drift past the 1e-5 tolerance at N = 4096.
#include <cuda.h>
#include <cuda_bf16.h>

#include <vector>

#include "common.cuh"

namespace rlsb {

constexpr int kQM = 128, kQN = 256, kQK = 64, kQStages = 2, kQLimbs = 3;
constexpr int kQThreads = 192;                       // warp 0 TMA, warp 1 MMA + TMEM, warps 2-5 epilogue
constexpr uint32_t kTileA = kQM * kQK * 2;           // 16 KB
constexpr uint32_t kTileB = kQN * kQK * 2;           // 32 KB
constexpr uint32_t kStageBytes = kQLimbs * kTileA + kTileB;   // 80 KB

constexpr int kQChunk = 8;                           // k-blocks (512 of K) accumulated in tensor memory before a drain

struct QuboSmem {
  uint8_t tiles[kQStages][kStageBytes];              // [A_hi | A_mid | A_lo | B], each 1024-byte aligned
  float part[4][kQN];
  uint32_t negm[kQN / 32][128], zerom[kQN / 32][128];   // x[row][chain] of this tile as sign / zero bit masks
  uint64_t full_bar[kQStages], empty_bar[kQStages], tmem_full_bar[2], tmem_empty_bar[2];
  uint32_t tmem_base;
};

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// K-major, 128-byte swizzle, 8-row groups 1024 bytes apart (cute::UMMA::SmemDescriptor, version 1)
__device__ __forceinline__ uint64_t umma_desc(const void* tile) {
  const uint32_t addr = smem_u32(tile);
  return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_c, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_c),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}

// The tensor core adds every K=16 step into the fp32 accumulator with truncation, so a long K
// loop drifts (measured 2.7e-5 relative at K = 4096 x 3 limbs).  K is therefore cut into chunks of
// kQChunk k-blocks that accumulate in tensor memory; the two 256-column accumulators alternate, and
// while one fills the epilogue warps drain the other into round-to-nearest fp32 running sums.
__global__ void __launch_bounds__(kQThreads, 1)
qubo_energy_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const __nv_bfloat16* __restrict__ xt, int np, int cp, float* __restrict__ partial) {
  extern __shared__ uint8_t smem_raw[];
  QuboSmem& S = *reinterpret_cast<QuboSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * kQM, n0 = blockIdx.y * kQN;
  const int kblocks = np / kQK;
  const int chunks = (kblocks + kQChunk - 1) / kQChunk;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kQStages; ++s) mbar_init(&S.full_bar[s], 1), mbar_init(&S.empty_bar[s], 1);
    for (int b = 0; b < 2; ++b) mbar_init(&S.tmem_full_bar[b], 1), mbar_init(&S.tmem_empty_bar[b], 4);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {   // tensor memory: two accumulators of 256 fp32 columns x 128 lanes
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_base)),
                 "r"((uint32_t)(2 * kQN))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = S.tmem_base;

  if (warp == 0) {
    if (elect_one()) {                                   // ===== TMA producer
      for (int kb = 0; kb < kblocks; ++kb) {
        const int s = kb % kQStages, ph = (kb / kQStages) & 1;
        mbar_wait(&S.empty_bar[s], ph ^ 1);
        mbar_arrive_expect_tx(&S.full_bar[s], kStageBytes);
        for (int l = 0; l < kQLimbs; ++l)
          tma_load_2d(S.tiles[s] + l * kTileA, &tmA, &S.full_bar[s], kb * kQK, l * np + m0);
        tma_load_2d(S.tiles[s] + kQLimbs * kTileA, &tmB, &S.full_bar[s], kb * kQK, n0);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {                                   // ===== MMA issuer
      // instruction descriptor: D = f32, A = B = bf16, both K-major, N = 256, M = 128
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kQN >> 3) << 17) | ((uint32_t)(kQM >> 4) << 24);
      for (int ch = 0; ch < chunks; ++ch) {
        const int buf = ch & 1;
        mbar_wait(&S.tmem_empty_bar[buf], ((ch >> 1) & 1) ^ 1);     // the epilogue has drained this accumulator
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int kb_end = min(kblocks, (ch + 1) * kQChunk);
        for (int kb = ch * kQChunk; kb < kb_end; ++kb) {
          const int s = kb % kQStages, ph = (kb / kQStages) & 1;
          mbar_wait(&S.full_bar[s], ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t db = umma_desc(S.tiles[s] + kQLimbs * kTileA);
#pragma unroll
          for (int l = 0; l < kQLimbs; ++l) {
            const uint64_t da = umma_desc(S.tiles[s] + l * kTileA);
#pragma unroll
            for (int k = 0; k < kQK / 16; ++k)             // 32 bytes = 2 x 16-byte units along K per step
              umma_bf16(tmem + buf * kQN, da + 2 * k, db + 2 * k, idesc,
                        (kb != ch * kQChunk || l != 0 || k != 0) ? 1u : 0u);
          }
          umma_commit(&S.empty_bar[s]);                    // frees the stage when these MMAs retire
        }
        umma_commit(&S.tmem_full_bar[buf]);
      }
    }
  } else {                                               // ===== epilogue: warps 2..5
    const int q = warp & 3;                                // TMEM lane quarter this warp may read
    const int row = m0 + 32 * q + lane;
    // x[row][n0 .. n0+255] as two bit masks (negative / zero); 32 consecutive rows of one chain are 64 contiguous bytes
    const int et = threadIdx.x - 64;                       // 0..127 within the epilogue group
#pragma unroll 1
    for (int g = 0; g < kQN / 32; ++g) {
      uint32_t nm = 0, zm = 0;
#pragma unroll 8
      for (int j = 0; j < 32; ++j) {
        const float xv = __bfloat162float(xt[(size_t)(n0 + g * 32 + j) * np + row]);
        nm |= (uint32_t)(xv < 0.f) << j;
        zm |= (uint32_t)(xv == 0.f) << j;
      }
      S.negm[g][et] = nm, S.zerom[g][et] = zm;
      S.part[q][g * 32 + lane] = 0.f;                      // running sum of column g*32 + lane over this warp's rows
    }
    for (int ch = 0; ch < chunks; ++ch) {
      const int buf = ch & 1;
      mbar_wait(&S.tmem_full_bar[buf], (ch >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
      for (int g = 0; g < kQN / 32; ++g) {
        const uint32_t neg = S.negm[g][et], zero = S.zerom[g][et];
        uint32_t r[32];
        const uint32_t taddr = tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(buf * kQN + g * 32);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
              "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
              "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {       // y * x with x in {-1, 0, +1}: flip the sign bit / clear
          const uint32_t sgn = ((neg >> j) & 1u) << 31;
          const uint32_t keep = ((zero >> j) & 1u) ? 0u : kFull;
          v[j] = __uint_as_float((r[j] ^ sgn) & keep);
        }
        // sum over the 32 lanes (rows) of every column; lane j ends up with column g*32 + j
#pragma unroll
        for (int h = 16; h >= 1; h >>= 1) {
#pragma unroll
          for (int j = 0; j < h; ++j) {
            const bool upper = (lane & h) != 0;
            const float send = upper ? v[j] : v[j + h];
            const float kept = upper ? v[j + h] : v[j];
            v[j] = kept + __shfl_xor_sync(kFull, send, h);
          }
        }
        S.part[q][g * 32 + lane] += v[0];
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&S.tmem_empty_bar[buf]);
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");        // the four epilogue warps
    for (int c = et; c < kQN; c += 128)
      partial[(size_t)blockIdx.x * cp + n0 + c] = (S.part[0][c] + S.part[1][c]) + (S.part[2][c] + S.part[3][c]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)(2 * kQN)) : "memory");
}

// ---------------------------------------------------------------- coordinate-ascent sweeps
// Local-search sweeps of mcpg_sampling_qubo / mcpg_sampling_qubo_bin
// (rlsolver/methods/MCPG/sampling.py:331-337, 356-362): for index in 0..N-1 (Gauss-Seidel):
//     x[index] = 0; res = Q[index] . x; x[index] = (res > 0) ? +1 : -1        (+-1 form)
//     x[index] = 0; res = Q[index] . x; x[index] = (res > -Q[index][index]/2)  (0/1 form)
// for every chain -- N dependent GEMVs in the reference.  Blocked form: for a block of 64 rows
//     F^T = X^T Q[block, :]^T     (128 chains x 64 rows x N: one tensor-core GEMM tile as above, with
//                                  the chains on the M side so that a TMEM lane = a chain)
// and then the 64 decisions of the block in order, each one correcting the rows after it,
//     res_i = F_i - Q_ii x_i ;  x_i' = rule(res_i) ;  F_j += Q_ji (x_i' - x_i)  for j > i in the block,
// which is algebraically the reference's sequence.  Chains are independent and a thread of the four
// epilogue warps owns one chain: F[64] and x[64] live in its registers, so the in-block
// corrections (64*63/2 FMAs against the broadcast diagonal block of Q in shared memory) need no
// barrier and no shuffles.  One launch per row block (the next block's GEMM reads the X this one wrote).
constexpr int kSM = 128;                              // chains per CTA (MMA M)
constexpr int kSB = 64;                               // rows of Q per block (MMA N)
constexpr int kSStages = 3;
constexpr int kSplitMaxOwn = 32;                      // split-K: chains a CTA decides (128 / ks), ks in {4, 8, 16, 32}
constexpr uint32_t kSTileX = kSM * kQK * 2;           // 16 KB
constexpr uint32_t kSTileQ = kSB * kQK * 2;           // 8 KB
constexpr uint32_t kSStageBytes = kSTileX + kQLimbs * kSTileQ;   // 40 KB

struct SweepSmem {
  uint8_t tiles[kSStages][kSStageBytes];              // [X | Q_hi | Q_mid | Q_lo]
  float qd[kSB][kSB];                                 // qd[j][i] = Q[m0 + j][m0 + i]
  float sf[kSplitMaxOwn][kSB];                        // split-K: the summed F of the chains this CTA decides
  uint64_t full_bar[kSStages], empty_bar[kSStages], tmem_full_bar[2], tmem_empty_bar[2];
  uint32_t tmem_base;
};

// The 64 in-order decisions of a row block for ONE chain (f = F of the block, xv = its current x values) and the
// write-back of the new x: the bf16 chain-major copy the next block's GEMM reads, and float32 [N][C].
__device__ __forceinline__ void sweep_decide_block(const float (&qd)[kSB][kSB], float (&f)[kSB], float (&xv)[kSB], int m0, int n,
                                                   int np, int binary, int64_t chain, int64_t num_chains,
                                                   __nv_bfloat16* __restrict__ xt, float* __restrict__ x) {
  const float lo = binary ? 0.f : -1.f;
#pragma unroll
  for (int i = 0; i < kSB; ++i) {
    const float qii = qd[i][i];
    const float res = f[i] - qii * xv[i];                // Q[i] . x with x_i zeroed
    const float xn = res > (binary ? -qii * 0.5f : 0.f) ? 1.f : lo;
    const float d = (m0 + i < n) ? xn - xv[i] : 0.f;     // padding rows never move
    xv[i] += d;
#pragma unroll
    for (int j = i + 1; j < kSB; ++j) f[j] = fmaf(qd[j][i], d, f[j]);
  }
  uint4* dst = reinterpret_cast<uint4*>(xt + (size_t)chain * np + m0);
#pragma unroll
  for (int v8 = 0; v8 < kSB / 8; ++v8) {
    uint32_t ws[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)       // values are -1, 0, +1: exact in bf16 (upper half of the f32 pattern)
      ws[k] = (__float_as_uint(xv[v8 * 8 + 2 * k]) >> 16) | (__float_as_uint(xv[v8 * 8 + 2 * k + 1]) & 0xffff0000u);
    dst[v8] = make_uint4(ws[0], ws[1], ws[2], ws[3]);
  }
  if (chain < num_chains) {
#pragma unroll
    for (int i = 0; i < kSB; ++i)
      if (m0 + i < n) x[(int64_t)(m0 + i) * num_chains + chain] = xv[i];
  }
}
// current x of the block for one chain from the bf16 chain-major copy (128 contiguous bytes)
__device__ __forceinline__ void sweep_load_x(const __nv_bfloat16* __restrict__ xt, int64_t chain, int np, int m0, float (&xv)[kSB]) {
  const uint4* src = reinterpret_cast<const uint4*>(xt + (size_t)chain * np + m0);
#pragma unroll
  for (int v8 = 0; v8 < kSB / 8; ++v8) {
    const uint4 w = src[v8];
    const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      xv[v8 * 8 + 2 * k] = __uint_as_float(ws[k] << 16);              // bf16 -> f32
      xv[v8 * 8 + 2 * k + 1] = __uint_as_float(ws[k] & 0xffff0000u);
    }
  }
}

__global__ void __launch_bounds__(kQThreads, 1)
qubo_sweep_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmQ,
                  const float* __restrict__ q, int n, int np, int m0, int64_t num_chains, int binary,
                  __nv_bfloat16* __restrict__ xt, float* __restrict__ x, int ks, float* __restrict__ part,
                  unsigned* __restrict__ counter, unsigned target) {
  extern __shared__ uint8_t smem_raw[];
  SweepSmem& S = *reinterpret_cast<SweepSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * kSM;                        // first chain of this CTA
  // split-K (ks > 1, few chains): blockIdx.y owns the k-blocks [kb0, kb0 + kblocks) of the row block's dot products;
  // the ks CTAs of a chain group add their partial F through global memory, then each decides 128 / ks chains
  const int kblocks = np / kQK / ks;
  const int kb0 = (int)blockIdx.y * kblocks;
  const int chunks = (kblocks + kQChunk - 1) / kQChunk;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kSStages; ++s) mbar_init(&S.full_bar[s], 1), mbar_init(&S.empty_bar[s], 1);
    for (int b = 0; b < 2; ++b) mbar_init(&S.tmem_full_bar[b], 1), mbar_init(&S.tmem_empty_bar[b], 4);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_base)),
                 "r"((uint32_t)(2 * kSB))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = S.tmem_base;

  if (warp == 0) {
    if (elect_one()) {                                   // ===== TMA producer
      for (int kb = 0; kb < kblocks; ++kb) {
        const int s = kb % kSStages, ph = (kb / kSStages) & 1;
        mbar_wait(&S.empty_bar[s], ph ^ 1);
        mbar_arrive_expect_tx(&S.full_bar[s], kSStageBytes);
        tma_load_2d(S.tiles[s], &tmX, &S.full_bar[s], (kb0 + kb) * kQK, n0);
        for (int l = 0; l < kQLimbs; ++l)
          tma_load_2d(S.tiles[s] + kSTileX + l * kSTileQ, &tmQ, &S.full_bar[s], (kb0 + kb) * kQK, l * np + m0);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {                                   // ===== MMA issuer (M128 chains, N64 rows, K16)
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kSB >> 3) << 17) | ((uint32_t)(kSM >> 4) << 24);
      for (int ch = 0; ch < chunks; ++ch) {
        const int buf = ch & 1;
        mbar_wait(&S.tmem_empty_bar[buf], ((ch >> 1) & 1) ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int kb_end = min(kblocks, (ch + 1) * kQChunk);
        for (int kb = ch * kQChunk; kb < kb_end; ++kb) {
          const int s = kb % kSStages, ph = (kb / kSStages) & 1;
          mbar_wait(&S.full_bar[s], ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint64_t da = umma_desc(S.tiles[s]);
#pragma unroll
          for (int l = 0; l < kQLimbs; ++l) {
            const uint64_t db = umma_desc(S.tiles[s] + kSTileX + l * kSTileQ);
#pragma unroll
            for (int k = 0; k < kQK / 16; ++k)
              umma_bf16(tmem + buf * kSB, da + 2 * k, db + 2 * k, idesc,
                        (kb != ch * kQChunk || l != 0 || k != 0) ? 1u : 0u);
          }
          umma_commit(&S.empty_bar[s]);
        }
        umma_commit(&S.tmem_full_bar[buf]);
      }
    }
  } else {                                               // ===== warps 2..5: one thread per chain
    const int qw = warp & 3;                               // TMEM lane quarter this warp may read
    const int64_t chain = n0 + 32 * qw + lane;             // < cp (padded chains carry zeros)
    const int et = threadIdx.x - 64;
    // diagonal block of Q in full precision: half a row (128 B) per thread as eight independent 16-byte loads -- one
    // trip to L2 (a scalar loop serialised 32 of them: ~10 us per launch, visible once the GEMM is split over K)
    if (n % 4 == 0 && m0 + kSB <= n) {
      const int j = et >> 1, i0 = (et & 1) * (kSB / 2);
      const float4* src = reinterpret_cast<const float4*>(q + (int64_t)(m0 + j) * n + m0 + i0);
      float4 v[kSB / 8];
#pragma unroll
      for (int k = 0; k < kSB / 8; ++k) v[k] = __ldg(src + k);
#pragma unroll
      for (int k = 0; k < kSB / 8; ++k)
        S.qd[j][i0 + 4 * k] = v[k].x, S.qd[j][i0 + 4 * k + 1] = v[k].y, S.qd[j][i0 + 4 * k + 2] = v[k].z,
        S.qd[j][i0 + 4 * k + 3] = v[k].w;
    } else {
      for (int idx = et; idx < kSB * kSB; idx += 128) {
        const int j = idx / kSB, i = idx % kSB;
        S.qd[j][i] = (m0 + j < n && m0 + i < n) ? __ldg(q + (int64_t)(m0 + j) * n + m0 + i) : 0.f;
      }
    }
    float xv[kSB], f[kSB];
    sweep_load_x(xt, chain, np, m0, xv);
#pragma unroll
    for (int c = 0; c < kSB; ++c) f[c] = 0.f;
    for (int ch = 0; ch < chunks; ++ch) {
      const int buf = ch & 1;
      mbar_wait(&S.tmem_full_bar[buf], (ch >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int g = 0; g < kSB / 32; ++g) {
        uint32_t v[32];
        const uint32_t taddr = tmem + ((uint32_t)(32 * qw) << 16) + (uint32_t)(buf * kSB + g * 32);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
              "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
              "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
              "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; ++j) f[g * 32 + j] += __uint_as_float(v[j]);     // round-to-nearest running sums
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&S.tmem_empty_bar[buf]);
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");        // qd complete
    if (ks == 1) {
      sweep_decide_block(S.qd, f, xv, m0, n, np, binary, chain, num_chains, xt, x);
    } else {
      // partial F of this K slice -> global; barrier over the ks CTAs of the chain group (all resident: the host keeps
      // groups * ks <= SMs, and the counter only grows: `target` = ks * launches so far); then every CTA adds the
      // ks partials of ITS 128 / ks chains in slice order (deterministic) and decides them
      const int own = kSM / ks;
      float4* dst = reinterpret_cast<float4*>(part + (((size_t)blockIdx.x * ks + blockIdx.y) * kSM + (32 * qw + lane)) * kSB);
#pragma unroll
      for (int v4 = 0; v4 < kSB / 4; ++v4) dst[v4] = make_float4(f[4 * v4], f[4 * v4 + 1], f[4 * v4 + 2], f[4 * v4 + 3]);
      __threadfence();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (et == 0) {
        atomicAdd(counter + blockIdx.x, 1u);
        unsigned v;
        do {
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter + blockIdx.x) : "memory");
        } while (v < target);
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      for (int item = et; item < own * (kSB / 4); item += 128) {
        const int c = item / (kSB / 4), v4 = item % (kSB / 4);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int sl = 0; sl < ks; ++sl) {
          const float4 p4 = __ldcg(reinterpret_cast<const float4*>(
              part + (((size_t)blockIdx.x * ks + sl) * kSM + (blockIdx.y * own + c)) * kSB) + v4);
          acc.x += p4.x, acc.y += p4.y, acc.z += p4.z, acc.w += p4.w;
        }
        S.sf[c][4 * v4] = acc.x, S.sf[c][4 * v4 + 1] = acc.y, S.sf[c][4 * v4 + 2] = acc.z, S.sf[c][4 * v4 + 3] = acc.w;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (et < own) {
        const int64_t mine = n0 + blockIdx.y * own + et;
        sweep_load_x(xt, mine, np, m0, xv);
#pragma unroll
        for (int c = 0; c < kSB; ++c) f[c] = S.sf[et][c];
        sweep_decide_block(S.qd, f, xv, m0, n, np, binary, mine, num_chains, xt, x);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)(2 * kSB)) : "memory");
}

// Q fp32 [n][n] -> three bf16 limbs [3][np][np], zero padded; hi + mid + lo == q exactly
__global__ void qubo_split_kernel(const float* __restrict__ q, int n, int np, __nv_bfloat16* __restrict__ limbs) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)np * np) return;
  const int i = (int)(idx / np), j = (int)(idx % np);
  const float v = (i < n && j < n) ? q[(int64_t)i * n + j] : 0.f;
  const __nv_bfloat16 hi = __float2bfloat16_rn(v);
  const float r1 = v - __bfloat162float(hi);
  const __nv_bfloat16 mid = __float2bfloat16_rn(r1);
  const float r2 = r1 - __bfloat162float(mid);
  const __nv_bfloat16 lo = __float2bfloat16_rn(r2);
  const int64_t plane = (int64_t)np * np;
  limbs[idx] = hi, limbs[plane + idx] = mid, limbs[2 * plane + idx] = lo;
}

// X float32 [n][c] (node-major) -> Xt bf16 [cp][np] (chain-major, zero padded), 32x32 tiles through smem
__global__ void qubo_xt_kernel(const float* __restrict__ x, int n, int64_t c, int np, int64_t cp,
                               __nv_bfloat16* __restrict__ xt) {
  __shared__ float tile[32][33];
  const int64_t c0 = (int64_t)blockIdx.x * 32;
  const int i0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int i = i0 + r;
    const int64_t cc = c0 + threadIdx.x;
    tile[r][threadIdx.x] = (i < n && cc < c) ? x[(int64_t)i * c + cc] : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int64_t cc = c0 + r;
    const int i = i0 + threadIdx.x;
    if (cc < cp && i < np) xt[cc * np + i] = __float2bfloat16_rn(tile[threadIdx.x][r]);
  }
}

__global__ void qubo_reduce_kernel(const float* __restrict__ partial, int mblocks, int64_t cp, int64_t c,
                                   float* __restrict__ energy) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c) return;
  float acc = 0.f;
  for (int mb = 0; mb < mblocks; ++mb) acc += partial[(size_t)mb * cp + i];
  energy[i] = acc;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// bf16 row-major [rows][cols] -> tensor map with a [box_rows][64] box, 128-byte swizzle
static int make_map(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  EncodeTiledFn fn = encode_tiled_fn();
  RLSB_REQUIRE(fn != nullptr, RLSB_ERR_CUDA, "qubo: cuTensorMapEncodeTiled is not available from this driver");
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {cols * 2};
  const cuuint32_t box[2] = {(cuuint32_t)kQK, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  RLSB_REQUIRE(r == CUDA_SUCCESS, RLSB_ERR_CUDA, "qubo: cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return RLSB_OK;
}

}  // namespace rlsb

struct rlsb_qubo {
  int32_t n = 0, np = 0, device = -1;
  __nv_bfloat16* limbs = nullptr;      // [3][np][np]
  CUtensorMap map_a;
};

extern "C" {

int rlsb_qubo_create(const float* q, int32_t num_vars, int32_t device, rlsb_qubo_t** out, void* stream) {
  using namespace rlsb;
  RLSB_REQUIRE(out != nullptr, RLSB_ERR_INVALID, "qubo_create: out is null");
  *out = nullptr;
  RLSB_REQUIRE(q != nullptr && num_vars > 0 && device >= 0, RLSB_ERR_INVALID, "qubo_create: bad argument");
  auto* h = new rlsb_qubo();
  h->n = num_vars, h->np = (num_vars + kQM - 1) / kQM * kQM, h->device = device;
  const size_t bytes = (size_t)3 * h->np * h->np * sizeof(__nv_bfloat16);
  cudaError_t e = cudaMalloc(&h->limbs, bytes);
  if (e != cudaSuccess) {
    set_error("qubo_create: cudaMalloc(%zu) -> %s", bytes, cudaGetErrorString(e));
    delete h;
    return RLSB_ERR_CUDA;
  }
  const int64_t total = (int64_t)h->np * h->np;
  qubo_split_kernel<<<(unsigned)((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(q, num_vars, h->np,
                                                                                                   h->limbs);
  int rc = RLSB_OK;
  if (cudaGetLastError() != cudaSuccess) rc = RLSB_ERR_CUDA, set_error("qubo_create: split kernel launch failed");
  if (rc == RLSB_OK) rc = make_map(&h->map_a, h->limbs, (uint64_t)3 * h->np, (uint64_t)h->np, kQM);
  if (rc != RLSB_OK) {
    cudaFree(h->limbs);
    delete h;
    return rc;
  }
  *out = h;
  return RLSB_OK;
}

int rlsb_qubo_destroy(rlsb_qubo_t* h) {
  if (!h) return RLSB_OK;
  if (h->limbs) cudaFree(h->limbs);
  delete h;
  return RLSB_OK;
}

int32_t rlsb_qubo_padded_vars(const rlsb_qubo_t* h) { return h ? h->np : 0; }

int64_t rlsb_qubo_workspace_bytes(const rlsb_qubo_t* h, int64_t num_chains) {
  using namespace rlsb;
  if (!h || num_chains < 0) return -1;
  const int64_t cp = (num_chains + kQN - 1) / kQN * kQN;
  // bf16 chain-major X, energy partials, split-K partials of the sweeps (<= one CTA per SM) + group counters
  return cp * h->np * 2 + (int64_t)(h->np / kQM) * cp * 4 + (int64_t)kNumSMs * kSM * kSB * 4 + 1024 + 512;
}

int rlsb_qubo_energy(const rlsb_qubo_t* h, const float* x, int64_t num_chains, float* energy, void* workspace,
                     void* stream) {
  using namespace rlsb;
  RLSB_REQUIRE(h != nullptr, RLSB_ERR_INVALID, "qubo_energy: null handle");
  RLSB_REQUIRE(num_chains >= 0, RLSB_ERR_INVALID, "qubo_energy: negative num_chains");
  if (num_chains == 0) return RLSB_OK;
  RLSB_REQUIRE(x && energy && workspace, RLSB_ERR_INVALID, "qubo_energy: null pointer");
  RLSB_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255u) == 0, RLSB_ERR_INVALID,
               "qubo_energy: workspace must be 256-byte aligned");
  auto st = static_cast<cudaStream_t>(stream);
  const int64_t cp = (num_chains + kQN - 1) / kQN * kQN;
  auto* xt = static_cast<__nv_bfloat16*>(workspace);
  auto* partial = reinterpret_cast<float*>(static_cast<char*>(workspace) + ((cp * h->np * 2 + 255) / 256 * 256));
  dim3 tb(32, 8), tg((unsigned)(cp / 32), (unsigned)(h->np / 32));
  qubo_xt_kernel<<<tg, tb, 0, st>>>(x, h->n, num_chains, h->np, cp, xt);
  RLSB_LAUNCH_OK();
  CUtensorMap map_b;
  if (int rc = make_map(&map_b, xt, (uint64_t)cp, (uint64_t)h->np, kQN)) return rc;
  const size_t smem = sizeof(QuboSmem) + 1024;
  RLSB_CUDA_OK(cudaFuncSetAttribute(qubo_energy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)(h->np / kQM), (unsigned)(cp / kQN));
  qubo_energy_kernel<<<grid, kQThreads, smem, st>>>(h->map_a, map_b, xt, h->np, (int)cp, partial);
  RLSB_LAUNCH_OK();
  qubo_reduce_kernel<<<(unsigned)((num_chains + 255) / 256), 256, 0, st>>>(partial, h->np / kQM, cp, num_chains, energy);
  RLSB_LAUNCH_OK();
  return RLSB_OK;
}

int rlsb_qubo_sweeps(const rlsb_qubo_t* h, const float* q, float* x, int64_t num_chains, int32_t num_sweeps,
                     int32_t binary, void* workspace, void* stream) {
  using namespace rlsb;
  RLSB_REQUIRE(h != nullptr, RLSB_ERR_INVALID, "qubo_sweeps: null handle");
  RLSB_REQUIRE(num_chains >= 0 && num_sweeps >= 0, RLSB_ERR_INVALID, "qubo_sweeps: negative size");
  if (num_chains == 0 || num_sweeps == 0) return RLSB_OK;
  RLSB_REQUIRE(q && x && workspace, RLSB_ERR_INVALID, "qubo_sweeps: null pointer");
  RLSB_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255u) == 0, RLSB_ERR_INVALID,
               "qubo_sweeps: workspace must be 256-byte aligned");
  auto st = static_cast<cudaStream_t>(stream);
  const int64_t cp = (num_chains + kQN - 1) / kQN * kQN;
  auto* xt = static_cast<__nv_bfloat16*>(workspace);
  dim3 tb(32, 8), tg((unsigned)(cp / 32), (unsigned)(h->np / 32));
  qubo_xt_kernel<<<tg, tb, 0, st>>>(x, h->n, num_chains, h->np, cp, xt);
  RLSB_LAUNCH_OK();
  CUtensorMap map_x, map_q;
  if (int rc = make_map(&map_x, xt, (uint64_t)cp, (uint64_t)h->np, kSM)) return rc;
  if (int rc = make_map(&map_q, h->limbs, (uint64_t)3 * h->np, (uint64_t)h->np, kSB)) return rc;
  const size_t smem = sizeof(SweepSmem) + 1024;
  RLSB_CUDA_OK(cudaFuncSetAttribute(qubo_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const unsigned groups = (unsigned)((num_chains + kSM - 1) / kSM);
  // Few chains (config 5 hands each GPU 1024): one CTA per 128 chains would leave most SMs idle, so the K range of
  // every row block is split over ks CTAs per chain group, groups * ks <= SMs (they synchronise inside the launch)
  int ks = 1;
  if (!(debug_flags() & RLSB_DEBUG_QUBO_NO_SPLITK))
    for (int cand : {32, 16, 8, 4})
      if ((int64_t)groups * cand <= kNumSMs && (h->np / kQK) % cand == 0 && (h->np / kQK) / cand >= 2) {
        ks = cand;
        break;
      }
  char* tail = static_cast<char*>(workspace) + cp * h->np * 2 + (int64_t)(h->np / kQM) * cp * 4;
  tail = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(tail) + 255) & ~uintptr_t(255));
  float* part = reinterpret_cast<float*>(tail);
  unsigned* counter = reinterpret_cast<unsigned*>(tail + (size_t)kNumSMs * kSM * kSB * 4);
  if (ks > 1) RLSB_CUDA_OK(cudaMemsetAsync(counter, 0, 1024, st));
  unsigned launches = 0;
  for (int sweep = 0; sweep < num_sweeps; ++sweep)
    for (int m0 = 0; m0 < h->n; m0 += kSB) {
      ++launches;
      qubo_sweep_kernel<<<dim3(groups, (unsigned)ks), kQThreads, smem, st>>>(map_x, map_q, q, h->n, h->np, m0, num_chains,
                                                                            binary ? 1 : 0, xt, x, ks, part, counter,
                                                                            (unsigned)ks * launches);
      RLSB_LAUNCH_OK();
    }
  return RLSB_OK;
}

}  // extern "C"
