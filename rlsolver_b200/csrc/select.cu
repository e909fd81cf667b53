// Row-select operators on the reference's bool [E][N] layout.
//   select_rows : update_xs_by_vs   rlsolver/methods/util_read_data.py:190-202
//                 (the reference uses boolean-mask index_put, which syncs the host for nonzero())
//   pick_best   : pick_xs_by_vs     rlsolver/methods/util_read_data.py:204-216
// Both are pure data movement (HBM-bound: at most 2*N bytes per selected row).
#include "common.cuh"

namespace rlsb {

__device__ __forceinline__ void copy_row(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src, int n, int tid,
                                         int nthreads) {
  const bool vec = ((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 15u) == 0;
  if (vec) {
    const int n16 = n >> 4;
    for (int i = tid; i < n16; i += nthreads)
      reinterpret_cast<uint4*>(dst)[i] = __ldg(reinterpret_cast<const uint4*>(src) + i);
    for (int i = (n16 << 4) + tid; i < n; i += nthreads) dst[i] = src[i];
  } else {
    for (int i = tid; i < n; i += nthreads) dst[i] = src[i];
  }
}

// one warp per row
__global__ void __launch_bounds__(256) select_rows_kernel(uint8_t* __restrict__ xs0, int64_t* __restrict__ vs0,
                                                          const uint8_t* __restrict__ xs1,
                                                          const int64_t* __restrict__ vs1, int64_t num_envs, int n,
                                                          int maximize) {
  const int lane = threadIdx.x & 31;
  const int64_t env = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (env >= num_envs) return;
  const int64_t a = vs0[env], b = vs1[env];
  const bool take = maximize ? (b >= a) : (b <= a);
  if (!take) return;
  copy_row(xs0 + env * (int64_t)n, xs1 + env * (int64_t)n, n, lane, 32);
  if (lane == 0) vs0[env] = b;
}

// rows dst[i] <- rows src[i] of the same batch (evolutionary_replacement, rlsolver/methods/util.py:87-94: the index
// sets are disjoint -- the replaced rows come from the non-elite part of the argsort, the sources from the elite part)
__global__ void __launch_bounds__(256) copy_rows_kernel(uint8_t* __restrict__ xs, int64_t* __restrict__ vs,
                                                        const int64_t* __restrict__ dst, const int64_t* __restrict__ src,
                                                        int64_t count, int64_t num_envs, int n, int32_t* __restrict__ bad) {
  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (i >= count) return;
  const int64_t d = dst[i], s = src[i];
  if (d < 0 || d >= num_envs || s < 0 || s >= num_envs) {
    if (lane == 0) atomicAdd(bad, 1);
    return;
  }
  copy_row(xs + d * (int64_t)n, xs + s * (int64_t)n, n, lane, 32);
  if (lane == 0) vs[d] = vs[s];
}

// one CTA per sim: warp 0 finds the best repeat (lowest index on ties), all threads copy the row
__global__ void __launch_bounds__(128) pick_best_kernel(const uint8_t* __restrict__ xs, const int64_t* __restrict__ vs,
                                                        int num_repeats, int64_t num_sims, int n, int maximize,
                                                        uint8_t* __restrict__ out_xs, int64_t* __restrict__ out_vs) {
  __shared__ int sBest;
  const int64_t sim = blockIdx.x;
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    int64_t best = 0;
    int arg = -1;
    for (int r = lane; r < num_repeats; r += 32) {     // ascending r per lane: strict compare keeps the lowest index
      const int64_t v = vs[(int64_t)r * num_sims + sim];
      if (arg < 0 || (maximize ? v > best : v < best)) best = v, arg = r;
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
      const int64_t ob = __shfl_xor_sync(kFull, best, off);
      const int oa = __shfl_xor_sync(kFull, arg, off);
      const bool better = oa >= 0 && (arg < 0 || (maximize ? ob > best : ob < best) || (ob == best && oa < arg));
      if (better) best = ob, arg = oa;
    }
    if (lane == 0) {
      sBest = arg;
      out_vs[sim] = best;
    }
  }
  __syncthreads();
  const int r = sBest;
  copy_row(out_xs + sim * (int64_t)n, xs + ((int64_t)r * num_sims + sim) * (int64_t)n, n, threadIdx.x, blockDim.x);
}

// The per-rank record of the multi-GPU best-cut exchange: 64-bit key of the best local row followed by that
// row.  key = (value + 2^31) << 32 | (0xFFFFFFFF - global env id), compared UNSIGNED: the bias makes negative
// values (weighted cuts, QUBO energies) order below positive ones, ties go to the lowest env id.  Values are
// saturated to the int32 range (the Python mirror applies the same rule).  One CTA: block-wide max, row copy.
// (best_key / best_key_value: common.cuh)
__global__ void __launch_bounds__(1024) best_record_kernel(const int64_t* __restrict__ vs, const uint8_t* __restrict__ xs,
                                                           int64_t num_envs, int n, int64_t env_offset,
                                                           uint8_t* __restrict__ record) {
  __shared__ unsigned long long sBest[32];
  unsigned long long best = 0;
  for (int64_t e = threadIdx.x; e < num_envs; e += blockDim.x) {
    const unsigned long long key = best_key(vs[e], (unsigned long long)(env_offset + e));
    best = key > best ? key : best;
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const unsigned long long o = __shfl_xor_sync(kFull, best, off);
    best = o > best ? o : best;
  }
  if ((threadIdx.x & 31) == 0) sBest[threadIdx.x >> 5] = best;
  __syncthreads();
  best = sBest[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w) best = sBest[w] > best ? sBest[w] : best;
  const int64_t local = (int64_t)(0xFFFFFFFFull - (best & 0xFFFFFFFFull)) - env_offset;
  if (threadIdx.x == 0) *reinterpret_cast<unsigned long long*>(record) = best;
  const uint8_t* row = xs + local * (int64_t)n;
  for (int i = threadIdx.x; i < n; i += blockDim.x) record[8 + i] = row[i];
}

// the same record from packed tiles (uint32 [ceil(E/32)][Np], bit b of word [t][i] = node i of env 32t + b):
// the row is expanded to the record's one-byte-per-node form, so best_pick_kernel serves both layouts
__global__ void __launch_bounds__(1024) best_record_packed_kernel(const int64_t* __restrict__ vs,
                                                                  const uint32_t* __restrict__ packed, int64_t num_envs,
                                                                  int n, int np, int64_t env_offset,
                                                                  uint8_t* __restrict__ record) {
  __shared__ unsigned long long sBest[32];
  unsigned long long best = 0;
  for (int64_t e = threadIdx.x; e < num_envs; e += blockDim.x) {
    const unsigned long long key = best_key(vs[e], (unsigned long long)(env_offset + e));
    best = key > best ? key : best;
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const unsigned long long o = __shfl_xor_sync(kFull, best, off);
    best = o > best ? o : best;
  }
  if ((threadIdx.x & 31) == 0) sBest[threadIdx.x >> 5] = best;
  __syncthreads();
  best = sBest[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w) best = sBest[w] > best ? sBest[w] : best;
  const int64_t local = (int64_t)(0xFFFFFFFFull - (best & 0xFFFFFFFFull)) - env_offset;
  if (threadIdx.x == 0) *reinterpret_cast<unsigned long long*>(record) = best;
  const uint32_t* tile = packed + (local >> 5) * (int64_t)np;
  const int bit = (int)(local & 31);
  for (int i = threadIdx.x; i < n; i += blockDim.x) record[8 + i] = (uint8_t)((tile[i] >> bit) & 1u);
}

// winner of the gathered records: out[0] = cut, out[1] = global env id, row = its spins
__global__ void __launch_bounds__(256) best_pick_kernel(const uint8_t* __restrict__ gathered, int world, int n,
                                                        int64_t stride, int64_t* __restrict__ out,
                                                        uint8_t* __restrict__ row) {
  unsigned long long best = 0;
  int win = 0;
  for (int r = 0; r < world; ++r) {
    unsigned long long key = 0;
    for (int b = 7; b >= 0; --b) key = (key << 8) | gathered[r * stride + b];     // records are only byte aligned
    if (r == 0 || key > best) best = key, win = r;
  }
  if (threadIdx.x == 0) {
    out[0] = best_key_value(best);
    out[1] = (int64_t)(0xFFFFFFFFull - (best & 0xFFFFFFFFull));
  }
  for (int i = threadIdx.x; i < n; i += blockDim.x) row[i] = gathered[win * stride + 8 + i];
}

}  // namespace rlsb

extern "C" {

int rlsb_best_pick(const uint8_t* gathered, int32_t world, int32_t num_nodes, int64_t* out2, uint8_t* row, void* stream) {
  using namespace rlsb;
  RLSB_REQUIRE(world > 0 && num_nodes >= 0, RLSB_ERR_INVALID, "best_pick: bad shape");
  RLSB_REQUIRE(gathered && out2 && row, RLSB_ERR_INVALID, "best_pick: null pointer");
  best_pick_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(gathered, world, num_nodes, 8 + (int64_t)num_nodes,
                                                                     out2, row);
  RLSB_LAUNCH_OK();
  return RLSB_OK;
}

int rlsb_best_pick_strided(const uint8_t* gathered, int32_t world, int32_t num_nodes, int64_t stride, int64_t* out2,
                           uint8_t* row, void* stream) {
  using namespace rlsb;
  RLSB_REQUIRE(world > 0 && num_nodes >= 0 && stride >= 8 + (int64_t)num_nodes, RLSB_ERR_INVALID,
               "best_pick_strided: bad shape");
  RLSB_REQUIRE(gathered && out2 && row, RLSB_ERR_INVALID, "best_pick_strided: null pointer");
  best_pick_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(gathered, world, num_nodes, stride, out2, row);
  RLSB_LAUNCH_OK();
  return RLSB_OK;
}

int rlsb_best_record(const int64_t* vs, const uint8_t* xs, int64_t num_envs, int32_t num_nodes, int64_t env_offset,
                     uint8_t* record, void* stream) {
  using namespace rlsb;
  RLSB_REQUIRE(num_envs > 0 && num_nodes >= 0 && env_offset >= 0 && env_offset + num_envs <= 0xFFFFFFFFll,
               RLSB_ERR_INVALID, "best_record: bad shape (global env ids must fit 32 bits, at least one env)");
  RLSB_REQUIRE(vs && xs && record, RLSB_ERR_INVALID, "best_record: null pointer");
  RLSB_REQUIRE((reinterpret_cast<uintptr_t>(record) & 7u) == 0, RLSB_ERR_INVALID, "best_record: record must be 8-byte aligned");
  best_record_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(vs, xs, num_envs, num_nodes, env_offset, record);
  RLSB_LAUNCH_OK();
  return RLSB_OK;
}

int rlsb_best_record_packed(const int64_t* vs, const uint32_t* packed, int64_t num_envs, int32_t num_nodes,
                            int32_t padded_nodes, int64_t env_offset, uint8_t* record, void* stream) {
  using namespace rlsb;
  RLSB_REQUIRE(num_envs > 0 && num_nodes >= 0 && padded_nodes >= num_nodes && env_offset >= 0 &&
                   env_offset + num_envs <= 0xFFFFFFFFll,
               RLSB_ERR_INVALID, "best_record_packed: bad shape (global env ids must fit 32 bits, at least one env)");
  RLSB_REQUIRE(vs && packed && record, RLSB_ERR_INVALID, "best_record_packed: null pointer");
  RLSB_REQUIRE((reinterpret_cast<uintptr_t>(record) & 7u) == 0, RLSB_ERR_INVALID,
               "best_record_packed: record must be 8-byte aligned");
  best_record_packed_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(vs, packed, num_envs, num_nodes,
                                                                               padded_nodes, env_offset, record);
  RLSB_LAUNCH_OK();
  return RLSB_OK;
}

int rlsb_select_rows(uint8_t* xs0, int64_t* vs0, const uint8_t* xs1, const int64_t* vs1, int64_t num_envs,
                     int32_t num_nodes, int32_t maximize, void* stream) {
  using namespace rlsb;
  RLSB_REQUIRE(num_envs >= 0 && num_nodes >= 0, RLSB_ERR_INVALID, "select_rows: negative size");
  if (num_envs == 0) return RLSB_OK;
  RLSB_REQUIRE(xs0 && vs0 && xs1 && vs1, RLSB_ERR_INVALID, "select_rows: null pointer");
  select_rows_kernel<<<(unsigned)((num_envs + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      xs0, vs0, xs1, vs1, num_envs, num_nodes, maximize);
  RLSB_LAUNCH_OK();
  return RLSB_OK;
}

int rlsb_copy_rows(uint8_t* xs, int64_t* vs, const int64_t* dst_ids, const int64_t* src_ids, int64_t count, int64_t num_envs,
                   int32_t num_nodes, int32_t* bad_ids, void* stream) {
  using namespace rlsb;
  RLSB_REQUIRE(count >= 0 && num_envs >= 0 && num_nodes >= 0, RLSB_ERR_INVALID, "copy_rows: negative size");
  if (count == 0) return RLSB_OK;
  RLSB_REQUIRE(xs && vs && dst_ids && src_ids && bad_ids, RLSB_ERR_INVALID, "copy_rows: null pointer");
  copy_rows_kernel<<<(unsigned)((count + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(xs, vs, dst_ids, src_ids, count,
                                                                                              num_envs, num_nodes, bad_ids);
  RLSB_LAUNCH_OK();
  return RLSB_OK;
}

int rlsb_pick_best(const uint8_t* xs, const int64_t* vs, int32_t num_repeats, int64_t num_sims, int32_t num_nodes,
                   int32_t maximize, uint8_t* out_xs, int64_t* out_vs, void* stream) {
  using namespace rlsb;
  RLSB_REQUIRE(num_repeats >= 1 && num_sims >= 0 && num_nodes >= 0, RLSB_ERR_INVALID, "pick_best: bad shape");
  if (num_sims == 0) return RLSB_OK;
  RLSB_REQUIRE(xs && vs && out_xs && out_vs, RLSB_ERR_INVALID, "pick_best: null pointer");
  pick_best_kernel<<<(unsigned)num_sims, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      xs, vs, num_repeats, num_sims, num_nodes, maximize, out_xs, out_vs);
  RLSB_LAUNCH_OK();
  return RLSB_OK;
}

}  // extern "C"
