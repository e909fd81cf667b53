// Shared device/host helpers for the rlsolver_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/rlsolver_b200.h"

namespace rlsb {

constexpr int kTileEnvs = 32;      // environments per packed word
constexpr int kNumSMs = 148;       // B200
constexpr unsigned kFull = 0xffffffffu;
constexpr int kMaxTileNodes = 49152;   // padded nodes a shared-memory tile may hold (uint16 ids, <= 192 KB of words)

void set_error(const char* fmt, ...);
int debug_flags();   // rlsb_debug_flags word (environment read once at load time)

#define RLSB_CUDA_OK(expr)                                                              \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) {                                                            \
      rlsb::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return RLSB_ERR_CUDA;                                                             \
    }                                                                                   \
  } while (0)

#define RLSB_REQUIRE(cond, code, ...)  \
  do {                                 \
    if (!(cond)) {                     \
      rlsb::set_error(__VA_ARGS__);    \
      return (code);                   \
    }                                  \
  } while (0)

#define RLSB_LAUNCH_OK() RLSB_CUDA_OK(cudaGetLastError())

// RLSB_REQUIRE that runs `cleanup` before returning
#define RLSB_REQUIRE_CLEAN(cond, cleanup, code, ...) \
  do {                                               \
    if (!(cond)) {                                   \
      rlsb::set_error(__VA_ARGS__);                  \
      cleanup;                                       \
      return (code);                                 \
    }                                                \
  } while (0)

// Sliced-ELL neighbour lists ("SELL-32"): a slice is 32 node slots (one per lane).  Column ids
// are stored in blocks of 4 rounds, lane-major inside a block, so a lane reads 4 neighbour ids
// with one 8-byte load and a warp reads 256 contiguous bytes.  Short rows are padded with the
// node's own id (word ^ word == 0: padding never counts).
struct SellDev {
  int32_t num_slices;
  const int32_t* off;     // [num_slices+1] offsets into col, in blocks (4 rounds x 32 lanes = 128 ids)
  const uint16_t* node;   // [num_slices*32] node id of every slot, 0xFFFF = inactive slot
  const uint16_t* half;   // [num_slices*32] floor(degree / 2) of the slot's node (sweep only)
  const uint16_t* col;    // [off[num_slices]*128]
};

// Device-side view of the graph store (all pointers device memory).
struct GraphDev {
  int32_t n;        // nodes
  int32_t np;       // padded nodes (multiple of 32)
  int32_t m;        // original edges (len(mygraph))
  int32_t md;       // listed edges (m or 2m)
  int32_t mf;       // full-neighbourhood slots
  int32_t levels;   // dependency levels of the in-order sweep
  int32_t bidir;
  int32_t max_listed_deg;
  int32_t max_full_deg;
  const uint32_t* edge_pair;  // [ceil4(m)] u | v << 16 (original edge list, padded with 0; tile kernels)
  const int32_t* listed_ptr;  // [n+1]
  const int32_t* listed_col;  // [md]  == n1_ids of the reference
  const int32_t* listed_row;  // [md]  == n0_ids of the reference
  const int32_t* listed_deg;  // [np]  == n0_num_n1 (0 for padding nodes)
  const int32_t* full_ptr;    // [n+1]
  const int32_t* full_col;    // [mf]
  SellDev listed;             // listed neighbours, natural node order (slot = node)
  // full neighbours in sweep order (dependency level, then degree-descending) as ONE contiguous
  // blob {level_slice i32[levels+1], off i32[S+1], node u16[32S], half u16[32S], col u16[..]},
  // every part 16-byte aligned, so a CTA can stage it into shared memory with one bulk copy
  const char* sweep_blob;
  int32_t sweep_blob_bytes, num_sweep_slices, max_level_slices;
  int32_t sweep_lvs, sweep_off, sweep_node, sweep_half, sweep_col;   // byte offsets inside the blob
  // weighted objective (null / 0 when every weight is 1): see rlsb_graph::wpair / wmeta / full_w / wdeg
  const uint32_t* wpair;
  const int4* wmeta;           // [wbuckets] {first quad, quads, signed scale, edges}
  const int32_t* full_w;       // [mf]
  const int32_t* wdeg;         // [np]
  int32_t wbuckets;
};

// the sweep structure seen through a base pointer (global blob or its shared-memory copy)
struct SweepView {
  const int32_t* level_slice;
  SellDev sell;
};
__host__ __device__ inline SweepView sweep_view(const GraphDev& g, const char* base) {
  SweepView v;
  v.level_slice = reinterpret_cast<const int32_t*>(base + g.sweep_lvs);
  v.sell.num_slices = g.num_sweep_slices;
  v.sell.off = reinterpret_cast<const int32_t*>(base + g.sweep_off);
  v.sell.node = reinterpret_cast<const uint16_t*>(base + g.sweep_node);
  v.sell.half = reinterpret_cast<const uint16_t*>(base + g.sweep_half);
  v.sell.col = reinterpret_cast<const uint16_t*>(base + g.sweep_col);
  return v;
}

const GraphDev* graph_dev(const rlsb_graph_t* g);   // nullptr if host-only or not tileable
int graph_device_id(const rlsb_graph_t* g);
// validates the handle for the tile kernels; sets the error text and returns the status
int graph_check(const rlsb_graph_t* g, const GraphDev** out, const char* what);
int cut_warps_for(int64_t m, int max_warps);

#ifdef __CUDACC__
// 32x32 bit-matrix transpose across a warp.  In: lane r holds row r (bit c = B[r][c]).
// Out: lane r holds column r (bit c = B[c][r]).  5 shuffle stages.
__device__ __forceinline__ uint32_t transpose32(uint32_t x, int lane) {
#pragma unroll
  for (int j = 16; j >= 1; j >>= 1) {
    // m selects bit positions whose j-bit is clear
    const uint32_t m = (j == 16) ? 0x0000ffffu : (j == 8) ? 0x00ff00ffu : (j == 4) ? 0x0f0f0f0fu
                     : (j == 2) ? 0x33333333u : 0x55555555u;
    const uint32_t y = __shfl_xor_sync(kFull, x, j);
    x = (lane & j) ? ((x & ~m) | ((y >> j) & m)) : ((x & m) | ((y & m) << j));
  }
  return x;
}

// ---- bulk asynchronous copy global -> shared (TMA engine, 1-D) completing on an mbarrier
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// dst/src 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  }
}

// key of the multi-GPU best-cut exchange (select.cu, peer_exchange.cu): (value + 2^31) << 32 | (0xFFFFFFFF - global env
// id), compared unsigned; values saturate to the int32 range
__device__ __forceinline__ unsigned long long best_key(int64_t v, unsigned long long gid) {
  v = v > 2147483647ll ? 2147483647ll : (v < -2147483648ll ? -2147483648ll : v);
  return ((unsigned long long)(v + 2147483648ll) << 32) | (0xFFFFFFFFull - gid);
}
__device__ __forceinline__ int64_t best_key_value(unsigned long long key) {
  return (int64_t)(key >> 32) - 2147483648ll;
}

__device__ __forceinline__ int warp_sum(int v) {
  return __reduce_add_sync(kFull, v);
}

// streaming (read-once) global loads: do not allocate in L1
__device__ __forceinline__ float4 ldg_stream4(const float* p) {
  float4 v;
  asm("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float ldg_stream(const float* p) {
  float v;
  asm("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
#endif

}  // namespace rlsb
