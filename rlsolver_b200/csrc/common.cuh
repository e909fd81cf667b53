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

void set_error(const char* fmt, ...);

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

// Device-side view of the graph store (all pointers device memory).
struct GraphDev {
  int32_t n;        // nodes
  int32_t np;       // padded nodes (multiple of 32)
  int32_t m;        // original edges (len(mygraph))
  int32_t md;       // listed edges (m or 2m)
  int32_t mf;       // full-neighbourhood slots
  int32_t levels;   // dependency levels of the in-order sweep
  int32_t bidir;
  const int32_t* edge_u;      // [m] original edge list
  const int32_t* edge_v;
  const int32_t* listed_ptr;  // [n+1]
  const int32_t* listed_col;  // [md]  == n1_ids of the reference
  const int32_t* listed_row;  // [md]  == n0_ids of the reference
  const int32_t* full_ptr;    // [n+1]
  const int32_t* full_col;    // [mf]
  const int32_t* level_ptr;   // [levels+1]
  const int32_t* level_nodes; // [n]
};

const GraphDev* graph_dev(const rlsb_graph_t* g);   // nullptr if host-only
int graph_device_id(const rlsb_graph_t* g);

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

__device__ __forceinline__ int warp_sum(int v) {
  return __reduce_add_sync(kFull, v);
}
#endif

}  // namespace rlsb
