// Layout change between the reference's bool [E][N] rows and packed env tiles.
// HBM-bound: reads E*N bytes, writes E*N/8 (pack) or the reverse (unpack).
#include "tile_ops.cuh"

namespace rlsb {

template <int VEC>
__global__ void __launch_bounds__(256) pack_kernel(const uint8_t* __restrict__ xs, int64_t num_envs, int32_t n,
                                                   int32_t np, uint32_t* __restrict__ packed) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int strips = (np + 32 * VEC - 1) / (32 * VEC);
  const int64_t tiles = (num_envs + kTileEnvs - 1) / kTileEnvs;
  const int64_t task = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
  if (task >= tiles * strips) return;
  const int64_t tile = task / strips;
  const int node0 = (int)(task % strips) * 32 * VEC + lane * VEC;
  uint32_t w[VEC];
  pack_strip<VEC>(xs, num_envs, n, tile * kTileEnvs, node0, w);
  uint32_t* dst = packed + tile * np + node0;
  if (VEC == 4) {
    if (node0 < np) *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
  } else {
    if (node0 < np) dst[0] = w[0];
  }
}

template <int VEC>
__global__ void __launch_bounds__(256) unpack_kernel(const uint32_t* __restrict__ packed, int64_t num_envs, int32_t n,
                                                     int32_t np, uint8_t* __restrict__ xs) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int strips = (np + 32 * VEC - 1) / (32 * VEC);
  const int64_t tiles = (num_envs + kTileEnvs - 1) / kTileEnvs;
  const int64_t task = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
  if (task >= tiles * strips) return;
  const int64_t tile = task / strips;
  const int node0 = (int)(task % strips) * 32 * VEC + lane * VEC;
  uint32_t w[VEC];
#pragma unroll
  for (int b = 0; b < VEC; ++b) w[b] = (node0 + b < np) ? __ldg(packed + tile * np + node0 + b) : 0u;
  unpack_strip<VEC>(w, xs, num_envs, n, tile * kTileEnvs, node0);
}

}  // namespace rlsb

extern "C" {

int rlsb_pack_spins(const uint8_t* xs, int64_t num_envs, int32_t num_nodes, int32_t padded_nodes, uint32_t* packed,
                    void* stream) {
  using namespace rlsb;
  RLSB_REQUIRE(num_envs >= 0 && num_nodes >= 0 && padded_nodes >= num_nodes && padded_nodes % 32 == 0,
               RLSB_ERR_INVALID, "pack_spins: bad shape E=%lld N=%d Np=%d", (long long)num_envs, num_nodes,
               padded_nodes);
  if (num_envs == 0 || num_nodes == 0) return RLSB_OK;
  RLSB_REQUIRE(xs && packed, RLSB_ERR_INVALID, "pack_spins: null pointer");
  const int64_t tiles = (num_envs + kTileEnvs - 1) / kTileEnvs;
  auto st = static_cast<cudaStream_t>(stream);
  if (rows_vec4_ok(xs, num_nodes)) {
    const int64_t tasks = tiles * ((padded_nodes + 127) / 128);
    pack_kernel<4><<<(unsigned)((tasks + 7) / 8), 256, 0, st>>>(xs, num_envs, num_nodes, padded_nodes, packed);
  } else {
    const int64_t tasks = tiles * (padded_nodes / 32);
    pack_kernel<1><<<(unsigned)((tasks + 7) / 8), 256, 0, st>>>(xs, num_envs, num_nodes, padded_nodes, packed);
  }
  RLSB_LAUNCH_OK();
  return RLSB_OK;
}

int rlsb_unpack_spins(const uint32_t* packed, int64_t num_envs, int32_t num_nodes, int32_t padded_nodes, uint8_t* xs,
                      void* stream) {
  using namespace rlsb;
  RLSB_REQUIRE(num_envs >= 0 && num_nodes >= 0 && padded_nodes >= num_nodes && padded_nodes % 32 == 0,
               RLSB_ERR_INVALID, "unpack_spins: bad shape E=%lld N=%d Np=%d", (long long)num_envs, num_nodes,
               padded_nodes);
  if (num_envs == 0 || num_nodes == 0) return RLSB_OK;
  RLSB_REQUIRE(xs && packed, RLSB_ERR_INVALID, "unpack_spins: null pointer");
  const int64_t tiles = (num_envs + kTileEnvs - 1) / kTileEnvs;
  auto st = static_cast<cudaStream_t>(stream);
  if (rows_vec4_ok(xs, num_nodes)) {
    const int64_t tasks = tiles * ((padded_nodes + 127) / 128);
    unpack_kernel<4><<<(unsigned)((tasks + 7) / 8), 256, 0, st>>>(packed, num_envs, num_nodes, padded_nodes, xs);
  } else {
    const int64_t tasks = tiles * (padded_nodes / 32);
    unpack_kernel<1><<<(unsigned)((tasks + 7) / 8), 256, 0, st>>>(packed, num_envs, num_nodes, padded_nodes, xs);
  }
  RLSB_LAUNCH_OK();
  return RLSB_OK;
}

}  // extern "C"
