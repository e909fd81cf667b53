// The best-cut exchange of the sharded environment batch as ONE kernel over NVLink peer memory.
//
// rlsolver_b200/dist.py's first form is record kernel -> ncclAllGather -> pick kernel: three launches and a
// collective whose latency (~20 us at two GPUs) is paid behind every 0.24 ms step.  The payload is tiny -- world
// records of 8 + N bytes -- so the exchange is pure latency, and a kernel that stores its record straight into
// every peer's mailbox and then polls its own mailbox needs one launch and one NVLink round trip.
//
// Every rank owns a mailbox in its own HBM (cudaMalloc; the other ranks map it through CUDA IPC, which also enables
// peer access): two banks (call parity) of `world` record slots, the call counter and a time-out counter.  A record
// is 2 + ceil(N / 4) words -- the 64-bit key (common.cuh: best_key), then four spins per word -- and every word
// travels as ONE 8-byte store {word, call number}: 8-byte stores are single-copy atomic over NVLink, so a reader that
// sees the call number in the upper half has the data in the lower half, and the exchange needs no fence and no
// separate flag -- one NVLink traversal (the "LL" scheme of NCCL's latency-bound protocols).  One call, one CTA:
//   1. block-wide max of the keys -> the local record;
//   2. the record is stored into slot [bank][rank] of EVERY mailbox (peer stores over NVLink / NVSwitch);
//   3. two threads per rank poll the key words of slot [bank][r] of the OWN mailbox until they carry this call's number;
//   4. the winner is picked, and only ITS spin words are polled and copied out.
// Two banks suffice: a rank can only enter call s + 2 after it saw every peer's record of call s + 1, and a peer sends
// that only after its kernel of call s -- the last reader of bank s % 2 -- has finished.  The call counter lives
// in the mailbox, so the launch has no per-call arguments and can sit inside a captured CUDA graph.
//
// Like every collective it needs all ranks to make the same sequence of calls.  The polls are bounded (default 20 s,
// RLSB_PEER_TIMEOUT_MS): a peer that never arrives costs a counted time-out (rlsb_peer_exchange_status) and a result
// made of whatever the slots held, never a hung GPU.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace rlsb {

constexpr int kPeerMaxWorld = 64;
constexpr int kPeerHeaderBytes = 128;      // {calls, time-outs, pad}

struct PeerBoxes {
  uint8_t* box[kPeerMaxWorld];             // box[rank] = the own mailbox
};

__device__ __forceinline__ uint2* box_slot(uint8_t* box, int bank, int r, int world, int64_t stride) {
  return reinterpret_cast<uint2*>(box + kPeerHeaderBytes + ((int64_t)bank * world + r) * stride);
}
__device__ __forceinline__ void st_pair_sys(uint2* p, uint32_t data, uint32_t flag) {
  asm volatile("st.relaxed.sys.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(data), "r"(flag) : "memory");
}
__device__ __forceinline__ uint2 ld_pair_sys(const uint2* p) {
  uint2 v;
  asm volatile("ld.relaxed.sys.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint64_t global_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// the data half of *p once its flag half shows `call` (bounded: counts a time-out and returns what is there)
__device__ __forceinline__ uint32_t poll_word(const uint2* p, uint32_t call, uint64_t timeout_ns, uint32_t* timeouts) {
  uint2 v = ld_pair_sys(p);
  if (v.y == call) return v.x;
  const uint64_t t0 = global_ns();
  for (int spin = 0;; ++spin) {
    v = ld_pair_sys(p);
    if (v.y == call) return v.x;
    if ((spin & 63) == 63 && global_ns() - t0 > timeout_ns) {
      atomicAdd(timeouts, 1u);
      return v.x;
    }
  }
}

// xs: bool rows [E][n] or null; packed: tiles uint32 [ceil(E/32)][np] (used when xs is null)
__global__ void __launch_bounds__(1024) peer_best_kernel(PeerBoxes boxes, int rank, int world, const int64_t* __restrict__ vs,
                                                         const uint8_t* __restrict__ xs, const uint32_t* __restrict__ packed,
                                                         int64_t num_envs, int n, int np, int64_t env_offset, int64_t stride,
                                                         uint64_t timeout_ns, int64_t* __restrict__ out,
                                                         uint8_t* __restrict__ row_out) {
  __shared__ unsigned long long sBest[32];
  __shared__ uint32_t sKey[kPeerMaxWorld][2];
  __shared__ int sWin;
  uint8_t* own = boxes.box[rank];
  uint32_t* header = reinterpret_cast<uint32_t*>(own);
  const uint32_t call = header[0] + 1u;            // written back by thread 0 behind a barrier
  const int bank = (int)(call & 1u);

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

  // the record: word 0..1 = key, then four spins per word; every word goes out as {word, call}
  const int words = 2 + (n + 3) / 4;
  const uint8_t* row = xs ? xs + local * (int64_t)n : nullptr;
  const uint32_t* tile = xs ? nullptr : packed + (local >> 5) * (int64_t)np;
  const int bit = (int)(local & 31);
  for (int i = threadIdx.x; i < words; i += blockDim.x) {
    uint32_t w;
    if (i == 0) {
      w = (uint32_t)best;
    } else if (i == 1) {
      w = (uint32_t)(best >> 32);
    } else {
      w = 0;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int node = (i - 2) * 4 + b;
        if (node < n) w |= (uint32_t)(row ? (row[node] != 0) : ((tile[node] >> bit) & 1u)) << (8 * b);
      }
    }
    for (int p = 0; p < world; ++p) st_pair_sys(box_slot(boxes.box[p], bank, rank, world, stride) + i, w, call);
  }

  // every rank's key (two threads per rank), then the winner
  if ((int)threadIdx.x < 2 * world) {
    const int r = threadIdx.x >> 1, w = threadIdx.x & 1;
    sKey[r][w] = poll_word(box_slot(own, bank, r, world, stride) + w, call, timeout_ns, header + 1);
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    unsigned long long key = 0;
    int win = 0;
    for (int r = threadIdx.x; r < world; r += 32) {
      const unsigned long long k = ((unsigned long long)sKey[r][1] << 32) | sKey[r][0];
      if (k > key) key = k, win = r;           // keys are distinct across ranks (they embed the global env id)
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
      const unsigned long long ok = __shfl_xor_sync(kFull, key, off);
      const int ow = __shfl_xor_sync(kFull, win, off);
      if (ok > key) key = ok, win = ow;
    }
    if (threadIdx.x == 0) {
      sWin = win;
      out[0] = best_key_value(key);
      out[1] = (int64_t)(0xFFFFFFFFull - (key & 0xFFFFFFFFull));
      header[0] = call;
    }
  }
  __syncthreads();
  const uint2* src = box_slot(own, bank, sWin, world, stride) + 2;
  for (int i = threadIdx.x; i < (n + 3) / 4; i += blockDim.x) {
    const uint32_t w = poll_word(src + i, call, timeout_ns, header + 1);
#pragma unroll
    for (int b = 0; b < 4; ++b)
      if (4 * i + b < n) row_out[4 * i + b] = (uint8_t)((w >> (8 * b)) & 0xffu);
  }
}

}  // namespace rlsb

struct rlsb_peer_exchange {
  int rank, world, n, device;
  int64_t stride, bytes;
  uint64_t timeout_ns;
  bool connected;
  rlsb::PeerBoxes boxes;
  bool opened[rlsb::kPeerMaxWorld];
};

extern "C" {

int64_t rlsb_peer_exchange_handle_bytes(void) { return (int64_t)sizeof(cudaIpcMemHandle_t); }

int rlsb_peer_exchange_create(int32_t rank, int32_t world, int32_t num_nodes, rlsb_peer_exchange_t** out, uint8_t* handle_out) {
  using namespace rlsb;
  RLSB_REQUIRE(out && handle_out, RLSB_ERR_INVALID, "peer_exchange_create: null pointer");
  RLSB_REQUIRE(world >= 1 && world <= kPeerMaxWorld && rank >= 0 && rank < world && num_nodes >= 0, RLSB_ERR_INVALID,
               "peer_exchange_create: bad shape (rank %d of %d, at most %d ranks)", rank, world, kPeerMaxWorld);
  auto* ex = new rlsb_peer_exchange();
  memset(ex, 0, sizeof(*ex));
  ex->rank = rank, ex->world = world, ex->n = num_nodes;
  ex->stride = (2 + ((int64_t)num_nodes + 3) / 4) * 8;       // 8 bytes per record word: {word, call number}
  ex->bytes = kPeerHeaderBytes + 2 * (int64_t)world * ex->stride;
  const char* ms = getenv("RLSB_PEER_TIMEOUT_MS");
  const long long t = ms ? atoll(ms) : 0;
  ex->timeout_ns = (uint64_t)(t > 0 ? t : 20000) * 1000000ull;
  cudaIpcMemHandle_t h;
  void* p = nullptr;
  if (cudaGetDevice(&ex->device) != cudaSuccess || cudaMalloc(&p, (size_t)ex->bytes) != cudaSuccess ||
      cudaMemset(p, 0, (size_t)ex->bytes) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess ||
      cudaIpcGetMemHandle(&h, p) != cudaSuccess) {
    set_error("peer_exchange_create: %s", cudaGetErrorString(cudaGetLastError()));
    if (p) cudaFree(p);
    delete ex;
    return RLSB_ERR_CUDA;
  }
  ex->boxes.box[rank] = static_cast<uint8_t*>(p);
  memcpy(handle_out, &h, sizeof(h));
  *out = ex;
  return RLSB_OK;
}

// handles: world x rlsb_peer_exchange_handle_bytes(), rank-major (an all-gather of the create() outputs)
int rlsb_peer_exchange_connect(rlsb_peer_exchange_t* ex, const uint8_t* handles) {
  using namespace rlsb;
  RLSB_REQUIRE(ex && handles, RLSB_ERR_INVALID, "peer_exchange_connect: null pointer");
  RLSB_REQUIRE(!ex->connected, RLSB_ERR_INVALID, "peer_exchange_connect: already connected");
  for (int r = 0; r < ex->world; ++r) {
    if (r == ex->rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + (size_t)r * sizeof(h), sizeof(h));
    void* p = nullptr;
    const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      (void)cudaGetLastError();
      set_error("peer_exchange_connect: cannot map the mailbox of rank %d (%s)", r, cudaGetErrorString(e));
      return RLSB_ERR_CUDA;
    }
    ex->boxes.box[r] = static_cast<uint8_t*>(p);
    ex->opened[r] = true;
  }
  ex->connected = true;
  return RLSB_OK;
}

static int peer_launch(rlsb_peer_exchange_t* ex, const int64_t* vs, const uint8_t* xs, const uint32_t* packed, int64_t num_envs,
                       int32_t padded_nodes, int64_t env_offset, int64_t* out2, uint8_t* row, void* stream, const char* what) {
  using namespace rlsb;
  RLSB_REQUIRE(ex && ex->connected, RLSB_ERR_INVALID, "%s: the exchange is not connected", what);
  RLSB_REQUIRE(num_envs > 0 && env_offset >= 0 && env_offset + num_envs <= 0xFFFFFFFFll, RLSB_ERR_INVALID,
               "%s: bad shape (global env ids must fit 32 bits, at least one env)", what);
  RLSB_REQUIRE(vs && (xs || packed) && out2 && row, RLSB_ERR_INVALID, "%s: null pointer", what);
  RLSB_REQUIRE(xs || padded_nodes >= ex->n, RLSB_ERR_INVALID, "%s: padded_nodes below the node count", what);
  peer_best_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(ex->boxes, ex->rank, ex->world, vs, xs, packed, num_envs,
                                                                      ex->n, padded_nodes, env_offset, ex->stride,
                                                                      ex->timeout_ns, out2, row);
  RLSB_LAUNCH_OK();
  return RLSB_OK;
}

int rlsb_peer_exchange_best(rlsb_peer_exchange_t* ex, const int64_t* vs, const uint8_t* xs, int64_t num_envs, int64_t env_offset,
                            int64_t* out2, uint8_t* row, void* stream) {
  return peer_launch(ex, vs, xs, nullptr, num_envs, 0, env_offset, out2, row, stream, "peer_exchange_best");
}

int rlsb_peer_exchange_best_packed(rlsb_peer_exchange_t* ex, const int64_t* vs, const uint32_t* packed, int64_t num_envs,
                                   int32_t padded_nodes, int64_t env_offset, int64_t* out2, uint8_t* row, void* stream) {
  return peer_launch(ex, vs, nullptr, packed, num_envs, padded_nodes, env_offset, out2, row, stream,
                     "peer_exchange_best_packed");
}

// out[0] = calls completed, out[1] = polls that ran into the time-out (synchronises the stream)
int rlsb_peer_exchange_status(rlsb_peer_exchange_t* ex, uint32_t* out, void* stream) {
  using namespace rlsb;
  RLSB_REQUIRE(ex && out, RLSB_ERR_INVALID, "peer_exchange_status: null pointer");
  auto st = static_cast<cudaStream_t>(stream);
  RLSB_CUDA_OK(cudaMemcpyAsync(out, ex->boxes.box[ex->rank], 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  RLSB_CUDA_OK(cudaStreamSynchronize(st));
  return RLSB_OK;
}

int rlsb_peer_exchange_destroy(rlsb_peer_exchange_t* ex) {
  if (!ex) return RLSB_OK;
  (void)cudaDeviceSynchronize();       // the last call's kernel: once it is over no peer touches this mailbox again
  for (int r = 0; r < ex->world; ++r)
    if (ex->opened[r]) (void)cudaIpcCloseMemHandle(ex->boxes.box[r]);
  if (ex->boxes.box[ex->rank]) (void)cudaFree(ex->boxes.box[ex->rank]);
  (void)cudaGetLastError();
  delete ex;
  return RLSB_OK;
}

}  // extern "C"
