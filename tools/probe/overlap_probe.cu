// Probe: can a kernel on a non-blocking side stream run NEXT TO a kernel that spins on a flag the side kernel sets?
// Variants: spinner on the legacy default stream / on a created stream; side kernel gated by an event or not;
// big spinner CTAs (512 threads, 96 regs, ~175 KB smem) that leave room for a 256-thread, 64-reg block or not.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <chrono>
#include <thread>
#include <unistd.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__global__ void __maxnreg__(96) spinner(volatile unsigned* flag, unsigned want, unsigned* out) {
  extern __shared__ unsigned sm[];
  if (threadIdx.x == 0) {
    unsigned v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
      if (v < want) __nanosleep(200);
    } while (v < want);
    sm[0] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) atomicAdd(out, 1u);
}

__global__ void __maxnreg__(64) setter(unsigned* flag) {
  __shared__ unsigned q[1280];
  q[threadIdx.x] = threadIdx.x;
  __syncthreads();
  if (threadIdx.x == 0) { __threadfence(); atomicAdd(flag, 1u); }
}

static bool wait_done(cudaStream_t a, cudaStream_t b, double secs) {
  auto t0 = std::chrono::steady_clock::now();
  for (;;) {
    cudaError_t ea = cudaStreamQuery(a), eb = cudaStreamQuery(b);
    if (ea == cudaSuccess && eb == cudaSuccess) return true;
    if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > secs) return false;
    std::this_thread::sleep_for(std::chrono::milliseconds(5));
  }
}

int main(int argc, char** argv) {
  unsigned *flag, *out;
  CK(cudaMalloc(&flag, 8)); out = flag + 1;
  cudaStream_t side, main_s;
  int lo, hi;
  CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  CK(cudaStreamCreateWithPriority(&side, cudaStreamNonBlocking, hi));
  CK(cudaStreamCreateWithFlags(&main_s, cudaStreamNonBlocking));
  cudaEvent_t fork;
  CK(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming));
  CK(cudaFuncSetAttribute(spinner, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  struct V { const char* name; bool legacy; bool event; int spin_ctas; int smem; int set_blocks; };
  V vs[] = {
    {"legacy spinner 128 CTAs x 175KB, setter 148 blocks, event gate", true, true, 128, 175 * 1024, 148},
    {"legacy spinner 128 CTAs x 175KB, setter 148 blocks, no event", true, false, 128, 175 * 1024, 148},
    {"created-stream spinner 128 CTAs x 175KB, setter 148, event gate", false, true, 128, 175 * 1024, 148},
    {"legacy spinner 148 CTAs x 175KB (every SM taken), setter 148, event", true, true, 148, 175 * 1024, 148},
    {"created-stream spinner 148 CTAs x 175KB, setter 148, event", false, true, 148, 175 * 1024, 148},
    {"created-stream spinner 148 CTAs x 20KB, setter 148, event", false, true, 148, 20 * 1024, 148},
    {"created-stream spinner 128 CTAs x 175KB, setter 148, no event", false, false, 128, 175 * 1024, 148},
    {"created-stream spinner 16 CTAs x 20KB, setter 16, no event", false, false, 16, 20 * 1024, 16},
    {"legacy spinner 16 CTAs x 20KB, setter 16, no event", true, false, 16, 20 * 1024, 16},
    {"legacy spinner 16 CTAs x 20KB, setter 16, event", true, true, 16, 20 * 1024, 16},
  };
  const int only = argc > 1 ? atoi(argv[1]) : -1;
  const bool preload = argc > 2 && atoi(argv[2]) != 0;      // launch both kernels once before the concurrent use
  if (preload) {                                             // (CUDA lazy module loading: a first launch may need a context sync)
    CK(cudaMemset(flag, 0, 8));
    setter<<<1, 256, 0, side>>>(flag);
    spinner<<<1, 512, 20 * 1024, main_s>>>(flag, 1u, out);
    CK(cudaDeviceSynchronize());
    printf("[kernels preloaded] ");
  }
  int vi = -1;
  for (auto& v : vs) {
    if (++vi != only && only >= 0) continue;
    CK(cudaMemset(flag, 0, 8));
    CK(cudaDeviceSynchronize());
    cudaStream_t st = v.legacy ? (cudaStream_t)0 : main_s;
    if (v.event) CK(cudaEventRecord(fork, st));
    spinner<<<v.spin_ctas, 512, v.smem, st>>>(flag, (unsigned)v.set_blocks, out);
    CK(cudaGetLastError());
    if (v.event) CK(cudaStreamWaitEvent(side, fork, 0));
    setter<<<v.set_blocks, 256, 0, side>>>(flag);
    CK(cudaGetLastError());
    const bool ok = wait_done(st, side, 3.0);
    printf("%-75s : %s\n", v.name, ok ? "completes" : "DEADLOCK");
    fflush(stdout);
    if (!ok) { printf("(stopping: the context is wedged)\n"); fflush(stdout); _exit(0); }
  }
  return 0;
}
