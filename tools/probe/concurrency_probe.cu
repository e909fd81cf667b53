// Probe 2: do two INDEPENDENT kernels on two streams overlap on this box at all?
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__global__ void burn(long long cycles, unsigned* out) {
  const long long t0 = clock64();
  while (clock64() - t0 < cycles) {}
  if (threadIdx.x == 0) atomicAdd(out, 1u);
}

int main() {
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, 0));
  int cm = -1;
  CK(cudaDeviceGetAttribute(&cm, cudaDevAttrComputeMode, 0));
  printf("device %s, SMs %d, concurrentKernels %d, asyncEngineCount %d, computeMode %d, computePreemption %d\n", p.name,
         p.multiProcessorCount, p.concurrentKernels, p.asyncEngineCount, cm, p.computePreemptionSupported);
  const char* vars[] = {"CUDA_DEVICE_MAX_CONNECTIONS", "CUDA_LAUNCH_BLOCKING", "CUDA_MPS_PIPE_DIRECTORY", "CUDA_VISIBLE_DEVICES",
                        "NVIDIA_VISIBLE_DEVICES", "CUDA_MODULE_LOADING", "CUDA_INJECTION64_PATH", "NSYS_PROFILING_SESSION_ID"};
  for (const char* v : vars) printf("  %s=%s\n", v, getenv(v) ? getenv(v) : "(unset)");
  unsigned* out;
  CK(cudaMalloc(&out, 4));
  cudaStream_t a, b;
  CK(cudaStreamCreateWithFlags(&a, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&b, cudaStreamNonBlocking));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const long long cyc = 40000000;   // ~20 ms
  burn<<<8, 64, 0, a>>>(1000, out);
  burn<<<8, 64, 0, b>>>(1000, out);
  CK(cudaDeviceSynchronize());
  for (int trial = 0; trial < 2; ++trial) {
    CK(cudaEventRecord(e0, a));
    burn<<<8, 64, 0, a>>>(cyc, out);
    if (trial == 0) burn<<<8, 64, 0, a>>>(cyc, out); else burn<<<8, 64, 0, b>>>(cyc, out);
    CK(cudaStreamSynchronize(b));
    CK(cudaEventRecord(e1, a));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("two 8-block kernels %s: %.1f ms\n", trial == 0 ? "on ONE stream (serial reference)" : "on TWO streams", ms);
  }
  return 0;
}
