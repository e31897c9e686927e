// Micro-benchmark: cost of a dependent kernel->kernel edge inside a CUDA graph, with and without programmatic
// dependent launch (griddepcontrol).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pdl_chain pdl_chain.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s failed: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

template <bool PDL>
__global__ void k_work(float* buf, int n, int iters) {
  if (PDL) asm volatile("griddepcontrol.wait;" ::: "memory");
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    float v = buf[i];
    for (int q = 0; q < iters; q++) v = v * 1.0001f + 0.5f;
    buf[i] = v;
  }
  if (PDL) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
template <bool PDL>
__global__ void k_work_early(float* buf, int n, int iters) {
  if (PDL) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (PDL) asm volatile("griddepcontrol.wait;" ::: "memory");
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    float v = buf[i];
    for (int q = 0; q < iters; q++) v = v * 1.0001f + 0.5f;
    buf[i] = v;
  }
}

template <class K>
static int run(const char* name, K kern, bool pdl, int nodes, int n, int iters, float* buf, cudaStream_t st) {
  cudaGraph_t g;
  cudaGraphExec_t ex;
  CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
  for (int i = 0; i < nodes; i++) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((n + 127) / 128);
    cfg.blockDim = dim3(128);
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl ? 1 : 0;
    CK(cudaLaunchKernelEx(&cfg, kern, buf, n, iters));
  }
  CK(cudaStreamEndCapture(st, &g));
  CK(cudaGraphInstantiate(&ex, g, 0));
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  for (int w = 0; w < 3; w++) CK(cudaGraphLaunch(ex, st));
  CK(cudaStreamSynchronize(st));
  CK(cudaEventRecord(a, st));
  for (int w = 0; w < 5; w++) CK(cudaGraphLaunch(ex, st));
  CK(cudaEventRecord(b, st));
  CK(cudaStreamSynchronize(st));
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  printf("%-28s n=%6d iters=%5d : %.3f us per node\n", name, n, iters, ms * 1000.f / (5.f * nodes));
  cudaGraphExecDestroy(ex); cudaGraphDestroy(g);
  return 0;
}
int main() {
  float* buf;
  CK(cudaMalloc(&buf, 1 << 24));
  CK(cudaMemset(buf, 0, 1 << 24));
  cudaStream_t st;
  CK(cudaStreamCreate(&st));
  int nodes = 2000;
  for (int n : {128, 15707, 128 * 148 * 4}) {
    for (int iters : {0, 500, 2000}) {
      if (run("plain", k_work<false>, false, nodes, n, iters, buf, st)) return 1;
      if (run("pdl (trigger at end)", k_work<true>, true, nodes, n, iters, buf, st)) return 1;
      if (run("pdl (trigger at start)", k_work_early<true>, true, nodes, n, iters, buf, st)) return 1;
    }
  }
  return 0;
}
