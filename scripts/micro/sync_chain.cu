// Micro-benchmark behind the "one persistent kernel per env step" decision: the cost of a phase
// boundary when a substep is  (a) a chain of dependent kernels replayed from a CUDA graph  vs  (b) phases of ONE
// co-resident kernel separated by a grid-wide barrier (cooperative-groups grid.sync, and a hand-written sense-reversing
// barrier on one global counter with ld.acquire spinning).  Every phase does the same dependent FMA chain of `iters`
// steps per thread (a stand-in for the latency-bound particle math of a small scene) plus one read and one write of a
// buffer the next phase reads, so the barrier has real memory ordering to do.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -rdc=false -o sync_chain sync_chain.cu
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s failed: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

__device__ __forceinline__ void phase_work(float* buf, int n, int iters, int ph) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    float v = buf[(i + ph * 97) % n];
    for (int q = 0; q < iters; q++) v = v * 1.0001f + 0.5f;
    buf[i] = v;
  }
}
__global__ void k_phase(float* buf, int n, int iters, int ph) { phase_work(buf, n, iters, ph); }

__global__ void k_coop(float* buf, int n, int iters, int phases) {
  cg::grid_group g = cg::this_grid();
  for (int ph = 0; ph < phases; ph++) {
    phase_work(buf, n, iters, ph);
    g.sync();
  }
}
// sense-reversing barrier: one arrival counter, generation flag polled with ld.acquire.gpu
__device__ __forceinline__ void grid_barrier(unsigned* count, volatile unsigned* gen, unsigned& my_gen) {
  __syncthreads();
  if (threadIdx.x == 0) {
    my_gen += 1;
    __threadfence();
    if (atomicAdd(count, 1) == gridDim.x - 1) {
      *count = 0;
      __threadfence();
      atomicExch((unsigned*)gen, my_gen);
    } else {
      unsigned g;
      do {
        asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(g) : "l"(gen) : "memory");
      } while (g != my_gen);
    }
  }
  __syncthreads();
}
__global__ void k_custom(float* buf, int n, int iters, int phases, unsigned* count, unsigned* gen) {
  unsigned my_gen = *(volatile unsigned*)gen;
  for (int ph = 0; ph < phases; ph++) {
    phase_work(buf, n, iters, ph);
    grid_barrier(count, gen, my_gen);
  }
}

int main() {
  float* buf;
  unsigned* sync;
  CK(cudaMalloc(&buf, 1 << 24));
  CK(cudaMemset(buf, 0, 1 << 24));
  CK(cudaMalloc(&sync, 256));
  CK(cudaMemset(sync, 0, 256));
  cudaStream_t st;
  CK(cudaStreamCreate(&st));
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  const int phases = 1000;
  int nsm = 0;
  CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0));
  printf("SMs %d; %d phases per measurement\n", nsm, phases);
  for (int threads : {192, 256}) {
    for (int ctas_per_sm : {1, 2, 4}) {
      int grid = nsm * ctas_per_sm;
      for (int n : {15707, 47121}) {
        for (int iters : {0, 200, 1000}) {
          float ms;
          // (a) graph of dependent kernels, grid sized like the engine's kernels (one thread per element)
          {
            cudaGraph_t g;
            cudaGraphExec_t ex;
            CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
            for (int ph = 0; ph < phases; ph++) k_phase<<<(n + threads - 1) / threads, threads, 0, st>>>(buf, n, iters, ph);
            CK(cudaStreamEndCapture(st, &g));
            CK(cudaGraphInstantiate(&ex, g, 0));
            for (int w = 0; w < 2; w++) CK(cudaGraphLaunch(ex, st));
            CK(cudaStreamSynchronize(st));
            CK(cudaEventRecord(a, st));
            for (int w = 0; w < 3; w++) CK(cudaGraphLaunch(ex, st));
            CK(cudaEventRecord(b, st));
            CK(cudaStreamSynchronize(st));
            cudaEventElapsedTime(&ms, a, b);
            printf("thr %3d cta/sm %d n %6d iters %4d | graph edge %.3f us/phase", threads, ctas_per_sm, n, iters, ms * 1000.f / (3.f * phases));
            cudaGraphExecDestroy(ex);
            cudaGraphDestroy(g);
          }
          // (b) cooperative kernel, grid.sync
          {
            int ph = phases;
            void* args[] = {&buf, (void*)&n, (void*)&iters, &ph};
            for (int w = 0; w < 2; w++) CK(cudaLaunchCooperativeKernel((void*)k_coop, dim3(grid), dim3(threads), args, 0, st));
            CK(cudaStreamSynchronize(st));
            CK(cudaEventRecord(a, st));
            for (int w = 0; w < 3; w++) CK(cudaLaunchCooperativeKernel((void*)k_coop, dim3(grid), dim3(threads), args, 0, st));
            CK(cudaEventRecord(b, st));
            CK(cudaStreamSynchronize(st));
            cudaEventElapsedTime(&ms, a, b);
            printf(" | grid.sync %.3f", ms * 1000.f / (3.f * phases));
          }
          // (c) hand-written barrier (same co-residency requirement: launched cooperatively)
          {
            int ph = phases;
            unsigned* cnt = sync;
            unsigned* gen = sync + 32;
            void* args[] = {&buf, (void*)&n, (void*)&iters, &ph, &cnt, &gen};
            for (int w = 0; w < 2; w++) CK(cudaLaunchCooperativeKernel((void*)k_custom, dim3(grid), dim3(threads), args, 0, st));
            CK(cudaStreamSynchronize(st));
            CK(cudaEventRecord(a, st));
            for (int w = 0; w < 3; w++) CK(cudaLaunchCooperativeKernel((void*)k_custom, dim3(grid), dim3(threads), args, 0, st));
            CK(cudaEventRecord(b, st));
            CK(cudaStreamSynchronize(st));
            cudaEventElapsedTime(&ms, a, b);
            printf(" | custom barrier %.3f\n", ms * 1000.f / (3.f * phases));
          }
        }
      }
    }
  }
  return 0;
}
