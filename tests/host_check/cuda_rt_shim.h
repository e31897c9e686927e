// TEST INFRASTRUCTURE: the slice of the CUDA runtime API that diffskill_b200/csrc/engine.cu uses, for the CPU emulation
// build of the engine (tests/host_check/make_emu.py -> libdiffskill_mpm_emu.so).  "Device" memory is host memory, every
// stream is synchronous, events are timestamps, graphs are never captured (the emulation build runs with DSK_NO_GRAPHS).
#pragma once
#include <chrono>
#include <cstdlib>
#include <cstring>

typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorUnknown = 999 };
typedef struct hc_stream_* cudaStream_t;
typedef struct hc_event_ { double t; }* cudaEvent_t;
typedef struct hc_graph_* cudaGraph_t;
typedef struct hc_graph_exec_* cudaGraphExec_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaStreamCaptureModeThreadLocal = 1 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
struct cudaDeviceProp {
  char name[256];
  int major, minor, multiProcessorCount;
  size_t totalGlobalMem;
};
static inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated CUDA error"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
  std::memset(p, 0, sizeof *p);
  std::strcpy(p->name, "CPU emulation of sm_100a");
  p->major = 10;
  p->multiProcessorCount = 148;
  p->totalGlobalMem = (size_t)8 << 30;
  return cudaSuccess;
}
template <class T>
static inline cudaError_t cudaMalloc(T** p, size_t n) {
  *p = (T*)std::malloc(n ? n : 1);
  return *p ? cudaSuccess : cudaErrorUnknown;
}
static inline cudaError_t cudaFree(void* p) { std::free(p); return cudaSuccess; }
static inline cudaError_t cudaMemset(void* p, int v, size_t n) { std::memset(p, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t = nullptr) { std::memset(p, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) {
  std::memmove(d, s, n);
  return cudaSuccess;
}
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (cudaStream_t)std::malloc(1); return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t s) { std::free(s); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new hc_event_{0.0}; return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr) {
  e->t = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
  return cudaSuccess;
}
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)(b->t - a->t); return cudaSuccess; }
template <class F>
static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
// graphs: never used by the emulation build (use_graphs is forced off); present so that the capture code compiles
static inline cudaError_t cudaStreamBeginCapture(cudaStream_t, int) { return cudaErrorUnknown; }
static inline cudaError_t cudaStreamEndCapture(cudaStream_t, cudaGraph_t*) { return cudaErrorUnknown; }
static inline cudaError_t cudaGraphInstantiate(cudaGraphExec_t*, cudaGraph_t, unsigned long long) { return cudaErrorUnknown; }
static inline cudaError_t cudaGraphInstantiate(cudaGraphExec_t*, cudaGraph_t, void*, void*, size_t) { return cudaErrorUnknown; }
static inline cudaError_t cudaGraphLaunch(cudaGraphExec_t, cudaStream_t) { return cudaErrorUnknown; }
static inline cudaError_t cudaGraphDestroy(cudaGraph_t) { return cudaSuccess; }
static inline cudaError_t cudaGraphExecDestroy(cudaGraphExec_t) { return cudaSuccess; }
