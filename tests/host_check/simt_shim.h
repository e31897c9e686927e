// TEST INFRASTRUCTURE: a lock-step emulation of CUDA thread blocks on the CPU, enough to run the device code of
// diffskill_b200/csrc (warp scatters, the substep kernels, the engine's eager launch path) under g++.
//   * every CUDA thread of the block being executed is a fiber (ucontext) of the calling host thread, resumed round-robin;
//     blocks of a launch run one after the other
//   * every *_sync warp intrinsic is a rendezvous of the live lanes of the warp through two barriers (publish, read);
//     __syncthreads a rendezvous of the live threads of the block; a thread that returns from the kernel leaves both
//   * __shared__ is `static` (one block at a time), dynamic shared memory one process-wide buffer (DSK_DYN_SMEM)
//   * deterministic: one host thread, so atomics are plain read-modify-writes
// All call sites in the product use the full mask from converged code, which is what this supports.
#pragma once
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <functional>
#include <mutex>
#include <setjmp.h>
#include <ucontext.h>
#include <vector>

// All CUDA threads of the block being executed are FIBERS (ucontext) of one host thread, resumed round-robin: a barrier
// is "yield until everybody alive has arrived", no locks, deterministic.
struct SimtBarrier {   // rendezvous with a shrinking set of participants
  int alive = 0, waiting = 0;
  uint64_t generation = 0;
  void reset(int n) { alive = n; waiting = 0; }
  inline void arrive_and_wait();
  void drop() {   // the calling fiber has left the kernel
    alive--;
    if (alive > 0 && waiting >= alive) {
      waiting = 0;
      generation++;
    }
  }
};
struct SimtWarp {
  SimtBarrier bar;
  uint32_t slot[32];
  unsigned live = 0;   // lanes that exist in this warp
  void barrier() { bar.arrive_and_wait(); }
};
struct SimtIdx {
  unsigned x, y, z;
};
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
static SimtIdx threadIdx = {0, 0, 0}, blockIdx = {0, 0, 0};   // of the fiber that is running
static dim3 blockDim, gridDim;
static SimtWarp* simt_warp = nullptr;
static SimtBarrier* simt_cta = nullptr;
static int simt_lane_id = 0;
struct SimtNoLock {
  void lock() {}
  void unlock() {}
};
static SimtNoLock simt_atomic_mutex;   // fibers of one host thread: atomics need no lock
static inline void simt_yield();
inline void SimtBarrier::arrive_and_wait() {
  uint64_t g = generation;
  if (++waiting >= alive) {
    waiting = 0;
    generation++;
  } else {
    while (generation == g) simt_yield();
  }
}
static std::vector<char> simt_dyn_smem;   // dynamic shared memory of the running launch

static inline int simt_lane() { return simt_lane_id; }
template <class T>
static inline uint32_t simt_bits(T v) {
  static_assert(sizeof(T) == 4, "32-bit values only");
  uint32_t u;
  std::memcpy(&u, &v, 4);
  return u;
}
template <class T>
static inline T simt_from(uint32_t u) {
  T v;
  std::memcpy(&v, &u, 4);
  return v;
}
// publish one 32-bit value per lane, let f read all of them, then release the slots
template <class T, class F>
static inline auto simt_exchange(T v, F f) -> decltype(f((const uint32_t*)nullptr)) {
  SimtWarp* w = simt_warp;
  w->slot[simt_lane()] = simt_bits(v);
  w->barrier();
  auto r = f((const uint32_t*)w->slot);
  w->barrier();
  return r;
}
static inline unsigned __ballot_sync(unsigned, int pred) {
  unsigned live = simt_warp->live;
  return simt_exchange<int>(pred ? 1 : 0, [live](const uint32_t* s) {
    unsigned b = 0;
    for (int l = 0; l < 32; l++)
      if ((live >> l) & 1u) b |= (s[l] ? 1u : 0u) << l;
    return b;
  });
}
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
static inline int __all_sync(unsigned m, int pred) { return __ballot_sync(m, !pred) == 0; }
static inline unsigned __match_any_sync(unsigned, int key) {
  int lane = simt_lane();
  unsigned live = simt_warp->live;
  return simt_exchange<int>(key, [lane, live](const uint32_t* s) {
    unsigned b = 0;
    for (int l = 0; l < 32; l++)
      if ((live >> l) & 1u) b |= (s[l] == s[lane] ? 1u : 0u) << l;
    return b;
  });
}
static inline int __reduce_max_sync(unsigned, int v) {
  unsigned live = simt_warp->live;
  return simt_exchange<int>(v, [live](const uint32_t* s) {
    int m = INT32_MIN;
    for (int l = 0; l < 32; l++)
      if ((live >> l) & 1u) m = std::max(m, simt_from<int>(s[l]));
    return m;
  });
}
template <class T>
static inline T __shfl_sync(unsigned, T v, int src) {
  return simt_exchange<T>(v, [src](const uint32_t* s) { return simt_from<T>(s[src & 31]); });
}
template <class T>
static inline T __shfl_xor_sync(unsigned, T v, int mask) {
  int lane = simt_lane();
  return simt_exchange<T>(v, [lane, mask](const uint32_t* s) { return simt_from<T>(s[(lane ^ mask) & 31]); });
}
template <class T>
static inline T __shfl_up_sync(unsigned, T v, int d) {
  int lane = simt_lane();
  return simt_exchange<T>(v, [lane, d](const uint32_t* s) { return simt_from<T>(s[lane - d >= 0 ? lane - d : lane]); });
}
template <class T>
static inline T __shfl_down_sync(unsigned, T v, int d) {
  int lane = simt_lane();
  return simt_exchange<T>(v, [lane, d](const uint32_t* s) { return simt_from<T>(s[lane + d < 32 ? lane + d : lane]); });
}
static inline void __syncwarp(unsigned = 0xffffffffu) { simt_warp->barrier(); }
static inline void __syncthreads() { simt_cta->arrive_and_wait(); }
static inline void __threadfence() {}
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __clz(unsigned v) { return v ? __builtin_clz(v) : 32; }
static inline float4 atomicAdd(float4* a, float4 v) {
  std::lock_guard<SimtNoLock> g(simt_atomic_mutex);
  float4 old = *a;
  a->x += v.x; a->y += v.y; a->z += v.z; a->w += v.w;
  return old;
}
template <class T>
static inline T atomicAdd(T* a, T v) {
  std::lock_guard<SimtNoLock> g(simt_atomic_mutex);
  T old = *a;
  *a = old + v;
  return old;
}
template <class T>
static inline T atomicExch(T* a, T v) {
  std::lock_guard<SimtNoLock> g(simt_atomic_mutex);
  T old = *a;
  *a = v;
  return old;
}
template <class T>
static inline T atomicMin(T* a, T v) {
  std::lock_guard<SimtNoLock> g(simt_atomic_mutex);
  T old = *a;
  *a = std::min(old, v);
  return old;
}
template <class T>
static inline T atomicMax(T* a, T v) {
  std::lock_guard<SimtNoLock> g(simt_atomic_mutex);
  T old = *a;
  *a = std::max(old, v);
  return old;
}
#define __global__
#define __shared__ static
#define __launch_bounds__(...)
#define __grid_constant__

// Runs `kernel` for every thread of every block of the launch.  The threads of a block are host threads that live for
// the whole launch and walk the blocks together.
static inline void simt_launch_impl(dim3 grid, dim3 block, size_t smem, const std::function<void()>& kernel);
static inline void simt_launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& kernel) {
  if (!std::getenv("HC_PROFILE")) return simt_launch_impl(grid, block, smem, kernel);
  auto t0 = std::chrono::steady_clock::now();
  simt_launch_impl(grid, block, smem, kernel);
  double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  std::fprintf(stderr, "[simt] grid %u x %u, block %u x %u: %.1f ms\n", grid.x, grid.y, block.x, block.y, ms);
}
struct SimtFiber {
  ucontext_t ctx;      // only to enter the fiber the first time; afterwards _setjmp / _longjmp (no signal-mask system calls)
  jmp_buf jb;
  bool started = false;
  char* stack = nullptr;
  bool done = true;
  SimtIdx tid;
  int lane = 0, warp = 0;
};
static const size_t SIMT_STACK = 256 << 10;
static std::vector<SimtFiber> simt_fibers;
static jmp_buf simt_sched_jb;
static int simt_current = -1;
static const std::function<void()>* simt_kernel = nullptr;
static inline void simt_yield() {
  if (!_setjmp(simt_fibers[simt_current].jb)) _longjmp(simt_sched_jb, 1);
}
static void simt_fiber_main() {
  (*simt_kernel)();
  SimtFiber& f = simt_fibers[simt_current];
  simt_warp->bar.drop();
  simt_cta->drop();
  f.done = true;
  _longjmp(simt_sched_jb, 1);
}
static inline void simt_launch_impl(dim3 grid, dim3 block, size_t smem, const std::function<void()>& kernel) {
  const int nt = (int)(block.x * block.y * block.z), nw = (nt + 31) / 32;
  if (smem > simt_dyn_smem.size()) simt_dyn_smem.resize(smem);
  if ((int)simt_fibers.size() < nt) simt_fibers.resize(nt);
  std::vector<SimtWarp> warps(nw);
  SimtBarrier cta;
  blockDim = block;
  gridDim = grid;
  simt_kernel = &kernel;
  simt_cta = &cta;
  for (unsigned bz = 0; bz < grid.z; bz++)
    for (unsigned by = 0; by < grid.y; by++)
      for (unsigned bx = 0; bx < grid.x; bx++) {
        blockIdx.x = bx; blockIdx.y = by; blockIdx.z = bz;
        cta.reset(nt);
        for (int w = 0; w < nw; w++) {
          int lanes = std::min(32, nt - 32 * w);
          warps[w].bar.reset(lanes);
          warps[w].live = lanes == 32 ? 0xffffffffu : ((1u << lanes) - 1u);
        }
        for (int t = 0; t < nt; t++) {
          SimtFiber& f = simt_fibers[t];
          if (!f.stack) f.stack = (char*)std::malloc(SIMT_STACK);
          getcontext(&f.ctx);
          f.ctx.uc_stack.ss_sp = f.stack;
          f.ctx.uc_stack.ss_size = SIMT_STACK;
          f.ctx.uc_link = nullptr;
          makecontext(&f.ctx, simt_fiber_main, 0);
          f.done = false;
          f.started = false;
          f.tid.x = t % block.x;
          f.tid.y = (t / block.x) % block.y;
          f.tid.z = t / (block.x * block.y);
          f.lane = t & 31;
          f.warp = t >> 5;
        }
        int left = nt;
        while (left > 0) {
          for (int t = 0; t < nt; t++) {
            SimtFiber& f = simt_fibers[t];
            if (f.done) continue;
            simt_current = t;
            threadIdx = f.tid;
            simt_lane_id = f.lane;
            simt_warp = &warps[f.warp];
            if (!_setjmp(simt_sched_jb)) {
              if (f.started) {
                _longjmp(f.jb, 1);
              } else {
                f.started = true;
                setcontext(&f.ctx);
              }
            }
            if (simt_fibers[t].done) left--;
          }
        }
      }
  simt_current = -1;
}
// one warp of 32 lanes (tests of warp-level functions)
static inline void simt_run_warp(const std::function<void(int)>& f) {
  simt_launch(dim3(1), dim3(32), 0, [&] { f((int)threadIdx.x); });
}
