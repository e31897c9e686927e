// TEST INFRASTRUCTURE: a lock-step emulation of ONE warp on the CPU, enough to run the warp-level device functions of
// diffskill_b200/csrc/kernels_common.cuh (warp_scatter27 / warp_scatter9: ballots, match.any, shuffles, vector reductions)
// under g++.  Every lane is a host thread; every *_sync intrinsic is a rendezvous of the 32 lanes through two barriers
// (publish, read).  All call sites in the product use the full mask from converged code, which is what this supports.
#pragma once
#include <condition_variable>
#include <cstdint>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

struct SimtWarp {
  std::mutex m;
  std::condition_variable cv;
  int waiting = 0;
  uint64_t generation = 0;
  uint32_t slot[32];
  void barrier() {
    std::unique_lock<std::mutex> lk(m);
    uint64_t g = generation;
    if (++waiting == 32) {
      waiting = 0;
      generation++;
      cv.notify_all();
    } else {
      cv.wait(lk, [&] { return generation != g; });
    }
  }
};
struct SimtIdx {
  int x, y, z;
};
static thread_local SimtIdx threadIdx = {0, 0, 0};
static thread_local SimtWarp* simt_warp = nullptr;
static std::mutex simt_atomic_mutex;

static inline int simt_lane() { return threadIdx.x & 31; }
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
  return simt_exchange<int>(pred ? 1 : 0, [](const uint32_t* s) {
    unsigned b = 0;
    for (int l = 0; l < 32; l++) b |= (s[l] ? 1u : 0u) << l;
    return b;
  });
}
static inline unsigned __match_any_sync(unsigned, int key) {
  int lane = simt_lane();
  return simt_exchange<int>(key, [lane](const uint32_t* s) {
    unsigned b = 0;
    for (int l = 0; l < 32; l++) b |= (s[l] == s[lane] ? 1u : 0u) << l;
    return b;
  });
}
static inline int __reduce_max_sync(unsigned, int v) {
  return simt_exchange<int>(v, [](const uint32_t* s) {
    int m = simt_from<int>(s[0]);
    for (int l = 1; l < 32; l++) m = std::max(m, simt_from<int>(s[l]));
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
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __clz(unsigned v) { return v ? __builtin_clz(v) : 32; }
static inline float4 atomicAdd(float4* a, float4 v) {
  std::lock_guard<std::mutex> g(simt_atomic_mutex);
  float4 old = *a;
  a->x += v.x; a->y += v.y; a->z += v.z; a->w += v.w;
  return old;
}
static inline int atomicAdd(int* a, int v) {
  std::lock_guard<std::mutex> g(simt_atomic_mutex);
  int old = *a;
  *a += v;
  return old;
}
static inline int atomicExch(int* a, int v) {
  std::lock_guard<std::mutex> g(simt_atomic_mutex);
  int old = *a;
  *a = v;
  return old;
}
static inline unsigned long long atomicMin(unsigned long long* a, unsigned long long v) { return *a = std::min(*a, v); }
static inline unsigned long long atomicMax(unsigned long long* a, unsigned long long v) { return *a = std::max(*a, v); }
#define __global__
// runs f(lane) on 32 lock-step lanes
static inline void simt_run_warp(const std::function<void(int)>& f) {
  SimtWarp w;
  std::vector<std::thread> th;
  for (int l = 0; l < 32; l++)
    th.emplace_back([&w, &f, l] {
      simt_warp = &w;
      threadIdx.x = l;
      f(l);
    });
  for (auto& t : th) t.join();
}
