// Shared definitions of the substep kernels: constants, memory layout, B-spline stencil.
//
// HBM layout (DESIGN.md "Data layout"):
//   particle frame : SoA, 24 components (x3 v3 C9 F9) x [n_envs * Npad] floats; component c
//                    of slot gid lives at frame[c * stride + gid]  -> every load/store of a
//                    warp is one 128-byte line.
//   grid           : per env n^3 nodes of float4 (vx, vy, vz, m) stored TILE-MAJOR: a 4x4x4
//                    tile is one contiguous 1 KB block, so grid kernels read/write whole
//                    lines and a particle's 27-node stencil touches at most 8 blocks.
#pragma once
#include "mpm_math.cuh"
#include "tools.cuh"

#include "particle_math.cuh"

// dynamic shared memory of a kernel (the CPU emulation of tests/host_check keeps it in one process-wide buffer)
#ifdef DSK_HOST_SIMT
#define DSK_DYN_SMEM(type, name) type* name = (type*)simt_dyn_smem.data()
#else
#define DSK_DYN_SMEM(type, name) extern __shared__ type name[]
#endif

// Profiling build (-DDSK_TIMELINE, libdiffskill_mpm_tl.so): every kernel stamps %globaltimer at its first and last
// warp into the record of its launch, which gives per-kernel start/end times INSIDE replayed CUDA graphs -- where
// CUDA events cannot be placed and ncu serialises the launches.  Compiled out of the product library.
#ifdef DSK_TIMELINE
struct TlRec {
  unsigned long long t0, t1;
};
struct TlScope {
  TlRec* r;
  __device__ TlScope(const SimConst& k) {
    r = (k.tl && k.tl_slot >= 0) ? k.tl + k.tl_slot : nullptr;
    if (r && threadIdx.x == 0 && threadIdx.y == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      atomicMin(&r->t0, t);
    }
  }
  __device__ ~TlScope() {
    if (r && (threadIdx.x & 31) == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      atomicMax(&r->t1, t);
    }
  }
};
#define DSK_TL(k) TlScope tl_scope__(k)
#else
#define DSK_TL(k)
#endif

DSK_DEV int node_offset(int X, int Y, int Z, int nt) {
  int tile = ((X >> 2) * nt + (Y >> 2)) * nt + (Z >> 2);
  return (tile << 6) | ((X & 3) << 4) | ((Y & 3) << 2) | (Z & 3);
}

// vector reduction to global memory: one RED.E.ADD.F32x4 (sm_90+) instead of four scalar REDs
DSK_DEV void red_add4(float4* addr, float4 v) { atomicAdd(addr, v); }

// Step-dependent pointers live in device memory so that ONE captured CUDA graph per step slot can be replayed
// for every env step: boundary kernels dereference these instead of taking the pointers as launch parameters.
struct StepArgs {
  float* ck_src;        // particle checkpoint the step starts from
  float* ck_dst;        // particle checkpoint the step writes
  float* tool_src;      // tool states [B][K][8] at the source checkpoint
  float* tool_dst;
  const float* action;  // clipped actions [B][A] of the step (or null)
  float* adj_in;        // adjoint checkpoint step+1
  float* adj_out;       // adjoint checkpoint step (accumulated into)
  float* tool_adj_in;
  float* tool_adj_out;
  float* action_grad;   // [B][A] of the step (accumulated into)
  int epoch_base;       // multiple of 4; substep q of the sequence runs as epoch epoch_base + q + 1
};
__global__ void k_set_args(StepArgs* dst, StepArgs v) { *dst = v; }

// Compact per-step record of the active grid tiles of every substep (see k_grid / k_tape_restore)
struct GridTape {
  int* base;      // [S+1] prefix offsets into list/data (null: taping off)
  int* list;      // [cap] packed env*ntile + tile
  float4* data;   // [cap][2][64]: (momentum, mass) then (velocity, mass)
  int* overflow;  // set when a step needed more than cap tiles
  int cap;
};

// Per-substep sparse-grid bookkeeping: tiles touched by a stencil are appended (once) to the
// active list of the current epoch.
struct TileTrack {
  int* epoch;  // [B*ntile] last epoch in which the tile was appended
  int* list;   // [B*ntile] packed env*ntile + tile
  int* count;  // scalar
};
DSK_DEV void mark_tile(const TileTrack& t, int gtile, int epoch) {
  if (t.epoch[gtile] != epoch) {
    if (atomicExch(&t.epoch[gtile], epoch) != epoch) t.list[atomicAdd(t.count, 1)] = gtile;
  }
}
DSK_DEV void mark_stencil_tiles(const SimConst& k, const TileTrack& t, int env, const Stencil& s, int epoch) {
  int tx0 = s.bx >> 2, tx1 = (s.bx + 2) >> 2;
  int ty0 = s.by >> 2, ty1 = (s.by + 2) >> 2;
  int tz0 = s.bz >> 2, tz1 = (s.bz + 2) >> 2;
  int base = env * k.ntile;
  for (int a = tx0; a <= tx1; a++)
    for (int b = ty0; b <= ty1; b++)
      for (int c = tz0; c <= tz1; c++) mark_tile(t, base + (a * k.nt + b) * k.nt + c, epoch);
}


// ---- warp-aggregated 27-node scatter ------------------------------------------------------------------------
// Particles are sorted by cell, so the lanes of a warp fall into a few groups that share one stencil.  For each
// group the 27 float4 contributions of every lane are summed across the warp with a recursive-halving butterfly
// (124 shuffles instead of 27*4*5): after the five steps lane l holds the complete sum for stencil node l, and
// lanes 0..26 issue ONE vector reduction each -- 27 RED.128 per group instead of 27 per particle.  Groups of
// fewer than SCATTER_MIN_GROUP lanes fall back to per-lane reductions.
#define SCATTER_MIN_GROUP 4
DSK_DEV float4 f4add(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
DSK_DEV float4 f4sel(bool c, float4 a, float4 b) { return c ? a : b; }
DSK_DEV float4 f4shfl_xor(float4 v, int m) {
  return make_float4(__shfl_xor_sync(0xffffffffu, v.x, m), __shfl_xor_sync(0xffffffffu, v.y, m),
                     __shfl_xor_sync(0xffffffffu, v.z, m), __shfl_xor_sync(0xffffffffu, v.w, m));
}
template <int Q, class F>
DSK_DEV float4 slot_value(bool mine, F& val) {
  if (Q >= 27) return make_float4(0.f, 0.f, 0.f, 0.f);
  float4 v = val(Q / 9, (Q / 3) % 3, Q % 3);
  return mine ? v : make_float4(0.f, 0.f, 0.f, 0.f);
}
template <int Q, class F>
DSK_DEV void butterfly_first(bool mine, bool hi, F& val, float4* acc) {
  float4 a = slot_value<Q>(mine, val), b = slot_value<Q + 16>(mine, val);
  acc[Q] = f4add(f4sel(hi, b, a), f4shfl_xor(f4sel(hi, a, b), 16));
}
// val(i,j,l) -> float4 contribution of this lane to stencil node (i,j,l); must be callable by every lane
DSK_DEV float4 f4shfl_down(float4 v, int d) {
  return make_float4(__shfl_down_sync(0xffffffffu, v.x, d), __shfl_down_sync(0xffffffffu, v.y, d),
                     __shfl_down_sync(0xffffffffu, v.z, d), __shfl_down_sync(0xffffffffu, v.w, d));
}
// val(i,j,l) -> float4 contribution of this lane to stencil node (i,j,l); must be callable by every lane.
// Regimes, chosen per warp from the groups of lanes with equal cell keys (match.any -- adjacent or not, because the
// particles are sorted once per env step and a few lanes per warp have strayed into a neighbour cell by the last
// substeps):
//   <= 2 groups of >= SCATTER_MIN_GROUP lanes : recursive-halving butterfly per group (dense dough: a whole warp
//               shares one cell); lanes in smaller groups issue their own reductions
//   > 2 such groups : ONE segmented shuffle-down reduction over all runs of adjacent equal keys at once
//               (log2(longest run) steps per node), run heads issue the vector reductions
template <class F>
DSK_DEV void warp_scatter27(const SimConst& k, bool active, const Stencil& s, float4* __restrict__ Ge,
                            const TileTrack& tt, bool mark, int env, int epoch, F val) {
  const int lane = threadIdx.x & 31;
  int key = active ? node_offset(s.bx, s.by, s.bz, k.nt) : -1 - lane;
  unsigned act = __ballot_sync(0xffffffffu, active);
  if (!act) return;
  // groups of equal keys, adjacent or not: the sort is per env step, so late substeps see a few strays per warp
  unsigned same = __match_any_sync(0xffffffffu, key);
  bool first = active && lane == __ffs(same) - 1;
  bool small = active && __popc(same) < SCATTER_MIN_GROUP;
  unsigned leaders = __ballot_sync(0xffffffffu, first && !small);
  if (mark && first) mark_stencil_tiles(k, tt, env, s, epoch);
  if (__popc(leaders) > 2) {
    int prev = __shfl_up_sync(0xffffffffu, key, 1);
    bool head = active && (lane == 0 || prev != key);
    unsigned heads = __ballot_sync(0xffffffffu, head);
    unsigned above = lane == 31 ? 0u : (heads & (0xffffffffu << (lane + 1)));
    int next_head = above ? (__ffs(above) - 1) : 32;
    int end = min(next_head - 1, 31 - __clz(act));   // last lane of my run
    int maxlen = __reduce_max_sync(0xffffffffu, head ? end - lane + 1 : 0);
    // nine stencil nodes (one x-slab) at a time: 36 independent shuffle chains per reduction step
#pragma unroll
    for (int i = 0; i < 3; i++) {
      float4 v[9];
#pragma unroll
      for (int q = 0; q < 9; q++) v[q] = active ? val(i, q / 3, q % 3) : make_float4(0.f, 0.f, 0.f, 0.f);
      for (int d = 1; d < maxlen; d <<= 1) {
        bool take = lane + d <= end;
#pragma unroll
        for (int q = 0; q < 9; q++) {
          float4 t = f4shfl_down(v[q], d);
          if (take) v[q] = f4add(v[q], t);
        }
      }
      if (head) {
#pragma unroll
        for (int q = 0; q < 9; q++) red_add4(&Ge[s.ox[i] + s.oy[q / 3] + s.oz[q % 3]], v[q]);
      }
    }
    return;
  }
  // strays (groups of fewer than SCATTER_MIN_GROUP lanes): per-lane reductions, all of them in one pass
  if (small) {
#pragma unroll
    for (int i = 0; i < 3; i++)
#pragma unroll
      for (int j = 0; j < 3; j++)
#pragma unroll
        for (int l = 0; l < 3; l++) red_add4(&Ge[s.ox[i] + s.oy[j] + s.oz[l]], val(i, j, l));
  }
  unsigned todo = leaders;
  while (todo) {
    int leader = __ffs(todo) - 1;
    todo &= todo - 1;
    unsigned grp = __shfl_sync(0xffffffffu, same, leader);
    bool mine = (grp >> lane) & 1u;
    float4 acc[16];
    {
      bool hi = lane & 16;
      butterfly_first<0>(mine, hi, val, acc);  butterfly_first<1>(mine, hi, val, acc);
      butterfly_first<2>(mine, hi, val, acc);  butterfly_first<3>(mine, hi, val, acc);
      butterfly_first<4>(mine, hi, val, acc);  butterfly_first<5>(mine, hi, val, acc);
      butterfly_first<6>(mine, hi, val, acc);  butterfly_first<7>(mine, hi, val, acc);
      butterfly_first<8>(mine, hi, val, acc);  butterfly_first<9>(mine, hi, val, acc);
      butterfly_first<10>(mine, hi, val, acc); butterfly_first<11>(mine, hi, val, acc);
      butterfly_first<12>(mine, hi, val, acc); butterfly_first<13>(mine, hi, val, acc);
      butterfly_first<14>(mine, hi, val, acc); butterfly_first<15>(mine, hi, val, acc);
    }
#pragma unroll
    for (int half = 8; half >= 1; half >>= 1) {
      bool hi = lane & half;
#pragma unroll
      for (int q = 0; q < half; q++)
        acc[q] = f4add(f4sel(hi, acc[q + half], acc[q]), f4shfl_xor(f4sel(hi, acc[q], acc[q + half]), half));
    }
    // lane l now owns stencil node l of the group's cell
    int bx = __shfl_sync(0xffffffffu, s.bx, leader), by = __shfl_sync(0xffffffffu, s.by, leader),
        bz = __shfl_sync(0xffffffffu, s.bz, leader);
    if (lane < 27) {
      int i = lane / 9, j = (lane / 3) % 3, l = lane % 3;
      red_add4(&Ge[node_offset(bx + i, by + j, bz + l, k.nt)], acc[0]);
    }
  }
}

// ---- plane-split variant for small engines ------------------------------------------------------------------------
// A single scene has fewer particle warps than the GPU has warp schedulers, so its particle kernels are bound by
// the instruction chain of ONE thread.  The *_pl kernels use three threads per particle -- one per x-plane of the
// stencil -- in three different warps with the same lane <-> particle mapping: each gathers / scatters 9 nodes, and
// partial sums are exchanged through shared memory.  warp_scatter9 is warp_scatter27 for one plane (16-slot
// butterfly: 64 shuffles).  oxp = s.ox[plane]; val(j, l) is the contribution to node (plane, j, l).
#define PL_PARTICLES 64   // particles per CTA of a plane-split kernel: blockDim = (64, 3)
template <int Q, class F>
DSK_DEV float4 slot_value9(bool mine, F& val) {
  if (Q >= 9) return make_float4(0.f, 0.f, 0.f, 0.f);
  float4 v = val(Q / 3, Q % 3);
  return mine ? v : make_float4(0.f, 0.f, 0.f, 0.f);
}
template <int Q, class F>
DSK_DEV void butterfly9_first(bool mine, bool hi, F& val, float4* acc) {
  float4 a = slot_value9<Q>(mine, val), b = slot_value9<Q + 8>(mine, val);
  acc[Q] = f4add(f4sel(hi, b, a), f4shfl_xor(f4sel(hi, a, b), 16));
}
template <class F>
DSK_DEV void warp_scatter9(const SimConst& k, bool active, const Stencil& s, int plane, int oxp,
                           float4* __restrict__ Ge, const TileTrack& tt, bool mark, int env, int epoch, F val) {
  const int lane = threadIdx.x & 31;
  int key = active ? node_offset(s.bx, s.by, s.bz, k.nt) : -1 - lane;
  unsigned act = __ballot_sync(0xffffffffu, active);
  if (!act) return;
  unsigned same = __match_any_sync(0xffffffffu, key);
  bool first = active && lane == __ffs(same) - 1;
  bool small = active && __popc(same) < SCATTER_MIN_GROUP;
  unsigned leaders = __ballot_sync(0xffffffffu, first && !small);
  if (mark && first) mark_stencil_tiles(k, tt, env, s, epoch);
  if (__popc(leaders) > 2) {
    int prev = __shfl_up_sync(0xffffffffu, key, 1);
    bool head = active && (lane == 0 || prev != key);
    unsigned heads = __ballot_sync(0xffffffffu, head);
    unsigned above = lane == 31 ? 0u : (heads & (0xffffffffu << (lane + 1)));
    int next_head = above ? (__ffs(above) - 1) : 32;
    int end = min(next_head - 1, 31 - __clz(act));   // last lane of my run
    int maxlen = __reduce_max_sync(0xffffffffu, head ? end - lane + 1 : 0);
    float4 v[9];
#pragma unroll
    for (int q = 0; q < 9; q++) v[q] = active ? val(q / 3, q % 3) : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int d = 1; d < maxlen; d <<= 1) {
      bool take = lane + d <= end;
#pragma unroll
      for (int q = 0; q < 9; q++) {
        float4 t = f4shfl_down(v[q], d);
        if (take) v[q] = f4add(v[q], t);
      }
    }
    if (head) {
#pragma unroll
      for (int q = 0; q < 9; q++) red_add4(&Ge[oxp + s.oy[q / 3] + s.oz[q % 3]], v[q]);
    }
    return;
  }
  if (small) {
#pragma unroll
    for (int q = 0; q < 9; q++) red_add4(&Ge[oxp + s.oy[q / 3] + s.oz[q % 3]], val(q / 3, q % 3));
  }
  unsigned todo = leaders;
  while (todo) {
    int leader = __ffs(todo) - 1;
    todo &= todo - 1;
    unsigned grp = __shfl_sync(0xffffffffu, same, leader);
    bool mine = (grp >> lane) & 1u;
    float4 acc[8];
    {
      bool hi = lane & 16;
      butterfly9_first<0>(mine, hi, val, acc); butterfly9_first<1>(mine, hi, val, acc);
      butterfly9_first<2>(mine, hi, val, acc); butterfly9_first<3>(mine, hi, val, acc);
      butterfly9_first<4>(mine, hi, val, acc); butterfly9_first<5>(mine, hi, val, acc);
      butterfly9_first<6>(mine, hi, val, acc); butterfly9_first<7>(mine, hi, val, acc);
    }
#pragma unroll
    for (int half = 4; half >= 1; half >>= 1) {
      bool hi = lane & (half << 1);
#pragma unroll
      for (int q = 0; q < half; q++)
        acc[q] = f4add(f4sel(hi, acc[q + half], acc[q]), f4shfl_xor(f4sel(hi, acc[q], acc[q + half]), half << 1));
    }
    acc[0] = f4add(acc[0], f4shfl_xor(acc[0], 1));
    // lanes 2m and 2m+1 now hold the complete sum for node m of the plane
    int bx = __shfl_sync(0xffffffffu, s.bx, leader), by = __shfl_sync(0xffffffffu, s.by, leader),
        bz = __shfl_sync(0xffffffffu, s.bz, leader);
    int m = lane >> 1;
    if (!(lane & 1) && m < 9) red_add4(&Ge[node_offset(bx + plane, by + m / 3, bz + m % 3, k.nt)], acc[0]);
  }
}
DSK_DEV float pick3(const float* a, int i) { return i == 0 ? a[0] : (i == 1 ? a[1] : a[2]); }
DSK_DEV int pick3(const int* a, int i) { return i == 0 ? a[0] : (i == 1 ? a[1] : a[2]); }
