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
// Tile marking of a scatter, warp-collective and in three pieces so that no atomic's round trip is waited for in place
// (r02j stall samples: the per-lane chain "tag load -> atomicExch -> atomicAdd(count) -> store" of the first version cost
// 20 % of k_g2p2g's warp time, up to 16 dependent L2 round trips per warp and one same-address atomicAdd per tile):
//   claim_tiles   before the scatter's value build: `epoch` is swapped into the tags of the (up to) eight tiles the stencils of
//                 a run of lanes touch, all atomics issued back to back;
//   append_begin  after the shared-memory stores: a tag that came back != epoch means this lane claimed the tile FIRST; the
//                 winners of the warp are counted with ballots and lane 0 reserves their list slots with ONE atomicAdd;
//   append_finish after the run reduction: the winners store their tiles.
// All lanes of the warp must call all three.
struct TileClaim {
  int tag[8];     // claim_tiles: previous tag of tile c (epoch: nothing won)
  unsigned won;   // append_begin: bit c set if this lane appends tile c; bits 8.. = its first slot relative to the warp's base
  int base;       // append_begin, lane 0: the warp's first slot in the list
};
// 32-bit global accesses the optimiser may not move (a plain load / atomic would be sunk to its first use)
DSK_DEV int load_int_here(const int* p) {
#if defined(__CUDA_ARCH__)
  int v;
  asm volatile("ld.global.b32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
#else
  return *p;
#endif
}
DSK_DEV int exch_int_here(int* p, int v) {
#if defined(__CUDA_ARCH__)
  int old;
  asm volatile("atom.global.exch.b32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
  return old;
#else
  return atomicExch(p, v);
#endif
}
DSK_DEV int add_int_here(int* p, int v) {
#if defined(__CUDA_ARCH__)
  int old;
  asm volatile("atom.global.add.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
  return old;
#else
  return atomicAdd(p, v);
#endif
}
// tile c of the 2x2x2 block of tiles at a stencil's base tile: base + (c>>2 & 1, c>>1 & 1, c & 1)
DSK_DEV int base_tile(const SimConst& k, int env, const Stencil& s) {
  return env * k.ntile + ((s.bx >> 2) * k.nt + (s.by >> 2)) * k.nt + (s.bz >> 2);
}
DSK_DEV int block_tile(const SimConst& k, int tile0, int c) {
  return tile0 + ((c & 4) ? k.nt * k.nt : 0) + ((c & 2) ? k.nt : 0) + (c & 1);
}
// The particles are sorted by tile-major cell key, so the lanes of a warp that share a base tile are adjacent: only the
// first lane of each such run claims, for the whole run (the OR of the lanes' crossing patterns: tile c = base tile +
// (c>>2 & 1, c>>1 & 1, c & 1) is touched by a stencil that crosses the tile border along exactly those axes).  A dense warp
// issues ~4-8 atomics instead of 3.4 per cell run.
DSK_DEV void claim_tiles(const SimConst& k, const TileTrack& t, int env, const Stencil& s, int epoch, bool active, TileClaim& g) {
  const int lane = threadIdx.x & 31;
  const int tile0 = active ? base_tile(k, env, s) : -1 - lane;
  const int prev = __shfl_up_sync(0xffffffffu, tile0, 1);
  const bool head = lane == 0 || prev != tile0;
  const unsigned heads = __ballot_sync(0xffffffffu, head);
  const unsigned above = lane == 31 ? 0u : (heads & (0xffffffffu << (lane + 1)));
  const int end = above ? __ffs(above) - 2 : 31;   // last lane of my run
  const unsigned cx = ((s.bx + 2) >> 2) != (s.bx >> 2), cy = ((s.by + 2) >> 2) != (s.by >> 2), cz = ((s.bz + 2) >> 2) != (s.bz >> 2);
  unsigned need = active ? (1u | (cz << 1) | (cy << 2) | ((cy & cz) << 3) | (cx << 4) | ((cx & cz) << 5) | ((cx & cy) << 6) |
                            ((cx & cy & cz) << 7))
                         : 0u;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    unsigned o = __shfl_down_sync(0xffffffffu, need, d);
    if (lane + d <= end) need |= o;
  }
  if (!(head && active)) need = 0u;
#pragma unroll
  for (int c = 0; c < 8; c++) {
    int v = epoch;
    if (need & (1u << c)) v = exch_int_here(t.epoch + block_tile(k, tile0, c), epoch);
    g.tag[c] = v;
  }
}
DSK_DEV void append_begin(const TileTrack& t, int epoch, TileClaim& g) {
  const int lane = threadIdx.x & 31;
  unsigned won = 0;
#pragma unroll
  for (int c = 0; c < 8; c++)
    if (g.tag[c] != epoch) won |= 1u << c;
  g.won = won;
  g.base = 0;
  if (!__any_sync(0xffffffffu, won != 0u)) return;
  // a lane stores its wins in consecutive slots: exclusive prefix of the per-lane win counts
  const int mine = __popc(won);
  int incl = mine;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int up = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += up;
  }
  const int total = __shfl_sync(0xffffffffu, incl, 31);
  g.won = won | ((unsigned)(incl - mine) << 8);
  if (lane == 0) g.base = add_int_here(t.count, total);
}
DSK_DEV void append_finish(const SimConst& k, const TileTrack& t, int env, const Stencil& s, const TileClaim& g) {
  if (!__any_sync(0xffffffffu, g.won & 0xffu)) return;
  int at = __shfl_sync(0xffffffffu, g.base, 0) + (int)(g.won >> 8);
#pragma unroll
  for (int c = 0; c < 8; c++)
    if (g.won & (1u << c)) t.list[at++] = block_tile(k, base_tile(k, env, s), c);
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
DSK_DEV void warp_scatter27_groups(const SimConst& k, bool active, const Stencil& s, float4* __restrict__ Ge, int key,
                                   unsigned act, unsigned same, F val) {
  const int lane = threadIdx.x & 31;
  bool first = active && lane == __ffs(same) - 1;
  bool small = active && __popc(same) < SCATTER_MIN_GROUP;
  unsigned leaders = __ballot_sync(0xffffffffu, first && !small);
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
DSK_DEV void warp_scatter9_groups(const SimConst& k, bool active, const Stencil& s, int plane, int oxp,
                                  float4* __restrict__ Ge, int key, unsigned act, unsigned same, F val) {
  const int lane = threadIdx.x & 31;
  bool first = active && lane == __ffs(same) - 1;
  bool small = active && __popc(same) < SCATTER_MIN_GROUP;
  unsigned leaders = __ballot_sync(0xffffffffu, first && !small);
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
// front halves of the two butterflies: groups of equal keys; the first lane of each group claims the stencil's tiles before
// the reduction and appends the ones it won after it (claim_tiles)
template <class F>
DSK_DEV void warp_scatter27(const SimConst& k, bool active, const Stencil& s, float4* __restrict__ Ge,
                            const TileTrack& tt, bool mark, int env, int epoch, F val) {
  const int lane = threadIdx.x & 31;
  int key = active ? node_offset(s.bx, s.by, s.bz, k.nt) : -1 - lane;
  unsigned act = __ballot_sync(0xffffffffu, active);
  if (!act) return;
  // groups of equal keys, adjacent or not: the sort is per env step, so late substeps see a few strays per warp
  unsigned same = __match_any_sync(0xffffffffu, key);
  TileClaim claim;
  if (mark) claim_tiles(k, tt, env, s, epoch, active, claim);
  warp_scatter27_groups(k, active, s, Ge, key, act, same, val);
  if (mark) {
    append_begin(tt, epoch, claim);
    append_finish(k, tt, env, s, claim);
  }
}
template <class F>
DSK_DEV void warp_scatter9(const SimConst& k, bool active, const Stencil& s, int plane, int oxp,
                           float4* __restrict__ Ge, const TileTrack& tt, bool mark, int env, int epoch, F val) {
  const int lane = threadIdx.x & 31;
  int key = active ? node_offset(s.bx, s.by, s.bz, k.nt) : -1 - lane;
  unsigned act = __ballot_sync(0xffffffffu, active);
  if (!act) return;
  unsigned same = __match_any_sync(0xffffffffu, key);
  TileClaim claim;
  if (mark) claim_tiles(k, tt, env, s, epoch, active, claim);
  warp_scatter9_groups(k, active, s, plane, oxp, Ge, key, act, same, val);
  if (mark) {
    append_begin(tt, epoch, claim);
    append_finish(k, tt, env, s, claim);
  }
}
// ---- transposed shared-memory scatter (round 2) -------------------------------------------------------------------------
// The butterfly above costs ~900 instructions per lane and group (124 shuffles, 248 selects, the 27 values re-evaluated
// for every group) and its segmented fallback ~1 300.  Here every lane writes its 27 float4 contributions ONCE into a
// per-warp shared-memory tile laid out [node][lane] (row stride 33 float4: the STS.128 of a warp and the LDS.128 of 27
// lanes walking 27 different rows are both conflict-free), and for every RUN of adjacent lanes with equal cell keys
// lanes 0..26 add up their node's row segment and issue one RED.128: 27 STS + 32 LDS + 128 FADD per lane whatever the
// number of runs.  Strays (the sort is per env step) simply form runs of length one; inactive lanes store zeros.
#define TS_ROW 33
#define TS_WARP_FLOAT4 (27 * TS_ROW)      // 14 256 bytes per warp
#define TS9_WARP_FLOAT4 (9 * TS_ROW)      //  4 752 bytes per warp (plane-split kernels)
// run bookkeeping shared by the transposed scatters: heads of the runs of adjacent equal keys (an inactive lane is a dead
// run of its own, so that no live run ever covers a row element an inactive lane wrote), the stencil tiles marked once per
// live run; returns the ballot of run heads (0: no active lane) and the ballot of active lanes
DSK_DEV unsigned ts_run_heads(const SimConst& k, bool active, const Stencil& s, const TileTrack& tt, bool mark, int env,
                              int epoch, unsigned& act, TileClaim& claim) {
  const int lane = threadIdx.x & 31;
  int key = active ? node_offset(s.bx, s.by, s.bz, k.nt) : -1 - lane;
  act = __ballot_sync(0xffffffffu, active);
  int prev = __shfl_up_sync(0xffffffffu, key, 1);
  bool head = lane == 0 || prev != key;
  unsigned heads = __ballot_sync(0xffffffffu, head);
  if (mark) claim_tiles(k, tt, env, s, epoch, active, claim);   // results used after the tile is written
  return act ? heads : 0u;
}
// second half: for every run lanes 0..26 add up their node's row segment and issue one RED.128
DSK_DEV void ts_reduce_runs27(const SimConst& k, const Stencil& s, float4* __restrict__ Ge, const float4* wbuf, unsigned todo,
                              unsigned act) {
  const int lane = threadIdx.x & 31;
  const int i = lane / 9, j = (lane / 3) % 3, l = lane % 3;   // stencil node owned by lanes 0..26
  const float4* row = wbuf + (lane < 27 ? lane : 0) * TS_ROW;
  while (todo) {
    int h = __ffs(todo) - 1;
    todo &= todo - 1;
    int e = todo ? __ffs(todo) - 1 : 32;   // the run is [h, e)
    if (!((act >> h) & 1u)) continue;      // dead run (inactive lane)
    int bx = __shfl_sync(0xffffffffu, s.bx, h), by = __shfl_sync(0xffffffffu, s.by, h),
        bz = __shfl_sync(0xffffffffu, s.bz, h);
    if (lane < 27) {
      float4 acc = row[h];
#pragma unroll 4
      for (int t = h + 1; t < e; t++) acc = f4add(acc, row[t]);
      red_add4(&Ge[node_offset(bx + i, by + j, bz + l, k.nt)], acc);
    }
  }
}
template <class F>
DSK_DEV void warp_scatter27_ts(const SimConst& k, bool active, const Stencil& s, float4* __restrict__ Ge,
                               const TileTrack& tt, bool mark, int env, int epoch, float4* wbuf, F val) {
  const int lane = threadIdx.x & 31;
  unsigned act;
  TileClaim claim;
  unsigned todo = ts_run_heads(k, active, s, tt, mark, env, epoch, act, claim);
  if (!todo) return;
#pragma unroll
  for (int q = 0; q < 27; q++) wbuf[q * TS_ROW + lane] = val(q / 9, (q / 3) % 3, q % 3);
  __syncwarp();
  if (mark) append_begin(tt, epoch, claim);
  ts_reduce_runs27(k, s, Ge, wbuf, todo, act);
  if (mark) append_finish(k, tt, env, s, claim);
  __syncwarp();   // the tile is rewritten by the warp's next scatter
}
// Both scatters of a substep have AFFINE contributions: node (i, j, l) receives w_ijl * (A0 + i AX + j AY + l AZ) with
// float4 coefficients (p2g: A0 = (p_mass v - dx affine fx, p_mass), AX.. = dx * columns of affine; g2p.grad: A0 = (b0, 0),
// AX.. = (c_C * columns of gC, 0)).  The 27 values are built incrementally, (x, y) and (z, w) as packed pairs: ~150
// instructions instead of 27 x 15.  (A three-pass variant with a 9-row tile -- 4.75 KB per warp, no occupancy limit from shared
// memory -- was measured in r02c: the run loop executed three times costs far more than the occupancy returns, 1 M-particle
// k_g2p2g 165 -> 257 us at 8 CTAs/SM; profiles/r02c_*.json.)
DSK_DEV void warp_scatter27_ts_affine(const SimConst& k, bool active, const Stencil& s, float4* __restrict__ Ge,
                                      const TileTrack& tt, bool mark, int env, int epoch, float4* wbuf, float4 A0, float3 AX,
                                      float3 AY, float3 AZ) {
  const int lane = threadIdx.x & 31;
  unsigned act;
  TileClaim claim;
  unsigned todo = ts_run_heads(k, active, s, tt, mark, env, epoch, act, claim);
  if (!todo) return;
  {
    const float2 axl = f2(AX.x, AX.y), axh = f2(AX.z, 0.f), ayl = f2(AY.x, AY.y), ayh = f2(AY.z, 0.f);
    const float2 azl = f2(AZ.x, AZ.y), azh = f2(AZ.z, 0.f);
    float2 lo_i = f2(A0.x, A0.y), hi_i = f2(A0.z, A0.w);
    float4* col = wbuf + lane;
#pragma unroll
    for (int i = 0; i < 3; i++) {
      float2 lo_ij = lo_i, hi_ij = hi_i;
#pragma unroll
      for (int j = 0; j < 3; j++) {
        const float wxy = s.wx[i] * s.wy[j];
        float2 lo = lo_ij, hi = hi_ij;
#pragma unroll
        for (int l = 0; l < 3; l++) {
          const float2 w2 = bc2(wxy * s.wz[l]);
          float2 vlo = mul2(w2, lo), vhi = mul2(w2, hi);
          col[((i * 3 + j) * 3 + l) * TS_ROW] = make_float4(vlo.x, vlo.y, vhi.x, vhi.y);
          if (l < 2) {
            lo = add2(lo, azl);
            hi = add2(hi, azh);
          }
        }
        if (j < 2) {
          lo_ij = add2(lo_ij, ayl);
          hi_ij = add2(hi_ij, ayh);
        }
      }
      if (i < 2) {
        lo_i = add2(lo_i, axl);
        hi_i = add2(hi_i, axh);
      }
    }
  }
  __syncwarp();
  if (mark) append_begin(tt, epoch, claim);
  ts_reduce_runs27(k, s, Ge, wbuf, todo, act);
  if (mark) append_finish(k, tt, env, s, claim);
  __syncwarp();
}
// one x-plane (9 nodes) per thread, for the plane-split kernels: lanes (part, node) = (lane / 9, lane % 9) add every third
// element of the run's row segment, two shuffle-downs combine the three parts
template <class F>
DSK_DEV void warp_scatter9_ts(const SimConst& k, bool active, const Stencil& s, int plane, float4* __restrict__ Ge,
                              const TileTrack& tt, bool mark, int env, int epoch, float4* wbuf, F val) {
  const int lane = threadIdx.x & 31;
  unsigned act;
  TileClaim claim;
  unsigned todo = ts_run_heads(k, active, s, tt, mark, env, epoch, act, claim);
  if (!todo) return;
#pragma unroll
  for (int q = 0; q < 9; q++) wbuf[q * TS_ROW + lane] = val(q / 3, q % 3);
  __syncwarp();
  if (mark) append_begin(tt, epoch, claim);
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const int part = lane / 9, q = lane - part * 9;
  const float4* row = wbuf + (lane < 27 ? q : 0) * TS_ROW;
  while (todo) {
    int h = __ffs(todo) - 1;
    todo &= todo - 1;
    int e = todo ? __ffs(todo) - 1 : 32;
    if (!((act >> h) & 1u)) continue;      // dead run (inactive lane)
    int bx = __shfl_sync(0xffffffffu, s.bx, h), by = __shfl_sync(0xffffffffu, s.by, h),
        bz = __shfl_sync(0xffffffffu, s.bz, h);
    float4 acc = z4;
    if (lane < 27)
      for (int t = h + part; t < e; t += 3) acc = f4add(acc, row[t]);
    float4 a1 = f4shfl_down(acc, 9), a2 = f4shfl_down(acc, 18);
    if (lane < 9) red_add4(&Ge[node_offset(bx + plane, by + q / 3, bz + q % 3, k.nt)], f4add(f4add(acc, a1), a2));
  }
  if (mark) append_finish(k, tt, env, s, claim);
  __syncwarp();
}
// scatter front ends of the kernels: TS selects the transposed shared-memory version (dynamic shared memory:
// TS_WARP_FLOAT4 / TS9_WARP_FLOAT4 float4 per warp of the CTA)
template <bool TS, class F>
DSK_DEV void scatter27(const SimConst& k, bool active, const Stencil& s, float4* __restrict__ Ge, const TileTrack& tt,
                       bool mark, int env, int epoch, F val) {
  if (TS) {
    DSK_DYN_SMEM(float4, ts_buf);
    warp_scatter27_ts(k, active, s, Ge, tt, mark, env, epoch, ts_buf + (threadIdx.x >> 5) * TS_WARP_FLOAT4, val);
  } else {
    warp_scatter27(k, active, s, Ge, tt, mark, env, epoch, val);
  }
}
// affine contributions (see warp_scatter27_ts_affine); the butterfly path evaluates the same values through a lambda
template <bool TS>
DSK_DEV void scatter27_affine(const SimConst& k, bool active, const Stencil& s, float4* __restrict__ Ge, const TileTrack& tt,
                              bool mark, int env, int epoch, float4 A0, float3 AX, float3 AY, float3 AZ) {
  if (TS) {
    DSK_DYN_SMEM(float4, ts_buf);
    warp_scatter27_ts_affine(k, active, s, Ge, tt, mark, env, epoch, ts_buf + (threadIdx.x >> 5) * TS_WARP_FLOAT4, A0, AX, AY, AZ);
  } else {
    warp_scatter27(k, active, s, Ge, tt, mark, env, epoch, [&](int i, int j, int l) {
      float w = s.wx[i] * s.wy[j] * s.wz[l];
      float3 a = f3(A0.x, A0.y, A0.z) + (float)i * AX + (float)j * AY + (float)l * AZ;
      return make_float4(w * a.x, w * a.y, w * a.z, w * A0.w);
    });
  }
}
template <bool TS, class F>
DSK_DEV void scatter9(const SimConst& k, bool active, const Stencil& s, int plane, int oxp, float4* __restrict__ Ge,
                      const TileTrack& tt, bool mark, int env, int epoch, F val) {
  if (TS) {
    DSK_DYN_SMEM(float4, ts_buf);
    int warp = (threadIdx.y * blockDim.x + threadIdx.x) >> 5;
    warp_scatter9_ts(k, active, s, plane, Ge, tt, mark, env, epoch, ts_buf + warp * TS9_WARP_FLOAT4, val);
  } else {
    warp_scatter9(k, active, s, plane, oxp, Ge, tt, mark, env, epoch, val);
  }
}
DSK_DEV float pick3(const float* a, int i) { return i == 0 ? a[0] : (i == 1 ? a[1] : a[2]); }
DSK_DEV int pick3(const int* a, int i) { return i == 0 ? a[0] : (i == 1 ? a[1] : a[2]); }
