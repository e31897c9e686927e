// Forward substep kernels (MPMSimulator.substep, plb/engine/mpm_simulator.py:307-323).
//   k_p2g   : compute_F_tmp + svd + p2g (mpm_simulator.py:121-129,198-225) fused; per particle
//   k_grid  : grid_op (mpm_simulator.py:230-262) over the ACTIVE tiles only, in place
//   k_g2p   : g2p (mpm_simulator.py:264-283); per particle
// clear_grid (mpm_simulator.py:100-110) is folded into k_grid: it zeroes the tiles the
// previous substep touched in the other grid buffer.
#pragma once
#include "kernels_common.cuh"
#include "svd3.cuh"

// MINB = min CTAs/SM the register allocation must allow: 1 for small (latency-bound) problems -- all registers, no
// spills; 3 for large batches where occupancy hides the scatter latency
template <bool WRITE_F, int MINB, bool TS>
__global__ void __launch_bounds__(128, MINB)
    k_p2g(SimConst k, const float* __restrict__ fin, float* __restrict__ fout, const float* __restrict__ mat,
          const int* __restrict__ npart, float4* __restrict__ G, TileTrack tt, const StepArgs* __restrict__ args,
          int q, const int* __restrict__ run_if, float* __restrict__ svd_out) {
  DSK_TL(k);
  const int epoch = load_int_here(&args->epoch_base) + q + 1;   // loaded first: the tile tags are compared with it
  if (run_if && *run_if == 0) return;   // adjoint recompute is skipped when the grid tape of the step is complete
  int gid = blockIdx.x * blockDim.x + threadIdx.x;
  int env = min(gid / k.Npad, k.B - 1), p = gid - env * k.Npad;   // a warp never straddles envs (Npad % 128 == 0)
  bool active = gid < k.stride && p < npart[env];
  int g = active ? gid : env * k.Npad;                            // inactive lanes shadow slot 0 (loads stay in bounds)
  float3 x = load_v3(fin, CX, k.stride, g);
  float3 v = load_v3(fin, CV, k.stride, g);
  M3 C = load_m3(fin, CC, k.stride, g);
  M3 F = load_m3(fin, CF, k.stride, g);
  float mu, lam, ys;
  load_mat(k, mat, g, mu, lam, ys);
  P2GParticle o;
  p2g_particle(k, C, F, mu, lam, ys, o);
  if (WRITE_F && active) {
    store_m3(fout, CF, k.stride, gid, o.newF);
    if (svd_out) store_svd(svd_out, k.stride, gid, o);
  }
  Stencil s;
  make_stencil(k, x.x, x.y, x.z, s);
  float4* Ge = G + (size_t)env * k.nnode;
  // w * (p_mass v + affine (offset - fx) dx) = w * (a0 + i ax + j ay + l az)
  float3 fxv = f3(s.fx, s.fy, s.fz);
  float3 a0 = k.p_mass * v - k.dx * mv(o.affine, fxv);
  float3 ax = f3(k.dx * o.affine.m[0], k.dx * o.affine.m[3], k.dx * o.affine.m[6]);
  float3 ay = f3(k.dx * o.affine.m[1], k.dx * o.affine.m[4], k.dx * o.affine.m[7]);
  float3 az = f3(k.dx * o.affine.m[2], k.dx * o.affine.m[5], k.dx * o.affine.m[8]);
  scatter27_affine<TS>(k, active, s, Ge, tt, true, env, epoch, make_float4(a0.x, a0.y, a0.z, k.p_mass), ax, ay, az);
}

#define GRID_CTA 64
// poses: [B][S+1][K][8]; grid kernels use frames j (P0) and j+1 (P1) of the tile's env
// The grid kernels begin with a handful of scalars that live in device memory (tile counts, the heads of the tile lists, the
// tape offset, the graph's run_if flag).  Issue is in order: read where they are needed, every one of them stalls the warp for
// a full L2 round trip at its first use (4-5 in a row at the top of k_grid, ~1.5 us of a 6 us kernel on a single scene).  The
// kernels therefore load ALL of them first, back to back, with loads the optimiser cannot sink (load_int_here); list heads
// are read speculatively (clamped index: entries past the count are stale but readable).
DSK_DEV int list_head(const SimConst& k, const int* list, int i) { return load_int_here(list + min(i, k.B * k.ntile - 1)); }

// `first` = list[blockIdx.x], already loaded (see above)
DSK_DEV void clear_tiles(const SimConst& k, const int* __restrict__ list, int count, int first, float4* c0, float4* c1,
                         float4* c2) {
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i = blockIdx.x; i < count; i += gridDim.x) {
    size_t o = ((size_t)(i == (int)blockIdx.x ? first : list[i]) << 6) + threadIdx.x;  // list holds env*ntile+tile and nnode == ntile*64
    if (c0) c0[o] = z;
    if (c1) c1[o] = z;
    if (c2) c2[o] = z;
  }
}

// end of a step sequence: zero the tiles of the last substep and reset the active-tile counters, so that every
// sequence (and every replay of its CUDA graph) starts from all-zero grids
__global__ void __launch_bounds__(GRID_CTA)
    k_end_clear(SimConst k, const int* __restrict__ list, const int* __restrict__ count, float4* c0, float4* c1,
                float4* c2, int* counts4, int* done) {
  DSK_TL(k);
  const int first = list_head(k, list, blockIdx.x);
  clear_tiles(k, list, load_int_here(count), first, c0, c1, c2);
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(done, 1) == (int)gridDim.x - 1) {  // last CTA: nobody reads the counters any more
      counts4[0] = counts4[1] = counts4[2] = counts4[3] = 0;
      *done = 0;
    }
  }
}

// ---- grid_op, parallel over (node, contact frame) --------------------------------------------------------------
// A "contact frame" is one rigid SDF body: a tool, or one jaw of a gripper (primitives.py:507-511 applies the
// two jaws sequentially).  The expensive geometry of a contact (signed distance, finite-difference normal,
// collider velocity) depends on the node and the pose only -- not on the velocity flowing through the chain of
// tools -- so it is evaluated by one thread per (node, frame); the velocity chain itself is a few dozen flops.
// Frames whose SDF at the tile centre exceeds the tile radius plus the soft-contact cut-off are culled per tile.
// thread (0, y) of the (node x frame) kernels
DSK_DEV void prepare_tile_frame(const SimConst& k, const ToolParams* sT, const FrameTable& ft, int y,
                                const float* __restrict__ poses, int env, int j, int tx, int ty, int tz,
                                TileFrames& tf) {
  tf.active[y] = prepare_frame(k, sT, ft, y, poses, env, j, tx, ty, tz, tf.F0[y], tf.F1[y]);
}
#define GRID_NODES 64
__global__ void __launch_bounds__(GRID_CTA)
    k_clear_set(SimConst k, const int* __restrict__ list, const int* __restrict__ count, float4* c0, float4* c1, float4* c2) {
  DSK_TL(k);
  const int first = list_head(k, list, blockIdx.x);
  clear_tiles(k, list, load_int_here(count), first, c0, c1, c2);
}

// grid_op over active tiles.  Gin holds (momentum, mass); Gout receives (velocity, mass) and may alias Gin.
// blockDim = (64, n_frames).
__global__ void __launch_bounds__(GRID_NODES * MAX_FRAMES)
    k_grid(SimConst k, const __grid_constant__ GridTools tp, const float* __restrict__ poses, int j,
           const float4* Gin, float4* Gout, const int* __restrict__ list, const int* __restrict__ count,
           // tiles of the previous substep to zero (clear_grid), may be null
           const int* __restrict__ clr_list, const int* __restrict__ clr_count, float4* clr0, float4* clr1,
           float4* clr2, int* zero_count, GridTape tape, const int* __restrict__ run_if) {
  DSK_TL(k);
  // every scalar of the prologue in one round trip (see list_head)
  const int run = run_if ? load_int_here(run_if) : 1;
  const int n_clr = clr_list ? load_int_here(clr_count) : 0;
  const int clr_first = clr_list ? list_head(k, clr_list, blockIdx.x) : 0;
  const int n_active = load_int_here(count);
  int gt_next = list_head(k, list, blockIdx.x);
  // grid tape: (momentum, mass) and velocity of every active tile are kept so that substep_grad need not
  // recompute p2g + grid_op (mpm_simulator.py:330-333 does); tape.base[j] = first tape slot of substep j
  const int tb = (tape.base && j > 0) ? load_int_here(tape.base + j) : 0;
  if (!run) return;
  const ToolParams* sT = tp.T;   // tool parameters and the frame table arrive as kernel parameters: no setup barrier
  const FrameTable& ft = tp.ft;
  __shared__ TileFrames tf;
  __shared__ ContactGeom geo[MAX_FRAMES][GRID_NODES];
  const int l = threadIdx.x, y = threadIdx.y, tid = y * GRID_NODES + l;
  if (blockIdx.x == 0 && tid == 0 && zero_count) *zero_count = 0;
  if (clr_list && y == 0) clear_tiles(k, clr_list, n_clr, clr_first, clr0, clr1, clr2);
  if (tape.base) {
    if (blockIdx.x == 0 && tid == 0) {
      tape.base[j + 1] = tb + n_active;
      bool over = tb + n_active > tape.cap;
      if (j == 0) *tape.overflow = over ? 1 : 0;
      else if (over) *tape.overflow = 1;
    }
  }
  for (int it = blockIdx.x; it < n_active; it += gridDim.x) {
    int gt = gt_next;
    if (it + (int)gridDim.x < n_active) gt_next = list[it + gridDim.x];   // prefetch: shortens the dependent-load chain
    int env = gt / k.ntile, tile = gt - env * k.ntile;
    int tz = tile % k.nt, ty = (tile / k.nt) % k.nt, tx = tile / (k.nt * k.nt);
    size_t o = ((size_t)gt << 6) + l;
    float4 g = Gin[o];
    bool live = g.w > k.m_eps;
    int I0 = tx * 4 + (l >> 4), I1 = ty * 4 + ((l >> 2) & 3), I2 = tz * 4 + (l & 3);
    float3 gp = f3(mul_rn((float)I0, k.dx), mul_rn((float)I1, k.dx), mul_rn((float)I2, k.dx));
    if (l == 0 && y < ft.n) prepare_tile_frame(k, sT, ft, y, poses, env, j, tx, ty, tz, tf);
    __syncthreads();
    const bool any = tile_any_active(tf, ft.n);
    if (any) {
      if (y < ft.n) {
        if (live && tf.active[y]) {
          const ToolParams& T = sT[ft.tool[y]];
          contact_geometry(T, sdf_kind(T.type), tf.F0[y], tf.F1[y], gp, k.dt, geo[y][l]);
        } else {
          geo[y][l].influence = -1.f;
        }
      }
      __syncthreads();
    }
    if (y == 0) {
      float3 vout = f3(0.f, 0.f, 0.f);
      if (live) {
        float inv = 1.f / g.w;
        float3 v = f3(inv * g.x + k.grav[0], inv * g.y + k.grav[1], inv * g.z + k.grav[2]);
        for (int f = 0; any && f < ft.n; f++) {
          const ContactGeom& c = geo[f][l];
          if (c.influence >= 0.f) v = contact_response(v, c.D, c.cv, c.influence, sT[ft.tool[f]].friction, ft.flag[f] != 0.f);
        }
        vout = grid_boundary(k, I0, I1, I2, v);
      }
      float4 go = make_float4(vout.x, vout.y, vout.z, g.w);
      Gout[o] = go;
      if (tape.base && tb + it < tape.cap) {
        if (l == 0) tape.list[tb + it] = gt;
        float4* d = tape.data + ((size_t)(tb + it) << 7);
        d[l] = g;
        d[64 + l] = go;
      }
    }
    __syncthreads();
  }
}

// ---- grid_op, throughput layout for batched engines ------------------------------------------------------------------
// One thread per node, one warp per half tile, four tiles per CTA and no CTA-wide barrier inside the tile loop: with
// thousands of active tiles the (node x frame) layout above leaves most of its warps idle (few tiles touch a tool), so
// tiles in flight per SM -- not threads per tile -- is what hides the dependent-load latency.  Lanes < n_frames
// prepare and cull the frames of the warp's tile; the surviving frames are applied in order by every lane.
#define FLAT_THREADS 256
#define FLAT_TILES (FLAT_THREADS / GRID_NODES)
struct WarpFrames {   // per warp, shared memory
  Frame F0[MAX_FRAMES], F1[MAX_FRAMES];
};
// Tile of the list a warp pair starts with; it advances by gridDim.x * FLAT_TILES.  Neighbours in the list go to DIFFERENT
// CTAs: the tiles a tool touches are neighbours in the list (it is built in particle order) and cost ~10x the others, and
// with four neighbours per CTA the SMs that got them ran 3x longer than the average (r03d: sm__cycles_active max 43.7 k,
// avg 15.0 k in k_grid_adj_flat with the gripper in contact; 13.0 k / 11.5 k without contact).
DSK_DEV int flat_first_tile(int pair) { return (int)blockIdx.x + pair * (int)gridDim.x; }
DSK_DEV void clear_tiles_flat(const int* __restrict__ list, int count, int first, float4* c0, float4* c1, float4* c2) {
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  const int i0 = flat_first_tile(threadIdx.x >> 6);
  for (int i = i0; i < count; i += gridDim.x * FLAT_TILES) {
    size_t o = ((size_t)(i == i0 ? first : list[i]) << 6) + (threadIdx.x & 63);
    if (c0) c0[o] = z;
    if (c1) c1[o] = z;
    if (c2) c2[o] = z;
  }
}
// returns the warp-uniform mask of frames that may touch the tile
DSK_DEV unsigned warp_prepare_frames(const SimConst& k, const ToolParams* sT, const FrameTable& ft,
                                     const float* __restrict__ poses, int env, int j, int tx, int ty, int tz,
                                     WarpFrames& wf, int lane) {
  int act = 0;
  if (lane < ft.n) act = prepare_frame(k, sT, ft, lane, poses, env, j, tx, ty, tz, wf.F0[lane], wf.F1[lane]);
  unsigned mask = __ballot_sync(0xffffffffu, act);   // also orders the shared-memory writes before the reads below
  __syncwarp();
  return mask;
}

__global__ void __launch_bounds__(FLAT_THREADS, 4)
    k_grid_flat(SimConst k, const __grid_constant__ GridTools tp, const float* __restrict__ poses, int j,
                const float4* Gin, float4* Gout, const int* __restrict__ list, const int* __restrict__ count,
                const int* __restrict__ clr_list, const int* __restrict__ clr_count, float4* clr0, float4* clr1,
                float4* clr2, int* zero_count, GridTape tape, const int* __restrict__ run_if) {
  DSK_TL(k);
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  const int it0 = flat_first_tile(w >> 1), stride = gridDim.x * FLAT_TILES;
  // every scalar of the prologue in one round trip (see list_head); the tile list runs TWO iterations ahead of the tile
  // data, which runs one ahead of the arithmetic
  const int run = run_if ? load_int_here(run_if) : 1;
  const int n_clr = clr_list ? load_int_here(clr_count) : 0;
  const int clr_first = clr_list ? list_head(k, clr_list, it0) : 0;
  const int n_active = load_int_here(count);
  int gt_n = list_head(k, list, it0), gt_nn = list_head(k, list, it0 + stride);
  const int tb = (tape.base && j > 0) ? load_int_here(tape.base + j) : 0;
  if (!run) return;
  const ToolParams* sT = tp.T;   // tool parameters and the frame table arrive as kernel parameters: no setup barrier
  const FrameTable& ft = tp.ft;
  __shared__ WarpFrames wf[FLAT_THREADS / 32];
  if (blockIdx.x == 0 && tid == 0 && zero_count) *zero_count = 0;
  if (clr_list) clear_tiles_flat(clr_list, n_clr, clr_first, clr0, clr1, clr2);
  if (tape.base) {   // as in k_grid
    if (blockIdx.x == 0 && tid == 0) {
      tape.base[j + 1] = tb + n_active;
      bool over = tb + n_active > tape.cap;
      if (j == 0) *tape.overflow = over ? 1 : 0;
      else if (over) *tape.overflow = 1;
    }
  }
  const int l = tid & 63;   // node within the tile
  // software pipeline over the tile loop: the list entry and the tile's 1 KB of the NEXT iteration are loaded before the
  // current tile is processed (list -> grid is a dependent L2 + HBM round trip of ~1 us per iteration otherwise; a batch of
  // 64 envs has ~4 iterations per warp pair)
  float4 g_n = it0 < n_active ? Gin[((size_t)gt_n << 6) + l] : make_float4(0.f, 0.f, 0.f, 0.f);
  for (int it = it0; it < n_active; it += stride) {
    const int gt = gt_n;
    const float4 g = g_n;
    gt_n = gt_nn;
    if (it + stride < n_active) g_n = Gin[((size_t)gt_n << 6) + l];
    if (it + 2 * stride < n_active) gt_nn = list[it + 2 * stride];
    int env = gt / k.ntile, tile = gt - env * k.ntile;
    int tz = tile % k.nt, ty = (tile / k.nt) % k.nt, tx = tile / (k.nt * k.nt);
    size_t o = ((size_t)gt << 6) + l;
    bool live = g.w > k.m_eps;
    int I0 = tx * 4 + (l >> 4), I1 = ty * 4 + ((l >> 2) & 3), I2 = tz * 4 + (l & 3);
    unsigned mask = warp_prepare_frames(k, sT, ft, poses, env, j, tx, ty, tz, wf[w], lane);
    float3 vout = f3(0.f, 0.f, 0.f);
    if (live) {
      float3 gp = f3(mul_rn((float)I0, k.dx), mul_rn((float)I1, k.dx), mul_rn((float)I2, k.dx));
      float inv = 1.f / g.w;
      float3 v = f3(inv * g.x + k.grav[0], inv * g.y + k.grav[1], inv * g.z + k.grav[2]);
      for (unsigned m = mask; m; m &= m - 1) {
        int f = __ffs(m) - 1;
        const ToolParams& T = sT[ft.tool[f]];
        ContactGeom c;
        contact_geometry(T, sdf_kind(T.type), wf[w].F0[f], wf[w].F1[f], gp, k.dt, c);
        if (c.influence >= 0.f) v = contact_response(v, c.D, c.cv, c.influence, T.friction, ft.flag[f] != 0.f);
      }
      vout = grid_boundary(k, I0, I1, I2, v);
    }
    float4 go = make_float4(vout.x, vout.y, vout.z, g.w);
    Gout[o] = go;
    if (tape.base && tb + it < tape.cap) {
      if (l == 0) tape.list[tb + it] = gt;
      float4* d = tape.data + ((size_t)(tb + it) << 7);
      d[l] = g;
      d[64 + l] = go;
    }
    __syncwarp();   // wf[w] is rewritten by the next tile
  }
}

// substep_grad with a valid grid tape: put the taped (momentum, mass) and velocity tiles of substep j back into
// the dense grids and rebuild the active-tile list -- replaces the p2g + grid_op recompute
__global__ void __launch_bounds__(GRID_NODES)
    k_tape_restore(SimConst k, GridTape tape, int j, float4* __restrict__ G0, float4* __restrict__ Gv,
                   int* __restrict__ list, int* __restrict__ count,
                   const int* __restrict__ clr_list, const int* __restrict__ clr_count, float4* clr0, float4* clr1,
                   float4* clr2, int* zero_count) {
  DSK_TL(k);
  // the scalars of the prologue in one round trip (see list_head); this kernel sits on the side branch of the adjoint graph
  const int over = load_int_here(tape.overflow);
  const int n_clr = clr_list ? load_int_here(clr_count) : 0;
  const int clr_first = clr_list ? list_head(k, clr_list, blockIdx.x) : 0;
  const int tb = j == 0 ? 0 : load_int_here(tape.base + j);
  const int n = load_int_here(tape.base + j + 1) - tb;
  if (over) return;   // incomplete tape: the recompute kernels that follow take over
  if (blockIdx.x == 0 && threadIdx.x == 0 && zero_count) *zero_count = 0;
  if (clr_list) clear_tiles(k, clr_list, n_clr, clr_first, clr0, clr1, clr2);
  if (blockIdx.x == 0 && threadIdx.x == 0) *count = n;
  for (int it = blockIdx.x; it < n; it += gridDim.x) {
    int gt = tape.list[tb + it];
    if (threadIdx.x == 0) list[it] = gt;
    size_t o = ((size_t)gt << 6) + threadIdx.x;
    const float4* d = tape.data + ((size_t)(tb + it) << 7);
    G0[o] = d[threadIdx.x];
    Gv[o] = d[64 + threadIdx.x];
  }
}

__global__ void __launch_bounds__(128)
    k_g2p(SimConst k, const float* __restrict__ fin, float* __restrict__ fout, const int* __restrict__ npart,
          const float4* __restrict__ G) {
  DSK_TL(k);
  int gid = blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= k.stride) return;
  int env = gid / k.Npad, p = gid - env * k.Npad;
  if (p >= npart[env]) return;
  float3 x = load_v3(fin, CX, k.stride, gid);
  Stencil s;
  make_stencil(k, x.x, x.y, x.z, s);
  float3 nx, nv;
  M3 nC;
  g2p_particle(k, s, G + (size_t)env * k.nnode, x, nx, nv, nC);
  store_v3(fout, CX, k.stride, gid, nx);
  store_v3(fout, CV, k.stride, gid, nv);
  store_m3(fout, CC, k.stride, gid, nC);
}

// Fused g2p of substep q and p2g of substep q+1 ("G2P2G"): the particle state of frame q+1 is produced and consumed
// in registers (it is still stored: the adjoint needs the frame), one launch fewer per substep.  Gprev = grid set of
// substep q (velocities), Gnext = the other set (scatter target, tiles tracked for substep q+1).
template <int MINB, bool TS>
__global__ void __launch_bounds__(128, MINB)
    k_g2p2g(SimConst k, const float* __restrict__ fprev, float* __restrict__ fcur, float* __restrict__ fnext,
            const float* __restrict__ mat, const int* __restrict__ npart, const float4* __restrict__ Gprev,
            float4* __restrict__ Gnext, TileTrack tt, const StepArgs* __restrict__ args, int qnext,
            float* __restrict__ svd_out) {
  DSK_TL(k);
  const int epoch = load_int_here(&args->epoch_base) + qnext + 1;   // loaded first: the tile tags are compared with it
  int gid = blockIdx.x * blockDim.x + threadIdx.x;
  int env = min(gid / k.Npad, k.B - 1), p = gid - env * k.Npad;
  bool active = gid < k.stride && p < npart[env];
  int g = active ? gid : env * k.Npad;
  float3 x = load_v3(fprev, CX, k.stride, g);
  Stencil s;
  make_stencil(k, x.x, x.y, x.z, s);
  float3 nx, nv;
  M3 C;
  g2p_particle(k, s, Gprev + (size_t)env * k.nnode, x, nx, nv, C);
  M3 F = load_m3(fcur, CF, k.stride, g);   // written by the p2g of substep q
  if (active) {
    store_v3(fcur, CX, k.stride, gid, nx);
    store_v3(fcur, CV, k.stride, gid, nv);
    store_m3(fcur, CC, k.stride, gid, C);
  }
  float mu, lam, ys;
  load_mat(k, mat, g, mu, lam, ys);
  P2GParticle o;
  p2g_particle(k, C, F, mu, lam, ys, o);
  if (active) {
    store_m3(fnext, CF, k.stride, gid, o.newF);
    if (svd_out) store_svd(svd_out, k.stride, gid, o);
  }
  make_stencil(k, nx.x, nx.y, nx.z, s);
  float3 fxv = f3(s.fx, s.fy, s.fz);
  float3 a0 = k.p_mass * nv - k.dx * mv(o.affine, fxv);
  float3 ax = f3(k.dx * o.affine.m[0], k.dx * o.affine.m[3], k.dx * o.affine.m[6]);
  float3 ay = f3(k.dx * o.affine.m[1], k.dx * o.affine.m[4], k.dx * o.affine.m[7]);
  float3 az = f3(k.dx * o.affine.m[2], k.dx * o.affine.m[5], k.dx * o.affine.m[8]);
  scatter27_affine<TS>(k, active, s, Gnext + (size_t)env * k.nnode, tt, true, env, epoch,
                       make_float4(a0.x, a0.y, a0.z, k.p_mass), ax, ay, az);
}

// plane-split fused g2p(q) + p2g(q+1) for small engines (see warp_scatter9): blockDim = (PL_PARTICLES, 3).  Each of
// the three threads of a particle gathers one x-plane of the stencil; the partial sums are exchanged through shared
// memory and added in a fixed order, so the three threads continue with bit-identical state (SVD, return map --
// redundantly, the machine is idle anyway) and each scatters one x-plane of the new stencil.
template <bool TS>
__global__ void __launch_bounds__(PL_PARTICLES * 3)
    k_g2p2g_pl(SimConst k, const float* __restrict__ fprev, float* __restrict__ fcur, float* __restrict__ fnext,
               const float* __restrict__ mat, const int* __restrict__ npart, const float4* __restrict__ Gprev,
               float4* __restrict__ Gnext, TileTrack tt, const StepArgs* __restrict__ args, int qnext,
               float* __restrict__ svd_out) {
  DSK_TL(k);
  const int epoch = load_int_here(&args->epoch_base) + qnext + 1;   // loaded first: the tile tags are compared with it
  __shared__ float ex[3][9][PL_PARTICLES];
  const int tx = threadIdx.x, pl = threadIdx.y;
  int gid = blockIdx.x * PL_PARTICLES + tx;
  int env = min(gid / k.Npad, k.B - 1), p = gid - env * k.Npad;
  bool active = gid < k.stride && p < npart[env];
  int g = active ? gid : env * k.Npad;
  float3 x = load_v3(fprev, CX, k.stride, g);
  M3 F = load_m3(fcur, CF, k.stride, g);   // written by the p2g of substep q
  float mu, lam, ys;
  load_mat(k, mat, g, mu, lam, ys);
  Stencil s;
  make_stencil(k, x.x, x.y, x.z, s);
  {
    const float4* Ge = Gprev + (size_t)env * k.nnode;
    const float wxp = pick3(s.wx, pl);
    const int oxp = pick3(s.ox, pl);
    float3 nvp = f3(0.f, 0.f, 0.f), m1 = f3(0.f, 0.f, 0.f), m2 = f3(0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
      for (int l = 0; l < 3; l++) {
        float4 gv = Ge[oxp + s.oy[j] + s.oz[l]];
        float w = wxp * s.wy[j] * s.wz[l];
        float3 wg = f3(w * gv.x, w * gv.y, w * gv.z);
        nvp += wg;
        if (j) m1 += (float)j * wg;
        if (l) m2 += (float)l * wg;
      }
    float part[9] = {nvp.x, nvp.y, nvp.z, m1.x, m1.y, m1.z, m2.x, m2.y, m2.z};
#pragma unroll
    for (int q = 0; q < 9; q++) ex[pl][q][tx] = part[q];
  }
  __syncthreads();
  float3 nx, nv;
  M3 C;
  {
    float t[3][9];
#pragma unroll
    for (int a = 0; a < 3; a++)
#pragma unroll
      for (int q = 0; q < 9; q++) t[a][q] = ex[a][q][tx];
    nv = f3(t[0][0] + t[1][0] + t[2][0], t[0][1] + t[1][1] + t[2][1], t[0][2] + t[1][2] + t[2][2]);
    float3 m0 = f3(t[1][0] + 2.f * t[2][0], t[1][1] + 2.f * t[2][1], t[1][2] + 2.f * t[2][2]);
    float3 m1 = f3(t[0][3] + t[1][3] + t[2][3], t[0][4] + t[1][4] + t[2][4], t[0][5] + t[1][5] + t[2][5]);
    float3 m2 = f3(t[0][6] + t[1][6] + t[2][6], t[0][7] + t[1][7] + t[2][7], t[0][8] + t[1][8] + t[2][8]);
    g2p_finish(k, s, x, nv, m0, m1, m2, nx, C);
  }
  if (active && pl == 0) {
    store_v3(fcur, CX, k.stride, gid, nx);
    store_v3(fcur, CV, k.stride, gid, nv);
    store_m3(fcur, CC, k.stride, gid, C);
  }
  P2GParticle o;
  p2g_particle(k, C, F, mu, lam, ys, o);
  if (active && pl == 0) {
    store_m3(fnext, CF, k.stride, gid, o.newF);
    if (svd_out) store_svd(svd_out, k.stride, gid, o);
  }
  make_stencil(k, nx.x, nx.y, nx.z, s);
  const float wxp = pick3(s.wx, pl);
  const int oxp = pick3(s.ox, pl);
  float3 fxv = f3(s.fx, s.fy, s.fz);
  float3 ax = f3(k.dx * o.affine.m[0], k.dx * o.affine.m[3], k.dx * o.affine.m[6]);
  float3 ay = f3(k.dx * o.affine.m[1], k.dx * o.affine.m[4], k.dx * o.affine.m[7]);
  float3 az = f3(k.dx * o.affine.m[2], k.dx * o.affine.m[5], k.dx * o.affine.m[8]);
  float3 a0 = k.p_mass * nv - k.dx * mv(o.affine, fxv) + (float)pl * ax;   // plane term folded in
  scatter9<TS>(k, active, s, pl, oxp, Gnext + (size_t)env * k.nnode, tt, pl == 0, env, epoch,
                [&](int j, int l) {
                  float w = wxp * s.wy[j] * s.wz[l];
                  float3 a = a0 + (float)j * ay + (float)l * az;
                  return make_float4(w * a.x, w * a.y, w * a.z, w * k.p_mass);
                });
}
