// Hand-derived adjoint kernels of one substep (MPMSimulator.substep_grad,
// plb/engine/mpm_simulator.py:325-345).  The reference gets these from Taichi's autodiff
// (`g2p.grad`, `grid_op.grad`, `p2g.grad`, `compute_F_tmp.grad`, `forward_kinematics.grad`,
// `apply_collision_projection.grad`, `set_surface_points.grad`) plus the hand-written
// `svd_grad`; here every one is written out by hand and checked against the oracle's tape AD.
//
//   schedule of substep_grad(j):  k_p2g<false> (recompute v_in,m)  ->  k_grid (recompute v_out, out of place)
//        -> k_g2p_adj -> k_grid_adj -> k_p2g_adj       (tool adjoints: k_kinematics_adj once per env step)
#pragma once
#include "kernels_aux.cuh"
#include "kernels_fwd.cuh"

// g2p.grad : reads adjoints of (x,v,C)[j+1], scatters adjoint of grid_v_out, writes the g2p part of x.grad[j]
template <int MINB, bool TS>
__global__ void __launch_bounds__(128, MINB)
    k_g2p_adj(SimConst k, const float* __restrict__ fin, const float* __restrict__ fnext,
              const float* __restrict__ adj_in, float* __restrict__ adj_out, const int* __restrict__ npart,
              const float4* __restrict__ Gv, float4* __restrict__ Ga) {
  DSK_TL(k);
  int gid = blockIdx.x * blockDim.x + threadIdx.x;
  int env = min(gid / k.Npad, k.B - 1), p = gid - env * k.Npad;
  bool active = gid < k.stride && p < npart[env];
  int gi = active ? gid : env * k.Npad;
  float3 x = load_v3(fin, CX, k.stride, gi);
  float3 xn = load_v3(fnext, CX, k.stride, gi);
  float3 gxn = load_v3(adj_in, CX, k.stride, gi);
  float3 gvn = load_v3(adj_in, CV, k.stride, gi);
  M3 gC = load_m3(adj_in, CC, k.stride, gi);
  Stencil s;
  G2PAdj c;
  g2p_adj_begin(k, x, xn, gxn, gvn, gC, s, c);
  const float4* Gve = Gv + (size_t)env * k.nnode;
  float4* Gae = Ga + (size_t)env * k.nnode;
  // adjoint of grid_v_out: warp-aggregated scatter
  TileTrack none{nullptr, nullptr, nullptr};
  scatter27_affine<TS>(k, active, s, Gae, none, false, env, 0, make_float4(c.b0.x, c.b0.y, c.b0.z, 0.f), c.cx, c.cy, c.cz);
  if (!active) return;
  float3 gx = g2p_adj_finish(k, s, Gve, gC, c);
  store_v3(adj_out, CX, k.stride, gid, gx);
}

// plane-split g2p.grad for small engines (see warp_scatter9): blockDim = (PL_PARTICLES, 3)
template <bool TS>
__global__ void __launch_bounds__(PL_PARTICLES * 3)
    k_g2p_adj_pl(SimConst k, const float* __restrict__ fin, const float* __restrict__ fnext,
                 const float* __restrict__ adj_in, float* __restrict__ adj_out, const int* __restrict__ npart,
                 const float4* __restrict__ Gv, float4* __restrict__ Ga) {
  DSK_TL(k);
  __shared__ float ex[3][10][PL_PARTICLES];
  const int tx = threadIdx.x, pl = threadIdx.y;
  int gid = blockIdx.x * PL_PARTICLES + tx;
  int env = min(gid / k.Npad, k.B - 1), p = gid - env * k.Npad;
  bool active = gid < k.stride && p < npart[env];
  int gi = active ? gid : env * k.Npad;
  float3 x = load_v3(fin, CX, k.stride, gi);
  float3 xn = load_v3(fnext, CX, k.stride, gi);
  float3 gxn = load_v3(adj_in, CX, k.stride, gi);
  float3 gvn = load_v3(adj_in, CV, k.stride, gi);
  M3 gC = load_m3(adj_in, CC, k.stride, gi);
  float3 gt = f3((k.x_lo < xn.x && xn.x < k.x_hi) ? gxn.x : 0.f, (k.x_lo < xn.y && xn.y < k.x_hi) ? gxn.y : 0.f,
                 (k.x_lo < xn.z && xn.z < k.x_hi) ? gxn.z : 0.f);
  gvn += k.dt * gt;
  Stencil s;
  make_stencil(k, x.x, x.y, x.z, s);
  const float4* Gve = Gv + (size_t)env * k.nnode;
  float4* Gae = Ga + (size_t)env * k.nnode;
  TileTrack none{nullptr, nullptr, nullptr};
  const float wxp = pick3(s.wx, pl);
  const int oxp = pick3(s.ox, pl);
  float3 cx = f3(k.c_C * gC.m[0], k.c_C * gC.m[3], k.c_C * gC.m[6]);
  float3 cy = f3(k.c_C * gC.m[1], k.c_C * gC.m[4], k.c_C * gC.m[7]);
  float3 cz = f3(k.c_C * gC.m[2], k.c_C * gC.m[5], k.c_C * gC.m[8]);
  float3 b0 = gvn - k.c_C * mv(gC, f3(s.fx, s.fy, s.fz)) + (float)pl * cx;   // plane term folded in
  scatter9<TS>(k, active, s, pl, oxp, Gae, none, false, env, 0, [&](int j, int l) {
    float w = wxp * s.wy[j] * s.wz[l];
    float3 a = b0 + (float)j * cy + (float)l * cz;
    return make_float4(w * a.x, w * a.y, w * a.z, 0.f);
  });
  // this plane's part of the weight / fx adjoints
  float gwxp = 0.f, gwy[3] = {0, 0, 0}, gwz[3] = {0, 0, 0};
  float3 sg = f3(0, 0, 0);
#pragma unroll
  for (int j = 0; j < 3; j++)
#pragma unroll
    for (int l = 0; l < 3; l++) {
      float4 g4 = Gve[oxp + s.oy[j] + s.oz[l]];
      float3 g = f3(g4.x, g4.y, g4.z);
      float3 a = b0 + (float)j * cy + (float)l * cz;
      float gw = dot(g, a);
      sg += (wxp * s.wy[j] * s.wz[l]) * g;
      gwxp += gw * s.wy[j] * s.wz[l];
      gwy[j] += gw * wxp * s.wz[l];
      gwz[l] += gw * wxp * s.wy[j];
    }
  float part[10] = {sg.x, sg.y, sg.z, gwxp, gwy[0], gwy[1], gwy[2], gwz[0], gwz[1], gwz[2]};
#pragma unroll
  for (int q = 0; q < 10; q++) ex[pl][q][tx] = part[q];
  __syncthreads();
  if (pl != 0 || !active) return;
  float t[3][10];
#pragma unroll
  for (int a = 0; a < 3; a++)
#pragma unroll
    for (int q = 0; q < 10; q++) t[a][q] = ex[a][q][tx];
  sg = f3(t[0][0] + t[1][0] + t[2][0], t[0][1] + t[1][1] + t[2][1], t[0][2] + t[1][2] + t[2][2]);
  float3 gf = (-k.c_C) * mTv(gC, sg);
  float dw[3];
  bspline1_grad(s.fx, dw);
  gf.x += t[0][3] * dw[0] + t[1][3] * dw[1] + t[2][3] * dw[2];
  bspline1_grad(s.fy, dw);
  gf.y += (t[0][4] + t[1][4] + t[2][4]) * dw[0] + (t[0][5] + t[1][5] + t[2][5]) * dw[1] + (t[0][6] + t[1][6] + t[2][6]) * dw[2];
  bspline1_grad(s.fz, dw);
  gf.z += (t[0][7] + t[1][7] + t[2][7]) * dw[0] + (t[0][8] + t[1][8] + t[2][8]) * dw[1] + (t[0][9] + t[1][9] + t[2][9]) * dw[2];
  store_v3(adj_out, CX, k.stride, gid, gt + k.inv_dx * gf);
}

struct ContactGeomAdj {   // per (node, frame), shared memory
  float3 gD, gcv;
  float gdist;
};

// phase C of grid_op.grad: geometry adjoints of one (node, frame), reduced over the tile's 64 nodes, then the pose
// adjoints of the frame's tool at substeps j and j+1.  Called by every thread of the CTA (one barrier inside).
DSK_DEV void frame_pose_adjoint(const SimConst& k, const ToolParams* sT, const FrameTable& ft, const TileFrames& tf, int y,
                                int l, float3 gp, const ContactGeom& c, bool contact_frame, float3 gD, float3 gcv,
                                float gdist, float (*red)[2][15], const float* __restrict__ poses, int env, int j,
                                float* __restrict__ pose_adj) {
  if (contact_frame) {
    FrameAdj a0 = frame_adj_zero(), a1 = frame_adj_zero();
    if (c.influence >= 0.f) {
      const ToolParams& T = sT[ft.tool[y]];
      int kind = sdf_kind(T.type);
      // D = qrot(q0, n/L)
      float3 Nl = (1.f / c.L) * c.nraw, gNl = f3(0, 0, 0);
      qrot_adj(tf.F0[y].q, Nl, gD, a0.q, gNl);
      float3 gpl = contact_local_adj(T, kind, tf.F0[y].aux, c.pl, c.nraw, c.L, gNl, gdist, a0.aux);   // + dist = sdf(pl)
      // cv = (qrot(q1, pl) + o1 - p) / dt
      float3 gnp = (1.f / k.dt) * gcv;
      a1.o += gnp;
      qrot_adj(tf.F1[y].q, c.pl, gnp, a1.q, gpl);
      // pl = inv_trans(F0, p)
      float3 unused = f3(0, 0, 0);
      inv_trans_adj(tf.F0[y], gp, gpl, a0, unused);
    }
    float vals[15] = {a0.o.x, a0.o.y, a0.o.z, a0.q.w, a0.q.x, a0.q.y, a0.q.z,
                      a1.o.x, a1.o.y, a1.o.z, a1.q.w, a1.q.x, a1.q.y, a1.q.z, a0.aux};
#pragma unroll
    for (int q = 0; q < 15; q++) {
      float s = warp_sum(vals[q]);
      if ((l & 31) == 0) red[y][l >> 5][q] = s;
    }
  }
  __syncthreads();
  if (l == 0 && contact_frame) {
    FrameAdj a0, a1;
    float r[15];
    for (int q = 0; q < 15; q++) r[q] = red[y][0][q] + red[y][1][q];
    a0.o = f3(r[0], r[1], r[2]); a0.q.w = r[3]; a0.q.x = r[4]; a0.q.y = r[5]; a0.q.z = r[6];
    a1.o = f3(r[7], r[8], r[9]); a1.q.w = r[10]; a1.q.x = r[11]; a1.q.y = r[12]; a1.q.z = r[13];
    a0.aux = r[14];
    a1.aux = 0.f;
    int t = ft.tool[y];
    const float* pa = poses + ((size_t)(env * (k.S + 1) + j) * k.K + t) * 8;
    PoseAdj g0 = pose_adj_zero(), g1 = pose_adj_zero();
    if (ft.flag[y] != 0.f) {   // jaw_frame_adj is linear in the frame adjoint: apply it after the reduction
      jaw_frame_adj(load_pose(pa), ft.flag[y], a0, g0);
      jaw_frame_adj(load_pose(pa + (size_t)k.K * 8), ft.flag[y], a1, g1);
    } else {
      tool_frame_adj(a0, g0);
      tool_frame_adj(a1, g1);
    }
    float* adj0 = pose_adj + ((size_t)(env * (k.S + 1) + j) * k.K + t) * 8;
    float* adj1 = adj0 + (size_t)k.K * 8;
    float v0[8] = {g0.p.x, g0.p.y, g0.p.z, g0.q.w, g0.q.x, g0.q.y, g0.q.z, g0.gap};
    float v1[8] = {g1.p.x, g1.p.y, g1.p.z, g1.q.w, g1.q.x, g1.q.y, g1.q.z, g1.gap};
    for (int q = 0; q < 8; q++) {
      if (v0[q] != 0.f) atomicAdd(adj0 + q, v0[q]);
      if (v1[q] != 0.f) atomicAdd(adj1 + q, v1[q]);
    }
  }
}

// Only the adjoint of (grid_v_in, grid_m) is on the critical path of substep_grad; the tool-pose adjoints are needed
// once per env step (k_kinematics_adj).  With a scratch buffer k_grid_adj parks the per-(node, frame) adjoints of
// the contact geometry (gD, gcv, gdist) there and k_grid_adj_tools turns them into pose adjoints on a side branch.
struct GridAdjScratch {
  float* data;   // [cap][n_frames][7][64], indexed by position in the active-tile list (null: inline pose adjoints)
  int* flags;    // [cap][MAX_FRAMES][2] frame had a contact in the tile (two writers: the half tiles of the flat kernel)
  int cap;
};
// grid_op.grad over the active tiles, blockDim = (64, n_frames).  G0: (momentum, mass) of the recomputed p2g.
// Ga: in = adjoint of grid_v_out (xyz), out = adjoint of (grid_v_in, grid_m), in place.
// pose_adj: [B][S+1][K][8] accumulators.
__global__ void __launch_bounds__(GRID_NODES * MAX_FRAMES)
    k_grid_adj(SimConst k, const __grid_constant__ GridTools tp, const float* __restrict__ poses, int j,
               const float4* __restrict__ G0, float4* __restrict__ Ga, const int* __restrict__ list,
               const int* __restrict__ count, float* __restrict__ pose_adj, GridAdjScratch sc) {
  DSK_TL(k);
  const ToolParams* sT = tp.T;   // tool parameters and the frame table arrive as kernel parameters: no setup barrier
  const FrameTable& ft = tp.ft;
  __shared__ TileFrames tf;
  __shared__ ContactGeom geo[MAX_FRAMES][GRID_NODES];
  __shared__ ContactGeomAdj gadj[MAX_FRAMES][GRID_NODES];
  __shared__ float red[MAX_FRAMES][2][15];
  __shared__ int any_contact[MAX_FRAMES];
  const int l = threadIdx.x, y = threadIdx.y;
  const int n_active = load_int_here(count);    // with the list head: one round trip (list_head, kernels_fwd.cuh)
  int gt_next = list_head(k, list, blockIdx.x);
  for (int it = blockIdx.x; it < n_active; it += gridDim.x) {
    int gt = gt_next;
    if (it + (int)gridDim.x < n_active) gt_next = list[it + gridDim.x];   // prefetch: shortens the dependent-load chain
    int env = gt / k.ntile, tile = gt - env * k.ntile;
    int tz = tile % k.nt, ty = (tile / k.nt) % k.nt, tx = tile / (k.nt * k.nt);
    size_t o = ((size_t)gt << 6) + l;
    float4 gin = G0[o];
    float4 ga4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (y == 0) ga4 = Ga[o];
    bool live = gin.w > k.m_eps;
    int I0 = tx * 4 + (l >> 4), I1 = ty * 4 + ((l >> 2) & 3), I2 = tz * 4 + (l & 3);
    float3 gp = f3(mul_rn((float)I0, k.dx), mul_rn((float)I1, k.dx), mul_rn((float)I2, k.dx));
    if (l == 0 && y < ft.n) {
      prepare_tile_frame(k, sT, ft, y, poses, env, j, tx, ty, tz, tf);
      any_contact[y] = 0;
    }
    __syncthreads();
    const bool any = tile_any_active(tf, ft.n);
    // phase A: contact geometry per (node, frame)
    if (any) {
      if (y < ft.n) {
        if (live && tf.active[y]) {
          const ToolParams& T = sT[ft.tool[y]];
          contact_geometry(T, sdf_kind(T.type), tf.F0[y], tf.F1[y], gp, k.dt, geo[y][l]);
          if (geo[y][l].influence >= 0.f) any_contact[y] = 1;
        } else {
          geo[y][l].influence = -1.f;
        }
      }
      __syncthreads();
    }
    // phase B: velocity chain forward and backward (one thread per node)
    if (y == 0) {
      float4 outv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (live) {
        float inv = 1.f / gin.w;
        float3 vs[MAX_FRAMES];
        float3 v = f3(inv * gin.x + k.grav[0], inv * gin.y + k.grav[1], inv * gin.z + k.grav[2]);
        for (int f = 0; any && f < ft.n; f++) {
          vs[f] = v;
          const ContactGeom& c = geo[f][l];
          if (c.influence >= 0.f) v = contact_response(v, c.D, c.cv, c.influence, sT[ft.tool[f]].friction, ft.flag[f] != 0.f);
        }
        float3 g = grid_boundary_adj(k, I0, I1, I2, v, f3(ga4.x, ga4.y, ga4.z));
        for (int f = ft.n - 1; any && f >= 0; f--) {
          const ContactGeom& c = geo[f][l];
          if (c.influence >= 0.f) {
            float ginfl;
            ContactGeomAdj& a = gadj[f][l];
            g = contact_response_adj(vs[f], c.D, c.cv, c.influence, sT[ft.tool[f]].friction, ft.flag[f] != 0.f, g, a.gD,
                                     a.gcv, ginfl);
            // influence = min(exp(-dist*softness), 1): exp(..) = influence when it is < 1
            float soft = sT[ft.tool[f]].softness;
            a.gdist = (c.influence < 1.f) ? (-soft * c.influence * ginfl) : 0.f;
          }
        }
        outv = make_float4(inv * g.x, inv * g.y, inv * g.z, -(inv * inv) * (gin.x * g.x + gin.y * g.y + gin.z * g.z));
      }
      Ga[o] = outv;
    }
    const bool park = sc.data != nullptr && it < sc.cap;   // uniform
    if (park && l == 0 && y < ft.n && !any) sc.flags[(it * MAX_FRAMES + y) * 2] = sc.flags[(it * MAX_FRAMES + y) * 2 + 1] = 0;
    if (!any) {   // no tool near this tile: nothing to differentiate through (uniform branch)
      __syncthreads();
      continue;
    }
    __syncthreads();
    if (park) {
      if (y < ft.n) {
        if (l == 0) {
          sc.flags[(it * MAX_FRAMES + y) * 2] = any_contact[y];
          sc.flags[(it * MAX_FRAMES + y) * 2 + 1] = 0;
        }
        if (any_contact[y]) {
          const ContactGeomAdj& a = gadj[y][l];
          bool hit = geo[y][l].influence >= 0.f;
          float* d = sc.data + ((size_t)(it * ft.n + y) * 7) * GRID_NODES + l;
          d[0 * GRID_NODES] = hit ? a.gD.x : 0.f;  d[1 * GRID_NODES] = hit ? a.gD.y : 0.f;  d[2 * GRID_NODES] = hit ? a.gD.z : 0.f;
          d[3 * GRID_NODES] = hit ? a.gcv.x : 0.f; d[4 * GRID_NODES] = hit ? a.gcv.y : 0.f; d[5 * GRID_NODES] = hit ? a.gcv.z : 0.f;
          d[6 * GRID_NODES] = hit ? a.gdist : 0.f;
        }
      }
      __syncthreads();
      continue;
    }
    frame_pose_adjoint(k, sT, ft, tf, y, l, gp, geo[y][l], y < ft.n && any_contact[y], gadj[y][l].gD, gadj[y][l].gcv,
                       gadj[y][l].gdist, red, poses, env, j, pose_adj);
    __syncthreads();
  }
}

// second half of the split k_grid_adj (side branch): pose adjoints from the parked (gD, gcv, gdist); the contact
// geometry is recomputed, which is free off the critical path
__global__ void __launch_bounds__(GRID_NODES * MAX_FRAMES)
    k_grid_adj_tools(SimConst k, const __grid_constant__ GridTools tp, const float* __restrict__ poses, int j,
                     const float4* __restrict__ G0, const int* __restrict__ list, const int* __restrict__ count,
                     float* __restrict__ pose_adj, GridAdjScratch sc) {
  DSK_TL(k);
  const ToolParams* sT = tp.T;   // tool parameters and the frame table arrive as kernel parameters: no setup barrier
  const FrameTable& ft = tp.ft;
  __shared__ TileFrames tf;
  __shared__ float red[MAX_FRAMES][2][15];
  const int l = threadIdx.x, y = threadIdx.y;
  int n_active = min(*count, sc.cap);
  for (int it = blockIdx.x; it < n_active; it += gridDim.x) {
    int anyf = 0;
    for (int f = 0; f < ft.n; f++) anyf |= sc.flags[(it * MAX_FRAMES + f) * 2] | sc.flags[(it * MAX_FRAMES + f) * 2 + 1];
    if (!anyf) continue;   // uniform
    const bool fl = y < ft.n && (sc.flags[(it * MAX_FRAMES + y) * 2] | sc.flags[(it * MAX_FRAMES + y) * 2 + 1]) != 0;
    int gt = list[it];
    int env = gt / k.ntile, tile = gt - env * k.ntile;
    int tz = tile % k.nt, ty = (tile / k.nt) % k.nt, tx = tile / (k.nt * k.nt);
    size_t o = ((size_t)gt << 6) + l;
    bool live = G0[o].w > k.m_eps;
    int I0 = tx * 4 + (l >> 4), I1 = ty * 4 + ((l >> 2) & 3), I2 = tz * 4 + (l & 3);
    float3 gp = f3(mul_rn((float)I0, k.dx), mul_rn((float)I1, k.dx), mul_rn((float)I2, k.dx));
    if (l == 0 && fl) prepare_tile_frame(k, sT, ft, y, poses, env, j, tx, ty, tz, tf);
    __syncthreads();
    ContactGeom c;
    c.influence = -1.f;
    float3 gD = f3(0, 0, 0), gcv = f3(0, 0, 0);
    float gdist = 0.f;
    if (fl) {
      if (live && tf.active[y]) {
        const ToolParams& T = sT[ft.tool[y]];
        contact_geometry(T, sdf_kind(T.type), tf.F0[y], tf.F1[y], gp, k.dt, c);
      }
      const float* d = sc.data + ((size_t)(it * ft.n + y) * 7) * GRID_NODES + l;
      gD = f3(d[0 * GRID_NODES], d[1 * GRID_NODES], d[2 * GRID_NODES]);
      gcv = f3(d[3 * GRID_NODES], d[4 * GRID_NODES], d[5 * GRID_NODES]);
      gdist = d[6 * GRID_NODES];
    }
    frame_pose_adjoint(k, sT, ft, tf, y, l, gp, c, fl, gD, gcv, gdist, red, poses, env, j, pose_adj);
    __syncthreads();
  }
}

// grid_op.grad in the throughput layout of k_grid_flat: one thread per node keeps the geometry of its active frames
// in local memory, walks the velocity chain forward and backward, and the pose adjoints of a frame are reduced over
// the warp (half tile) before the atomics.
__global__ void __launch_bounds__(FLAT_THREADS, 4)
    k_grid_adj_flat(SimConst k, const __grid_constant__ GridTools tp, const float* __restrict__ poses, int j,
                    const float4* __restrict__ G0, float4* __restrict__ Ga, const int* __restrict__ list,
                    const int* __restrict__ count, float* __restrict__ pose_adj, GridAdjScratch sc) {
  DSK_TL(k);
  const ToolParams* sT = tp.T;   // tool parameters and the frame table arrive as kernel parameters: no setup barrier
  const FrameTable& ft = tp.ft;
  __shared__ WarpFrames wf[FLAT_THREADS / 32];
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  const int l = tid & 63;
  // software pipeline over the tile loop (see k_grid_flat): the tile list runs two iterations ahead, (momentum, mass) and
  // adjoint tile one iteration ahead of the arithmetic; count and both list heads come in one round trip
  const int it0 = flat_first_tile(w >> 1), stride = gridDim.x * FLAT_TILES;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const int n_active = load_int_here(count);
  int gt_n = list_head(k, list, it0), gt_nn = list_head(k, list, it0 + stride);
  float4 gin_n = it0 < n_active ? G0[((size_t)gt_n << 6) + l] : z4;
  float4 ga_n = it0 < n_active ? Ga[((size_t)gt_n << 6) + l] : z4;
  for (int it = it0; it < n_active; it += stride) {
    const int gt = gt_n;
    const float4 gin = gin_n, ga4 = ga_n;
    gt_n = gt_nn;
    if (it + stride < n_active) {
      gin_n = G0[((size_t)gt_n << 6) + l];
      ga_n = Ga[((size_t)gt_n << 6) + l];
    }
    if (it + 2 * stride < n_active) gt_nn = list[it + 2 * stride];
    int env = gt / k.ntile, tile = gt - env * k.ntile;
    int tz = tile % k.nt, ty = (tile / k.nt) % k.nt, tx = tile / (k.nt * k.nt);
    size_t o = ((size_t)gt << 6) + l;
    bool live = gin.w > k.m_eps;
    int I0 = tx * 4 + (l >> 4), I1 = ty * 4 + ((l >> 2) & 3), I2 = tz * 4 + (l & 3);
    float3 gp = f3(mul_rn((float)I0, k.dx), mul_rn((float)I1, k.dx), mul_rn((float)I2, k.dx));
    unsigned mask = warp_prepare_frames(k, sT, ft, poses, env, j, tx, ty, tz, wf[w], lane);
    ContactGeom geo[MAX_FRAMES];
    float3 vs[MAX_FRAMES];
    float inv = live ? 1.f / gin.w : 0.f;
    float3 v = f3(inv * gin.x + k.grav[0], inv * gin.y + k.grav[1], inv * gin.z + k.grav[2]);
    int na = 0;
    for (unsigned m = mask; m; m &= m - 1, na++) {
      int f = __ffs(m) - 1;
      const ToolParams& T = sT[ft.tool[f]];
      geo[na].influence = -1.f;
      if (live) {
        contact_geometry(T, sdf_kind(T.type), wf[w].F0[f], wf[w].F1[f], gp, k.dt, geo[na]);
        vs[na] = v;
        if (geo[na].influence >= 0.f)
          v = contact_response(v, geo[na].D, geo[na].cv, geo[na].influence, T.friction, ft.flag[f] != 0.f);
      }
    }
    float3 g = f3(0.f, 0.f, 0.f);
    if (live) g = grid_boundary_adj(k, I0, I1, I2, v, f3(ga4.x, ga4.y, ga4.z));
    // With a scratch the pose adjoints leave this kernel (and the critical path): the warp parks (gD, gcv, gdist) of its 32
    // nodes for every frame near the tile and k_grid_adj_tools reduces them on the side branch, as for k_grid_adj.  A tile in
    // contact is a ~4 000-instruction serial chain in a few divergent lanes otherwise: with the gripper in contact the
    // slowest SM was busy 3x longer than the average (r03d: 25 us per launch, 9.7 us without contact).
    const bool park = sc.data != nullptr && it < sc.cap;   // warp-uniform
    unsigned hitmask = 0;
    // frames in reverse order; every lane of the warp walks the same frames
    for (int a = na - 1; a >= 0; a--) {
      unsigned rest = mask;
      for (int q = 0; q < a; q++) rest &= rest - 1;
      int f = __ffs(rest) - 1;
      const ContactGeom& c = geo[a];
      bool hit = c.influence >= 0.f;
      const bool anyhit = __any_sync(0xffffffffu, hit);
      if (park) {
        float3 gD = f3(0, 0, 0), gcv = f3(0, 0, 0);
        float gdist = 0.f;
        if (hit) {
          float ginfl;
          const ToolParams& T = sT[ft.tool[f]];
          g = contact_response_adj(vs[a], c.D, c.cv, c.influence, T.friction, ft.flag[f] != 0.f, g, gD, gcv, ginfl);
          gdist = (c.influence < 1.f) ? (-T.softness * c.influence * ginfl) : 0.f;
        }
        if (anyhit) {   // the other half of the tile cannot know: both halves write their nodes whenever THEY had a hit, and
          hitmask |= 1u << f;   // the reader takes a (tile, frame) if either half flagged it -- see the zero fill below
        }
        float* d = sc.data + ((size_t)(it * ft.n + f) * 7) * GRID_NODES + l;
        d[0 * GRID_NODES] = gD.x;  d[1 * GRID_NODES] = gD.y;  d[2 * GRID_NODES] = gD.z;
        d[3 * GRID_NODES] = gcv.x; d[4 * GRID_NODES] = gcv.y; d[5 * GRID_NODES] = gcv.z;
        d[6 * GRID_NODES] = gdist;
        continue;
      }
      if (!anyhit) continue;
      const ToolParams& T = sT[ft.tool[f]];
      FrameAdj a0 = frame_adj_zero(), a1 = frame_adj_zero();
      if (hit) {
        float ginfl;
        float3 gD, gcv;
        g = contact_response_adj(vs[a], c.D, c.cv, c.influence, T.friction, ft.flag[f] != 0.f, g, gD, gcv, ginfl);
        // influence = min(exp(-dist*softness), 1): exp(..) = influence when it is < 1
        float gdist = (c.influence < 1.f) ? (-T.softness * c.influence * ginfl) : 0.f;
        int kind = sdf_kind(T.type);
        // D = qrot(q0, n/L)
        float3 Nl = (1.f / c.L) * c.nraw, gNl = f3(0, 0, 0);
        qrot_adj(wf[w].F0[f].q, Nl, gD, a0.q, gNl);
        float3 gpl = contact_local_adj(T, kind, wf[w].F0[f].aux, c.pl, c.nraw, c.L, gNl, gdist, a0.aux);   // + dist = sdf(pl)
        // cv = (qrot(q1, pl) + o1 - p) / dt
        float3 gnp = (1.f / k.dt) * gcv;
        a1.o += gnp;
        qrot_adj(wf[w].F1[f].q, c.pl, gnp, a1.q, gpl);
        // pl = inv_trans(F0, p)
        float3 unused = f3(0, 0, 0);
        inv_trans_adj(wf[w].F0[f], gp, gpl, a0, unused);
      }
      float vals[15] = {a0.o.x, a0.o.y, a0.o.z, a0.q.w, a0.q.x, a0.q.y, a0.q.z,
                        a1.o.x, a1.o.y, a1.o.z, a1.q.w, a1.q.x, a1.q.y, a1.q.z, a0.aux};
#pragma unroll
      for (int q = 0; q < 15; q++) vals[q] = warp_sum(vals[q]);
      if (lane == 0) {
        a0.o = f3(vals[0], vals[1], vals[2]); a0.q.w = vals[3]; a0.q.x = vals[4]; a0.q.y = vals[5]; a0.q.z = vals[6];
        a1.o = f3(vals[7], vals[8], vals[9]); a1.q.w = vals[10]; a1.q.x = vals[11]; a1.q.y = vals[12]; a1.q.z = vals[13];
        a0.aux = vals[14];
        a1.aux = 0.f;
        int t = ft.tool[f];
        const float* pa = poses + ((size_t)(env * (k.S + 1) + j) * k.K + t) * 8;
        PoseAdj g0 = pose_adj_zero(), g1 = pose_adj_zero();
        if (ft.flag[f] != 0.f) {   // jaw_frame_adj is linear in the frame adjoint: apply it after the reduction
          jaw_frame_adj(load_pose(pa), ft.flag[f], a0, g0);
          jaw_frame_adj(load_pose(pa + (size_t)k.K * 8), ft.flag[f], a1, g1);
        } else {
          tool_frame_adj(a0, g0);
          tool_frame_adj(a1, g1);
        }
        float* adj0 = pose_adj + ((size_t)(env * (k.S + 1) + j) * k.K + t) * 8;
        float* adj1 = adj0 + (size_t)k.K * 8;
        float v0[8] = {g0.p.x, g0.p.y, g0.p.z, g0.q.w, g0.q.x, g0.q.y, g0.q.z, g0.gap};
        float v1[8] = {g1.p.x, g1.p.y, g1.p.z, g1.q.w, g1.q.x, g1.q.y, g1.q.z, g1.gap};
        for (int q = 0; q < 8; q++) {
          if (v0[q] != 0.f) atomicAdd(adj0 + q, v0[q]);
          if (v1[q] != 0.f) atomicAdd(adj1 + q, v1[q]);
        }
      }
    }
    if (park && lane < ft.n) sc.flags[(it * MAX_FRAMES + lane) * 2 + (w & 1)] = (hitmask >> lane) & 1u;   // every flag, every substep
    float4 outv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live) outv = make_float4(inv * g.x, inv * g.y, inv * g.z, -(inv * inv) * (gin.x * g.x + gin.y * g.y + gin.z * g.z));
    Ga[o] = outv;
    __syncwarp();   // wf[w] is rewritten by the next tile
  }
}

// p2g.grad + svd_grad + compute_F_tmp.grad fused: gathers adjoints of (grid_v_in, grid_m), reads F.grad[j+1],
// writes x.grad (adding the g2p part already stored), v.grad, C.grad, F.grad of frame j.
template <int MINB>
__global__ void __launch_bounds__(128, MINB)
    k_p2g_adj(SimConst k, const float* __restrict__ fin, const float* __restrict__ adj_in, float* __restrict__ adj_out,
              const float* __restrict__ mat, const int* __restrict__ npart, const float4* __restrict__ Ga,
              const float* __restrict__ svd_in) {
  DSK_TL(k);
  int gid = blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= k.stride) return;
  int env = gid / k.Npad, p = gid - env * k.Npad;
  if (p >= npart[env]) return;
  p2g_adj_particle(k, gid, env, fin, adj_in, adj_out, mat, Ga, svd_in);
}

// plane-split p2g.grad for small engines: three threads per particle gather one x-plane of the stencil each
__global__ void __launch_bounds__(PL_PARTICLES * 3)
    k_p2g_adj_pl(SimConst k, const float* __restrict__ fin, const float* __restrict__ adj_in, float* __restrict__ adj_out,
                 const float* __restrict__ mat, const int* __restrict__ npart, const float4* __restrict__ Ga,
                 const float* __restrict__ svd_in) {
  DSK_TL(k);
  __shared__ float ex[3][16][PL_PARTICLES];
  const int tx = threadIdx.x, pl = threadIdx.y;
  int gid = blockIdx.x * PL_PARTICLES + tx;
  int env = min(gid / k.Npad, k.B - 1), p = gid - env * k.Npad;
  bool active = gid < k.stride && p < npart[env];
  int gi = active ? gid : env * k.Npad;
  float3 x = load_v3(fin, CX, k.stride, gi);
  float3 v = load_v3(fin, CV, k.stride, gi);
  M3 C = load_m3(fin, CC, k.stride, gi);
  M3 F = load_m3(fin, CF, k.stride, gi);
  float mu, lam, ys;
  load_mat(k, mat, gi, mu, lam, ys);
  P2GParticle o;
  p2g_particle_adj(k, svd_in, gi, C, F, mu, lam, ys, o);
  Stencil s;
  make_stencil(k, x.x, x.y, x.z, s);
  const float4* Gae = Ga + (size_t)env * k.nnode;
  {
    const float wxp = pick3(s.wx, pl);
    const int oxp = pick3(s.ox, pl);
    float3 fxv = f3(s.fx, s.fy, s.fz);
    float3 ax = f3(k.dx * o.affine.m[0], k.dx * o.affine.m[3], k.dx * o.affine.m[6]);
    float3 ay = f3(k.dx * o.affine.m[1], k.dx * o.affine.m[4], k.dx * o.affine.m[7]);
    float3 az = f3(k.dx * o.affine.m[2], k.dx * o.affine.m[5], k.dx * o.affine.m[8]);
    float3 a0 = k.p_mass * v - k.dx * mv(o.affine, fxv) + (float)pl * ax;
    float gwxp = 0.f, gwy[3] = {0, 0, 0}, gwz[3] = {0, 0, 0};
    float3 S0 = f3(0, 0, 0), m1 = f3(0, 0, 0), m2 = f3(0, 0, 0);
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
      for (int l = 0; l < 3; l++) {
        float4 g4 = Gae[oxp + s.oy[j] + s.oz[l]];
        float3 G = f3(g4.x, g4.y, g4.z);
        float3 a = a0 + (float)j * ay + (float)l * az;
        float gw = dot(G, a) + g4.w * k.p_mass;
        float3 wG = (wxp * s.wy[j] * s.wz[l]) * G;
        S0 += wG;
        if (j) m1 += (float)j * wG;
        if (l) m2 += (float)l * wG;
        gwxp += gw * s.wy[j] * s.wz[l];
        gwy[j] += gw * wxp * s.wz[l];
        gwz[l] += gw * wxp * s.wy[j];
      }
    float part[16] = {S0.x, S0.y, S0.z, m1.x, m1.y, m1.z, m2.x, m2.y, m2.z, gwxp, gwy[0], gwy[1], gwy[2], gwz[0], gwz[1], gwz[2]};
#pragma unroll
    for (int q = 0; q < 16; q++) ex[pl][q][tx] = part[q];
  }
  __syncthreads();
  if (pl != 0 || !active) return;
  float t[3][16];
#pragma unroll
  for (int a = 0; a < 3; a++)
#pragma unroll
    for (int q = 0; q < 16; q++) t[a][q] = ex[a][q][tx];
  float3 S0 = f3(t[0][0] + t[1][0] + t[2][0], t[0][1] + t[1][1] + t[2][1], t[0][2] + t[1][2] + t[2][2]);
  float3 m0 = f3(t[1][0] + 2.f * t[2][0], t[1][1] + 2.f * t[2][1], t[1][2] + 2.f * t[2][2]);
  float3 m1 = f3(t[0][3] + t[1][3] + t[2][3], t[0][4] + t[1][4] + t[2][4], t[0][5] + t[1][5] + t[2][5]);
  float3 m2 = f3(t[0][6] + t[1][6] + t[2][6], t[0][7] + t[1][7] + t[2][7], t[0][8] + t[1][8] + t[2][8]);
  float gwx[3] = {t[0][9], t[1][9], t[2][9]};
  float gwy[3] = {t[0][10] + t[1][10] + t[2][10], t[0][11] + t[1][11] + t[2][11], t[0][12] + t[1][12] + t[2][12]};
  float gwz[3] = {t[0][13] + t[1][13] + t[2][13], t[0][14] + t[1][14] + t[2][14], t[0][15] + t[1][15] + t[2][15]};
  p2g_adj_finish(k, gid, s, o, mu, lam, C, F, adj_in, adj_out, S0, m0, m1, m2, gwx, gwy, gwz);
}

// Tool adjoints of one env step: for j = S-1..0: apply_collision_projection.grad, set_surface_points.grad,
// forward_kinematics.grad (tools in reverse order), then set_velocity.grad (primive_base.py:260-268).
// The kinematics of a substep is a tiny map pose_j, u -> pose_{j+1}; its adjoint is linear in the incoming
// adjoint, so phase 1 evaluates the 8x8 / 8x7 Jacobian blocks of ALL substeps in parallel (one thread per
// (substep, tool, output component), unit seeds through tool_fk_adj) and phase 2 walks the chain with 8
// multiply-adds per thread per substep.  One CTA per env.
#define KINADJ_CTA 256
DSK_DEV void projection_adj(const SimConst& k, const ToolParams* sT, const float* P1, const float* __restrict__ rand_num,
                            int c, int idx, float* a1) {
  int ti = k.pairs[c][0], tj = k.pairs[c][1];
  // tool i is evaluated at its POST-projection pose, as Taichi's grad kernel re-evaluates the forward on the
  // current field values; the obstacle pose is the post-step pose of frame j+1
  const float* rn = rand_num + ((size_t)c * DSK_NUM_COLLISION_POINTS + idx) * 3;
  Pose Pj = load_pose(P1 + tj * 8), Pi = load_pose(P1 + ti * 8);
  float3 pt = surface_point(sT[tj], Pj, rn);
  float d = tool_sdf(sT[ti], Pi, pt);
  float3 nr = tool_normal(sT[ti], Pi, pt);
  float n2 = dot(nr, nr);
  float inv = 1.f / sqrtf(n2);
  // new_pos = pos + (inv*nr)*d ; seed = current position.grad
  float3 gs = f3(a1[ti * 8 + 0], a1[ti * 8 + 1], a1[ti * 8 + 2]);
  float3 un = inv * nr;
  float gd = dot(gs, un);
  float3 gun = d * gs;
  float ginv = dot(gun, nr);
  float3 gnr = inv * gun;
  gnr += (2.f * (-0.5f * ginv * inv / n2)) * nr;
  PoseAdj gPi = pose_adj_zero();
  float3 gpt = f3(0, 0, 0);
  tool_sdf_adj(sT[ti], Pi, pt, gd, gPi, gpt);
  tool_normal_adj(sT[ti], Pi, pt, gnr, gPi, gpt);
  a1[ti * 8 + 0] += gPi.p.x; a1[ti * 8 + 1] += gPi.p.y; a1[ti * 8 + 2] += gPi.p.z;
  a1[ti * 8 + 3] += gPi.q.w; a1[ti * 8 + 4] += gPi.q.x; a1[ti * 8 + 5] += gPi.q.y; a1[ti * 8 + 6] += gPi.q.z;
  a1[ti * 8 + 7] += gPi.gap;
  // set_surface_points.grad: pt = qrot(rot_j, proj) + pos_j
  float3 q;
  q.x = tmax(tmin(rn[0] * sT[tj].size[0], sT[tj].size[0]), -sT[tj].size[0]);
  q.y = tmax(tmin(rn[1] * sT[tj].size[1], sT[tj].size[1]), -sT[tj].size[1]);
  q.z = tmax(tmin(rn[2] * sT[tj].size[2], sT[tj].size[2]), -sT[tj].size[2]);
  Q4 gq = {0, 0, 0, 0};
  float3 gdm = f3(0, 0, 0);
  qrot_adj(Pj.q, q, gpt, gq, gdm);
  a1[tj * 8 + 0] += gpt.x; a1[tj * 8 + 1] += gpt.y; a1[tj * 8 + 2] += gpt.z;
  a1[tj * 8 + 3] += gq.w; a1[tj * 8 + 4] += gq.x; a1[tj * 8 + 5] += gq.y; a1[tj * 8 + 6] += gq.z;
}
__global__ void __launch_bounds__(KINADJ_CTA)
    k_kinematics_adj(SimConst k, const ToolParams* __restrict__ tools, const float* __restrict__ poses,
                     const int* __restrict__ cidx, const float* __restrict__ rand_num,
                     const StepArgs* __restrict__ args, float* __restrict__ pose_adj) {
  DSK_TL(k);
  const float* __restrict__ action = args->action;
  float* __restrict__ action_grad = args->action_grad;  // [B][A] of this step, +=
  __shared__ ToolParams sT[DSK_MAX_TOOLS];
  __shared__ float s_gu[DSK_MAX_TOOLS][8];
  __shared__ int s_any_hit;
  DSK_DYN_SMEM(float, dyn);
  int env = blockIdx.x, tid = threadIdx.x;
  int tot = (k.S + 1) * k.K * 8;
  float* sadj = dyn;                          // [(S+1)][K][8]
  float* Jp = sadj + tot;                     // [S][K][8 out][8 in]
  float* Ju = Jp + (size_t)k.S * k.K * 64;    // [S][K][8 out][8 (7 used)]
  for (int i = tid; i < k.K * (int)(sizeof(ToolParams) / 4); i += blockDim.x) ((int*)sT)[i] = ((const int*)tools)[i];
  float* gadj = pose_adj + (size_t)env * tot;
  const float* P = poses + (size_t)env * tot;
  for (int i = tid; i < tot; i += blockDim.x) sadj[i] = gadj[i];
  if (tid < DSK_MAX_TOOLS * 8) s_gu[tid / 8][tid % 8] = 0.f;
  if (tid == 0) s_any_hit = 0;
  __syncthreads();
  // did any tool-tool projection fire in this step?  (almost never: then the chain below needs no projection adjoint)
  for (int i = tid; i < k.S * k.npairs; i += blockDim.x)
    if (cidx[((size_t)env * (k.S + 1) + 1) * k.npairs + i] >= 0) s_any_hit = 1;
  __syncthreads();
  const bool any_hit = s_any_hit != 0;
  int A = 0;
  for (int t = 0; t < k.K; t++) A += sT[t].action_dim;
  float zero[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  // phase 1: Jacobian blocks of every (substep, tool), one output component per thread
  for (int w = tid; w < k.S * k.K * 8; w += blockDim.x) {
    int r = w & 7, t = (w >> 3) % k.K, j = (w >> 3) / k.K;
    int off = 0;
    for (int q = 0; q < t; q++) off += sT[q].action_dim;
    ToolVel u = action_to_vel(sT[t], action ? action + (size_t)env * A + off : zero, k.S), gu;
    gu.v = f3(0, 0, 0);
    gu.w = f3(0, 0, 0);
    gu.gap_vel = 0.f;
    PoseAdj gN = pose_adj_zero(), gP = pose_adj_zero();
    float* gn = (float*)&gN;   // PoseAdj is 8 packed floats: p3, q4, gap
    gn[r] = 1.f;
    tool_fk_adj(sT[t], load_pose(P + ((size_t)j * k.K + t) * 8), u, gN, gP, gu);
    float* o = Jp + (size_t)w * 8;
    o[0] = gP.p.x; o[1] = gP.p.y; o[2] = gP.p.z; o[3] = gP.q.w; o[4] = gP.q.x; o[5] = gP.q.y; o[6] = gP.q.z; o[7] = gP.gap;
    float* ou = Ju + (size_t)w * 8;
    ou[0] = gu.v.x; ou[1] = gu.v.y; ou[2] = gu.v.z; ou[3] = gu.w.x; ou[4] = gu.w.y; ou[5] = gu.w.z; ou[6] = gu.gap_vel; ou[7] = 0.f;
  }
  __syncthreads();
  // phase 2: walk the chain
  int t = tid >> 3, c = tid & 7;
  float gu_acc = 0.f;
  for (int j = k.S - 1; j >= 0; j--) {
    float* a1 = sadj + (size_t)(j + 1) * k.K * 8;
    float* a0 = sadj + (size_t)j * k.K * 8;
    if (any_hit) {
      if (tid == 0) {
        for (int cc = k.npairs - 1; cc >= 0; cc--) {
          int idx = cidx[((size_t)env * (k.S + 1) + (j + 1)) * k.npairs + cc];
          if (idx >= 0) projection_adj(k, sT, P + (size_t)(j + 1) * k.K * 8, rand_num, cc, idx, a1);
        }
      }
      __syncthreads();
    }
    if (t < k.K) {
      const float* jp = Jp + ((size_t)j * k.K + t) * 64;
      const float* ju = Ju + ((size_t)j * k.K + t) * 64;
      const float* g = a1 + t * 8;
      float sp = 0.f, su = 0.f;
#pragma unroll
      for (int r = 0; r < 8; r++) {
        sp += jp[r * 8 + c] * g[r];
        su += ju[r * 8 + c] * g[r];
      }
      a0[t * 8 + c] += sp;
      gu_acc += su;
    }
    __syncthreads();
  }
  for (int i = tid; i < tot; i += blockDim.x) gadj[i] = sadj[i];
  if (t < k.K) s_gu[t][c] = gu_acc;
  __syncthreads();
  if (tid < k.K && sT[tid].action_dim > 0 && action_grad) {
    const ToolParams& T = sT[tid];
    int off = 0;
    for (int q = 0; q < tid; q++) off += sT[q].action_dim;
    float fs = (float)k.S;
    float* ga = action_grad + (size_t)env * A + off;
    const float* gu = s_gu[tid];
    ga[0] += gu[0] * (T.action_scale[0] / fs);
    ga[1] += gu[1] * (T.action_scale[1] / fs);
    ga[2] += gu[2] * (T.action_scale[2] / fs);
    if (T.action_dim > 3) {
      ga[3] += gu[3] * (T.action_scale[3] / fs);
      ga[4] += gu[4] * (T.action_scale[4] / fs);
      ga[5] += gu[5] * (T.action_scale[5] / fs);
    }
    if (has_gap(T.type)) ga[6] += gu[6] * (T.action_scale[6] / fs);
  }
}

// compute_min_dist (+grad), function.py:79-88 : one thread per particle, loops over tools.
// out [B,cap_out,ncols]
__global__ void k_min_dist(SimConst k, const ToolParams* __restrict__ tools, const float* __restrict__ frame,
                           const int* __restrict__ npart, const float* __restrict__ tool_state /*[B][K][8]*/,
                           int cap_out, int ncols, float* __restrict__ out) {
  DSK_TL(k);
  int gid = blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= k.stride) return;
  int env = gid / k.Npad, p = gid - env * k.Npad;
  if (p >= cap_out) return;
  float* o = out + ((size_t)env * cap_out + p) * ncols;
  if (p >= npart[env]) {
    for (int c = 0; c < ncols; c++) o[c] = 0.f;
    return;
  }
  float3 x = load_v3(frame, CX, k.stride, gid);
  int col = 0;
  for (int t = 0; t < k.K; t++) {
    ToolParams T = tools[t];
    Pose P = load_pose(tool_state + ((size_t)env * k.K + t) * 8);
    if (is_gripper(T.type)) {
      o[col++] = frame_sdf(T, sdf_kind(T.type), jaw_frame(P, -1.f), x);
      o[col++] = frame_sdf(T, sdf_kind(T.type), jaw_frame(P, 1.f), x);
    } else {
      o[col++] = tool_sdf(T, P, x);
    }
  }
}
__global__ void k_min_dist_adj(SimConst k, const ToolParams* __restrict__ tools, const float* __restrict__ frame,
                               const int* __restrict__ npart, const float* __restrict__ tool_state, int cap_in,
                               int ncols, const float* __restrict__ gout, float* __restrict__ adj_frame,
                               float* __restrict__ tool_adj /*[B][K][8]*/) {
  DSK_TL(k);
  int gid = blockIdx.x * blockDim.x + threadIdx.x;
  int env = gid < k.stride ? gid / k.Npad : 0, p = gid - env * k.Npad;
  bool live = gid < k.stride && p < npart[env] && p < cap_in;
  float3 x = live ? load_v3(frame, CX, k.stride, gid) : f3(0, 0, 0);
  const float* g = gout + ((size_t)env * cap_in + (live ? p : 0)) * ncols;
  float3 gx = f3(0, 0, 0);
  int col = 0;
  // all lanes of a warp belong to one env when Npad is a multiple of 32 (it is)
  for (int t = 0; t < k.K; t++) {
    ToolParams T = tools[t];
    PoseAdj gP = pose_adj_zero();
    if (live) {
      Pose P = load_pose(tool_state + ((size_t)env * k.K + t) * 8);
      if (is_gripper(T.type)) {
        FrameAdj fa = frame_adj_zero();
        frame_sdf_adj(T, sdf_kind(T.type), jaw_frame(P, -1.f), x, g[col], fa, gx);
        jaw_frame_adj(P, -1.f, fa, gP);
        fa = frame_adj_zero();
        frame_sdf_adj(T, sdf_kind(T.type), jaw_frame(P, 1.f), x, g[col + 1], fa, gx);
        jaw_frame_adj(P, 1.f, fa, gP);
      } else {
        tool_sdf_adj(T, P, x, g[col], gP, gx);
      }
    }
    col += is_gripper(T.type) ? 2 : 1;
    float vals[8] = {gP.p.x, gP.p.y, gP.p.z, gP.q.w, gP.q.x, gP.q.y, gP.q.z, gP.gap};
#pragma unroll
    for (int q = 0; q < 8; q++) {
      float s = warp_sum(vals[q]);
      if ((threadIdx.x & 31) == 0 && s != 0.f && gid < k.stride) atomicAdd(tool_adj + ((size_t)env * k.K + t) * 8 + q, s);
    }
  }
  if (live) {
    adj_frame[(CX + 0) * k.stride + gid] += gx.x;
    adj_frame[(CX + 1) * k.stride + gid] += gx.y;
    adj_frame[(CX + 2) * k.stride + gid] += gx.z;
  }
}

// compute_grid_m_kernel (+grad), mpm_simulator.py:456-466.  dense [B,n,n,n] output (row-major x,y,z)
__global__ void k_grid_m(SimConst k, const float* __restrict__ frame, const int* __restrict__ npart,
                         float* __restrict__ out) {
  DSK_TL(k);
  int gid = blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= k.stride) return;
  int env = gid / k.Npad, p = gid - env * k.Npad;
  if (p >= npart[env]) return;
  float3 x = load_v3(frame, CX, k.stride, gid);
  Stencil s;
  make_stencil(k, x.x, x.y, x.z, s);
  float* o = out + (size_t)env * k.nnode;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      for (int l = 0; l < 3; l++)
        atomicAdd(&o[((size_t)(s.bx + i) * k.n + (s.by + j)) * k.n + (s.bz + l)],
                  s.wx[i] * s.wy[j] * s.wz[l] * k.p_mass);
}
__global__ void k_grid_m_adj(SimConst k, const float* __restrict__ frame, const int* __restrict__ npart,
                             const float* __restrict__ gm, float* __restrict__ adj_frame) {
  DSK_TL(k);
  int gid = blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= k.stride) return;
  int env = gid / k.Npad, p = gid - env * k.Npad;
  if (p >= npart[env]) return;
  float3 x = load_v3(frame, CX, k.stride, gid);
  Stencil s;
  make_stencil(k, x.x, x.y, x.z, s);
  const float* g = gm + (size_t)env * k.nnode;
  float gwx[3] = {0, 0, 0}, gwy[3] = {0, 0, 0}, gwz[3] = {0, 0, 0};
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      for (int l = 0; l < 3; l++) {
        float gw = g[((size_t)(s.bx + i) * k.n + (s.by + j)) * k.n + (s.bz + l)] * k.p_mass;
        gwx[i] += gw * s.wy[j] * s.wz[l];
        gwy[j] += gw * s.wx[i] * s.wz[l];
        gwz[l] += gw * s.wx[i] * s.wy[j];
      }
  float dw[3];
  float3 gf;
  bspline1_grad(s.fx, dw);
  gf.x = gwx[0] * dw[0] + gwx[1] * dw[1] + gwx[2] * dw[2];
  bspline1_grad(s.fy, dw);
  gf.y = gwy[0] * dw[0] + gwy[1] * dw[1] + gwy[2] * dw[2];
  bspline1_grad(s.fz, dw);
  gf.z = gwz[0] * dw[0] + gwz[1] * dw[1] + gwz[2] * dw[2];
  adj_frame[(CX + 0) * k.stride + gid] += k.inv_dx * gf.x;
  adj_frame[(CX + 1) * k.stride + gid] += k.inv_dx * gf.y;
  adj_frame[(CX + 2) * k.stride + gid] += k.inv_dx * gf.z;
}

// tile-major grid -> dense [n,n,n,*] for the debug getters
__global__ void k_grid_to_dense(SimConst k, const float4* __restrict__ G, int env, float* v3, float* m,
                                unsigned char* occ, const int* __restrict__ tile_epoch, int epoch) {
  DSK_TL(k);
  int node = blockIdx.x * blockDim.x + threadIdx.x;
  if (node >= k.nnode) return;
  int Z = node % k.n, Y = (node / k.n) % k.n, X = node / (k.n * k.n);
  int o = node_offset(X, Y, Z, k.nt);
  float4 g = G ? G[(size_t)env * k.nnode + o] : make_float4(0, 0, 0, 0);
  if (v3) {
    v3[(size_t)node * 3] = g.x;
    v3[(size_t)node * 3 + 1] = g.y;
    v3[(size_t)node * 3 + 2] = g.z;
  }
  if (m) m[node] = g.w;
  if (occ) occ[node] = tile_epoch[env * k.ntile + (o >> 6)] == epoch ? 1 : 0;
}
