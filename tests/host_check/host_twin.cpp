// TEST INFRASTRUCTURE: a CPU twin of ONE adjoint substep of the CUDA engine (substep_grad, plain kernel family), built from
// the product's own device functions (diffskill_b200/csrc/particle_math.cuh + tools.cuh compiled by g++): g2p.grad ->
// grid_op.grad (boundary, contact chain, tool-pose adjoints) -> p2g.grad + svd_grad + compute_F_tmp.grad.  What it does
// not contain is the launch geometry: the warp-aggregated scatters are plain loops over a dense tile-major grid, the
// per-tile culling and the shared-memory staging are absent.  Tests feed it the oracle's state of a substep and compare
// the adjoints with the oracle's substep_grad, in regimes the small GPU parity cases do not reach.
#define DSK_HOST_CHECK 1
#include <cmath>
#include <cstring>
#include <vector>

#include "../../diffskill_b200/csrc/particle_math.cuh"

static int node_offset_h(int X, int Y, int Z, int nt) {   // kernels_common.cuh: tile-major node index
  int tile = ((X >> 2) * nt + (Y >> 2)) * nt + (Z >> 2);
  return (tile << 6) | ((X & 3) << 4) | ((Y & 3) << 2) | (Z & 3);
}
static ToolParams make_tool_h(const dsk_tool_desc& d) {   // fill_tool() of engine.cu
  ToolParams T;
  std::memset(&T, 0, sizeof T);
  T.type = d.type;
  T.action_dim = d.action_dim;
  for (int j = 0; j < 8; j++) T.action_scale[j] = (float)d.action_scale[j];
  T.friction = (float)d.friction;
  T.softness = (float)d.softness;
  for (int j = 0; j < 3; j++) {
    T.lo[j] = (float)d.lower_bound[j];
    T.hi[j] = (float)d.upper_bound[j];
    T.size[j] = (float)d.size[j];
  }
  T.h = (float)d.h;
  T.half_h = (float)(d.h / 2);
  T.r = (float)d.r;
  T.prism_h0 = (float)d.prism_h[0];
  T.prism_h1 = (float)d.prism_h[1];
  float w = (float)d.prot[0], x = (float)d.prot[1], y = (float)d.prot[2], z = (float)d.prot[3];
  T.prot = Q4{w, x, y, z};
  float n2 = w * w;
  n2 = n2 + x * x;
  n2 = n2 + y * y;
  n2 = n2 + z * z;
  float inv = 1.0f / std::sqrt(n2);
  T.prot_inv = Q4{inv * w, inv * -x, inv * -y, inv * -z};
  T.min_gap = (float)d.minimal_gap;
  T.max_gap = (float)d.maximal_gap;
  // radius of a sphere around the contact-frame origin that contains the shape (tile culling), as fill_tool()
  if (d.type == DSK_TOOL_CAPSULE || d.type == DSK_TOOL_ROLLINGPIN_EXT || d.type == DSK_TOOL_ROLLINGPIN ||
      d.type == DSK_TOOL_GRIPPER2)
    T.bound_r = (float)(d.h / 2 + d.r);
  else if (d.type == DSK_TOOL_SPHERE)
    T.bound_r = (float)d.r;
  else if (d.type == DSK_TOOL_CYLINDER)
    T.bound_r = (float)std::sqrt(d.h * d.h + d.r * d.r);
  else if (d.type == DSK_TOOL_TORUS)
    T.bound_r = (float)(d.h + d.r);
  else if (d.type == DSK_TOOL_CHOPSTICKS) {
    T.max_gap = 1e30f;
    double g = std::max(d.maximal_gap > 0 && d.maximal_gap < 1e3 ? d.maximal_gap : 0.0, 1.0);
    T.bound_r = (float)std::sqrt((g / 2 + d.r) * (g / 2 + d.r) + (d.h + d.r) * (d.h + d.r));
  } else
    T.bound_r = (float)std::sqrt(d.size[0] * d.size[0] + d.size[1] * d.size[1] + d.size[2] * d.size[2]);
  T.bound_r *= 1.001f;
  return T;
}

static SimConst make_const_h(const dsk_config* cfg, int N) {   // dsk_create() of engine.cu
  SimConst k;
  std::memset(&k, 0, sizeof k);
  k.n = cfg->n_grid;
  k.nt = k.n / 4;
  k.ntile = k.nt * k.nt * k.nt;
  k.nnode = k.n * k.n * k.n;
  k.B = 1;
  k.Npad = k.stride = N;
  k.S = 1;
  k.K = cfg->n_tools;
  k.gf_mode = cfg->ground_friction == 0.0 ? 0 : (cfg->ground_friction < 10.0 ? 1 : 2);
  k.dt = (float)cfg->dt;
  k.dx = (float)cfg->dx;
  k.inv_dx = (float)cfg->inv_dx;
  k.p_mass = (float)cfg->p_mass;
  k.c_stress = (float)(-cfg->dt * cfg->p_vol * 4 * cfg->inv_dx * cfg->inv_dx);
  k.c_C = (float)(4 * cfg->inv_dx);
  k.x_hi = (float)(1. - 3 * cfg->dx);
  k.x_lo = (float)(cfg->lower_bound * cfg->dx);
  k.m_eps = 1e-12f;
  k.ground_friction = (float)cfg->ground_friction;
  for (int d = 0; d < 3; d++) {
    volatile float t = k.dt * (float)cfg->gravity[d];
    k.grav[d] = t * 30.f;
  }
  return k;
}
// contact frames of all tools in table order, with the per-tile culling decision of the grid kernels (prepare_frame)
struct FrH {
  int t;
  float flag;
  Frame F0, F1;
  int active;
  ContactGeom c;
  float3 vs;
};
static std::vector<FrH> frames_h(const SimConst& k, const std::vector<ToolParams>& T, const float* poses, int X, int Y, int Z,
                                 bool cull) {
  GridTools gt;
  std::memset(&gt, 0, sizeof gt);
  for (int t = 0; t < k.K; t++) gt.T[t] = T[t];
  build_frame_table(k, gt.T, gt.ft);
  std::vector<FrH> fr(gt.ft.n);
  for (int y = 0; y < gt.ft.n; y++) {
    fr[y].t = gt.ft.tool[y];
    fr[y].flag = gt.ft.flag[y];
    // poses[2][K][8] is the [S+1 = 2][K][8] pose table of env 0, substep j = 0
    fr[y].active = prepare_frame(k, gt.T, gt.ft, y, poses, 0, 0, X >> 2, Y >> 2, Z >> 2, fr[y].F0, fr[y].F1);
    if (!cull) fr[y].active = 1;
  }
  return fr;
}

extern "C" {
// cfg: the dsk_config the engine would be created with (scene constants, tools).  N particles.
// particles (row-major, as the C ABI hands them out): x[N][3] v[N][3] C[N][9] F[N][9] at frame f; xn[N][3] at f+1;
// mat[3][N] = mu, lam, yield_stress.  Dense grids indexed (X*n+Y)*n+Z: g0[n^3][4] = (grid_v_in, grid_m) and
// gv[n^3][3] = grid_v_out of substep f.  poses[2][K][8] at f and f+1.  Adjoints at f+1: gxn gvn gCn gFn.
// Outputs: adjoints at f (gx gv gC gF), grid adjoints ga[n^3][4] = g(grid_v_in), g(grid_m), pose_adj[2][K][8].
void hc_substep_grad(const dsk_config* cfg, int N, const float* x, const float* v, const float* Cm, const float* F,
                     const float* xn, const float* mat, const float* g0, const float* gvout, const float* poses,
                     const float* gxn, const float* gvn, const float* gCn, const float* gFn, float* gx, float* gv,
                     float* gC, float* gF, float* ga, float* pose_adj) {
  SimConst k = make_const_h(cfg, N);
  const int K = k.K, n = k.n;
  std::vector<ToolParams> T(K > 0 ? K : 1);
  for (int t = 0; t < K; t++) T[t] = make_tool_h(cfg->tools[t]);
  // SoA particle frames and adjoint frames, tile-major grids
  std::vector<float> fin((size_t)FRAME_COMPS * N), fnext((size_t)FRAME_COMPS * N, 0.f), ain((size_t)FRAME_COMPS * N),
      aout((size_t)FRAME_COMPS * N, 0.f);
  for (int p = 0; p < N; p++) {
    for (int d = 0; d < 3; d++) {
      fin[(CX + d) * N + p] = x[p * 3 + d];
      fin[(CV + d) * N + p] = v[p * 3 + d];
      fnext[(CX + d) * N + p] = xn[p * 3 + d];
      ain[(CX + d) * N + p] = gxn[p * 3 + d];
      ain[(CV + d) * N + p] = gvn[p * 3 + d];
    }
    for (int d = 0; d < 9; d++) {
      fin[(CC + d) * N + p] = Cm[p * 9 + d];
      fin[(CF + d) * N + p] = F[p * 9 + d];
      ain[(CC + d) * N + p] = gCn[p * 9 + d];
      ain[(CF + d) * N + p] = gFn[p * 9 + d];
    }
  }
  std::vector<float4> G0(k.nnode), Gv(k.nnode), Ga(k.nnode, make_float4(0, 0, 0, 0));
  for (int X = 0; X < n; X++)
    for (int Y = 0; Y < n; Y++)
      for (int Z = 0; Z < n; Z++) {
        size_t g = ((size_t)X * n + Y) * n + Z;
        int o = node_offset_h(X, Y, Z, k.nt);
        G0[o] = make_float4(g0[g * 4], g0[g * 4 + 1], g0[g * 4 + 2], g0[g * 4 + 3]);
        Gv[o] = make_float4(gvout[g * 3], gvout[g * 3 + 1], gvout[g * 3 + 2], 0.f);
      }
  // ---- k_g2p_adj ---------------------------------------------------------------------------------------------------
  for (int p = 0; p < N; p++) {
    float3 xx = load_v3(fin.data(), CX, N, p), xxn = load_v3(fnext.data(), CX, N, p);
    float3 a_gx = load_v3(ain.data(), CX, N, p), a_gv = load_v3(ain.data(), CV, N, p);
    M3 a_gC = load_m3(ain.data(), CC, N, p);
    Stencil s;
    G2PAdj c;
    g2p_adj_begin(k, xx, xxn, a_gx, a_gv, a_gC, s, c);
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++)
        for (int l = 0; l < 3; l++) {
          float4 val = g2p_adj_node(s, c, i, j, l);
          float4& dst = Ga[s.ox[i] + s.oy[j] + s.oz[l]];
          dst.x += val.x; dst.y += val.y; dst.z += val.z; dst.w += val.w;
        }
    store_v3(aout.data(), CX, N, p, g2p_adj_finish(k, s, Gv.data(), a_gC, c));
  }
  // ---- k_grid_adj: per live node, velocity chain forward and backward, tool-pose adjoints -------------------------------
  std::vector<double> padj((size_t)2 * (K > 0 ? K : 1) * 8, 0.0);
  for (int X = 0; X < n; X++)
    for (int Y = 0; Y < n; Y++)
      for (int Z = 0; Z < n; Z++) {
        int o = node_offset_h(X, Y, Z, k.nt);
        float4 gin = G0[o];
        float4 ga4 = Ga[o];
        if (!(gin.w > k.m_eps)) {
          Ga[o] = make_float4(0, 0, 0, 0);
          continue;
        }
        float3 gp = f3(mul_rn((float)X, k.dx), mul_rn((float)Y, k.dx), mul_rn((float)Z, k.dx));
        float inv = 1.f / gin.w;
        float3 vv = f3(inv * gin.x + k.grav[0], inv * gin.y + k.grav[1], inv * gin.z + k.grav[2]);
        std::vector<FrH> fr = frames_h(k, T, poses, X, Y, Z, true);
        for (auto& f : fr) {
          f.vs = vv;
          f.c.influence = -1.f;
          if (f.active) contact_geometry(T[f.t], sdf_kind(T[f.t].type), f.F0, f.F1, gp, k.dt, f.c);
          if (f.c.influence >= 0.f) vv = contact_response(vv, f.c.D, f.c.cv, f.c.influence, T[f.t].friction, f.flag != 0.f);
        }
        float3 g = grid_boundary_adj(k, X, Y, Z, vv, f3(ga4.x, ga4.y, ga4.z));
        for (int q = (int)fr.size() - 1; q >= 0; q--) {
          FrH& f = fr[q];
          if (!(f.c.influence >= 0.f)) continue;
          const ToolParams& Tt = T[f.t];
          float ginfl;
          float3 gD, gcv;
          g = contact_response_adj(f.vs, f.c.D, f.c.cv, f.c.influence, Tt.friction, f.flag != 0.f, g, gD, gcv, ginfl);
          float gdist = (f.c.influence < 1.f) ? (-Tt.softness * f.c.influence * ginfl) : 0.f;
          // frame_pose_adjoint of kernels_bwd.cuh, per node
          int kind = sdf_kind(Tt.type);
          FrameAdj a0 = frame_adj_zero(), a1 = frame_adj_zero();
          float3 Nl = (1.f / f.c.L) * f.c.nraw, gNl = f3(0, 0, 0);
          qrot_adj(f.F0.q, Nl, gD, a0.q, gNl);
          float3 gpl = contact_local_adj(Tt, kind, f.F0.aux, f.c.pl, f.c.nraw, f.c.L, gNl, gdist, a0.aux);
          float3 gnp = (1.f / k.dt) * gcv;
          a1.o += gnp;
          qrot_adj(f.F1.q, f.c.pl, gnp, a1.q, gpl);
          float3 unused = f3(0, 0, 0);
          inv_trans_adj(f.F0, gp, gpl, a0, unused);
          PoseAdj g0p = pose_adj_zero(), g1p = pose_adj_zero();
          if (f.flag != 0.f) {
            jaw_frame_adj(load_pose(poses + (size_t)f.t * 8), f.flag, a0, g0p);
            jaw_frame_adj(load_pose(poses + (size_t)(K + f.t) * 8), f.flag, a1, g1p);
          } else {
            tool_frame_adj(a0, g0p);
            tool_frame_adj(a1, g1p);
          }
          float v0[8] = {g0p.p.x, g0p.p.y, g0p.p.z, g0p.q.w, g0p.q.x, g0p.q.y, g0p.q.z, g0p.gap};
          float v1[8] = {g1p.p.x, g1p.p.y, g1p.p.z, g1p.q.w, g1p.q.x, g1p.q.y, g1p.q.z, g1p.gap};
          for (int d = 0; d < 8; d++) {
            padj[(size_t)f.t * 8 + d] += v0[d];
            padj[(size_t)(K + f.t) * 8 + d] += v1[d];
          }
        }
        Ga[o] = make_float4(inv * g.x, inv * g.y, inv * g.z, -(inv * inv) * (gin.x * g.x + gin.y * g.y + gin.z * g.z));
      }
  for (size_t q = 0; q < (size_t)2 * K * 8; q++) pose_adj[q] = (float)padj[q];
  // ---- k_p2g_adj ---------------------------------------------------------------------------------------------------
  for (int p = 0; p < N; p++) p2g_adj_particle(k, p, 0, fin.data(), ain.data(), aout.data(), mat, Ga.data(), nullptr);
  for (int p = 0; p < N; p++) {
    for (int d = 0; d < 3; d++) {
      gx[p * 3 + d] = aout[(CX + d) * N + p];
      gv[p * 3 + d] = aout[(CV + d) * N + p];
    }
    for (int d = 0; d < 9; d++) {
      gC[p * 9 + d] = aout[(CC + d) * N + p];
      gF[p * 9 + d] = aout[(CF + d) * N + p];
    }
  }
  for (int X = 0; X < n; X++)
    for (int Y = 0; Y < n; Y++)
      for (int Z = 0; Z < n; Z++) {
        size_t g = ((size_t)X * n + Y) * n + Z;
        float4 a = Ga[node_offset_h(X, Y, Z, k.nt)];
        ga[g * 4] = a.x; ga[g * 4 + 1] = a.y; ga[g * 4 + 2] = a.z; ga[g * 4 + 3] = a.w;
      }
}
// One forward substep (k_p2g -> k_grid -> k_g2p, plain family) for N particles on a dense grid.  cull != 0: frames are
// culled per tile as in the kernels.  Outputs frame f+1 (xn vn Cn Fn), g0[n^3][4] = (grid_v_in, grid_m), gv[n^3][3].
void hc_substep(const dsk_config* cfg, int N, const float* x, const float* v, const float* Cm, const float* F,
                const float* mat, const float* poses, int cull, float* xn, float* vn, float* Cn, float* Fn, float* g0,
                float* gvout) {
  SimConst k = make_const_h(cfg, N);
  const int K = k.K, n = k.n;
  std::vector<ToolParams> T(K > 0 ? K : 1);
  for (int t = 0; t < K; t++) T[t] = make_tool_h(cfg->tools[t]);
  std::vector<float4> G(k.nnode, make_float4(0, 0, 0, 0));
  std::vector<Stencil> st(N);
  for (int p = 0; p < N; p++) {
    M3 C_, F_;
    for (int d = 0; d < 9; d++) {
      C_.m[d] = Cm[p * 9 + d];
      F_.m[d] = F[p * 9 + d];
    }
    float3 xx = f3(x[p * 3], x[p * 3 + 1], x[p * 3 + 2]), vv = f3(v[p * 3], v[p * 3 + 1], v[p * 3 + 2]);
    P2GParticle o;
    p2g_particle(k, C_, F_, mat[p], mat[N + p], mat[2 * N + p], o);
    for (int d = 0; d < 9; d++) Fn[p * 9 + d] = o.newF.m[d];
    Stencil& s = st[p];
    make_stencil(k, xx.x, xx.y, xx.z, s);
    // scatter values of k_p2g: w * (p_mass v + affine (offset - fx) dx, p_mass)
    float3 fxv = f3(s.fx, s.fy, s.fz);
    float3 a0 = k.p_mass * vv - k.dx * mv(o.affine, fxv);
    float3 ax = f3(k.dx * o.affine.m[0], k.dx * o.affine.m[3], k.dx * o.affine.m[6]);
    float3 ay = f3(k.dx * o.affine.m[1], k.dx * o.affine.m[4], k.dx * o.affine.m[7]);
    float3 az = f3(k.dx * o.affine.m[2], k.dx * o.affine.m[5], k.dx * o.affine.m[8]);
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++)
        for (int l = 0; l < 3; l++) {
          float w = s.wx[i] * s.wy[j] * s.wz[l];
          float3 a = a0 + (float)i * ax + (float)j * ay + (float)l * az;
          float4& dst = G[s.ox[i] + s.oy[j] + s.oz[l]];
          dst.x += w * a.x; dst.y += w * a.y; dst.z += w * a.z; dst.w += w * k.p_mass;
        }
  }
  std::vector<float4> Gv(k.nnode, make_float4(0, 0, 0, 0));
  for (int X = 0; X < n; X++)
    for (int Y = 0; Y < n; Y++)
      for (int Z = 0; Z < n; Z++) {
        int o = node_offset_h(X, Y, Z, k.nt);
        size_t g = ((size_t)X * n + Y) * n + Z;
        float4 gin = G[o];
        g0[g * 4] = gin.x; g0[g * 4 + 1] = gin.y; g0[g * 4 + 2] = gin.z; g0[g * 4 + 3] = gin.w;
        gvout[g * 3] = gvout[g * 3 + 1] = gvout[g * 3 + 2] = 0.f;
        if (!(gin.w > k.m_eps)) continue;
        float3 gp = f3(mul_rn((float)X, k.dx), mul_rn((float)Y, k.dx), mul_rn((float)Z, k.dx));
        float inv = 1.f / gin.w;
        float3 vv = f3(inv * gin.x + k.grav[0], inv * gin.y + k.grav[1], inv * gin.z + k.grav[2]);
        std::vector<FrH> fr = frames_h(k, T, poses, X, Y, Z, cull != 0);
        for (auto& f : fr) {
          if (!f.active) continue;
          contact_geometry(T[f.t], sdf_kind(T[f.t].type), f.F0, f.F1, gp, k.dt, f.c);
          if (f.c.influence >= 0.f) vv = contact_response(vv, f.c.D, f.c.cv, f.c.influence, T[f.t].friction, f.flag != 0.f);
        }
        vv = grid_boundary(k, X, Y, Z, vv);
        Gv[o] = make_float4(vv.x, vv.y, vv.z, gin.w);
        gvout[g * 3] = vv.x; gvout[g * 3 + 1] = vv.y; gvout[g * 3 + 2] = vv.z;
      }
  for (int p = 0; p < N; p++) {
    float3 xx = f3(x[p * 3], x[p * 3 + 1], x[p * 3 + 2]), nx, nv;
    M3 nC;
    g2p_particle(k, st[p], Gv.data(), xx, nx, nv, nC);
    xn[p * 3] = nx.x; xn[p * 3 + 1] = nx.y; xn[p * 3 + 2] = nx.z;
    vn[p * 3] = nv.x; vn[p * 3 + 1] = nv.y; vn[p * 3 + 2] = nv.z;
    for (int d = 0; d < 9; d++) Cn[p * 9 + d] = nC.m[d];
  }
}
}
extern "C" {
// debug: for every grid node and frame, does the tile culling drop a frame whose contact is active at the node?
// out[0] = number of (node, frame) pairs culled although active, out[1] = active pairs, out[2..] details of the first
int hc_culling_check(const dsk_config* cfg, const float* poses, float* out) {
  SimConst k = make_const_h(cfg, 1);
  const int K = k.K, n = k.n;
  std::vector<ToolParams> T(K > 0 ? K : 1);
  for (int t = 0; t < K; t++) T[t] = make_tool_h(cfg->tools[t]);
  int bad = 0, act = 0;
  for (int X = 0; X < n; X++)
    for (int Y = 0; Y < n; Y++)
      for (int Z = 0; Z < n; Z++) {
        float3 gp = f3(mul_rn((float)X, k.dx), mul_rn((float)Y, k.dx), mul_rn((float)Z, k.dx));
        std::vector<FrH> fr = frames_h(k, T, poses, X, Y, Z, true);
        for (size_t y = 0; y < fr.size(); y++) {
          ContactGeom c;
          contact_geometry(T[fr[y].t], sdf_kind(T[fr[y].t].type), fr[y].F0, fr[y].F1, gp, k.dt, c);
          if (c.influence >= 0.f) {
            act++;
            if (!fr[y].active) {
              if (bad == 0) {
                out[2] = X; out[3] = Y; out[4] = Z; out[5] = (float)y; out[6] = c.influence;
                out[7] = frame_sdf(T[fr[y].t], sdf_kind(T[fr[y].t].type), fr[y].F0, gp);
                out[8] = T[fr[y].t].bound_r;
              }
              bad++;
            }
          }
        }
      }
  out[0] = (float)bad;
  out[1] = (float)act;
  return bad;
}
}
