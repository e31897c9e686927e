// TEST INFRASTRUCTURE: the warp-aggregated scatters of the particle kernels (kernels_common.cuh: warp_scatter27 of the
// one-thread-per-particle family, warp_scatter9 of the plane-split family) executed on an emulated warp (simt_shim.h) and
// compared by the tests with a plain per-particle accumulation, for key patterns that drive each of their regimes.
#define DSK_HOST_CHECK 1
#define DSK_HOST_SIMT 1
#include <cstring>
#include "../../diffskill_b200/csrc/kernels_common.cuh"

extern "C" {
// n_grid: grid nodes per axis (multiple of 4).  x[N][3]: particle positions (32 consecutive particles = one warp, in the
// given order); active[N]; a[N][4][4]: per particle a0, ax, ay, az (xyz + mass weight in .w of a0) so that the
// contribution to stencil node (i,j,l) is w_ijl * (a0 + i ax + j ay + l az) -- the affine form of every scatter in the
// engine.  mode 27: warp_scatter27; mode 9: warp_scatter9 called for the three planes; 127 / 227 / 109: the transposed
// shared-memory scatters of round 2 (warp_scatter27_ts, warp_scatter27_ts_affine, warp_scatter9_ts).
// Outputs: grid[n^3][4] tile-major order undone (dense (X*n+Y)*n+Z), tiles[ntile] = 1 if the tile was appended to the list.
void hc_scatter(int n_grid, float inv_dx, int N, const float* x, const int* active, const float* a, int mode, float* grid,
                int* tiles) {
  SimConst k;
  std::memset(&k, 0, sizeof k);
  k.n = n_grid;
  k.nt = n_grid / 4;
  k.ntile = k.nt * k.nt * k.nt;
  k.nnode = n_grid * n_grid * n_grid;
  k.B = 1;
  k.inv_dx = inv_dx;
  k.dx = 1.f / inv_dx;
  std::vector<float4> G(k.nnode, make_float4(0, 0, 0, 0));
  std::vector<int> epoch(k.ntile, 0), list(k.ntile, 0);
  int count = 0;
  TileTrack tt{epoch.data(), list.data(), &count};
  for (int w0 = 0; w0 < N; w0 += 32) {
    simt_run_warp([&](int lane) {
      int p = w0 + lane;
      bool act = p < N && active[p];
      int g = p < N ? p : w0;   // inactive lanes shadow a valid slot, as the kernels do
      Stencil s;
      make_stencil(k, x[g * 3], x[g * 3 + 1], x[g * 3 + 2], s);
      const float* c = a + (size_t)g * 16;
      auto contrib = [&](int i, int j, int l) {
        float w = s.wx[i] * s.wy[j] * s.wz[l];
        return make_float4(w * (c[0] + i * c[4] + j * c[8] + l * c[12]), w * (c[1] + i * c[5] + j * c[9] + l * c[13]),
                           w * (c[2] + i * c[6] + j * c[10] + l * c[14]), w * c[3]);
      };
      static std::vector<float4> wbuf(TS_WARP_FLOAT4);   // the per-warp shared-memory tile of the transposed scatters
      if (mode == 27) {
        warp_scatter27(k, act, s, G.data(), tt, true, 0, 1, contrib);
      } else if (mode == 127) {   // transposed shared-memory scatter, values through the lambda
        warp_scatter27_ts(k, act, s, G.data(), tt, true, 0, 1, wbuf.data(), contrib);
      } else if (mode == 227) {   // ... affine coefficients, values built incrementally (packed pairs)
        warp_scatter27_ts_affine(k, act, s, G.data(), tt, true, 0, 1, wbuf.data(), make_float4(c[0], c[1], c[2], c[3]),
                                 make_float3(c[4], c[5], c[6]), make_float3(c[8], c[9], c[10]), make_float3(c[12], c[13], c[14]));
      } else if (mode == 109) {   // transposed scatter of the plane-split kernels
        for (int pl = 0; pl < 3; pl++)
          warp_scatter9_ts(k, act, s, pl, G.data(), tt, pl == 0, 0, 1, wbuf.data(), [&](int j, int l) { return contrib(pl, j, l); });
      } else {
        for (int pl = 0; pl < 3; pl++)
          warp_scatter9(k, act, s, pl, s.ox[pl], G.data(), tt, pl == 0, 0, 1, [&](int j, int l) { return contrib(pl, j, l); });
      }
    });
  }
  for (int X = 0; X < n_grid; X++)
    for (int Y = 0; Y < n_grid; Y++)
      for (int Z = 0; Z < n_grid; Z++) {
        float4 v = G[node_offset(X, Y, Z, k.nt)];
        float* o = grid + (((size_t)X * n_grid + Y) * n_grid + Z) * 4;
        o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
      }
  for (int t = 0; t < k.ntile; t++) tiles[t] = 0;
  for (int i = 0; i < count; i++) tiles[list[i]] += 1;
}
}
