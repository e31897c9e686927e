// Per-env-step kernels around the substep loop: counting sort by cell, tool kinematics
// (forward_kinematics + tool-tool collision projection), state IO, observation helpers.
#pragma once
#include "kernels_common.cuh"

// ---- counting sort by cell (key = tile-major linear cell index of the stencil base) -------------------
DSK_DEV int cell_key(const SimConst& k, float x, float y, float z, int& bx, int& by, int& bz) {
  float fx, w[3];
  bspline1(x, k.inv_dx, k.n, bx, fx, w);
  bspline1(y, k.inv_dx, k.n, by, fx, w);
  bspline1(z, k.inv_dx, k.n, bz, fx, w);
  return node_offset(bx, by, bz, k.nt);
}
// The cell histogram is dense (B * n^3 counters: 67 MB for 64 envs on 64^3) but the dough occupies a few per cent of it, and
// the keys are tile-major, so the occupied cells cluster in a few SCAN_CHUNK-cell chunks (16 tiles).  k_sort_bin flags the
// chunks it touches; the two scan passes and the clean-up only work on flagged chunks (r02p timeline, GatherMove x64: the
// dense scan cost 15 + 64 us per env step and the 67 MB memset before it ~12 us).  Invariant between env steps: all counters
// and all flags are zero (k_sort_clear restores it after the scatter; the arrays are zero-initialised).
#define SCAN_CHUNK 1024
#define SCAN_CTA 256
__global__ void k_sort_bin(SimConst k, const StepArgs* __restrict__ args, const int* __restrict__ npart,
                           int* __restrict__ cell_count, int* __restrict__ key, int* __restrict__ rank,
                           int* __restrict__ chunk_flag /*[B][chunks]*/) {
  DSK_TL(k);
  int gid = blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= k.stride) return;
  const float* __restrict__ ck = args->ck_src;
  int env = gid / k.Npad, p = gid - env * k.Npad;
  if (p >= npart[env]) return;
  int bx, by, bz;
  int kk = cell_key(k, ck[gid], ck[k.stride + gid], ck[2 * k.stride + gid], bx, by, bz);
  key[gid] = kk;
  rank[gid] = atomicAdd(&cell_count[(size_t)env * k.nnode + kk], 1);
  chunk_flag[env * ((k.nnode + SCAN_CHUNK - 1) / SCAN_CHUNK) + kk / SCAN_CHUNK] = 1;   // idempotent plain store
}
// exclusive scan of cell_count per env, in place, in two multi-CTA passes (a single CTA per env is bound by one
// SM's bandwidth on the n^3-cell histogram), both skipping the chunks no particle fell into:
//   k_scan_partial: sum of each flagged SCAN_CHUNK-cell chunk (0 for the others)
//   k_scan_chunks : offset = sum of the preceding chunk sums of the env, then an in-place chunk scan
//   k_sort_clear  : after the scatter: zeroes the flagged chunks and their flags
DSK_DEV int block_sum(int v, int* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  int t = 0;
  for (int w = 0; w < SCAN_CTA / 32; w++) t += sh[w];
  __syncthreads();
  return t;
}
// A CTA takes SCAN_GROUP consecutive (env, chunk) pairs (linear index env * chunks + chunk), collects the flagged ones in
// shared memory -- one flag load per thread, one round trip -- and works through them; launching one CTA per pair costs more
// in CTA launches (16 384 pairs, ~90 % of them empty) than the whole scan.
#define SCAN_GROUP 64
DSK_DEV int flagged_chunks(const int* __restrict__ chunk_flag, int total, int* lst, int* n, int* partial_zero) {
  if (threadIdx.x == 0) *n = 0;
  __syncthreads();
  const int idx = blockIdx.x * SCAN_GROUP + threadIdx.x;
  if (threadIdx.x < SCAN_GROUP && idx < total) {
    if (chunk_flag[idx]) lst[atomicAdd(n, 1)] = idx;
    else if (partial_zero) partial_zero[idx] = 0;
  }
  __syncthreads();
  return *n;
}
// grid: cdiv(B * chunks, SCAN_GROUP)
__global__ void __launch_bounds__(SCAN_CTA) k_scan_partial(SimConst k, const int* __restrict__ cell_count, int* __restrict__ partial,
                                                           const int* __restrict__ chunk_flag, int chunks) {
  DSK_TL(k);
  __shared__ int sh[SCAN_CTA / 32], lst[SCAN_GROUP], nl;
  const int n = flagged_chunks(chunk_flag, k.B * chunks, lst, &nl, partial);
  for (int q = 0; q < n; q++) {
    const int idx = lst[q], env = idx / chunks, chunk = idx - env * chunks;
    const int* c = cell_count + (size_t)env * k.nnode;
    int lo = chunk * SCAN_CHUNK, hi = min(lo + SCAN_CHUNK, k.nnode);
    int v = 0;
    for (int i = lo + threadIdx.x; i < hi; i += SCAN_CTA) v += c[i];
    int t = block_sum(v, sh);
    if (threadIdx.x == 0) partial[idx] = t;
  }
}
__global__ void __launch_bounds__(SCAN_CTA) k_scan_chunks(SimConst k, int* __restrict__ cell_count, const int* __restrict__ partial,
                                                          const int* __restrict__ chunk_flag, int chunks) {
  DSK_TL(k);
  __shared__ int sh[SCAN_CTA / 32], wtot[SCAN_CTA / 32], lst[SCAN_GROUP], nl;
  const int n = flagged_chunks(chunk_flag, k.B * chunks, lst, &nl, nullptr);   // empty chunks: nobody reads their offsets
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int q = 0; q < n; q++) {
    const int idx = lst[q], env = idx / chunks, chunk = idx - env * chunks;
    int* c = cell_count + (size_t)env * k.nnode;
    int v = 0;
    for (int i = threadIdx.x; i < chunk; i += SCAN_CTA) v += partial[env * chunks + i];
    int carry = block_sum(v, sh);
    int lo = chunk * SCAN_CHUNK, hi = min(lo + SCAN_CHUNK, k.nnode);
    for (int i0 = lo; i0 < hi; i0 += SCAN_CTA) {
      int i = i0 + threadIdx.x;
      int x = i < hi ? c[i] : 0, sc = x;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, sc, o);
        if (lane >= o) sc += t;
      }
      if (lane == 31) wtot[warp] = sc;
      __syncthreads();
      int woff = 0, tot = 0;
      for (int w = 0; w < SCAN_CTA / 32; w++) {
        if (w < warp) woff += wtot[w];
        tot += wtot[w];
      }
      if (i < hi) c[i] = carry + woff + sc - x;
      carry += tot;
      __syncthreads();
    }
  }
}
__global__ void __launch_bounds__(SCAN_CTA) k_sort_clear(SimConst k, int* __restrict__ cell_count, int* __restrict__ chunk_flag, int chunks) {
  DSK_TL(k);
  __shared__ int lst[SCAN_GROUP], nl;
  const int n = flagged_chunks(chunk_flag, k.B * chunks, lst, &nl, nullptr);
  for (int q = 0; q < n; q++) {
    const int idx = lst[q], env = idx / chunks, chunk = idx - env * chunks;
    int* c = cell_count + (size_t)env * k.nnode;
    int lo = chunk * SCAN_CHUNK, hi = min(lo + SCAN_CHUNK, k.nnode);
    for (int i = lo + threadIdx.x; i < hi; i += SCAN_CTA) c[i] = 0;
    if (threadIdx.x == 0) chunk_flag[idx] = 0;   // every flag of the group was read before the barrier in flagged_chunks
  }
}
// scatter checkpoint (canonical order) -> work frame 0 (sorted order); also permutes the material arrays
__global__ void k_sort_scatter(SimConst k, const StepArgs* __restrict__ args, const float* __restrict__ mat,
                               const int* __restrict__ npart, const int* __restrict__ cell_start,
                               const int* __restrict__ key, const int* __restrict__ rank, int use_sort,
                               float* __restrict__ w0, float* __restrict__ mat_sorted, int* __restrict__ perm) {
  DSK_TL(k);
  int gid = blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= k.stride) return;
  const float* __restrict__ ck = args->ck_src;
  int env = gid / k.Npad, p = gid - env * k.Npad;
  if (p >= npart[env]) return;
  int dst = use_sort ? env * k.Npad + cell_start[(size_t)env * k.nnode + key[gid]] + rank[gid] : gid;
  perm[dst] = p;
#pragma unroll
  for (int c = 0; c < FRAME_COMPS; c++) w0[c * k.stride + dst] = ck[c * k.stride + gid];
#pragma unroll
  for (int c = 0; c < 3; c++) mat_sorted[c * k.stride + dst] = mat[c * k.stride + gid];
}
// re-use of an earlier sort: gather the checkpoint into sorted order with a stored permutation (particles move a small
// fraction of a cell per env step, so the cell order stays nearly sorted for a few steps)
__global__ void k_apply_perm(SimConst k, const StepArgs* __restrict__ args, const float* __restrict__ mat,
                             const int* __restrict__ npart, const int* __restrict__ perm_cache,
                             float* __restrict__ w0, float* __restrict__ mat_sorted, int* __restrict__ perm) {
  DSK_TL(k);
  int gid = blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= k.stride) return;
  int env = gid / k.Npad, d = gid - env * k.Npad;
  if (d >= npart[env]) return;
  const float* __restrict__ ck = args->ck_src;
  int p = perm_cache[gid];
  perm[gid] = p;
  int src = env * k.Npad + p;
#pragma unroll
  for (int c = 0; c < FRAME_COMPS; c++) w0[c * k.stride + gid] = ck[c * k.stride + src];
#pragma unroll
  for (int c = 0; c < 3; c++) mat_sorted[c * k.stride + gid] = mat[c * k.stride + src];
}
// sorted frame -> canonical checkpoint.  accumulate=1: += (adjoint checkpoints)
__global__ void k_unsort(SimConst k, const float* __restrict__ w, const int* __restrict__ npart,
                         const int* __restrict__ perm, float* const* __restrict__ pck, int accumulate) {
  DSK_TL(k);
  int gid = blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= k.stride) return;
  float* __restrict__ ck = *pck;
  int env = gid / k.Npad, d = gid - env * k.Npad;
  if (d >= npart[env]) return;
  int dst = env * k.Npad + perm[gid];
  if (accumulate) {
#pragma unroll
    for (int c = 0; c < FRAME_COMPS; c++) ck[c * k.stride + dst] += w[c * k.stride + gid];
  } else {
#pragma unroll
    for (int c = 0; c < FRAME_COMPS; c++) ck[c * k.stride + dst] = w[c * k.stride + gid];
  }
}
// canonical (adjoint) checkpoint -> sorted frame
__global__ void k_gather_sorted(SimConst k, float* const* __restrict__ pck, const int* __restrict__ npart,
                                const int* __restrict__ perm, float* __restrict__ w) {
  DSK_TL(k);
  int gid = blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= k.stride) return;
  const float* __restrict__ ck = *pck;
  int env = gid / k.Npad, d = gid - env * k.Npad;
  if (d >= npart[env]) return;
  int src = env * k.Npad + perm[gid];
#pragma unroll
  for (int c = 0; c < FRAME_COMPS; c++) w[c * k.stride + gid] = ck[c * k.stride + src];
}

// ---- permutation of whole frames through shared memory (batched engines) ---------------------------------------------------
// The four kernels above move 24 rows between canonical and sorted order with one RANDOM 4-byte access per element on one
// side: every such access moves a 32-byte sector (8x the useful bytes), and the read-modify-write of the adjoint un-sort is
// worse (r02x timeline, GatherMove x64: 55 us per adjoint reorder, 32 us for the sort's scatter -- 24 MB of useful traffic).
// Here a CTA takes `rows` rows of ONE env, brings the randomly-addressed side through shared memory and touches global
// memory with coalesced accesses only.  grid (nrows / ROWS, B), dynamic shared memory ROWS * Npad floats.
//   MODE 0 gather          out[i]        = in[perm[i]]     (canonical -> sorted: sort, cached sort, adjoint checkpoint)
//   MODE 1 scatter         out[perm[i]]  = in[i]           (sorted -> canonical checkpoint)
//   MODE 2 scatter-add     out[perm[i]] += in[i]           (sorted adjoint -> canonical adjoint checkpoint)
// pin / pout: when non-null the array is *pin / *pout (device-resident step arguments, see StepArgs).
#define PERM_CTA 256
template <int MODE, int ROWS>
__global__ void __launch_bounds__(PERM_CTA)
    k_permute_rows(SimConst k, const float* in, const float* const* pin, float* out, float* const* pout,
                   const int* __restrict__ npart, const int* __restrict__ perm) {
  DSK_TL(k);
  DSK_DYN_SMEM(float, sm);
  const float* __restrict__ src = pin ? *pin : in;
  float* __restrict__ dst = pout ? *pout : out;
  const int env = blockIdx.y, r0 = blockIdx.x * ROWS, n = npart[env], base = env * k.Npad;
  // ROWS independent accesses per loop iteration (unrolled): the loads of an iteration are all in flight together
  if (MODE == 0) {
    for (int p = threadIdx.x; p < n; p += PERM_CTA) {
      float v[ROWS];
#pragma unroll
      for (int r = 0; r < ROWS; r++) v[r] = src[(size_t)(r0 + r) * k.stride + base + p];
#pragma unroll
      for (int r = 0; r < ROWS; r++) sm[r * k.Npad + p] = v[r];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += PERM_CTA) {
      const int pi = perm[base + i];
#pragma unroll
      for (int r = 0; r < ROWS; r++) dst[(size_t)(r0 + r) * k.stride + base + i] = sm[r * k.Npad + pi];
    }
  } else {
    for (int i = threadIdx.x; i < n; i += PERM_CTA) {
      const int pi = perm[base + i];
      float v[ROWS];
#pragma unroll
      for (int r = 0; r < ROWS; r++) v[r] = src[(size_t)(r0 + r) * k.stride + base + i];
#pragma unroll
      for (int r = 0; r < ROWS; r++) sm[r * k.Npad + pi] = v[r];
    }
    __syncthreads();
    for (int p = threadIdx.x; p < n; p += PERM_CTA) {
      float v[ROWS];
      if (MODE == 2) {
#pragma unroll
        for (int r = 0; r < ROWS; r++) v[r] = dst[(size_t)(r0 + r) * k.stride + base + p];
      }
#pragma unroll
      for (int r = 0; r < ROWS; r++)
        dst[(size_t)(r0 + r) * k.stride + base + p] = (MODE == 2 ? v[r] : 0.f) + sm[r * k.Npad + p];
    }
  }
}
// the sort's permutation alone: perm[slot in sorted order] = canonical index (k_sort_scatter without the data movement)
__global__ void k_sort_perm(SimConst k, const int* __restrict__ npart, const int* __restrict__ cell_start,
                            const int* __restrict__ key, const int* __restrict__ rank, int* __restrict__ perm) {
  DSK_TL(k);
  int gid = blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= k.stride) return;
  int env = gid / k.Npad, p = gid - env * k.Npad;
  if (p >= npart[env]) return;
  perm[env * k.Npad + cell_start[(size_t)env * k.nnode + key[gid]] + rank[gid]] = p;
}

// ---- particle-major <-> SoA conversion for the API --------------------------------------------------------
// aos: x[n,3] v[n,3] F[n,9] C[n,9] (reference layout); direction 0: aos -> frame, 1: frame -> aos, 2: frame += aos
__global__ void k_particles_io(SimConst k, float* __restrict__ frame, int env, int n, float* x, float* v, float* F,
                               float* C, int direction) {
  DSK_TL(k);
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  int gid = env * k.Npad + p;
  for (int c = 0; c < FRAME_COMPS; c++) {
    float* a;
    int idx;
    if (c < 3) { a = x; idx = p * 3 + c; }
    else if (c < 6) { a = v; idx = p * 3 + (c - 3); }
    else if (c < 15) { a = C; idx = p * 9 + (c - 6); }
    else { a = F; idx = p * 9 + (c - 15); }
    if (!a) continue;
    float* f = frame + c * k.stride + gid;
    if (direction == 0) *f = a[idx];
    else if (direction == 1) a[idx] = *f;
    else *f += a[idx];
  }
}
// batched [B,cap,3]-style adjoint injection (function.py:130-135): gx,gv [B,cap_in,3]; gF,gC [B,cap_in,9]
__global__ void k_add_particle_grad(SimConst k, float* __restrict__ frame, const int* __restrict__ npart, int cap_in,
                                    const float* gx, const float* gv, const float* gF, const float* gC) {
  DSK_TL(k);
  int gid = blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= k.stride) return;
  int env = gid / k.Npad, p = gid - env * k.Npad;
  if (p >= npart[env] || p >= cap_in) return;
  size_t row = (size_t)env * cap_in + p;
  if (gx) for (int c = 0; c < 3; c++) frame[(CX + c) * k.stride + gid] += gx[row * 3 + c];
  if (gv) for (int c = 0; c < 3; c++) frame[(CV + c) * k.stride + gid] += gv[row * 3 + c];
  if (gC) for (int c = 0; c < 9; c++) frame[(CC + c) * k.stride + gid] += gC[row * 9 + c];
  if (gF) for (int c = 0; c < 9; c++) frame[(CF + c) * k.stride + gid] += gF[row * 9 + c];
}
// observation: xv [B,cap,6] (function.py:90-95)
__global__ void k_get_obs(SimConst k, const float* __restrict__ frame, const int* __restrict__ npart, int cap_out,
                          float* __restrict__ xv) {
  DSK_TL(k);
  int gid = blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= k.stride) return;
  int env = gid / k.Npad, p = gid - env * k.Npad;
  if (p >= cap_out) return;
  size_t row = (size_t)env * cap_out + p;
  bool live = p < npart[env];
  for (int c = 0; c < 6; c++) xv[row * 6 + c] = live ? frame[c * k.stride + gid] : 0.f;
}
__global__ void k_scale(float* a, size_t n, float alpha) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] *= alpha;
}
__global__ void k_axpy(float* y, const float* x, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] += x[i];
}
__global__ void k_clip_actions(float* dst, const float* src, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = fminf(1.f, fmaxf(-1.f, src[i]));
}

// ---- tool kinematics: tool_fk / tool_fk_adj / action_to_vel live in tools.cuh ---------------------------------
DSK_DEV void store_pose(float* s, const Pose& P) {
  s[0] = P.p.x; s[1] = P.p.y; s[2] = P.p.z;
  s[3] = P.q.w; s[4] = P.q.x; s[5] = P.q.y; s[6] = P.q.z;
  s[7] = P.gap;
}
// get_surface_pos(project(rand*size)), mpm_simulator.py:291-292, primitives.py:369-372
DSK_DEV float3 surface_point(const ToolParams& Tj, const Pose& Pj, const float* rn) {
  float3 q;
  q.x = tmax(tmin(mul_rn(rn[0], Tj.size[0]), Tj.size[0]), -Tj.size[0]);
  q.y = tmax(tmin(mul_rn(rn[1], Tj.size[1]), Tj.size[1]), -Tj.size[1]);
  q.z = tmax(tmin(mul_rn(rn[2], Tj.size[2]), Tj.size[2]), -Tj.size[2]);
  return add3_rn(qrot_rn(Pj.q, q), Pj.p);
}
// tool states of the last substep frame -> destination checkpoint
__global__ void k_tool_store(SimConst k, const float* __restrict__ poses, const StepArgs* __restrict__ args) {
  DSK_TL(k);
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int per = k.K * 8;
  if (i >= k.B * per) return;
  int env = i / per, r = i - env * per;
  args->tool_dst[i] = poses[((size_t)env * (k.S + 1) + k.S) * per + r];
}
// pose_adj[B][S+1][K][8] = 0 except frame S = adjoint checkpoint step+1
__global__ void k_pose_adj_init(SimConst k, float* __restrict__ pose_adj, const StepArgs* __restrict__ args) {
  DSK_TL(k);
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int per = k.K * 8;
  if (i >= k.B * (k.S + 1) * per) return;
  int env = i / ((k.S + 1) * per), r = i - env * (k.S + 1) * per;
  int f = r / per, q = r - f * per;
  pose_adj[i] = (f == k.S && args) ? args->tool_adj_in[env * per + q] : 0.f;   // args == null: all zero, seeded later
}
// deferred tool adjoints (dsk_backward_steps): frame S of a zero-initialised pose_adj += adjoint checkpoint step+1
__global__ void k_pose_adj_seed(SimConst k, float* __restrict__ pose_adj, const StepArgs* __restrict__ args) {
  DSK_TL(k);
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int per = k.K * 8;
  if (i >= k.B * per) return;
  int env = i / per, q = i - env * per;
  pose_adj[((size_t)env * (k.S + 1) + k.S) * per + q] += args->tool_adj_in[i];
}
// adjoint checkpoint step += pose_adj frame 0
__global__ void k_tool_adj_accum(SimConst k, const float* __restrict__ pose_adj, const StepArgs* __restrict__ args) {
  DSK_TL(k);
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int per = k.K * 8;
  if (i >= k.B * per) return;
  int env = i / per, r = i - env * per;
  args->tool_adj_out[i] += pose_adj[(size_t)env * (k.S + 1) * per + r];
}
// sum_t w * mean_p |x - target|^2 per env (a deterministic stand-in for the reference's torch-side losses) and its
// gradient accumulated into an adjoint checkpoint.  target [B,cap,3]; loss [B] (+=)
__global__ void k_loss_l2(SimConst k, const float* __restrict__ frame, float* __restrict__ adj,
                          const int* __restrict__ npart, const float* __restrict__ target, int cap, float weight,
                          float* __restrict__ loss) {
  DSK_TL(k);
  int gid = blockIdx.x * blockDim.x + threadIdx.x;
  int env = gid < k.stride ? gid / k.Npad : 0, p = gid - env * k.Npad;
  int n = npart[env];
  float l = 0.f;
  if (gid < k.stride && p < n && p < cap) {
    float sc = weight / (float)n;
#pragma unroll
    for (int c = 0; c < 3; c++) {
      float d = frame[c * k.stride + gid] - target[((size_t)env * cap + p) * 3 + c];
      l += sc * d * d;
      adj[c * k.stride + gid] += 2.f * sc * d;
    }
  }
  l = warp_sum(l);
  if ((threadIdx.x & 31) == 0 && l != 0.f && gid < k.stride) atomicAdd(&loss[env], l);
}

#define KIN_CTA 1024   // largest CTA (one per env): its warps share the S*npairs collision queries of the optimistic pass; see kin_block()
// One CTA per env: S substeps of forward_kinematics for every tool, then (if any pair) set_surface_points,
// set_collision_idx (deterministic first minimum) and apply_collision_projection (mpm_simulator.py:286-305).
// Tool-tool projections are rare, so the kernel is optimistic: (1) the kinematics chain of all S substeps without
// projections, one thread per tool; (2) the collision query of every (substep, pair) in parallel, one warp per query;
// (3) only from the first substep that actually hit, the reference's sequential per-substep procedure.
__global__ void __launch_bounds__(KIN_CTA)
    k_kinematics(SimConst k, const ToolParams* __restrict__ tools, const StepArgs* __restrict__ args,
                 const float* __restrict__ rand_num, float* __restrict__ poses /*[B][S+1][K][8]*/,
                 int* __restrict__ cidx /*[B][S+1][npairs]*/) {
  DSK_TL(k);
  const float* __restrict__ state0 = args->tool_src;  // [B][K][8]
  const float* __restrict__ action = args->action;    // [B][A] or null
  __shared__ ToolParams sT[DSK_MAX_TOOLS];
  __shared__ float sP[DSK_MAX_TOOLS][8];    // current poses (frame j+1 under construction)
  __shared__ float sPre[DSK_MAX_TOOLS][8];  // poses after FK, before any projection
  __shared__ float red_d[KIN_CTA / 32];
  __shared__ int red_i[KIN_CTA / 32];
  __shared__ int s_idx[DSK_MAX_PAIRS];
  __shared__ int s_first;
  DSK_DYN_SMEM(float, sAll);           // [(S+1)][K][8] projection-free chain, then [npairs][600][3] collision samples
  int env = blockIdx.x, tid = threadIdx.x;
  const int per = k.K * 8;
  for (int i = tid; i < k.K * (int)(sizeof(ToolParams) / 4); i += blockDim.x) ((int*)sT)[i] = ((const int*)tools)[i];
  for (int i = tid; i < per; i += blockDim.x) sAll[i] = state0[(size_t)env * per + i];
  float* sRand = sAll + (size_t)(k.S + 1) * per;
  for (int i = tid; i < k.npairs * DSK_NUM_COLLISION_POINTS * 3; i += blockDim.x) sRand[i] = rand_num[i];
  if (tid == 0) s_first = k.S;
  __syncthreads();
  int A = 0;
  for (int t = 0; t < k.K; t++) A += sT[t].action_dim;
  if (tid < k.K) {   // (1) projection-free chain
    int off = 0;
    for (int t = 0; t < tid; t++) off += sT[t].action_dim;
    float zero[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    ToolVel u = action_to_vel(sT[tid], action ? action + (size_t)env * A + off : zero, k.S);
    ToolRotInc ri = tool_rot_inc(sT[tid], u);
    Pose Pc = load_pose(sAll + tid * 8);
    for (int j = 0; j < k.S; j++) {
      Pc = tool_fk_inc(sT[tid], Pc, u, ri);
      store_pose(sAll + (size_t)(j + 1) * per + tid * 8, Pc);
    }
  }
  __syncthreads();
  if (k.npairs > 0) {   // (2) all collision queries at once, one warp per (substep, pair)
    int warp = tid >> 5, lane = tid & 31;
    for (int it = warp; it < k.S * k.npairs; it += (int)(blockDim.x >> 5)) {
      int j = it / k.npairs, c = it - j * k.npairs;
      int ti = k.pairs[c][0], tj = k.pairs[c][1];
      Pose Pi = load_pose(sAll + (size_t)(j + 1) * per + ti * 8), Pj = load_pose(sAll + (size_t)(j + 1) * per + tj * 8);
      // the frames of tool i and their normalised inverse rotations are shared by the 600 points of the query
      const ToolParams& Ti = sT[ti];
      bool grip = is_gripper(Ti.type);
      Frame Fa = grip ? jaw_frame(Pi, -1.f) : tool_frame(Pi), Fb = grip ? jaw_frame(Pi, 1.f) : Fa;
      Q4 qa = qconj_normalized_rn(Fa.q);
      int kind = sdf_kind(Ti.type);
      float best = 0.f;
      int bi = -1;
      for (int q = lane; q < DSK_NUM_COLLISION_POINTS; q += 32) {
        float3 pt = surface_point(sT[tj], Pj, sRand + ((size_t)c * DSK_NUM_COLLISION_POINTS + q) * 3);
        float d = local_sdf(Ti, kind, qrot_rn(qa, sub3_rn(pt, Fa.o)));
        if (grip) d = tmin(d, local_sdf(Ti, kind, qrot_rn(qa, sub3_rn(pt, Fb.o))));
        if (d < best) {
          best = d;
          bi = q;
        }
      }
      for (int o = 16; o > 0; o >>= 1) {
        float od = __shfl_xor_sync(0xffffffffu, best, o);
        int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (oi >= 0 && (bi < 0 || od < best || (od == best && oi < bi))) {
          best = od;
          bi = oi;
        }
      }
      if (lane == 0) {
        cidx[((size_t)env * (k.S + 1) + (j + 1)) * k.npairs + c] = bi;
        if (bi >= 0) atomicMin(&s_first, j);
      }
    }
    __syncthreads();
  }
  const int j0 = s_first;   // first substep whose frame needs a projection (S: none)
  float* out = poses + (size_t)env * (k.S + 1) * k.K * 8;
  for (int i = tid; i < (j0 + 1) * per && i < (k.S + 1) * per; i += blockDim.x) out[i] = sAll[i];
  if (j0 >= k.S) return;
  for (int i = tid; i < per; i += blockDim.x) sP[i / 8][i % 8] = sAll[(size_t)j0 * per + i];
  __syncthreads();
  // (3) the reference's sequential procedure from substep j0 on
  ToolVel u3;
  ToolRotInc ri3;
  if (tid < k.K) {
    int off = 0;
    for (int t = 0; t < tid; t++) off += sT[t].action_dim;
    float zero[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    u3 = action_to_vel(sT[tid], action ? action + (size_t)env * A + off : zero, k.S);
    ri3 = tool_rot_inc(sT[tid], u3);
  }
  for (int j = j0; j < k.S; j++) {
    if (tid < k.K) {
      Pose N = tool_fk_inc(sT[tid], load_pose(sP[tid]), u3, ri3);   // same arithmetic as the optimistic chain
      store_pose(sP[tid], N);
      store_pose(sPre[tid], N);
    }
    __syncthreads();
    if (k.npairs > 0) {
      // all pairs at once: the warps are split evenly between the pairs (every query reads the pre-projection poses)
      const int wpp = (int)(blockDim.x >> 5) / k.npairs;   // warps per pair (>= 8 warps per CTA, DSK_MAX_PAIRS <= 8 -> at least 1)
      const int warp = tid >> 5, lane = tid & 31;
      const int c = min(warp / wpp, k.npairs - 1), wq = warp - c * wpp;
      const bool worker = warp < wpp * k.npairs;
      float best = 0.f;
      int bi = -1;
      if (worker) {
        int ti = k.pairs[c][0], tj = k.pairs[c][1];
        Pose Pi = load_pose(sPre[ti]), Pj = load_pose(sPre[tj]);
        const ToolParams& Ti = sT[ti];   // tool_sdf with the frames and their inverse rotation hoisted, as in (2)
        bool grip = is_gripper(Ti.type);
        Frame Fa = grip ? jaw_frame(Pi, -1.f) : tool_frame(Pi), Fb = grip ? jaw_frame(Pi, 1.f) : Fa;
        Q4 qa = qconj_normalized_rn(Fa.q);
        int kind = sdf_kind(Ti.type);
        for (int q = wq * 32 + lane; q < DSK_NUM_COLLISION_POINTS; q += wpp * 32) {
          float3 pt = surface_point(sT[tj], Pj, sRand + ((size_t)c * DSK_NUM_COLLISION_POINTS + q) * 3);
          float d = local_sdf(Ti, kind, qrot_rn(qa, sub3_rn(pt, Fa.o)));
          if (grip) d = tmin(d, local_sdf(Ti, kind, qrot_rn(qa, sub3_rn(pt, Fb.o))));
          if (d < best) {
            best = d;
            bi = q;
          }
        }
      }
      // arg-min, ties -> smaller index ("first minimum"): warp, then the first warp of each pair over its warps
      for (int o = 16; o > 0; o >>= 1) {
        float od = __shfl_xor_sync(0xffffffffu, best, o);
        int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (oi >= 0 && (bi < 0 || od < best || (od == best && oi < bi))) {
          best = od;
          bi = oi;
        }
      }
      if (lane == 0) {
        red_d[warp] = best;
        red_i[warp] = bi;
      }
      __syncthreads();
      if (worker && wq == 0) {
        best = lane < wpp ? red_d[c * wpp + lane] : 0.f;
        bi = lane < wpp ? red_i[c * wpp + lane] : -1;
        for (int o = 16; o > 0; o >>= 1) {
          float od = __shfl_xor_sync(0xffffffffu, best, o);
          int oi = __shfl_xor_sync(0xffffffffu, bi, o);
          if (oi >= 0 && (bi < 0 || od < best || (od == best && oi < bi))) {
            best = od;
            bi = oi;
          }
        }
        if (lane == 0) {
          s_idx[c] = bi;
          cidx[((size_t)env * (k.S + 1) + (j + 1)) * k.npairs + c] = bi;
        }
      }
      __syncthreads();
      // collision_projection, primive_base.py:145-150, in pair order for every moved tool: thread t handles the
      // pairs that move tool t (different tools are independent: a projection changes only its own tool's position)
      if (tid < k.K) {
        for (int c2 = 0; c2 < k.npairs; c2++) {
          int ti = k.pairs[c2][0], tj = k.pairs[c2][1];
          if (ti != tid || s_idx[c2] < 0) continue;
          float3 pt = surface_point(sT[tj], load_pose(sPre[tj]),
                                    sRand + ((size_t)c2 * DSK_NUM_COLLISION_POINTS + s_idx[c2]) * 3);
          Pose Pi = load_pose(sP[ti]);
          // d = tool_sdf, nr = tool_normal (primitives.py:489-496 picks the jaw with da <= db), frames built once
          const ToolParams& Ti = sT[ti];
          bool grip = is_gripper(Ti.type);
          int kind = sdf_kind(Ti.type);
          Frame Fa = grip ? jaw_frame(Pi, -1.f) : tool_frame(Pi);
          Q4 qa = qconj_normalized_rn(Fa.q);
          float3 pl = qrot_rn(qa, sub3_rn(pt, Fa.o));
          float d = local_sdf(Ti, kind, pl);
          Q4 qf = Fa.q;
          if (grip) {
            Frame Fb = jaw_frame(Pi, 1.f);
            float3 plb = qrot_rn(qconj_normalized_rn(Fb.q), sub3_rn(pt, Fb.o));
            float db = local_sdf(Ti, kind, plb);
            if (!(d <= db)) {
              pl = plb;
              qf = Fb.q;
            }
            d = tmin(d, db);
          }
          float3 nr = qrot_rn(qf, local_normal(Ti, kind, pl));
          float inv = __fdiv_rn(1.f, __fsqrt_rn(dot_rn(nr, nr)));
          sP[ti][0] = add_rn(Pi.p.x, mul_rn(mul_rn(inv, nr.x), d));
          sP[ti][1] = add_rn(Pi.p.y, mul_rn(mul_rn(inv, nr.y), d));
          sP[ti][2] = add_rn(Pi.p.z, mul_rn(mul_rn(inv, nr.z), d));
        }
      }
      __syncthreads();
    }
    for (int i = tid; i < k.K * 8; i += blockDim.x) out[(size_t)(j + 1) * k.K * 8 + i] = sP[i / 8][i % 8];
    __syncthreads();
  }
}
