// Host side of the C ABI (include/diffskill_mpm.h): owns all device state, schedules the
// kernels of kernels_{aux,fwd,bwd}.cuh on one CUDA stream.  No torch, no CPU fallback: every
// entry point either launches the sm_100a kernels or fails.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/diffskill_mpm.h"
#include "kernels_bwd.cuh"

static thread_local std::string g_err;
static int fail(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return -1;
}
#define CK(call)                                                                                    \
  do {                                                                                              \
    cudaError_t err__ = (call);                                                                     \
    if (err__ != cudaSuccess) return fail("%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(err__)); \
  } while (0)
#define CKE(e)                                                     \
  do {                                                             \
    if (!(e)) return fail("null engine");                          \
    CK(cudaSetDevice((e)->cfg.device));                            \
  } while (0)
#define LAUNCH_CHECK() CK(cudaGetLastError())

struct StepSlot {
  float* frames;     // (S+1) particle frames, sorted order
  float* mat;        // sorted material arrays [3][stride]
  int* perm;         // sorted slot -> particle id, [stride]
  float* poses;      // [B][S+1][K][8]
  int* cidx;         // [B][S+1][npairs]
  float* svd = nullptr;  // [S][SVD_COMPS][stride] U, sigma, V of every substep's F_tmp (null: adjoint recomputes the SVD)
  GridTape tape = {nullptr, nullptr, nullptr, nullptr, 0};
  bool tape_written = false;  // the slot's tape belongs to src_step
  int src_step = -1; // checkpoint the frames were simulated from (-1: invalid)
  int action_step = -1;
};

enum KernelId {
  KID_SORT = 0, KID_KINEMATICS, KID_P2G, KID_GRID, KID_G2P, KID_P2G_RECOMPUTE, KID_GRID_RECOMPUTE, KID_G2P_ADJ,
  KID_GRID_ADJ, KID_P2G_ADJ, KID_KINEMATICS_ADJ, KID_REORDER, KID_IO, KID_LOSS, KID_G2P2G, KID_GRID_ADJ_TOOLS, KID_COUNT
};
static const char* kKernelNames[KID_COUNT] = {
    "sort", "kinematics", "p2g", "grid_op", "g2p", "p2g_recompute", "grid_op_recompute", "g2p_adj",
    "grid_op_adj", "p2g_adj", "kinematics_adj", "reorder", "io", "loss", "g2p2g", "grid_op_adj_tools"};

struct ProfRec {
  int kid;
  cudaEvent_t a, b;
};

struct dsk_engine {
  dsk_config cfg;
  bool profiling = false;
  std::vector<ProfRec> prof;
  std::vector<cudaEvent_t> ev_pool;
  int64_t kid_launches[KID_COUNT] = {0};
  SimConst k;
  cudaStream_t stream = 0;
  int B, Npad, S, H, K, A, slots, ncols, n_frames = 1;
  size_t frame_floats, tool_floats;
  int64_t launches = 0, bytes = 0;
  std::vector<void*> allocs;
  // particles
  float* ckpt = nullptr;      // (H+1) frames, canonical order
  float* adj_ckpt = nullptr;  // (H+1) adjoint frames
  float* adjw[2] = {nullptr, nullptr};
  float* mat = nullptr;  // [3][stride] canonical
  int* npart = nullptr;
  std::vector<int> h_npart;
  std::vector<StepSlot> slot;
  // sort scratch
  int *cell_count = nullptr, *key = nullptr, *rank = nullptr, *scan_partial = nullptr, *chunk_flag = nullptr;
  // grids: set index = epoch & 1
  float4 *G0[2], *Gv[2], *Ga[2];
  int *tile_epoch[2], *tile_list[2], *tile_count = nullptr;  // tile_count[4] ring
  int dirty[2] = {0, 0};  // bit0 G0, bit1 Gv, bit2 Ga written in that set since its last clear
  int epoch = 0;
  // tools
  ToolParams* d_tools = nullptr;
  GridTools grid_tools;   // by-value copy for the grid kernels (refreshed by sync_grid_tools)
  std::vector<ToolParams> h_tools;
  float *tool_ckpt = nullptr, *tool_adj_ckpt = nullptr;  // [H+1][B][K][8]
  float* pose_adj = nullptr;                             // [B][S+1][K][8] (the buffer the sequence being enqueued uses)
  float* pose_adj_buf[2] = {nullptr, nullptr};           // two of them: the tool adjoints of step s run on tool_stream while
                                                         // the main stream already simulates step s-1 (dsk_backward_steps)
  cudaStream_t tool_stream = nullptr;
  cudaEvent_t ev_bwd_main[2] = {nullptr, nullptr}, ev_bwd_tools[2] = {nullptr, nullptr};
  bool tools_pending[2] = {false, false};
  StepArgs* d_args_bwd = nullptr;                        // [H] per-step args of the deferred tool adjoints (written once)
  bool seq_defer_tools = false;                          // backward sequence being enqueued leaves the tool adjoints out
  float *actions = nullptr, *action_grad = nullptr;      // [H][B][A]
  float* rand_num = nullptr;
  // io staging
  float* stage = nullptr;
  size_t stage_floats = 0;
  // fine-grained substep cursor
  int last_fwd_frame = -1, last_bwd_frame = -1, last_substep_slot = -1, last_substep_j = -1;
  bool last_was_backward = false;
  int bwd_cur = 0;  // adjw index holding the adjoint of the current frame
  // sequences / graphs
  struct GraphSet {
    cudaGraphExec_t ex[24] = {};  // [kind*2 + full_sort]
    int64_t n_launch[24] = {0};
    int64_t kid[24][KID_COUNT] = {{0}};
  };
  std::vector<GraphSet> graphs;
  bool use_graphs = true;
  cudaStream_t cap_stream = nullptr, qs = 0;  // capture stream; stream KL currently enqueues on
  StepArgs* d_args = nullptr;
  int epoch_base = 0;
  int* done = nullptr;
  bool seq_use_tape = false;  // the backward sequence being enqueued restores grids from the slot's tape
  bool seq_tape_trusted = false;  // ... and the host has verified that the tape did not overflow (no fallback kernels)
  int pending_q = -1;  // fine-grained mode: last substep's grids still hold data
  bool pending_bwd = false, grids_valid = false;
  float* loss = nullptr;  // [B]
  int* perm_cache = nullptr;   // permutation of the last full sort
  int sort_age = 1 << 30, resort_interval = 1;   // >1 re-uses the last permutation (cheaper sort, more fragmented warps)
  bool seq_full_sort = true;
  bool seq_skip_kin = false;          // sequence being enqueued must not run its own tool kinematics / tool store
  StepSlot* seq_next_slot = nullptr;  // ... and runs the kinematics of the next step (this slot) on a side branch
  int seq_next_step = -1;
  cudaEvent_t ev_join2 = nullptr;
  StepArgs* d_args_kin = nullptr;     // [H] device-resident args of the lookahead kinematics (pointers only: set once)
#ifdef DSK_TIMELINE
  TlRec* d_tl = nullptr;
  int tl_cap = 16384, tl_next = 0;
  bool tl_on = false;
  std::vector<int> tl_kid;
#endif
  bool big = false;   // enough particles to fill the machine: prefer occupancy over registers
  float* gadj_scratch[2] = {nullptr, nullptr};   // parked contact adjoints of k_grid_adj / k_grid_adj_flat
  int* gadj_flags[2] = {nullptr, nullptr};
  int gadj_cap = 0;
  // resident 128-thread CTAs per SM requested from the particle kernels of batched engines (register cap 65536/(128*n))
  int big_block = 128;
  int minb_g2p2g = 4, minb_g2p_adj = 4, minb_p2g_adj = 4;
  int grid_ctas_per_sm = 4;  // latency layout of the grid kernels: CTAs per SM walking the active-tile list (r02j: 8-env GatherMove 48.1 -> 46.1 ms from 2 to 4)
  bool mat_uniform = true;  // no per-particle material set: the particle kernels take (mu, lam, yield_stress) from SimConst
  bool flat_grid = false;   // many active tiles: throughput layout of the grid kernels
  bool ts = true;           // batched engines: transposed shared-memory scatter (warp_scatter27_ts_affine) instead of the shuffle butterfly
  int flat_fwd_ctas_per_sm = 4, flat_adj_ctas_per_sm = 4;   // grid of the throughput-layout grid kernels (4 CTAs are resident per SM)
  bool perm_smem = true;    // batched engines: frame permutations through shared memory (k_permute_rows)
  bool perm_smem_small = false;   // ... also for single scenes (DSK_PERM_SMEM_SMALL)
  bool perm_opt_in[3] = {false, false, false};
  bool ts_pl = false;       // ... in the plane-split kernels of single scenes (slower there: r02b liftspread 46.3 vs 42.3 ms)
  cudaStream_t cap_side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  std::vector<cudaEvent_t> ev_restored, ev_main;   // per backward position, used while capturing the pipelined adjoint
  bool kin_join = false;
  int tape_cap = 0;       // grid-tape capacity per step slot, in tiles (0: taping off)
  bool tape_flags_stale = true;
  int* tape_flags = nullptr;   // [slots] device overflow flags
  std::vector<int> tape_overflow;  // host copy of the slots' overflow flags

  float* frame_of(float* base, int i) { return base + (size_t)i * frame_floats; }
  float* tools_of(float* base, int step) { return base + (size_t)step * tool_floats; }
};

static cudaEvent_t prof_event(dsk_engine* e) {
  if (!e->ev_pool.empty()) {
    cudaEvent_t ev = e->ev_pool.back();
    e->ev_pool.pop_back();
    return ev;
  }
  cudaEvent_t ev;
  cudaEventCreate(&ev);
  return ev;
}
// every kernel launch of the engine goes through KL: counts it and, when profiling, brackets it with events
#ifdef DSK_TIMELINE
#define TL_ASSIGN(kid_)                                                  \
  do {                                                                   \
    if (e->tl_on && e->tl_next < e->tl_cap) {                            \
      e->k.tl_slot = e->tl_next++;                                       \
      e->tl_kid.push_back(kid_);                                         \
    } else {                                                             \
      e->k.tl_slot = -1;                                                 \
    }                                                                    \
  } while (0)
#else
#define TL_ASSIGN(kid_) do { } while (0)
#endif
#define KL(kid_, ...)                                          \
  do {                                                         \
    ProfRec pr__;                                              \
    if (e->profiling) {                                        \
      pr__.kid = (kid_);                                       \
      pr__.a = prof_event(e);                                  \
      pr__.b = prof_event(e);                                  \
      cudaEventRecord(pr__.a, e->qs);                      \
    }                                                          \
    TL_ASSIGN(kid_);                                           \
    __VA_ARGS__;                                               \
    if (e->profiling) {                                        \
      cudaEventRecord(pr__.b, e->qs);                      \
      e->prof.push_back(pr__);                                 \
    }                                                          \
    e->launches += 1;                                          \
    e->kid_launches[(kid_)] += 1;                              \
  } while (0)

template <class T>
static int dalloc(dsk_engine* e, T** p, size_t count, bool zero = true) {
  void* q = nullptr;
  size_t nb = std::max<size_t>(count, 1) * sizeof(T);
  CK(cudaMalloc(&q, nb));
  if (zero) CK(cudaMemset(q, 0, nb));
  e->allocs.push_back(q);
  e->bytes += (int64_t)nb;
  *p = (T*)q;
  return 0;
}
#define DA(p, n)                        \
  do {                                  \
    if (dalloc(e, &(p), (n))) return -1; \
  } while (0)

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static float* svd_at(dsk_engine* e, StepSlot& s, int j);
static size_t kin_smem(dsk_engine* e);
static int kin_block(dsk_engine* e, bool hidden);
static void drop_graphs(dsk_engine* e);
static int ts_opt_in(dsk_engine* e);

static void fill_tool(ToolParams& T, const dsk_tool_desc& d) {
  memset(&T, 0, sizeof T);
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
  // normalised conjugate in fp32 with the reference's operation order (primitives.py:713-714)
  volatile float n2 = w * w;
  n2 = n2 + x * x;
  n2 = n2 + y * y;
  n2 = n2 + z * z;
  volatile float inv = 1.0f / sqrtf(n2);
  T.prot_inv = Q4{inv * w, inv * -x, inv * -y, inv * -z};
  T.min_gap = (float)d.minimal_gap;
  T.max_gap = (float)d.maximal_gap;
  // radius of a sphere around the contact-frame origin (tool position, or a jaw's position) that contains the shape
  if (d.type == DSK_TOOL_CAPSULE || d.type == DSK_TOOL_ROLLINGPIN_EXT || d.type == DSK_TOOL_ROLLINGPIN ||
      d.type == DSK_TOOL_GRIPPER2)
    T.bound_r = (float)(d.h / 2 + d.r);
  else if (d.type == DSK_TOOL_SPHERE)
    T.bound_r = (float)d.r;
  else if (d.type == DSK_TOOL_CYLINDER)   // radial extent h, axial half extent r
    T.bound_r = (float)std::sqrt(d.h * d.h + d.r * d.r);
  else if (d.type == DSK_TOOL_TORUS)      // major + minor radius
    T.bound_r = (float)(d.h + d.r);
  else if (d.type == DSK_TOOL_CHOPSTICKS) {   // two capsules spanning local y in [-h, 0] at x = -+gap/2; no upper clamp on the gap
    T.max_gap = 1e30f;
    double g = std::max(d.maximal_gap > 0 && d.maximal_gap < 1e3 ? d.maximal_gap : 0.0, 1.0);   // a gap never exceeds the unit box
    T.bound_r = (float)std::sqrt((g / 2 + d.r) * (g / 2 + d.r) + (d.h + d.r) * (d.h + d.r));
  }
  else
    T.bound_r = (float)std::sqrt(d.size[0] * d.size[0] + d.size[1] * d.size[1] + d.size[2] * d.size[2]);
  T.bound_r *= 1.001f;
}
static bool host_is_gripper(int type) { return type == DSK_TOOL_GRIPPER || type == DSK_TOOL_GRIPPER2; }
// host mirror of build_frame_table: one contact frame per tool, two (the jaws) per gripper
static void sync_grid_tools(dsk_engine* e) {
  GridTools& g = e->grid_tools;
  memset(&g, 0, sizeof g);
  int n = 0;
  for (int t = 0; t < e->K; t++) {
    g.T[t] = e->h_tools[t];
    if (host_is_gripper(e->h_tools[t].type)) {
      if (n + 2 > MAX_FRAMES) break;
      g.ft.tool[n] = t; g.ft.flag[n++] = -1.f;
      g.ft.tool[n] = t; g.ft.flag[n++] = 1.f;
    } else {
      if (n + 1 > MAX_FRAMES) break;
      g.ft.tool[n] = t; g.ft.flag[n++] = 0.f;
    }
  }
  g.ft.n = n;
}

extern "C" {

const char* dsk_last_error(void) { return g_err.c_str(); }
int dsk_abi_version(void) { return DSK_ABI_VERSION; }
int dsk_sizeof_config(void) { return (int)sizeof(dsk_config); }
int dsk_sizeof_tool_desc(void) { return (int)sizeof(dsk_tool_desc); }

int dsk_create(const dsk_config* c, dsk_engine** out) {
  if (!c || !out) return fail("null argument");
  if (c->abi_version != DSK_ABI_VERSION) return fail("ABI version mismatch: header %d, library %d", c->abi_version, DSK_ABI_VERSION);
  if (c->n_grid % 4 != 0 || c->n_grid < 8) return fail("n_grid must be a multiple of 4 (got %d)", c->n_grid);
  if (c->n_envs < 1 || c->particle_capacity < 1 || c->substeps < 1 || c->max_steps < 1) return fail("bad sizes");
  if (c->n_tools < 0 || c->n_tools > DSK_MAX_TOOLS || c->n_pairs < 0 || c->n_pairs > DSK_MAX_PAIRS) return fail("too many tools/pairs");
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  if (c->device < 0 || c->device >= ndev) return fail("CUDA device %d not available (%d devices)", c->device, ndev);
  CK(cudaSetDevice(c->device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, c->device));
  if (prop.major < 10) return fail("this library contains sm_100a code only; device is sm_%d%d", prop.major, prop.minor);
  dsk_engine* e = new dsk_engine();
  e->cfg = *c;
  e->B = c->n_envs;
  e->Npad = (c->particle_capacity + 127) / 128 * 128;
  e->S = c->substeps;
  e->H = c->max_steps;
  e->K = c->n_tools;
  e->slots = std::max(1, std::min(c->step_slots, c->max_steps));
  if (c->grid_tape_mib > 0) {
    size_t per_slot = ((size_t)c->grid_tape_mib << 20) / e->slots;
    e->tape_cap = (int)std::min<size_t>(per_slot / 2052, (size_t)1 << 30);
    if (e->tape_cap < c->substeps * 8) e->tape_cap = 0;
  }
  e->tape_overflow.assign(e->slots, 1);
  SimConst& k = e->k;
  memset(&k, 0, sizeof k);
  k.n = c->n_grid;
  k.nt = k.n / 4;
  k.ntile = k.nt * k.nt * k.nt;
  k.nnode = k.n * k.n * k.n;
  k.B = e->B;
  k.Npad = e->Npad;
  k.stride = e->B * e->Npad;
  k.S = e->S;
  k.K = e->K;
  k.npairs = c->n_pairs;
  k.gf_mode = c->ground_friction == 0.0 ? 0 : (c->ground_friction < 10.0 ? 1 : 2);
  k.dt = (float)c->dt;
  k.dx = (float)c->dx;
  k.inv_dx = (float)c->inv_dx;
  k.p_mass = (float)c->p_mass;
  k.c_stress = (float)(-c->dt * c->p_vol * 4 * c->inv_dx * c->inv_dx);  // mpm_simulator.py:214
  k.c_C = (float)(4 * c->inv_dx);                                       // :279
  k.x_hi = (float)(1. - 3 * c->dx);                                     // :283
  k.x_lo = (float)(c->lower_bound * c->dx);
  k.m_eps = 1e-12f;
  k.ground_friction = (float)c->ground_friction;
  k.mu = (float)c->mu;
  k.lam = (float)c->lam;
  k.ys = (float)c->yield_stress;
  for (int d = 0; d < 3; d++) {
    volatile float t = k.dt * (float)c->gravity[d];
    k.grav[d] = t * 30.f;
  }
  for (int i = 0; i < c->n_pairs; i++) {
    k.pairs[i][0] = c->pairs[i][0];
    k.pairs[i][1] = c->pairs[i][1];
    if (c->pairs[i][0] < 0 || c->pairs[i][0] >= e->K || c->pairs[i][1] < 0 || c->pairs[i][1] >= e->K) {
      delete e;
      return fail("bad collision pair");
    }
  }
  e->A = 0;
  e->ncols = 0;
  e->h_tools.resize(std::max(1, e->K));
  for (int i = 0; i < e->K; i++) {
    fill_tool(e->h_tools[i], c->tools[i]);
    e->A += c->tools[i].action_dim;
    e->ncols += host_is_gripper(c->tools[i].type) ? 2 : 1;
    e->n_frames = std::max(1, e->ncols);
    if (c->tools[i].type < 0 || c->tools[i].type > DSK_TOOL_CHOPSTICKS) {
      delete e;
      return fail("unknown tool type %d", c->tools[i].type);
    }
  }
  e->frame_floats = (size_t)FRAME_COMPS * k.stride;
#ifdef DSK_TIMELINE
  e->k.tl = nullptr;
  e->k.tl_slot = -1;
#endif
  // measured crossover (GatherMove 8 vs 16 envs, LiftSpread 2 envs): ~24-32 k particles, i.e. ~1.5 warps per scheduler
  e->big = (size_t)c->n_envs * c->particle_capacity >= 24576;
  if (const char* v = getenv("DSK_FORCE_BIG")) e->big = atoi(v) != 0;
  if (const char* v = getenv("DSK_BIG_MINB")) e->minb_g2p2g = e->minb_g2p_adj = e->minb_p2g_adj = atoi(v);
  if (const char* v = getenv("DSK_BIG_BLOCK")) {   // threads per CTA of the batched particle kernels: 64, 96 or 128
    int b = atoi(v);
    e->big_block = (b == 64 || b == 96) ? b : 128;
  }
  if (const char* v = getenv("DSK_MINB_G2P2G")) e->minb_g2p2g = atoi(v);
  if (const char* v = getenv("DSK_MINB_G2P_ADJ")) e->minb_g2p_adj = atoi(v);
  if (const char* v = getenv("DSK_MINB_P2G_ADJ")) e->minb_p2g_adj = atoi(v);
  e->flat_grid = e->big;
  if (const char* v = getenv("DSK_GRID_CTAS_PER_SM")) e->grid_ctas_per_sm = std::max(1, atoi(v));
  if (const char* v = getenv("DSK_TS")) e->ts = atoi(v) != 0;
  if (const char* v = getenv("DSK_PERM_SMEM")) e->perm_smem = atoi(v) != 0;
  if (const char* v = getenv("DSK_FLAT_FWD_CTAS")) e->flat_fwd_ctas_per_sm = std::max(1, atoi(v));
  if (const char* v = getenv("DSK_FLAT_ADJ_CTAS")) e->flat_adj_ctas_per_sm = std::max(1, atoi(v));
  if (const char* v = getenv("DSK_PERM_SMEM_SMALL")) e->perm_smem_small = atoi(v) != 0;
  if (const char* v = getenv("DSK_TS_PL")) e->ts_pl = atoi(v) != 0;
  if (const char* v = getenv("DSK_FLAT_GRID")) e->flat_grid = atoi(v) != 0;
  e->tool_floats = (size_t)e->B * std::max(1, e->K) * 8;
  int rc = [&]() -> int {
    DA(e->ckpt, (size_t)(e->H + 1) * e->frame_floats);
    DA(e->adj_ckpt, (size_t)(e->H + 1) * e->frame_floats);
    DA(e->adjw[0], e->frame_floats);
    DA(e->adjw[1], e->frame_floats);
    DA(e->mat, (size_t)3 * k.stride);
    DA(e->npart, e->B);
    e->h_npart.assign(e->B, 0);
    e->slot.resize(e->slots);
    if (e->tape_cap > 0) DA(e->tape_flags, e->slots);
    for (auto& s : e->slot) {
      if (e->tape_cap > 0) s.tape.overflow = e->tape_flags + (&s - e->slot.data());
      DA(s.frames, (size_t)(e->S + 1) * e->frame_floats);
      DA(s.mat, (size_t)3 * k.stride);
      DA(s.perm, k.stride);
      DA(s.poses, (size_t)(e->S + 1) * e->tool_floats);
      DA(s.cidx, (size_t)e->B * (e->S + 1) * std::max(1, k.npairs));
      if (!getenv("DSK_NO_SVD_TAPE")) DA(s.svd, (size_t)e->S * SVD_COMPS * k.stride);
      if (e->tape_cap > 0) {
        s.tape.cap = e->tape_cap;
        DA(s.tape.base, e->S + 1);
        DA(s.tape.list, e->tape_cap);
        DA(s.tape.data, (size_t)e->tape_cap * 128);
      }
    }
    DA(e->cell_count, (size_t)e->B * k.nnode);
    DA(e->key, k.stride);
    DA(e->scan_partial, (size_t)e->B * cdiv(k.nnode, SCAN_CHUNK));
    DA(e->chunk_flag, (size_t)e->B * cdiv(k.nnode, SCAN_CHUNK));
    DA(e->rank, k.stride);
    for (int s = 0; s < 2; s++) {
      DA(e->G0[s], (size_t)e->B * k.nnode);
      DA(e->Gv[s], (size_t)e->B * k.nnode);
      DA(e->Ga[s], (size_t)e->B * k.nnode);
      DA(e->tile_epoch[s], (size_t)e->B * k.ntile);
      DA(e->tile_list[s], (size_t)e->B * k.ntile);
    }
    // throughput layout: parking is implemented (k_grid_adj_flat 18.3 -> 11.3 us on GatherMove x64) but the side-branch kernel
    // then costs 27 us per substep next to the particle kernels and the step gets slower (r03f: 95.8 -> 101.1 ms); opt-in
    if ((!e->flat_grid || getenv("DSK_FLAT_PARK")) && !getenv("DSK_NO_GRID_ADJ_SPLIT")) {
      e->gadj_cap = (int)std::min<size_t>((size_t)e->B * k.ntile, e->flat_grid ? 32768 : 4096);
      int nf = std::min(e->n_frames, MAX_FRAMES);
      for (int s = 0; s < 2; s++) {
        DA(e->gadj_scratch[s], (size_t)e->gadj_cap * nf * 7 * GRID_NODES);
        DA(e->gadj_flags[s], (size_t)e->gadj_cap * MAX_FRAMES * 2);
      }
    }
    if (kin_smem(e) > 48 * 1024) {
      if (kin_smem(e) > 200 * 1024) return fail("tool kinematics kernel needs %zu bytes of shared memory", kin_smem(e));
      CK(cudaFuncSetAttribute(k_kinematics, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kin_smem(e)));
    }
    if (ts_opt_in(e)) return -1;
    DA(e->tile_count, 4);
    DA(e->done, 1);
    DA(e->d_args, 1);
    DA(e->loss, e->B);
    e->graphs.resize(e->slots);
    CK(cudaStreamCreateWithFlags(&e->cap_stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&e->cap_side, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&e->ev_join, cudaEventDisableTiming));
    e->ev_restored.resize(e->S);
    e->ev_main.resize(e->S);
    for (int i = 0; i < e->S; i++) {
      CK(cudaEventCreateWithFlags(&e->ev_restored[i], cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&e->ev_main[i], cudaEventDisableTiming));
    }
    DA(e->perm_cache, k.stride);
    if (getenv("DSK_RESORT_INTERVAL")) e->resort_interval = std::max(1, atoi(getenv("DSK_RESORT_INTERVAL")));
    e->use_graphs = getenv("DSK_NO_GRAPHS") == nullptr;
    DA(e->d_tools, std::max(1, e->K));
    DA(e->tool_ckpt, (size_t)(e->H + 1) * e->tool_floats);
    DA(e->tool_adj_ckpt, (size_t)(e->H + 1) * e->tool_floats);
    DA(e->pose_adj_buf[0], (size_t)(e->S + 1) * e->tool_floats);
    DA(e->pose_adj_buf[1], (size_t)(e->S + 1) * e->tool_floats);
    e->pose_adj = e->pose_adj_buf[0];
    CK(cudaStreamCreateWithFlags(&e->tool_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; i++) {
      CK(cudaEventCreateWithFlags(&e->ev_bwd_main[i], cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&e->ev_bwd_tools[i], cudaEventDisableTiming));
    }
    DA(e->actions, (size_t)e->H * e->B * std::max(1, e->A));
    DA(e->action_grad, (size_t)e->H * e->B * std::max(1, e->A));
    DA(e->rand_num, (size_t)std::max(1, k.npairs) * DSK_NUM_COLLISION_POINTS * 3);
    e->stage_floats = std::max<size_t>((size_t)k.stride * 24, (size_t)e->B * k.nnode * 4);
    DA(e->stage, e->stage_floats);
    CK(cudaMemcpy(e->d_tools, e->h_tools.data(), sizeof(ToolParams) * std::max(1, e->K), cudaMemcpyHostToDevice));
    sync_grid_tools(e);
    // material fill, mpm_simulator.py:85-87
    std::vector<float> m((size_t)3 * k.stride);
    for (int i = 0; i < k.stride; i++) {
      m[i] = (float)c->mu;
      m[k.stride + i] = (float)c->lam;
      m[2 * (size_t)k.stride + i] = (float)c->yield_stress;
    }
    CK(cudaMemcpy(e->mat, m.data(), m.size() * 4, cudaMemcpyHostToDevice));
    return 0;
  }();
  if (rc) {
    for (void* p : e->allocs) cudaFree(p);
    delete e;
    return -1;
  }
  *out = e;
  return 0;
}

int dsk_destroy(dsk_engine* e) {
  if (!e) return 0;
  cudaSetDevice(e->cfg.device);
  cudaStreamSynchronize(e->stream);
  drop_graphs(e);
#ifdef DSK_TIMELINE
  if (e->d_tl) cudaFree(e->d_tl);
#endif
  if (e->cap_stream) cudaStreamDestroy(e->cap_stream);
  if (e->ev_join2) cudaEventDestroy(e->ev_join2);
  if (e->cap_side) cudaStreamDestroy(e->cap_side);
  if (e->tool_stream) cudaStreamDestroy(e->tool_stream);
  for (int i = 0; i < 2; i++) {
    if (e->ev_bwd_main[i]) cudaEventDestroy(e->ev_bwd_main[i]);
    if (e->ev_bwd_tools[i]) cudaEventDestroy(e->ev_bwd_tools[i]);
  }
  if (e->ev_fork) cudaEventDestroy(e->ev_fork);
  if (e->ev_join) cudaEventDestroy(e->ev_join);
  for (auto ev : e->ev_restored) cudaEventDestroy(ev);
  for (auto ev : e->ev_main) cudaEventDestroy(ev);
  for (auto& r : e->prof) {
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  for (auto ev : e->ev_pool) cudaEventDestroy(ev);
  for (void* p : e->allocs) cudaFree(p);
  delete e;
  return 0;
}
int dsk_set_stream(dsk_engine* e, void* s) {
  CKE(e);
  e->stream = (cudaStream_t)s;
  e->qs = e->stream;
  return 0;
}
int dsk_synchronize(dsk_engine* e) {
  CKE(e);
  CK(cudaStreamSynchronize(e->stream));
  return 0;
}
int dsk_set_rand_num(dsk_engine* e, const double* rn) {
  CKE(e);
  size_t n = (size_t)e->k.npairs * DSK_NUM_COLLISION_POINTS * 3;
  if (!n) return 0;
  std::vector<float> f(n);
  for (size_t i = 0; i < n; i++) f[i] = (float)rn[i];
  CK(cudaMemcpyAsync(e->rand_num, f.data(), n * 4, cudaMemcpyHostToDevice, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  return 0;
}
int dsk_launch_count(dsk_engine* e, int64_t* n) {
  if (!e) return fail("null engine");
  *n = e->launches;
  return 0;
}
int dsk_memory_bytes(dsk_engine* e, int64_t* n) {
  if (!e) return fail("null engine");
  *n = e->bytes;
  return 0;
}

}  // extern "C"

// -------------------------------------------------------------------------------------------------------
// internal scheduling
//
// A "sequence" is the kernel train of one env step: forward (sort, kinematics, S x {p2g, grid_op, g2p}, store),
// recompute (the same without the store) or backward (gather, S x {p2g, grid_op, g2p_adj, grid_op_adj,
// p2g_adj}, tool adjoints, scatter).  Every sequence starts and ends with all grids zero and all tile counters
// zero, and takes its step-dependent pointers from the device-resident StepArgs, so each of the three is
// captured ONCE per step slot into a CUDA graph and replayed for every env step.
// -------------------------------------------------------------------------------------------------------
static int check_step(dsk_engine* e, int step, const char* what) {
  if (step < 0 || step > e->H) return fail("%s: step %d outside [0, %d] (max_steps of this engine)", what, step, e->H);
  return 0;
}
static void invalidate_slots(dsk_engine* e, int step) {
  // a checkpoint was overwritten: substep frames simulated from it are stale
  for (auto& s : e->slot)
    if (s.src_step == step) s.src_step = -1;
}
static void invalidate_all(dsk_engine* e) {
  for (auto& s : e->slot) s.src_step = -1;
}
static void drop_graphs(dsk_engine* e) {
  for (auto& g : e->graphs)
    for (auto& x : g.ex)
      if (x) {
        cudaGraphExecDestroy(x);
        x = nullptr;
      }
}

// one staging copy in: host or device source -> device pointer valid on the stream
static int stage_in(dsk_engine* e, const float* src, size_t n, int on_device, size_t offset, const float** out) {
  if (!src) {
    *out = nullptr;
    return 0;
  }
  if (on_device) {
    *out = src;
    return 0;
  }
  if (offset + n > e->stage_floats) return fail("staging buffer too small");
  CK(cudaMemcpyAsync(e->stage + offset, src, n * 4, cudaMemcpyHostToDevice, e->stream));
  *out = e->stage + offset;
  return 0;
}

static int grid_ctas(dsk_engine* e) { return e->big ? 148 * 8 : 148 * e->grid_ctas_per_sm; }
static float* svd_at(dsk_engine* e, StepSlot& s, int j) {
  return s.svd ? s.svd + (size_t)j * SVD_COMPS * e->k.stride : nullptr;
}
// Threads per CTA of k_kinematics (one CTA per env).  Without tool-tool pairs the kernel is one pose chain per tool: two warps.
// With pairs the warps share the collision queries: 32 warps when the kernel is on the critical path; 8 on the lookahead
// branch, where it has a whole env step to finish and a 1024-thread CTA would hold a whole SM's register file next to the
// particle kernels for as long as it runs (~150 us per env step when a tool touches an obstacle from the first substep on
// and every substep takes the sequential path: GatherMove x64, r02z timeline).
static int kin_block(dsk_engine* e, bool hidden) { return e->k.npairs == 0 ? 64 : (hidden ? 256 : KIN_CTA); }
static size_t kin_smem(dsk_engine* e) {   // pose chain + the collision samples of every pair
  return (size_t)(e->S + 1) * e->K * 32 + (size_t)e->k.npairs * DSK_NUM_COLLISION_POINTS * 12;
}
static dim3 grid_block(dsk_engine* e) { return dim3(GRID_NODES, std::min(e->n_frames, MAX_FRAMES)); }
// batched engines use the throughput layout of the grid kernels (k_grid_flat / k_grid_adj_flat), single scenes the
// latency layout (node x frame)
#define GRID_FWD_LAUNCH(e, stream, ...)                                            \
  do {                                                                             \
    if ((e)->flat_grid) k_grid_flat<<<148 * (e)->flat_fwd_ctas_per_sm, FLAT_THREADS, 0, stream>>>(__VA_ARGS__); \
    else k_grid<<<grid_ctas(e), grid_block(e), 0, stream>>>(__VA_ARGS__);          \
  } while (0)
#define GRID_ADJ_LAUNCH(e, stream, sc, ...)                                            \
  do {                                                                                 \
    if ((e)->flat_grid) k_grid_adj_flat<<<148 * (e)->flat_adj_ctas_per_sm, FLAT_THREADS, 0, stream>>>(__VA_ARGS__, sc); \
    else k_grid_adj<<<grid_ctas(e), grid_block(e), 0, stream>>>(__VA_ARGS__, sc);      \
  } while (0)


// ---- launchers of the particle kernels: pick the kernel family (latency: plane-split, 3 threads per particle; throughput:
// one thread per particle), the register cap and the scatter variant (TS: transposed shared-memory scatter) -------------
static size_t ts_smem(dsk_engine* e, int threads) { return e->ts ? (size_t)(threads / 32) * TS_WARP_FLOAT4 * sizeof(float4) : 0; }
static size_t ts9_smem(dsk_engine* e) { return e->ts_pl ? (size_t)(PL_PARTICLES * 3 / 32) * TS9_WARP_FLOAT4 * sizeof(float4) : 0; }
static int launch_p2g(dsk_engine* e, bool write_f, const float* fin, float* fout, const float* mat, float4* G, TileTrack tt,
                      int q, const int* run_if, float* svd) {
  const SimConst& k = e->k;
  const int pb = e->big ? e->big_block : 128;
  const int nb = cdiv(k.stride, pb);
  const size_t sm = ts_smem(e, pb);
  if (!write_f) {
    if (e->ts) KL(KID_P2G_RECOMPUTE, k_p2g<false, 3, true><<<nb, pb, sm, e->qs>>>(k, fin, fout, mat, e->npart, G, tt, e->d_args, q, run_if, nullptr));
    else KL(KID_P2G_RECOMPUTE, k_p2g<false, 3, false><<<nb, pb, 0, e->qs>>>(k, fin, fout, mat, e->npart, G, tt, e->d_args, q, run_if, nullptr));
  } else if (e->big) {
    if (e->ts) KL(KID_P2G, k_p2g<true, 3, true><<<nb, pb, sm, e->qs>>>(k, fin, fout, mat, e->npart, G, tt, e->d_args, q, run_if, svd));
    else KL(KID_P2G, k_p2g<true, 3, false><<<nb, pb, 0, e->qs>>>(k, fin, fout, mat, e->npart, G, tt, e->d_args, q, run_if, svd));
  } else {
    if (e->ts) KL(KID_P2G, k_p2g<true, 1, true><<<nb, pb, sm, e->qs>>>(k, fin, fout, mat, e->npart, G, tt, e->d_args, q, run_if, svd));
    else KL(KID_P2G, k_p2g<true, 1, false><<<nb, pb, 0, e->qs>>>(k, fin, fout, mat, e->npart, G, tt, e->d_args, q, run_if, svd));
  }
  LAUNCH_CHECK();
  return 0;
}
static int minb_class(int minb) { return minb >= 5 ? 6 : 4; }
static int launch_g2p2g(dsk_engine* e, const float* fprev, float* fcur, float* fnext, const float* mat, const float4* Gprev,
                        float4* Gnext, TileTrack tt, int qnext, float* svd) {
  const SimConst& k = e->k;
  if (!e->big) {
    const int nb = cdiv(k.stride, PL_PARTICLES);
    if (e->ts_pl) KL(KID_G2P2G, k_g2p2g_pl<true><<<nb, dim3(PL_PARTICLES, 3), ts9_smem(e), e->qs>>>(k, fprev, fcur, fnext, mat, e->npart, Gprev, Gnext, tt, e->d_args, qnext, svd));
    else KL(KID_G2P2G, k_g2p2g_pl<false><<<nb, dim3(PL_PARTICLES, 3), 0, e->qs>>>(k, fprev, fcur, fnext, mat, e->npart, Gprev, Gnext, tt, e->d_args, qnext, svd));
  } else {
    const int pb = e->big_block, nb = cdiv(k.stride, pb);
    const size_t sm = ts_smem(e, pb);
    if (!e->ts) KL(KID_G2P2G, k_g2p2g<4, false><<<nb, pb, 0, e->qs>>>(k, fprev, fcur, fnext, mat, e->npart, Gprev, Gnext, tt, e->d_args, qnext, svd));
    else switch (minb_class(e->minb_g2p2g)) {
      case 6: KL(KID_G2P2G, k_g2p2g<6, true><<<nb, pb, sm, e->qs>>>(k, fprev, fcur, fnext, mat, e->npart, Gprev, Gnext, tt, e->d_args, qnext, svd)); break;
      default: KL(KID_G2P2G, k_g2p2g<4, true><<<nb, pb, sm, e->qs>>>(k, fprev, fcur, fnext, mat, e->npart, Gprev, Gnext, tt, e->d_args, qnext, svd)); break;
    }
  }
  LAUNCH_CHECK();
  return 0;
}
static int launch_g2p_adj(dsk_engine* e, const float* fin, const float* fnext, const float* ain, float* aout, const float4* Gv,
                          float4* Ga) {
  const SimConst& k = e->k;
  if (!e->big) {
    const int nb = cdiv(k.stride, PL_PARTICLES);
    if (e->ts_pl) KL(KID_G2P_ADJ, k_g2p_adj_pl<true><<<nb, dim3(PL_PARTICLES, 3), ts9_smem(e), e->qs>>>(k, fin, fnext, ain, aout, e->npart, Gv, Ga));
    else KL(KID_G2P_ADJ, k_g2p_adj_pl<false><<<nb, dim3(PL_PARTICLES, 3), 0, e->qs>>>(k, fin, fnext, ain, aout, e->npart, Gv, Ga));
  } else {
    const int pb = e->big_block, nb = cdiv(k.stride, pb);
    const size_t sm = ts_smem(e, pb);
    if (!e->ts) KL(KID_G2P_ADJ, k_g2p_adj<4, false><<<nb, pb, 0, e->qs>>>(k, fin, fnext, ain, aout, e->npart, Gv, Ga));
    else switch (minb_class(e->minb_g2p_adj)) {
      case 6: KL(KID_G2P_ADJ, k_g2p_adj<6, true><<<nb, pb, sm, e->qs>>>(k, fin, fnext, ain, aout, e->npart, Gv, Ga)); break;
      default: KL(KID_G2P_ADJ, k_g2p_adj<4, true><<<nb, pb, sm, e->qs>>>(k, fin, fnext, ain, aout, e->npart, Gv, Ga)); break;
    }
  }
  LAUNCH_CHECK();
  return 0;
}
static int launch_p2g_adj(dsk_engine* e, const float* fin, const float* ain, float* aout, const float* mat, const float4* Ga,
                          const float* svd) {
  const SimConst& k = e->k;
  if (!e->big) {
    KL(KID_P2G_ADJ, k_p2g_adj_pl<<<cdiv(k.stride, PL_PARTICLES), dim3(PL_PARTICLES, 3), 0, e->qs>>>(k, fin, ain, aout, mat, e->npart, Ga, svd));
  } else {
    const int pb = e->big_block, nb = cdiv(k.stride, pb);
    switch (minb_class(e->minb_p2g_adj)) {
      case 6: KL(KID_P2G_ADJ, k_p2g_adj<6><<<nb, pb, 0, e->qs>>>(k, fin, ain, aout, mat, e->npart, Ga, svd)); break;
      default: KL(KID_P2G_ADJ, k_p2g_adj<4><<<nb, pb, 0, e->qs>>>(k, fin, ain, aout, mat, e->npart, Ga, svd)); break;
    }
  }
  LAUNCH_CHECK();
  return 0;
}
// shared-memory opt-in of the TS kernel instantiations (57 KB for a 128-thread CTA)
static int ts_opt_in(dsk_engine* e) {
  const int sm = (int)(4 * TS_WARP_FLOAT4 * sizeof(float4));
  CK(cudaFuncSetAttribute(k_p2g<false, 3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
  CK(cudaFuncSetAttribute(k_p2g<true, 3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
  CK(cudaFuncSetAttribute(k_p2g<true, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
  CK(cudaFuncSetAttribute(k_g2p2g<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
  CK(cudaFuncSetAttribute(k_g2p2g<6, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
  CK(cudaFuncSetAttribute(k_g2p_adj<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
  CK(cudaFuncSetAttribute(k_g2p_adj<6, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
  (void)e;
  return 0;
}

// zero the grids of the last substep of a fine-grained (dsk_substep / dsk_substep_grad) sequence
static int flush_pending_clear(dsk_engine* e) {
  if (e->pending_q < 0) return 0;
  int q = e->pending_q, set = (q + 1) & 1;
  bool bwd = e->pending_bwd;
  KL(KID_GRID, k_end_clear<<<grid_ctas(e), GRID_CTA, 0, e->qs>>>(e->k, e->tile_list[set], e->tile_count + ((q + 1) & 3),
                                                                 e->G0[set], bwd ? e->Gv[set] : nullptr,
                                                                 bwd ? e->Ga[set] : nullptr, e->tile_count, e->done));
  LAUNCH_CHECK();
  e->pending_q = -1;
  e->grids_valid = false;
  return 0;
}

static StepArgs make_args(dsk_engine* e, int src, int dst, int action_step, int adj_step) {
  StepArgs a;
  memset(&a, 0, sizeof a);
  a.ck_src = e->frame_of(e->ckpt, src);
  a.ck_dst = e->frame_of(e->ckpt, dst);
  a.tool_src = e->tools_of(e->tool_ckpt, src);
  a.tool_dst = e->tools_of(e->tool_ckpt, dst);
  a.action = (e->A > 0 && action_step >= 0) ? e->actions + (size_t)action_step * e->B * e->A : nullptr;
  if (adj_step >= 0) {
    a.adj_in = e->frame_of(e->adj_ckpt, adj_step + 1);
    a.adj_out = e->frame_of(e->adj_ckpt, adj_step);
    a.tool_adj_in = e->tools_of(e->tool_adj_ckpt, adj_step + 1);
    a.tool_adj_out = e->tools_of(e->tool_adj_ckpt, adj_step);
    a.action_grad = e->A > 0 ? e->action_grad + (size_t)adj_step * e->B * e->A : nullptr;
  }
  e->epoch_base += 4 * ((e->S + 8) / 4);
  a.epoch_base = e->epoch_base;
  return a;
}
static int push_args(dsk_engine* e, const StepArgs& a) {
  KL(KID_IO, k_set_args<<<1, 1, 0, e->stream>>>(e->d_args, a));
  LAUNCH_CHECK();
  return 0;
}

// Frames are permuted through shared memory (k_permute_rows) when rows of one env fit there and the batch is large enough
// to fill the GPU; launch_permute returns 0 if it launched, 1 if the caller has to use the element-wise kernel, -1 on error.
// Rows per CTA: the largest of 8, 4, 2 that divides nrows and stays within 48 KB (several CTAs per SM: a batch of 64 envs
// gives 384 CTAs of 32 KB), else 1 row of up to 200 KB.
static int perm_rows_per_cta(dsk_engine* e, int nrows) {
  if (!e->perm_smem || (!e->big && !e->perm_smem_small)) return 0;
  const size_t row = (size_t)e->k.Npad * sizeof(float);
  for (int g = 8; g >= 2; g >>= 1)
    if (nrows % g == 0 && g * row <= 48 * 1024) return g;
  return row <= 200 * 1024 ? 1 : 0;
}
template <int MODE, int ROWS>
static int launch_permute_rows(dsk_engine* e, int kid, const float* in, const float* const* pin, float* out, float* const* pout,
                               const int* perm, int nrows) {
  const size_t sm = (size_t)ROWS * e->k.Npad * sizeof(float);
  if (sm > 48 * 1024 && !e->perm_opt_in[MODE]) {   // only ROWS == 1 gets here
    CK(cudaFuncSetAttribute(k_permute_rows<MODE, ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    e->perm_opt_in[MODE] = true;
  }
  KL(kid, (k_permute_rows<MODE, ROWS><<<dim3(nrows / ROWS, e->B), PERM_CTA, sm, e->qs>>>(e->k, in, pin, out, pout, e->npart, perm)));
  LAUNCH_CHECK();
  return 0;
}
template <int MODE>
static int launch_permute(dsk_engine* e, int kid, const float* in, const float* const* pin, float* out, float* const* pout,
                          const int* perm, int nrows) {
  switch (perm_rows_per_cta(e, nrows)) {
    case 8: return launch_permute_rows<MODE, 8>(e, kid, in, pin, out, pout, perm, nrows);
    case 4: return launch_permute_rows<MODE, 4>(e, kid, in, pin, out, pout, perm, nrows);
    case 2: return launch_permute_rows<MODE, 2>(e, kid, in, pin, out, pout, perm, nrows);
    case 1: return launch_permute_rows<MODE, 1>(e, kid, in, pin, out, pout, perm, nrows);
    default: return 1;
  }
}

// ---- pieces of a sequence (all enqueue on e->qs) --------------------------------------------------------------
static int seq_begin_forward(dsk_engine* e, StepSlot& s) {
  SimConst& k = e->k;
  int nb = cdiv(k.stride, 256);
  bool capturing = e->qs == e->cap_stream;
  // tool kinematics of the whole step depends on the tool state and the action only: in a captured graph it runs
  // on a parallel branch next to the sort and the first p2g, joined before the first grid_op
  const bool next_kin = capturing && e->seq_next_slot != nullptr;
  if (e->K > 0 && (!e->seq_skip_kin || next_kin)) {
    if (capturing) {
      CK(cudaEventRecord(e->ev_fork, e->qs));
      CK(cudaStreamWaitEvent(e->cap_side, e->ev_fork, 0));
      if (!e->seq_skip_kin) {
        KL(KID_KINEMATICS, k_kinematics<<<e->B, kin_block(e, false), kin_smem(e), e->cap_side>>>(k, e->d_tools, e->d_args, e->rand_num, s.poses, s.cidx));
        CK(cudaEventRecord(e->ev_join, e->cap_side));
        e->kin_join = true;
      }
      if (next_kin) {   // lookahead: this step's tool checkpoint first, then the whole kinematics of the next step
        StepSlot& nx = *e->seq_next_slot;
        const StepArgs* na = e->d_args_kin + e->seq_next_step;
        if (!e->seq_skip_kin) KL(KID_IO, k_tool_store<<<cdiv(e->B * e->K * 8, 128), 128, 0, e->cap_side>>>(k, s.poses, e->d_args));
        KL(KID_KINEMATICS, k_kinematics<<<e->B, kin_block(e, true), kin_smem(e), e->cap_side>>>(k, e->d_tools, na, e->rand_num, nx.poses, nx.cidx));
        KL(KID_IO, k_tool_store<<<cdiv(e->B * e->K * 8, 128), 128, 0, e->cap_side>>>(k, nx.poses, na));
        CK(cudaEventRecord(e->ev_join2, e->cap_side));
      }
    } else {
      KL(KID_KINEMATICS, k_kinematics<<<e->B, kin_block(e, false), kin_smem(e), e->qs>>>(k, e->d_tools, e->d_args, e->rand_num, s.poses, s.cidx));
    }
  }
  if (e->cfg.sort_particles && !e->seq_full_sort) {
    KL(KID_SORT, k_apply_perm<<<nb, 256, 0, e->qs>>>(k, e->d_args, e->mat, e->npart, e->perm_cache, s.frames, s.mat, s.perm));   // rare path (DSK_RESORT_INTERVAL > 1)
  } else {
    if (e->cfg.sort_particles) {
      // counters and chunk flags are all zero here (k_sort_clear below; zero-initialised)
      KL(KID_SORT, k_sort_bin<<<nb, 256, 0, e->qs>>>(k, e->d_args, e->npart, e->cell_count, e->key, e->rank, e->chunk_flag));
      const int chunks = cdiv(k.nnode, SCAN_CHUNK), sg = cdiv(e->B * chunks, SCAN_GROUP);
      KL(KID_SORT, k_scan_partial<<<sg, SCAN_CTA, 0, e->qs>>>(k, e->cell_count, e->scan_partial, e->chunk_flag, chunks));
      KL(KID_SORT, k_scan_chunks<<<sg, SCAN_CTA, 0, e->qs>>>(k, e->cell_count, e->scan_partial, e->chunk_flag, chunks));
    }
    bool moved = false;
    if (e->cfg.sort_particles && perm_rows_per_cta(e, FRAME_COMPS)) {
      // permutation first (ints only), then the frame -- and the material rows if there are any -- through shared memory
      KL(KID_SORT, k_sort_perm<<<nb, 256, 0, e->qs>>>(k, e->npart, e->cell_count, e->key, e->rank, s.perm));
      if (launch_permute<0>(e, KID_SORT, nullptr, &e->d_args->ck_src, s.frames, nullptr, s.perm, FRAME_COMPS) < 0) return -1;
      if (!e->mat_uniform && launch_permute<0>(e, KID_SORT, e->mat, nullptr, s.mat, nullptr, s.perm, 3) < 0) return -1;
      moved = true;
    }
    if (!moved)
      KL(KID_SORT, k_sort_scatter<<<nb, 256, 0, e->qs>>>(k, e->d_args, e->mat, e->npart, e->cell_count, e->key, e->rank,
                                                         e->cfg.sort_particles, s.frames, s.mat, s.perm));
    if (e->cfg.sort_particles) {
      const int chunks = cdiv(k.nnode, SCAN_CHUNK), sg = cdiv(e->B * chunks, SCAN_GROUP);
      KL(KID_SORT, k_sort_clear<<<sg, SCAN_CTA, 0, e->qs>>>(k, e->cell_count, e->chunk_flag, chunks));
      CK(cudaMemcpyAsync(e->perm_cache, s.perm, (size_t)k.stride * 4, cudaMemcpyDeviceToDevice, e->qs));
    }
  }
  LAUNCH_CHECK();
  return 0;
}
// forward substep: q = position in the sequence, j = substep index within the step (frames j -> j+1)
static int seq_substep(dsk_engine* e, StepSlot& s, int q, int j, bool write_state) {
  SimConst& k = e->k;
  int set = (q + 1) & 1, prev = set ^ 1;
  TileTrack tt{e->tile_epoch[set], e->tile_list[set], e->tile_count + ((q + 1) & 3)};
  const int pb = e->big ? e->big_block : 128;   // threads per CTA of the particle kernels
  int nb = cdiv(k.stride, pb);
  float* fin = s.frames + (size_t)j * e->frame_floats;
  float* fout = s.frames + (size_t)(j + 1) * e->frame_floats;
  if (launch_p2g(e, write_state, fin, fout, (e->mat_uniform ? nullptr : s.mat), e->G0[set], tt, q, nullptr, write_state ? svd_at(e, s, j) : nullptr)) return -1;
  bool clr = q > 0;
  if (e->kin_join) {
    CK(cudaStreamWaitEvent(e->qs, e->ev_join, 0));
    e->kin_join = false;
  }
  KL(KID_GRID, GRID_FWD_LAUNCH(e, e->qs,
                   k, e->grid_tools, s.poses, j, e->G0[set], e->G0[set], tt.list, tt.count,
                   clr ? e->tile_list[prev] : nullptr, e->tile_count + (q & 3), clr ? e->G0[prev] : nullptr, nullptr,
                   nullptr, e->tile_count + ((q + 3) & 3), write_state ? s.tape : GridTape{nullptr, nullptr, nullptr, nullptr, 0}, nullptr));
  if (write_state) KL(KID_G2P, k_g2p<<<nb, pb, 0, e->qs>>>(k, fin, fout, e->npart, e->G0[set]));
  LAUNCH_CHECK();
  return 0;
}
// grid_op of substep q (position q == substep index in forward sequences), with taping
static int seq_grid_fwd(dsk_engine* e, StepSlot& s, int q) {
  SimConst& k = e->k;
  int set = (q + 1) & 1, prev = set ^ 1;
  bool clr = q > 0;
  if (e->kin_join) {
    CK(cudaStreamWaitEvent(e->qs, e->ev_join, 0));
    e->kin_join = false;
  }
  KL(KID_GRID, GRID_FWD_LAUNCH(e, e->qs,
                   k, e->grid_tools, s.poses, q, e->G0[set], e->G0[set], e->tile_list[set], e->tile_count + ((q + 1) & 3),
                   clr ? e->tile_list[prev] : nullptr, e->tile_count + (q & 3), clr ? e->G0[prev] : nullptr, nullptr,
                   nullptr, e->tile_count + ((q + 3) & 3), s.tape, nullptr));
  LAUNCH_CHECK();
  return 0;
}
// forward substeps of a whole step with g2p(q) fused into p2g(q+1)
static int seq_forward_fused(dsk_engine* e, StepSlot& s) {
  SimConst& k = e->k;
  const int pb = e->big ? e->big_block : 128;   // threads per CTA of the particle kernels
  int nb = cdiv(k.stride, pb);
  auto frame = [&](int j) { return s.frames + (size_t)j * e->frame_floats; };
  {
    TileTrack tt{e->tile_epoch[1], e->tile_list[1], e->tile_count + 1};
    if (launch_p2g(e, true, frame(0), frame(1), (e->mat_uniform ? nullptr : s.mat), e->G0[1], tt, 0, nullptr, svd_at(e, s, 0))) return -1;
  }
  for (int q = 0; q < e->S; q++) {
    if (seq_grid_fwd(e, s, q)) return -1;
    int set = (q + 1) & 1, nset = set ^ 1;
    if (q + 1 < e->S) {
      TileTrack tt{e->tile_epoch[nset], e->tile_list[nset], e->tile_count + ((q + 2) & 3)};
      if (launch_g2p2g(e, frame(q), frame(q + 1), frame(q + 2), (e->mat_uniform ? nullptr : s.mat), e->G0[set], e->G0[nset], tt, q + 1, svd_at(e, s, q + 1))) return -1;
    } else {
      KL(KID_G2P, k_g2p<<<nb, pb, 0, e->qs>>>(k, frame(q), frame(q + 1), e->npart, e->G0[set]));
    }
  }
  LAUNCH_CHECK();
  return 0;
}
static int seq_end_forward(dsk_engine* e, StepSlot& s, bool store) {
  SimConst& k = e->k;
  if (store) {
    int rc = launch_permute<1>(e, KID_REORDER, s.frames + (size_t)e->S * e->frame_floats, nullptr, nullptr, &e->d_args->ck_dst,
                               s.perm, FRAME_COMPS);
    if (rc < 0) return -1;
    if (rc > 0)
      KL(KID_REORDER, k_unsort<<<cdiv(k.stride, 256), 256, 0, e->qs>>>(k, s.frames + (size_t)e->S * e->frame_floats,
                                                                      e->npart, s.perm, &e->d_args->ck_dst, 0));
    const bool next_kin = e->qs == e->cap_stream && e->seq_next_slot != nullptr && e->K > 0;
    if (e->K > 0 && !e->seq_skip_kin && !next_kin) KL(KID_IO, k_tool_store<<<cdiv(e->B * e->K * 8, 128), 128, 0, e->qs>>>(k, s.poses, e->d_args));
    if (next_kin) CK(cudaStreamWaitEvent(e->qs, e->ev_join2, 0));   // join the lookahead branch
  }
  LAUNCH_CHECK();
  return 0;
}
static int seq_clear(dsk_engine* e, int last_q, bool bwd) {
  int set = (last_q + 1) & 1;
  KL(KID_GRID, k_end_clear<<<grid_ctas(e), GRID_CTA, 0, e->qs>>>(e->k, e->tile_list[set], e->tile_count + ((last_q + 1) & 3),
                                                                 e->G0[set], bwd ? e->Gv[set] : nullptr,
                                                                 bwd ? e->Ga[set] : nullptr, e->tile_count, e->done));
  LAUNCH_CHECK();
  return 0;
}
static int seq_begin_backward(dsk_engine* e, StepSlot& s) {
  SimConst& k = e->k;
  e->bwd_cur = 0;
  {
    int rc = launch_permute<0>(e, KID_REORDER, nullptr, &e->d_args->adj_in, e->adjw[0], nullptr, s.perm, FRAME_COMPS);
    if (rc < 0) return -1;
    if (rc > 0)
      KL(KID_REORDER, k_gather_sorted<<<cdiv(k.stride, 256), 256, 0, e->qs>>>(k, &e->d_args->adj_in, e->npart, s.perm, e->adjw[0]));
  }
  if (e->K > 0)
    KL(KID_IO, k_pose_adj_init<<<cdiv(e->B * (e->S + 1) * e->K * 8, 128), 128, 0, e->qs>>>(k, e->pose_adj, e->seq_defer_tools ? nullptr : e->d_args));
  LAUNCH_CHECK();
  return 0;
}
// adjoint substep: adjoint of frame j+1 in adjw[cur] -> adjoint of frame j in adjw[cur^1]
static int seq_substep_grad(dsk_engine* e, StepSlot& s, int q, int j) {
  SimConst& k = e->k;
  int set = (q + 1) & 1, prev = set ^ 1;
  TileTrack tt{e->tile_epoch[set], e->tile_list[set], e->tile_count + ((q + 1) & 3)};
  const int pb = e->big ? e->big_block : 128;   // threads per CTA of the particle kernels
  int nb = cdiv(k.stride, pb);
  float* fin = s.frames + (size_t)j * e->frame_floats;
  float* fnext = s.frames + (size_t)(j + 1) * e->frame_floats;
  float* ain = e->adjw[e->bwd_cur];
  float* aout = e->adjw[e->bwd_cur ^ 1];
  bool clr = q > 0;
  const int* run_if = nullptr;
  if (e->seq_use_tape) {
    // complete tape: restore the grids; the two recompute kernels below then return at once
    KL(KID_GRID_RECOMPUTE, k_tape_restore<<<grid_ctas(e), GRID_NODES, 0, e->qs>>>(
                               k, s.tape, j, e->G0[set], e->Gv[set], tt.list, tt.count,
                               clr ? e->tile_list[prev] : nullptr, e->tile_count + (q & 3), clr ? e->G0[prev] : nullptr,
                               clr ? e->Gv[prev] : nullptr, clr ? e->Ga[prev] : nullptr, e->tile_count + ((q + 3) & 3)));
    run_if = s.tape.overflow;
  }
  if (!(e->seq_use_tape && e->seq_tape_trusted)) {
    if (launch_p2g(e, false, fin, nullptr, (e->mat_uniform ? nullptr : s.mat), e->G0[set], tt, q, run_if, nullptr)) return -1;
    KL(KID_GRID_RECOMPUTE, GRID_FWD_LAUNCH(e, e->qs,
                               k, e->grid_tools, s.poses, j, e->G0[set], e->Gv[set], tt.list, tt.count,
                               clr ? e->tile_list[prev] : nullptr, e->tile_count + (q & 3), clr ? e->G0[prev] : nullptr,
                               clr ? e->Gv[prev] : nullptr, clr ? e->Ga[prev] : nullptr, e->tile_count + ((q + 3) & 3),
                               GridTape{nullptr, nullptr, nullptr, nullptr, 0}, run_if));
  }
  if (launch_g2p_adj(e, fin, fnext, ain, aout, e->Gv[set], e->Ga[set])) return -1;
  {
    // same split as in the two-branch graph (the pose adjoints of the contacts are parked and reduced by a second kernel), here
    // back to back on one stream: the eager per-kernel profile of bench.py then times the kernels the graph replays
    const bool park = e->gadj_scratch[0] != nullptr;
    GridAdjScratch sc{park ? e->gadj_scratch[0] : nullptr, park ? e->gadj_flags[0] : nullptr, park ? e->gadj_cap : 0};
    KL(KID_GRID_ADJ, GRID_ADJ_LAUNCH(e, e->qs, sc, k, e->grid_tools, s.poses, j, e->G0[set], e->Ga[set], tt.list, tt.count, e->pose_adj));
    if (park)
      KL(KID_GRID_ADJ_TOOLS, k_grid_adj_tools<<<grid_ctas(e), grid_block(e), 0, e->qs>>>(k, e->grid_tools, s.poses, j, e->G0[set], tt.list,
                                                                                          tt.count, e->pose_adj, sc));
  }
  if (launch_p2g_adj(e, fin, ain, aout, (e->mat_uniform ? nullptr : s.mat), e->Ga[set], svd_at(e, s, j))) return -1;
  LAUNCH_CHECK();
  e->bwd_cur ^= 1;
  return 0;
}
static size_t kin_adj_smem(dsk_engine* e) { return ((size_t)(e->S + 1) * e->K * 8 + (size_t)e->S * e->K * 128) * 4; }
// tool adjoints of one env step from the pose adjoints the grid kernels accumulated: on e->qs with `args`
static int enqueue_tool_adjoints(dsk_engine* e, StepSlot& s, const StepArgs* args, bool seed) {
  SimConst& k = e->k;
  size_t sh = kin_adj_smem(e);
  if (sh > 48 * 1024) {
    if (sh > 200 * 1024) return fail("tool-adjoint kernel needs %zu bytes of shared memory (substeps x tools too large)", sh);
    CK(cudaFuncSetAttribute(k_kinematics_adj, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
  }
  if (seed) KL(KID_IO, k_pose_adj_seed<<<cdiv(e->B * e->K * 8, 128), 128, 0, e->qs>>>(k, e->pose_adj, args));
  KL(KID_KINEMATICS_ADJ, k_kinematics_adj<<<e->B, KINADJ_CTA, sh, e->qs>>>(k, e->d_tools, s.poses, s.cidx, e->rand_num, args, e->pose_adj));
  KL(KID_IO, k_tool_adj_accum<<<cdiv(e->B * e->K * 8, 128), 128, 0, e->qs>>>(k, e->pose_adj, args));
  LAUNCH_CHECK();
  return 0;
}
static int seq_end_backward(dsk_engine* e, StepSlot& s) {
  SimConst& k = e->k;
  if (e->K > 0 && !e->seq_defer_tools && enqueue_tool_adjoints(e, s, e->d_args, false)) return -1;
  {
    int rc = launch_permute<2>(e, KID_REORDER, e->adjw[e->bwd_cur], nullptr, nullptr, &e->d_args->adj_out, s.perm, FRAME_COMPS);
    if (rc < 0) return -1;
    if (rc > 0)
      KL(KID_REORDER, k_unsort<<<cdiv(k.stride, 256), 256, 0, e->qs>>>(k, e->adjw[e->bwd_cur], e->npart, s.perm, &e->d_args->adj_out, 1));
  }
  LAUNCH_CHECK();
  return 0;
}

// Adjoint sequence with a verified grid tape, captured as a two-branch graph: the side branch restores the grids of
// substep q (and clears the set position q-2 used) while the main branch still works on position q-1, so the
// restore leaves the critical path:   main:  g2p_adj(q) -> grid_op_adj(q) -> p2g_adj(q)
//                                     side:  clear(set of q-2) -> restore(q)      [waits for main(q-2)]
static int enqueue_backward_pipelined(dsk_engine* e, StepSlot& s) {
  SimConst& k = e->k;
  cudaStream_t mainq = e->qs, side = e->cap_side;
  if (seq_begin_backward(e, s)) return -1;
  CK(cudaEventRecord(e->ev_fork, mainq));
  CK(cudaStreamWaitEvent(side, e->ev_fork, 0));
  const int pb = e->big ? e->big_block : 128;   // threads per CTA of the particle kernels
  int nb = cdiv(k.stride, pb);
  for (int q = 0; q < e->S; q++) {
    int j = e->S - 1 - q, set = (q + 1) & 1;
    TileTrack tt{e->tile_epoch[set], e->tile_list[set], e->tile_count + ((q + 1) & 3)};
    if (q >= 2) {
      CK(cudaStreamWaitEvent(side, e->ev_main[q - 2], 0));
      if (e->gadj_scratch[0]) {   // position q-2 used this set and parked its contact adjoints
        GridAdjScratch sc{e->gadj_scratch[q & 1], e->gadj_flags[q & 1], e->gadj_cap};
        KL(KID_GRID_ADJ_TOOLS, k_grid_adj_tools<<<grid_ctas(e), grid_block(e), 0, side>>>(
                                   k, e->grid_tools, s.poses, e->S - 1 - (q - 2), e->G0[set], e->tile_list[set],
                                   e->tile_count + ((q - 1) & 3), e->pose_adj, sc));
      }
      KL(KID_GRID_RECOMPUTE, k_clear_set<<<grid_ctas(e), GRID_CTA, 0, side>>>(k, e->tile_list[set], e->tile_count + ((q - 1) & 3),
                                                                              e->G0[set], e->Gv[set], e->Ga[set]));
    }
    KL(KID_GRID_RECOMPUTE, k_tape_restore<<<grid_ctas(e), GRID_NODES, 0, side>>>(k, s.tape, j, e->G0[set], e->Gv[set], tt.list, tt.count,
                                                                                 nullptr, nullptr, nullptr, nullptr, nullptr, nullptr));
    CK(cudaEventRecord(e->ev_restored[q], side));
    CK(cudaStreamWaitEvent(mainq, e->ev_restored[q], 0));
    float* fin = s.frames + (size_t)j * e->frame_floats;
    float* fnext = s.frames + (size_t)(j + 1) * e->frame_floats;
    float* ain = e->adjw[e->bwd_cur];
    float* aout = e->adjw[e->bwd_cur ^ 1];
    if (launch_g2p_adj(e, fin, fnext, ain, aout, e->Gv[set], e->Ga[set])) return -1;   // on e->qs == mainq
    // the pose adjoints of the contacts leave the critical path: parked here, reduced on the side branch two
    // positions later (before this grid set is cleared); the last two positions do them inline
    bool park = e->gadj_scratch[0] && q + 2 < e->S;
    GridAdjScratch sc{park ? e->gadj_scratch[q & 1] : nullptr, park ? e->gadj_flags[q & 1] : nullptr, park ? e->gadj_cap : 0};
    KL(KID_GRID_ADJ, GRID_ADJ_LAUNCH(e, mainq, sc, k, e->grid_tools, s.poses, j, e->G0[set], e->Ga[set],
                                                                          tt.list, tt.count, e->pose_adj));
    if (launch_p2g_adj(e, fin, ain, aout, (e->mat_uniform ? nullptr : s.mat), e->Ga[set], svd_at(e, s, j))) return -1;
    CK(cudaEventRecord(e->ev_main[q], mainq));
    e->bwd_cur ^= 1;
  }
  LAUNCH_CHECK();
  if (seq_end_backward(e, s)) return -1;
  // both sets still hold data: positions S-2 and S-1
  if (e->S >= 2) {
    int q = e->S - 2, set = (q + 1) & 1;
    KL(KID_GRID, k_clear_set<<<grid_ctas(e), GRID_CTA, 0, mainq>>>(k, e->tile_list[set], e->tile_count + ((q + 1) & 3), e->G0[set],
                                                                   e->Gv[set], e->Ga[set]));
  }
  return seq_clear(e, e->S - 1, true);
}

// ---- whole sequences, eager or as a replayed graph ---------------------------------------------------------------
// SEQ_FWD_NOKIN: forward step whose tool kinematics (and tool checkpoint) were produced ahead of time on the lookahead
// stream by dsk_forward_steps
// as is, or (LOOK_*) with the kinematics of the NEXT step on a side branch of this step's graph
enum SeqKind { SEQ_FWD, SEQ_RECOMPUTE, SEQ_BWD, SEQ_BWD_TAPE, SEQ_BWD_TAPE_TRUSTED, SEQ_FWD_NOKIN, SEQ_FWD_LOOK_FIRST, SEQ_FWD_LOOK_MID,
               SEQ_BWD_TRUSTED_DEFER_A, SEQ_BWD_TRUSTED_DEFER_B /* tool adjoints left to tool_stream; pose_adj buffer 0 / 1 */, SEQ_KIND_COUNT };
static bool is_fwd_kind(SeqKind k) { return k == SEQ_FWD || k == SEQ_FWD_NOKIN || k == SEQ_FWD_LOOK_FIRST || k == SEQ_FWD_LOOK_MID; }
static int enqueue_sequence(dsk_engine* e, StepSlot& s, SeqKind kind) {
  if (kind == SEQ_BWD || kind == SEQ_BWD_TAPE || kind == SEQ_BWD_TAPE_TRUSTED || kind == SEQ_BWD_TRUSTED_DEFER_A ||
      kind == SEQ_BWD_TRUSTED_DEFER_B) {
    const bool defer = kind == SEQ_BWD_TRUSTED_DEFER_A || kind == SEQ_BWD_TRUSTED_DEFER_B;
    e->seq_use_tape = kind != SEQ_BWD;
    e->seq_tape_trusted = kind == SEQ_BWD_TAPE_TRUSTED || defer;
    e->seq_defer_tools = defer;
    e->pose_adj = e->pose_adj_buf[kind == SEQ_BWD_TRUSTED_DEFER_B ? 1 : 0];
    int rc;
    if (e->seq_tape_trusted && e->qs == e->cap_stream && !getenv("DSK_NO_PIPELINE")) {
      rc = enqueue_backward_pipelined(e, s);
    } else {
      rc = seq_begin_backward(e, s);
      for (int q = 0; !rc && q < e->S; q++) rc = seq_substep_grad(e, s, q, e->S - 1 - q);
      if (!rc) rc = seq_end_backward(e, s);
      if (!rc) rc = seq_clear(e, e->S - 1, true);
    }
    e->seq_defer_tools = false;
    return rc;
  }
  e->seq_skip_kin = kind == SEQ_FWD_NOKIN || kind == SEQ_FWD_LOOK_MID;
  if (!(kind == SEQ_FWD_LOOK_FIRST || kind == SEQ_FWD_LOOK_MID)) e->seq_next_slot = nullptr;
  if (seq_begin_forward(e, s)) return -1;
  if (getenv("DSK_NO_G2P2G")) {
    for (int q = 0; q < e->S; q++)
      if (seq_substep(e, s, q, q, true)) return -1;
  } else if (seq_forward_fused(e, s)) {
    return -1;
  }
  if (seq_end_forward(e, s, is_fwd_kind(kind))) return -1;
  e->seq_skip_kin = false;
  e->seq_next_slot = nullptr;
  return seq_clear(e, e->S - 1, false);
}
static int run_sequence(dsk_engine* e, int slot_idx, SeqKind kind) {
  StepSlot& s = e->slot[slot_idx];
  if (flush_pending_clear(e)) return -1;
  e->grids_valid = false;
  if (is_fwd_kind(kind) || kind == SEQ_RECOMPUTE) {
    e->seq_full_sort = !e->cfg.sort_particles || e->sort_age >= e->resort_interval;
    e->sort_age = e->seq_full_sort ? 1 : e->sort_age + 1;
  } else {
    e->seq_full_sort = false;
  }
  if (e->profiling || !e->use_graphs) {
    e->qs = e->stream;
    return enqueue_sequence(e, s, kind);
  }
  dsk_engine::GraphSet& g = e->graphs[slot_idx];
  int idx = (int)kind * 2 + (e->seq_full_sort ? 1 : 0);
  cudaGraphExec_t* ex = &g.ex[idx];
  if (!*ex) {
    int64_t l0 = e->launches;
    int64_t kl0[KID_COUNT];
    memcpy(kl0, e->kid_launches, sizeof kl0);
    CK(cudaStreamBeginCapture(e->cap_stream, cudaStreamCaptureModeThreadLocal));
    e->qs = e->cap_stream;
    int rc = enqueue_sequence(e, s, kind);
    e->qs = e->stream;
    cudaGraph_t graph = nullptr;
    cudaError_t err = cudaStreamEndCapture(e->cap_stream, &graph);
    if (rc) {
      if (graph) cudaGraphDestroy(graph);
      return -1;
    }
    if (err != cudaSuccess) return fail("cudaStreamEndCapture: %s", cudaGetErrorString(err));
    err = cudaGraphInstantiate(ex, graph, 0);
    cudaGraphDestroy(graph);
    if (err != cudaSuccess) return fail("cudaGraphInstantiate: %s", cudaGetErrorString(err));
    // launches counted during capture are the per-replay launch counts of this graph
    g.n_launch[idx] = e->launches - l0;
    for (int i = 0; i < KID_COUNT; i++) g.kid[idx][i] = e->kid_launches[i] - kl0[i];
    e->launches = l0;
    memcpy(e->kid_launches, kl0, sizeof kl0);
  }
  CK(cudaGraphLaunch(*ex, e->stream));
  e->launches += g.n_launch[idx];
  for (int i = 0; i < KID_COUNT; i++) e->kid_launches[i] += g.kid[idx][i];
  return 0;
}

extern "C" {

// ---- state ---------------------------------------------------------------------------------------------
int dsk_set_particles(dsk_engine* e, int step, int env, int n, const float* x, const float* v, const float* F,
                      const float* C, int on_device) {
  CKE(e);
  if (check_step(e, step, "dsk_set_particles")) return -1;
  if (env < 0 || env >= e->B) return fail("env %d outside [0,%d)", env, e->B);
  if (n < 0 || n > e->cfg.particle_capacity) return fail("n_particles %d exceeds capacity %d", n, e->cfg.particle_capacity);
  if (!x || !v || !F || !C) return fail("dsk_set_particles needs x, v, F, C");
  const float *dx, *dv, *dF, *dC;
  size_t o = 0;
  if (stage_in(e, x, (size_t)n * 3, on_device, o, &dx)) return -1;
  o += (size_t)n * 3;
  if (stage_in(e, v, (size_t)n * 3, on_device, o, &dv)) return -1;
  o += (size_t)n * 3;
  if (stage_in(e, F, (size_t)n * 9, on_device, o, &dF)) return -1;
  o += (size_t)n * 9;
  if (stage_in(e, C, (size_t)n * 9, on_device, o, &dC)) return -1;
  if (n > 0) {
    KL(KID_IO, k_particles_io<<<cdiv(n, 256), 256, 0, e->stream>>>(e->k, e->frame_of(e->ckpt, step), env, n, (float*)dx, (float*)dv,
                                                        (float*)dF, (float*)dC, 0));
    LAUNCH_CHECK();
  }
  e->h_npart[env] = n;
  e->sort_age = 1 << 30;
  CK(cudaMemcpyAsync(e->npart + env, &e->h_npart[env], 4, cudaMemcpyHostToDevice, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  invalidate_slots(e, step);
  return 0;
}

static int get_frame_aos(dsk_engine* e, float* frame, int env, float* x, float* v, float* F, float* C, int on_device) {
  int n = e->h_npart[env];
  if (n == 0) return 0;
  float *dx = x, *dv = v, *dF = F, *dC = C;
  size_t o = 0;
  if (!on_device) {
    dx = x ? e->stage + o : nullptr;
    o += (size_t)n * 3;
    dv = v ? e->stage + o : nullptr;
    o += (size_t)n * 3;
    dF = F ? e->stage + o : nullptr;
    o += (size_t)n * 9;
    dC = C ? e->stage + o : nullptr;
  }
  KL(KID_IO, k_particles_io<<<cdiv(n, 256), 256, 0, e->stream>>>(e->k, frame, env, n, dx, dv, dF, dC, 1));
  LAUNCH_CHECK();
  if (!on_device) {
    if (x) CK(cudaMemcpyAsync(x, dx, (size_t)n * 12, cudaMemcpyDeviceToHost, e->stream));
    if (v) CK(cudaMemcpyAsync(v, dv, (size_t)n * 12, cudaMemcpyDeviceToHost, e->stream));
    if (F) CK(cudaMemcpyAsync(F, dF, (size_t)n * 36, cudaMemcpyDeviceToHost, e->stream));
    if (C) CK(cudaMemcpyAsync(C, dC, (size_t)n * 36, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
  }
  return 0;
}

int dsk_get_particles(dsk_engine* e, int step, int env, float* x, float* v, float* F, float* C, int on_device) {
  CKE(e);
  if (check_step(e, step, "dsk_get_particles")) return -1;
  if (env < 0 || env >= e->B) return fail("env %d outside [0,%d)", env, e->B);
  return get_frame_aos(e, e->frame_of(e->ckpt, step), env, x, v, F, C, on_device);
}
int dsk_get_n_particles(dsk_engine* e, int env, int* n) {
  if (!e) return fail("null engine");
  if (env < 0 || env >= e->B) return fail("env %d outside [0,%d)", env, e->B);
  *n = e->h_npart[env];
  return 0;
}
int dsk_set_tool_state(dsk_engine* e, int step, int env, int tool, const float* st) {
  CKE(e);
  if (check_step(e, step, "dsk_set_tool_state")) return -1;
  if (env < 0 || env >= e->B || tool < 0 || tool >= e->K) return fail("bad env/tool index");
  CK(cudaMemcpyAsync(e->tools_of(e->tool_ckpt, step) + ((size_t)env * e->K + tool) * 8, st, 32, cudaMemcpyHostToDevice,
                     e->stream));
  CK(cudaStreamSynchronize(e->stream));
  invalidate_slots(e, step);
  return 0;
}
int dsk_get_tool_state(dsk_engine* e, int step, int env, int tool, float* st) {
  CKE(e);
  if (check_step(e, step, "dsk_get_tool_state")) return -1;
  if (env < 0 || env >= e->B || tool < 0 || tool >= e->K) return fail("bad env/tool index");
  CK(cudaMemcpyAsync(st, e->tools_of(e->tool_ckpt, step) + ((size_t)env * e->K + tool) * 8, 32, cudaMemcpyDeviceToHost,
                     e->stream));
  CK(cudaStreamSynchronize(e->stream));
  return 0;
}
int dsk_copy_step(dsk_engine* e, int src, int dst) {
  CKE(e);
  if (check_step(e, src, "dsk_copy_step") || check_step(e, dst, "dsk_copy_step")) return -1;
  if (src == dst) return 0;
  CK(cudaMemcpyAsync(e->frame_of(e->ckpt, dst), e->frame_of(e->ckpt, src), e->frame_floats * 4, cudaMemcpyDeviceToDevice,
                     e->stream));
  CK(cudaMemcpyAsync(e->tools_of(e->tool_ckpt, dst), e->tools_of(e->tool_ckpt, src), e->tool_floats * 4,
                     cudaMemcpyDeviceToDevice, e->stream));
  invalidate_slots(e, dst);
  return 0;
}
int dsk_set_material(dsk_engine* e, int env, const float* mu, const float* lam, const float* ys) {
  CKE(e);
  if (env < 0 || env >= e->B) return fail("env %d outside [0,%d)", env, e->B);
  int n = e->h_npart[env];
  const float* src[3] = {mu, lam, ys};
  for (int c = 0; c < 3; c++)
    if (src[c] && n > 0)
      CK(cudaMemcpyAsync(e->mat + (size_t)c * e->k.stride + (size_t)env * e->Npad, src[c], (size_t)n * 4,
                         cudaMemcpyHostToDevice, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  invalidate_all(e);
  if (e->mat_uniform) {   // from now on the kernels read the per-particle arrays (the pointer is baked into the graphs)
    e->mat_uniform = false;
    drop_graphs(e);
  }
  return 0;
}
int dsk_set_tool_param(dsk_engine* e, int tool, int which, double value) {
  CKE(e);
  if (tool < 0 || tool >= e->K) return fail("tool %d outside [0,%d)", tool, e->K);
  ToolParams& T = e->h_tools[tool];
  float v = (float)value;
  {   // unchanged value: keep the captured graphs and the resident substep tapes (TaichiEnv.set_state sets the softness
      // on every call, i.e. once per planner iteration)
    double cur = 0;
    if (dsk_get_tool_param(e, tool, which, &cur)) return -1;
    if ((float)cur == v) return 0;
  }
  switch (which) {
    case DSK_PARAM_FRICTION: T.friction = v; break;
    case DSK_PARAM_SOFTNESS: T.softness = v; break;
    case DSK_PARAM_LOWER_X: case DSK_PARAM_LOWER_Y: case DSK_PARAM_LOWER_Z: T.lo[which - DSK_PARAM_LOWER_X] = v; break;
    case DSK_PARAM_UPPER_X: case DSK_PARAM_UPPER_Y: case DSK_PARAM_UPPER_Z: T.hi[which - DSK_PARAM_UPPER_X] = v; break;
    default: return fail("unknown tool parameter %d", which);
  }
  CK(cudaStreamSynchronize(e->stream));
  CK(cudaMemcpy(e->d_tools + tool, &T, sizeof T, cudaMemcpyHostToDevice));
  sync_grid_tools(e);
  drop_graphs(e);   // the grid kernels take the tool parameters by value
  for (auto& s : e->slot) s.src_step = -1;
  return 0;
}
int dsk_get_tool_param(dsk_engine* e, int tool, int which, double* value) {
  if (!e) return fail("null engine");
  if (tool < 0 || tool >= e->K) return fail("tool %d outside [0,%d)", tool, e->K);
  const ToolParams& T = e->h_tools[tool];
  switch (which) {
    case DSK_PARAM_FRICTION: *value = T.friction; break;
    case DSK_PARAM_SOFTNESS: *value = T.softness; break;
    case DSK_PARAM_LOWER_X: case DSK_PARAM_LOWER_Y: case DSK_PARAM_LOWER_Z: *value = T.lo[which - DSK_PARAM_LOWER_X]; break;
    case DSK_PARAM_UPPER_X: case DSK_PARAM_UPPER_Y: case DSK_PARAM_UPPER_Z: *value = T.hi[which - DSK_PARAM_UPPER_X]; break;
    default: return fail("unknown tool parameter %d", which);
  }
  return 0;
}
int dsk_set_gravity(dsk_engine* e, const double* g) {
  CKE(e);
  for (int d = 0; d < 3; d++) {
    volatile float t = e->k.dt * (float)g[d];
    e->k.grav[d] = t * 30.f;
    e->cfg.gravity[d] = g[d];
  }
  invalidate_all(e);
  CK(cudaStreamSynchronize(e->stream));
  drop_graphs(e);  // SimConst is baked into the captured kernel parameters
  return 0;
}

// ---- stepping -------------------------------------------------------------------------------------------
int dsk_set_action(dsk_engine* e, int step, const float* actions, int on_device) {
  CKE(e);
  if (step < 0 || step >= e->H) return fail("dsk_set_action: step %d outside [0,%d)", step, e->H);
  if (e->A == 0) return 0;
  const float* src;
  int n = e->B * e->A;
  if (stage_in(e, actions, n, on_device, 0, &src)) return -1;
  KL(KID_IO, k_clip_actions<<<cdiv(n, 256), 256, 0, e->stream>>>(e->actions + (size_t)step * n, src, n));
  LAUNCH_CHECK();
  if (!on_device) CK(cudaStreamSynchronize(e->stream));  // the staging buffer may be reused by the next call
  for (auto& s : e->slot)
    if (s.action_step == step) s.src_step = -1;
  return 0;
}
static int forward_step_impl(dsk_engine* e, int src_step, int dst_step, int action_step, SeqKind kind) {
  if (check_step(e, src_step, "dsk_forward_step") || check_step(e, dst_step, "dsk_forward_step")) return -1;
  if (action_step >= e->H) return fail("action step %d outside [0,%d)", action_step, e->H);
  int si = src_step % e->slots;
  StepSlot& s = e->slot[si];
  if (push_args(e, make_args(e, src_step, dst_step, action_step, -1))) return -1;
  s.src_step = -1;
  s.action_step = action_step;
  if (run_sequence(e, si, kind)) return -1;
  s.tape_written = true;
  e->tape_flags_stale = true;
  invalidate_slots(e, dst_step);
  // the slot's substep frames can serve backward_step(src_step) iff the step was (src, src+?, src) shaped
  s.src_step = (dst_step != src_step && action_step == src_step) ? src_step : -1;
  e->last_fwd_frame = -1;
  e->last_substep_slot = si;
  return 0;
}
int dsk_forward_step(dsk_engine* e, int src_step, int dst_step, int action_step) {
  CKE(e);
  return forward_step_impl(e, src_step, dst_step, action_step, SEQ_FWD);
}
// join the deferred tool adjoints (dsk_backward_steps) back into the main stream
static int join_tool_stream(dsk_engine* e) {
  for (int b = 0; b < 2; b++)
    if (e->tools_pending[b]) {
      CK(cudaStreamWaitEvent(e->stream, e->ev_bwd_tools[b], 0));
      e->tools_pending[b] = false;
    }
  return 0;
}
// allow_defer: the caller (dsk_backward_steps) lets the tool adjoints of this step run on tool_stream, concurrently with
// the particle / grid adjoints of the NEXT (earlier) step: they are needed once per env step only (the pose-adjoint chain
// k_kinematics_adj, 30-210 us on one SM per env) and nothing on the particle chain depends on them
static int backward_step_impl(dsk_engine* e, int step, bool allow_defer) {
  if (step < 0 || step >= e->H) return fail("dsk_backward_step: step %d outside [0,%d)", step, e->H);
  int si = step % e->slots;
  StepSlot& s = e->slot[si];
  bool recomputed = false;
  if (s.src_step != step) {
    // per-step checkpointing: recompute the substep frames of this step from its checkpoint
    if (push_args(e, make_args(e, step, step, step, -1))) return -1;
    s.action_step = step;
    if (run_sequence(e, si, SEQ_RECOMPUTE)) return -1;
    s.src_step = step;
    s.tape_written = true;
    recomputed = true;
  }
  if (push_args(e, make_args(e, step, step, step, step))) return -1;
  SeqKind bk = SEQ_BWD;
  if (e->tape_cap > 0 && s.tape_written) {
    bk = SEQ_BWD_TAPE;
    if (!recomputed) {
      // full-tape mode: one host check of the overflow flags per backward pass lets the sequence drop its two
      // (normally empty) fallback recompute launches per substep
      if (e->tape_flags_stale) {
        CK(cudaMemcpyAsync(e->tape_overflow.data(), e->tape_flags, (size_t)e->slots * 4, cudaMemcpyDeviceToHost, e->stream));
        CK(cudaStreamSynchronize(e->stream));
        e->tape_flags_stale = false;
      }
      if (e->tape_overflow[si] == 0) bk = SEQ_BWD_TAPE_TRUSTED;
    }
  }
  const bool defer = allow_defer && bk == SEQ_BWD_TAPE_TRUSTED && e->K > 0 && e->use_graphs && !e->profiling && e->slots >= e->H;
  if (!defer) {
    if (join_tool_stream(e)) return -1;   // this step's own tool adjoints need the checkpoint an earlier deferred step writes
    if (run_sequence(e, si, bk)) return -1;
  } else {
    const int b = si & 1;
    if (e->tools_pending[b]) {   // the buffer's previous user (two steps ago) must be done with it
      CK(cudaStreamWaitEvent(e->stream, e->ev_bwd_tools[b], 0));
      e->tools_pending[b] = false;
    }
    if (!e->d_args_bwd) {   // pointers only, one entry per step: written once
      DA(e->d_args_bwd, e->H);
      int eb = e->epoch_base;
      for (int t = 0; t < e->H; t++) KL(KID_IO, k_set_args<<<1, 1, 0, e->stream>>>(e->d_args_bwd + t, make_args(e, t, t, t, t)));
      e->epoch_base = eb;   // these entries never reach a grid kernel: do not burn tile epochs
      LAUNCH_CHECK();
    }
    if (run_sequence(e, si, b ? SEQ_BWD_TRUSTED_DEFER_B : SEQ_BWD_TRUSTED_DEFER_A)) return -1;
    CK(cudaEventRecord(e->ev_bwd_main[b], e->stream));
    CK(cudaStreamWaitEvent(e->tool_stream, e->ev_bwd_main[b], 0));
    cudaStream_t keep = e->qs;
    e->qs = e->tool_stream;
    e->pose_adj = e->pose_adj_buf[b];
    int rc = enqueue_tool_adjoints(e, s, e->d_args_bwd + step, true);
    e->qs = keep;
    if (rc) return -1;
    CK(cudaEventRecord(e->ev_bwd_tools[b], e->tool_stream));
    e->tools_pending[b] = true;
  }
  e->last_bwd_frame = -1;
  e->last_substep_slot = si;
  return 0;
}
int dsk_backward_step(dsk_engine* e, int step) {
  CKE(e);
  return backward_step_impl(e, step, false);
}
int dsk_substep(dsk_engine* e, int f) {
  CKE(e);
  if (f < 0 || f >= e->H * e->S) return fail("dsk_substep: frame %d outside the horizon (%d steps x %d substeps)", f, e->H, e->S);
  int step = f / e->S, j = f % e->S;
  int si = step % e->slots;
  StepSlot& s = e->slot[si];
  e->qs = e->stream;
  if (flush_pending_clear(e)) return -1;
  if (j == 0) {
    if (push_args(e, make_args(e, step, step + 1, step, -1))) return -1;
    s.src_step = -1;
    s.action_step = step;
    e->seq_full_sort = true;
    e->sort_age = 1;
    if (seq_begin_forward(e, s)) return -1;
  } else if (e->last_fwd_frame != f - 1) {
    return fail("dsk_substep(%d): substeps of a step must run in ascending order starting at a step boundary (last was %d)", f, e->last_fwd_frame);
  } else {
    // every fine-grained substep runs at sequence position q = 0: it needs its own tile epoch (with the epoch of the
    // previous substep mark_tile would skip the tiles that one tagged and grid_op would miss them)
    if (push_args(e, make_args(e, step, step + 1, step, -1))) return -1;
  }
  // every fine-grained substep is its own one-substep sequence (q = 0) so its grids stay inspectable
  if (seq_substep(e, s, 0, j, true)) return -1;
  e->pending_q = 0;
  e->pending_bwd = false;
  e->grids_valid = true;
  e->last_was_backward = false;
  e->last_substep_slot = si;
  e->last_substep_j = j;
  e->last_fwd_frame = f;
  if (j == e->S - 1) {
    if (seq_end_forward(e, s, true)) return -1;
    invalidate_slots(e, step + 1);
    s.src_step = step;
  }
  return 0;
}
int dsk_substep_grad(dsk_engine* e, int f) {
  CKE(e);
  if (f < 0 || f >= e->H * e->S) return fail("dsk_substep_grad: frame %d outside the horizon", f);
  int step = f / e->S, j = f % e->S;
  int si = step % e->slots;
  StepSlot& s = e->slot[si];
  e->qs = e->stream;
  if (j == e->S - 1) {
    if (s.src_step != step) {
      if (push_args(e, make_args(e, step, step, step, -1))) return -1;
      s.action_step = step;
      if (run_sequence(e, si, SEQ_RECOMPUTE)) return -1;
      s.src_step = step;
    }
    if (flush_pending_clear(e)) return -1;
    if (push_args(e, make_args(e, step, step, step, step))) return -1;
    e->qs = e->stream;
    if (seq_begin_backward(e, s)) return -1;
  } else if (e->last_bwd_frame != f + 1) {
    return fail("dsk_substep_grad(%d): adjoint substeps of a step must run in descending order from the step's last substep (last was %d)", f, e->last_bwd_frame);
  } else {
    if (flush_pending_clear(e)) return -1;
    if (push_args(e, make_args(e, step, step, step, step))) return -1;   // fresh tile epoch, see dsk_substep
  }
  e->seq_use_tape = false;   // fine-grained adjoint substeps always recompute (their grids stay inspectable)
  if (seq_substep_grad(e, s, 0, j)) return -1;
  e->pending_q = 0;
  e->pending_bwd = true;
  e->grids_valid = true;
  e->last_was_backward = true;
  e->last_substep_slot = si;
  e->last_substep_j = j;
  e->last_bwd_frame = f;
  if (j == 0) {
    if (seq_end_backward(e, s)) return -1;
  }
  return 0;
}

// ---- adjoints --------------------------------------------------------------------------------------------
int dsk_zero_grad(dsk_engine* e) {
  CKE(e);
  CK(cudaMemsetAsync(e->adj_ckpt, 0, (size_t)(e->H + 1) * e->frame_floats * 4, e->stream));
  CK(cudaMemsetAsync(e->tool_adj_ckpt, 0, (size_t)(e->H + 1) * e->tool_floats * 4, e->stream));
  CK(cudaMemsetAsync(e->action_grad, 0, (size_t)e->H * e->B * std::max(1, e->A) * 4, e->stream));
  return 0;
}
int dsk_add_particle_grad(dsk_engine* e, int step, const float* gx, const float* gv, const float* gF, const float* gC,
                          int on_device) {
  CKE(e);
  if (check_step(e, step, "dsk_add_particle_grad")) return -1;
  int cap = e->cfg.particle_capacity;
  size_t n3 = (size_t)e->B * cap * 3, n9 = (size_t)e->B * cap * 9;
  const float *dx, *dv, *dF, *dC;
  size_t o = 0;
  if (stage_in(e, gx, n3, on_device, o, &dx)) return -1;
  o += gx ? n3 : 0;
  if (stage_in(e, gv, n3, on_device, o, &dv)) return -1;
  o += gv ? n3 : 0;
  if (stage_in(e, gF, n9, on_device, o, &dF)) return -1;
  o += gF ? n9 : 0;
  if (stage_in(e, gC, n9, on_device, o, &dC)) return -1;
  KL(KID_IO, k_add_particle_grad<<<cdiv(e->k.stride, 256), 256, 0, e->stream>>>(e->k, e->frame_of(e->adj_ckpt, step), e->npart, cap,
                                                                     dx, dv, dF, dC));
  LAUNCH_CHECK();
  if (!on_device) CK(cudaStreamSynchronize(e->stream));
  return 0;
}
int dsk_add_tool_grad(dsk_engine* e, int step, const float* g, int on_device) {
  CKE(e);
  if (check_step(e, step, "dsk_add_tool_grad")) return -1;
  if (e->K == 0) return 0;
  const float* d;
  if (stage_in(e, g, e->tool_floats, on_device, 0, &d)) return -1;
  KL(KID_IO, k_axpy<<<cdiv((int)e->tool_floats, 256), 256, 0, e->stream>>>(e->tools_of(e->tool_adj_ckpt, step), d, e->tool_floats));
  LAUNCH_CHECK();
  if (!on_device) CK(cudaStreamSynchronize(e->stream));
  return 0;
}
int dsk_get_particle_grad(dsk_engine* e, int step, int env, float* gx, float* gv, float* gF, float* gC, int on_device) {
  CKE(e);
  if (check_step(e, step, "dsk_get_particle_grad")) return -1;
  if (env < 0 || env >= e->B) return fail("env %d outside [0,%d)", env, e->B);
  return get_frame_aos(e, e->frame_of(e->adj_ckpt, step), env, gx, gv, gF, gC, on_device);
}
int dsk_get_tool_grad(dsk_engine* e, int step, int env, int tool, float* g8) {
  CKE(e);
  if (check_step(e, step, "dsk_get_tool_grad")) return -1;
  if (env < 0 || env >= e->B || tool < 0 || tool >= e->K) return fail("bad env/tool index");
  CK(cudaMemcpyAsync(g8, e->tools_of(e->tool_adj_ckpt, step) + ((size_t)env * e->K + tool) * 8, 32,
                     cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  return 0;
}
int dsk_scale_grad(dsk_engine* e, int step, double alpha) {
  CKE(e);
  if (check_step(e, step, "dsk_scale_grad")) return -1;
  KL(KID_IO, k_scale<<<cdiv((int)e->frame_floats, 256), 256, 0, e->stream>>>(e->frame_of(e->adj_ckpt, step), e->frame_floats, (float)alpha));
  // decay_kernel scales position.grad and rotation.grad (function.py:74-77); gap.grad is left alone
  if (e->K > 0) {
    std::vector<float> h(e->tool_floats);
    CK(cudaMemcpyAsync(h.data(), e->tools_of(e->tool_adj_ckpt, step), e->tool_floats * 4, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    for (size_t i = 0; i < e->tool_floats; i++)
      if (i % 8 != 7) h[i] *= (float)alpha;
    CK(cudaMemcpyAsync(e->tools_of(e->tool_adj_ckpt, step), h.data(), e->tool_floats * 4, cudaMemcpyHostToDevice, e->stream));
    CK(cudaStreamSynchronize(e->stream));
  }
  LAUNCH_CHECK();
  return 0;
}
int dsk_get_action_grad(dsk_engine* e, int step, float* out, int on_device) {
  CKE(e);
  if (step < 0 || step >= e->H) return fail("dsk_get_action_grad: step %d outside [0,%d)", step, e->H);
  if (e->A == 0) return 0;
  size_t n = (size_t)e->B * e->A;
  CK(cudaMemcpyAsync(out, e->action_grad + (size_t)step * n, n * 4, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, e->stream));
  if (!on_device) CK(cudaStreamSynchronize(e->stream));
  return 0;
}

// action gradients of steps [step0, step0+nsteps): [nsteps, n_envs, action_dim] in one copy
int dsk_get_action_grads(dsk_engine* e, int step0, int nsteps, float* out, int on_device) {
  CKE(e);
  if (step0 < 0 || nsteps < 0 || step0 + nsteps > e->H) return fail("dsk_get_action_grads: [%d,%d) outside [0,%d)", step0, step0 + nsteps, e->H);
  if (e->A == 0 || nsteps == 0) return 0;
  size_t n = (size_t)e->B * e->A;
  CK(cudaMemcpyAsync(out, e->action_grad + (size_t)step0 * n, n * nsteps * 4, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, e->stream));
  if (!on_device) CK(cudaStreamSynchronize(e->stream));
  return 0;
}

// ---- observations ----------------------------------------------------------------------------------------
int dsk_get_obs(dsk_engine* e, int step, float* xv, float* tools, int on_device) {
  CKE(e);
  if (check_step(e, step, "dsk_get_obs")) return -1;
  int cap = e->cfg.particle_capacity;
  size_t n = (size_t)e->B * cap * 6;
  if (xv) {
    float* d = on_device ? xv : e->stage;
    KL(KID_IO, k_get_obs<<<cdiv(e->k.stride, 256), 256, 0, e->stream>>>(e->k, e->frame_of(e->ckpt, step), e->npart, cap, d));
    LAUNCH_CHECK();
    if (!on_device) CK(cudaMemcpyAsync(xv, d, n * 4, cudaMemcpyDeviceToHost, e->stream));
  }
  if (tools && e->K > 0)
    CK(cudaMemcpyAsync(tools, e->tools_of(e->tool_ckpt, step), e->tool_floats * 4,
                       on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, e->stream));
  if (!on_device) CK(cudaStreamSynchronize(e->stream));
  return 0;
}
int dsk_min_dist_cols(dsk_engine* e, int* ncols) {
  if (!e) return fail("null engine");
  *ncols = e->ncols;
  return 0;
}
int dsk_compute_min_dist(dsk_engine* e, int step, float* out, int on_device) {
  CKE(e);
  if (check_step(e, step, "dsk_compute_min_dist")) return -1;
  if (e->ncols == 0) return 0;
  int cap = e->cfg.particle_capacity;
  size_t n = (size_t)e->B * cap * e->ncols;
  if (!on_device && n > e->stage_floats) return fail("staging buffer too small");
  float* d = on_device ? out : e->stage;
  KL(KID_IO, k_min_dist<<<cdiv(e->k.stride, 128), 128, 0, e->stream>>>(e->k, e->d_tools, e->frame_of(e->ckpt, step), e->npart,
                                                           e->tools_of(e->tool_ckpt, step), cap, e->ncols, d));
  LAUNCH_CHECK();
  if (!on_device) {
    CK(cudaMemcpyAsync(out, d, n * 4, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
  }
  return 0;
}
int dsk_compute_min_dist_grad(dsk_engine* e, int step, const float* g, int on_device) {
  CKE(e);
  if (check_step(e, step, "dsk_compute_min_dist_grad")) return -1;
  if (e->ncols == 0) return 0;
  int cap = e->cfg.particle_capacity;
  const float* d;
  if (stage_in(e, g, (size_t)e->B * cap * e->ncols, on_device, 0, &d)) return -1;
  KL(KID_IO, k_min_dist_adj<<<cdiv(e->k.stride, 128), 128, 0, e->stream>>>(e->k, e->d_tools, e->frame_of(e->ckpt, step), e->npart,
                                                               e->tools_of(e->tool_ckpt, step), cap, e->ncols, d,
                                                               e->frame_of(e->adj_ckpt, step),
                                                               e->tools_of(e->tool_adj_ckpt, step)));
  LAUNCH_CHECK();
  if (!on_device) CK(cudaStreamSynchronize(e->stream));
  return 0;
}
int dsk_compute_grid_m(dsk_engine* e, int step, float* out, int on_device) {
  CKE(e);
  if (check_step(e, step, "dsk_compute_grid_m")) return -1;
  size_t n = (size_t)e->B * e->k.nnode;
  float* d = on_device ? out : e->stage;
  CK(cudaMemsetAsync(d, 0, n * 4, e->stream));
  KL(KID_IO, k_grid_m<<<cdiv(e->k.stride, 128), 128, 0, e->stream>>>(e->k, e->frame_of(e->ckpt, step), e->npart, d));
  LAUNCH_CHECK();
  if (!on_device) {
    CK(cudaMemcpyAsync(out, d, n * 4, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
  }
  return 0;
}
int dsk_compute_grid_m_grad(dsk_engine* e, int step, const float* gm, int on_device) {
  CKE(e);
  if (check_step(e, step, "dsk_compute_grid_m_grad")) return -1;
  const float* d;
  if (stage_in(e, gm, (size_t)e->B * e->k.nnode, on_device, 0, &d)) return -1;
  KL(KID_IO, k_grid_m_adj<<<cdiv(e->k.stride, 128), 128, 0, e->stream>>>(e->k, e->frame_of(e->ckpt, step), e->npart, d,
                                                             e->frame_of(e->adj_ckpt, step)));
  LAUNCH_CHECK();
  if (!on_device) CK(cudaStreamSynchronize(e->stream));
  return 0;
}

// ---- introspection ---------------------------------------------------------------------------------------
__global__ void k_debug_cells(SimConst k, const float* __restrict__ ck, int env, int n, int* base, int* key) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  int gid = env * k.Npad + p;
  int bx, by, bz;
  int kk = cell_key(k, ck[gid], ck[k.stride + gid], ck[2 * k.stride + gid], bx, by, bz);
  base[p * 3] = bx;
  base[p * 3 + 1] = by;
  base[p * 3 + 2] = bz;
  key[p] = kk;
}
__global__ void k_debug_occupancy(SimConst k, const float* __restrict__ frame, const int* __restrict__ perm_unused,
                                  int env, int n, unsigned char* occ) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  int gid = env * k.Npad + p;
  Stencil s;
  make_stencil(k, frame[gid], frame[k.stride + gid], frame[2 * k.stride + gid], s);
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      for (int l = 0; l < 3; l++) occ[((size_t)(s.bx + i) * k.n + (s.by + j)) * k.n + (s.bz + l)] = 1;
}
__global__ void k_svd_probe(int n, const float* F, float* U, float* sg, float* V) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  M3 A, u, v;
  for (int q = 0; q < 9; q++) A.m[q] = F[i * 9 + q];
  float3 s;
  svd3(A, u, s, v);
  for (int q = 0; q < 9; q++) {
    U[i * 9 + q] = u.m[q];
    V[i * 9 + q] = v.m[q];
  }
  sg[i * 3] = s.x;
  sg[i * 3 + 1] = s.y;
  sg[i * 3 + 2] = s.z;
}

int dsk_debug_cell_index(dsk_engine* e, int step, int env, int32_t* base, int32_t* key) {
  CKE(e);
  if (check_step(e, step, "dsk_debug_cell_index")) return -1;
  if (env < 0 || env >= e->B) return fail("env %d outside [0,%d)", env, e->B);
  int n = e->h_npart[env];
  if (!n) return 0;
  int* d = (int*)e->stage;
  KL(KID_IO, k_debug_cells<<<cdiv(n, 256), 256, 0, e->stream>>>(e->k, e->frame_of(e->ckpt, step), env, n, d, d + (size_t)n * 3));
  LAUNCH_CHECK();
  if (base) CK(cudaMemcpyAsync(base, d, (size_t)n * 12, cudaMemcpyDeviceToHost, e->stream));
  if (key) CK(cudaMemcpyAsync(key, d + (size_t)n * 3, (size_t)n * 4, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  return 0;
}
int dsk_debug_sort_order(dsk_engine* e, int env, int32_t* perm) {
  CKE(e);
  if (env < 0 || env >= e->B) return fail("env %d outside [0,%d)", env, e->B);
  if (e->last_substep_slot < 0) return fail("no substep has run yet");
  int n = e->h_npart[env];
  CK(cudaMemcpyAsync(perm, e->slot[e->last_substep_slot].perm + (size_t)env * e->Npad, (size_t)n * 4,
                     cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  return 0;
}
static int grid_out(dsk_engine* e, const float4* G, int env, float* v3, float* m) {
  SimConst& k = e->k;
  float* dv = e->stage;
  float* dm = e->stage + (size_t)k.nnode * 3;
  KL(KID_IO, k_grid_to_dense<<<cdiv(k.nnode, 256), 256, 0, e->stream>>>(k, G, env, v3 ? dv : nullptr, m ? dm : nullptr, nullptr,
                                                             nullptr, 0));
  LAUNCH_CHECK();
  if (v3) CK(cudaMemcpyAsync(v3, dv, (size_t)k.nnode * 12, cudaMemcpyDeviceToHost, e->stream));
  if (m) CK(cudaMemcpyAsync(m, dm, (size_t)k.nnode * 4, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  return 0;
}
int dsk_debug_grid(dsk_engine* e, int env, float* v_in, float* v_out, float* m, uint8_t* occupied) {
  CKE(e);
  if (env < 0 || env >= e->B) return fail("env %d outside [0,%d)", env, e->B);
  if (!e->grids_valid)
    return fail("grids are only inspectable right after dsk_substep / dsk_substep_grad (whole-step calls clear them)");
  int set = (e->pending_q + 1) & 1;
  if (e->last_was_backward) {
    if (grid_out(e, e->G0[set], env, v_in, m)) return -1;
    if (v_out && grid_out(e, e->Gv[set], env, v_out, nullptr)) return -1;
  } else {
    if (v_in) return fail("v_in is overwritten in place by the forward grid kernel; run dsk_substep_grad to inspect it");
    if (grid_out(e, e->G0[set], env, v_out, m)) return -1;
  }
  if (occupied) {
    SimConst& k = e->k;
    int n = e->h_npart[env];
    unsigned char* d = (unsigned char*)e->stage;
    CK(cudaMemsetAsync(d, 0, k.nnode, e->stream));
    StepSlot& s = e->slot[e->last_substep_slot];
    if (n)
      KL(KID_IO, k_debug_occupancy<<<cdiv(n, 256), 256, 0, e->stream>>>(k, s.frames + (size_t)e->last_substep_j * e->frame_floats,
                                                             nullptr, env, n, d));
    LAUNCH_CHECK();
    CK(cudaMemcpyAsync(occupied, d, k.nnode, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
  }
  return 0;
}
int dsk_debug_grid_grad(dsk_engine* e, int env, float* g_v_in, float* g_v_out, float* g_m) {
  CKE(e);
  if (env < 0 || env >= e->B) return fail("env %d outside [0,%d)", env, e->B);
  if (!e->grids_valid || !e->last_was_backward) return fail("grid adjoints exist only right after dsk_substep_grad");
  if (g_v_out) return fail("the adjoint of grid_v_out is overwritten in place by k_grid_adj");
  return grid_out(e, e->Ga[(e->pending_q + 1) & 1], env, g_v_in, g_m);
}
int dsk_debug_frame(dsk_engine* e, int f, int env, float* x, float* v, float* F, float* C) {
  CKE(e);
  if (env < 0 || env >= e->B) return fail("env %d outside [0,%d)", env, e->B);
  if (f < 0 || f > e->H * e->S) return fail("frame %d outside the horizon", f);
  int step = f / e->S, j = f % e->S;
  if (j == 0) return dsk_get_particles(e, step, env, x, v, F, C, 0);
  StepSlot* s = &e->slot[step % e->slots];
  if (s->src_step != step && !(e->last_substep_slot == step % e->slots && e->last_fwd_frame >= f - 1 && e->last_fwd_frame / e->S == step))
    return fail("frame %d is not resident in the step-slot ring", f);
  // un-sort into a scratch frame (rewritten by the next adjoint substep anyway), then reuse the AoS reader
  SimConst& k = e->k;
  float* tmp = e->adjw[e->bwd_cur ^ 1];
  float** dptr = (float**)e->stage;  // device slot holding the destination pointer
  CK(cudaMemcpyAsync(dptr, &tmp, sizeof(float*), cudaMemcpyHostToDevice, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  KL(KID_IO, k_unsort<<<cdiv(k.stride, 256), 256, 0, e->stream>>>(k, s->frames + (size_t)j * e->frame_floats, e->npart, s->perm, dptr, 0));
  LAUNCH_CHECK();
  CK(cudaStreamSynchronize(e->stream));
  return get_frame_aos(e, tmp, env, x, v, F, C, 0);
}
int dsk_debug_tool_frame(dsk_engine* e, int f, int env, int tool, float* st, int32_t* cidx) {
  CKE(e);
  if (env < 0 || env >= e->B || tool < 0 || tool >= e->K) return fail("bad env/tool index");
  int step = f / e->S, j = f % e->S;
  if (j == 0 && step >= 1) {
    step -= 1;
    j = e->S;
  }
  StepSlot& s = e->slot[step % e->slots];
  if (st)
    CK(cudaMemcpyAsync(st, s.poses + (((size_t)env * (e->S + 1) + j) * e->K + tool) * 8, 32, cudaMemcpyDeviceToHost, e->stream));
  if (cidx && e->k.npairs > 0)
    CK(cudaMemcpyAsync(cidx, s.cidx + ((size_t)env * (e->S + 1) + j) * e->k.npairs, (size_t)e->k.npairs * 4,
                       cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  return 0;
}
int dsk_debug_tool_frame_grad(dsk_engine* e, int f, int env, int tool, float* g8) {
  CKE(e);
  if (env < 0 || env >= e->B || tool < 0 || tool >= e->K) return fail("bad env/tool index");
  int j = f % e->S;
  if (j == 0 && f > 0 && e->last_bwd_frame != f) j = e->S;
  CK(cudaMemcpyAsync(g8, e->pose_adj + (((size_t)env * (e->S + 1) + j) * e->K + tool) * 8, 32, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  return 0;
}
int dsk_debug_svd(dsk_engine* e, int n, const float* F, float* U, float* sig, float* V) {
  CKE(e);
  if ((size_t)n * 30 > e->stage_floats) return fail("too many matrices for the staging buffer");
  float* d = e->stage;
  CK(cudaMemcpyAsync(d, F, (size_t)n * 36, cudaMemcpyHostToDevice, e->stream));
  KL(KID_IO, k_svd_probe<<<cdiv(n, 128), 128, 0, e->stream>>>(n, d, d + (size_t)n * 9, d + (size_t)n * 18, d + (size_t)n * 21));
  LAUNCH_CHECK();
  CK(cudaMemcpyAsync(U, d + (size_t)n * 9, (size_t)n * 36, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaMemcpyAsync(sig, d + (size_t)n * 18, (size_t)n * 12, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaMemcpyAsync(V, d + (size_t)n * 21, (size_t)n * 36, cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  return 0;
}

// ---- measurement helpers ------------------------------------------------------------------------------------
int dsk_set_graphs(dsk_engine* e, int on) {
  CKE(e);
  e->use_graphs = on != 0;
  return 0;
}
int dsk_profile_enable(dsk_engine* e, int on) {
  CKE(e);
  CK(cudaStreamSynchronize(e->stream));
  e->profiling = on != 0;
  return 0;
}
// In-graph timeline (profiling build only; the product library reports "unavailable").
// enable: drops the cached graphs so that the next capture assigns one record per launch; reset: clears the
// stamps but keeps the assignment, so a following graph replay fills exactly its own records.
int dsk_timeline_enable(dsk_engine* e, int on) {
  CKE(e);
#ifdef DSK_TIMELINE
  CK(cudaStreamSynchronize(e->stream));
  if (!e->d_tl) CK(cudaMalloc(&e->d_tl, sizeof(TlRec) * e->tl_cap));
  e->tl_on = on != 0;
  e->tl_next = 0;
  e->tl_kid.clear();
  e->k.tl = e->d_tl;
  e->k.tl_slot = -1;
  drop_graphs(e);
  return dsk_timeline_reset(e);
#else
  (void)on;
  return fail("timeline: this library was built without -DDSK_TIMELINE");
#endif
}
int dsk_timeline_reset(dsk_engine* e) {
  CKE(e);
#ifdef DSK_TIMELINE
  CK(cudaStreamSynchronize(e->stream));
  std::vector<TlRec> z(e->tl_cap, TlRec{~0ull, 0ull});
  CK(cudaMemcpy(e->d_tl, z.data(), sizeof(TlRec) * e->tl_cap, cudaMemcpyHostToDevice));
  return 0;
#else
  return fail("timeline: this library was built without -DDSK_TIMELINE");
#endif
}
// returns the number of records written to (kid, t0_ns, t1_ns); records never stamped have t1 == 0
int dsk_timeline_read(dsk_engine* e, int* kid, unsigned long long* t0, unsigned long long* t1, int cap) {
  CKE(e);
#ifdef DSK_TIMELINE
  CK(cudaStreamSynchronize(e->stream));
  int n = std::min(cap, e->tl_next);
  std::vector<TlRec> r(n);
  if (n) CK(cudaMemcpy(r.data(), e->d_tl, sizeof(TlRec) * n, cudaMemcpyDeviceToHost));
  for (int i = 0; i < n; i++) {
    kid[i] = e->tl_kid[i];
    t0[i] = r[i].t0;
    t1[i] = r[i].t1;
  }
  return n;
#else
  (void)kid; (void)t0; (void)t1; (void)cap;
  return fail("timeline: this library was built without -DDSK_TIMELINE");
#endif
}
int dsk_kernel_class_count(void) { return KID_COUNT; }
const char* dsk_kernel_class_name(int i) { return (i >= 0 && i < KID_COUNT) ? kKernelNames[i] : ""; }
int dsk_profile_report(dsk_engine* e, double* ms, int64_t* launches, int n, int reset) {
  CKE(e);
  CK(cudaStreamSynchronize(e->stream));
  for (int i = 0; i < n; i++) {
    if (ms) ms[i] = 0.0;
    if (launches) launches[i] = 0;
  }
  for (auto& r : e->prof) {
    float t = 0.f;
    CK(cudaEventElapsedTime(&t, r.a, r.b));
    if (r.kid < n) {
      if (ms) ms[r.kid] += t;
      if (launches) launches[r.kid] += 1;
    }
  }
  if (reset) {
    for (auto& r : e->prof) {
      e->ev_pool.push_back(r.a);
      e->ev_pool.push_back(r.b);
    }
    e->prof.clear();
  }
  return 0;
}
int dsk_launch_counts(dsk_engine* e, int64_t* per_class, int n) {
  if (!e) return fail("null engine");
  for (int i = 0; i < n && i < KID_COUNT; i++) per_class[i] = e->kid_launches[i];
  return 0;
}
// deterministic stand-in for the reference's torch-side losses (taichi_env.py:246-275 needs geomloss):
// loss[env] += weight * mean_p |x_p - target_p|^2 at checkpoint `step`, gradient += into the adjoint checkpoint
int dsk_loss_reset(dsk_engine* e) {
  CKE(e);
  CK(cudaMemsetAsync(e->loss, 0, (size_t)e->B * 4, e->stream));
  return 0;
}
int dsk_loss_add_l2(dsk_engine* e, int step, const float* target, double weight, int on_device) {
  CKE(e);
  if (check_step(e, step, "dsk_loss_add_l2")) return -1;
  int cap = e->cfg.particle_capacity;
  const float* d;
  if (stage_in(e, target, (size_t)e->B * cap * 3, on_device, 0, &d)) return -1;
  KL(KID_LOSS, k_loss_l2<<<cdiv(e->k.stride, 128), 128, 0, e->stream>>>(e->k, e->frame_of(e->ckpt, step), e->frame_of(e->adj_ckpt, step),
                                                                        e->npart, d, cap, (float)weight, e->loss));
  LAUNCH_CHECK();
  if (!on_device) CK(cudaStreamSynchronize(e->stream));
  return 0;
}
/* ---- multi-step calls: the planner's fast path -- one host call instead of one per env step ---- */
int dsk_set_actions(dsk_engine* e, int step0, int nsteps, const float* actions, int on_device) {
  CKE(e);
  if (nsteps <= 0) return 0;
  if (step0 < 0 || step0 + nsteps > e->H) return fail("dsk_set_actions: steps [%d,%d) outside [0,%d)", step0, step0 + nsteps, e->H);
  if (e->A == 0) return 0;
  size_t n = (size_t)nsteps * e->B * e->A;
  const float* src;
  if (stage_in(e, actions, n, on_device, 0, &src)) return -1;
  KL(KID_IO, k_clip_actions<<<cdiv((int)n, 256), 256, 0, e->stream>>>(e->actions + (size_t)step0 * e->B * e->A, src, (int)n));
  LAUNCH_CHECK();
  if (!on_device) CK(cudaStreamSynchronize(e->stream));  // the staging buffer may be reused by the next call
  for (auto& s : e->slot)
    if (s.action_step >= step0 && s.action_step < step0 + nsteps) s.src_step = -1;
  return 0;
}
// The tool trajectory never depends on the dough, and here the actions of all the steps are known up front: the
// kinematics of every step (pose chain, tool-tool collision projections, tool checkpoints) runs ahead on its own
// stream while the main stream simulates earlier steps, so it leaves the critical path even when a collision forces
// the reference's sequential per-substep procedure.  Needs one step slot per env step (poses live in the slot).
int dsk_forward_steps(dsk_engine* e, int step0, int nsteps) {
  CKE(e);
  if (nsteps <= 0) return 0;
  if (step0 < 0 || step0 + nsteps > e->H) return fail("dsk_forward_steps: steps [%d,%d) outside [0,%d)", step0, step0 + nsteps, e->H);
  bool look = e->K > 0 && e->use_graphs && !e->profiling && e->slots >= e->H && nsteps >= 2 && !getenv("DSK_NO_KIN_LOOKAHEAD");
  if (!look) {
    for (int s = step0; s < step0 + nsteps; s++)
      if (forward_step_impl(e, s, s + 1, s, SEQ_FWD)) return -1;
    return 0;
  }
  if (!e->d_args_kin) {   // pointers only, one entry per step: written once
    DA(e->d_args_kin, e->H);
    CK(cudaEventCreateWithFlags(&e->ev_join2, cudaEventDisableTiming));
    for (int s = 0; s < e->H; s++) KL(KID_IO, k_set_args<<<1, 1, 0, e->stream>>>(e->d_args_kin + s, make_args(e, s, s + 1, s, -1)));
    LAUNCH_CHECK();
  }
  const int last = step0 + nsteps - 1;
  for (int s = step0; s <= last; s++) {
    SeqKind kind = s == last ? SEQ_FWD_NOKIN : (s == step0 ? SEQ_FWD_LOOK_FIRST : SEQ_FWD_LOOK_MID);
    if (s < last) {
      e->seq_next_slot = &e->slot[(s + 1) % e->slots];
      e->seq_next_step = s + 1;
    }
    int rc = forward_step_impl(e, s, s + 1, s, kind);
    e->seq_next_slot = nullptr;
    if (rc) return -1;
  }
  return 0;
}
int dsk_backward_steps(dsk_engine* e, int step_hi, int nsteps) {
  CKE(e);
  // worth its two events per step only where the pose-adjoint chain is long: many envs per launch or tool-tool collision
  // pairs (r02e: GatherMove x64 113.3 -> 105.1 ms, x8 50.1 -> 48.2 ms; LiftSpread, one env and one pair, 42.5 -> 43.4 ms)
  bool defer = nsteps >= 2 && e->B * std::max(1, e->k.npairs) >= 8;
  if (const char* v = getenv("DSK_TOOL_DEFER")) defer = nsteps >= 2 && atoi(v) != 0;
  for (int s = step_hi; s > step_hi - nsteps; s--)
    if (backward_step_impl(e, s, defer)) {
      join_tool_stream(e);
      return -1;
    }
  return join_tool_stream(e);   // action gradients and tool adjoints are complete when the call returns (stream order)
}
int dsk_loss_add_l2_steps(dsk_engine* e, int step0, int nsteps, const float* target, double weight, int on_device) {
  CKE(e);
  if (nsteps <= 0) return 0;
  if (check_step(e, step0, "dsk_loss_add_l2_steps") || check_step(e, step0 + nsteps - 1, "dsk_loss_add_l2_steps")) return -1;
  int cap = e->cfg.particle_capacity;
  const float* d;
  if (stage_in(e, target, (size_t)e->B * cap * 3, on_device, 0, &d)) return -1;
  for (int step = step0; step < step0 + nsteps; step++)
    KL(KID_LOSS, k_loss_l2<<<cdiv(e->k.stride, 128), 128, 0, e->stream>>>(e->k, e->frame_of(e->ckpt, step), e->frame_of(e->adj_ckpt, step),
                                                                          e->npart, d, cap, (float)weight, e->loss));
  LAUNCH_CHECK();
  if (!on_device) CK(cudaStreamSynchronize(e->stream));
  return 0;
}
int dsk_loss_get(dsk_engine* e, float* out, int on_device) {
  CKE(e);
  CK(cudaMemcpyAsync(out, e->loss, (size_t)e->B * 4, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, e->stream));
  if (!on_device) CK(cudaStreamSynchronize(e->stream));
  return 0;
}

}  // extern "C"
