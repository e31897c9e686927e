/* include/diffskill_mpm.h
 *
 * C ABI of the B200-native differentiable MLS-MPM engine (libdiffskill_mpm.so).
 *
 * The reference exposes this path as Python classes over Taichi fields, not as a
 * C ABI.  The entry points below are what a ctypes
 * binding inside the reference's own classes would call; each one cites the
 * reference interface it replaces.  No torch types cross this boundary: tensors
 * are raw pointers (`on_device` = 0: host memory, 1: device memory on the
 * engine's GPU), sizes are plain ints.  Every call returns 0 on success and a
 * negative code on failure; dsk_last_error() returns the message.  All work is
 * enqueued on the engine's stream (dsk_set_stream); calls that return data to
 * host memory synchronise that stream before returning.
 *
 * Batching: one engine simulates `n_envs` independent environments that share
 * one scene description (grid, material model, tool set) and differ in particle
 * state, tool state and actions.  `n_envs = 1` is the reference's single
 * environment.  Arrays carry the env index as their leading dimension.
 *
 * Frames: the reference indexes a 1024-frame tape by substep `f`
 * (mpm_simulator.py:40-43).  The engine keeps *checkpoints* at env-step
 * boundaries (`step s` == reference frame `s * substeps`) plus a ring of step
 * slots holding the substep frames of recently simulated steps; adjoints are
 * kept per checkpoint and accumulate (`+=`) exactly like the reference's
 * `x.grad[f]` fields until dsk_zero_grad.
 */
#ifndef DIFFSKILL_MPM_H
#define DIFFSKILL_MPM_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DSK_MAX_TOOLS 8
#define DSK_MAX_PAIRS 8
#define DSK_NUM_COLLISION_POINTS 600 /* mpm_simulator.py:59 */
#define DSK_ABI_VERSION 1

enum dsk_tool_type {
  DSK_TOOL_CAPSULE = 0,        /* primitives.py:43, base forward_kinematics primive_base.py:152 */
  DSK_TOOL_ROLLINGPIN_EXT = 1, /* primitives.py:120 */
  DSK_TOOL_BOX = 2,            /* primitives.py:359 */
  DSK_TOOL_GRIPPER = 3,        /* primitives.py:428 */
  DSK_TOOL_KNIFE = 4,          /* primitives.py:740 (Prism :700 + Box) */
  DSK_TOOL_SPHERE = 5,         /* primitives.py:23 (legacy PlasticineLab tool; radius in `r`) */
  DSK_TOOL_ROLLINGPIN = 6,     /* primitives.py:101 (capsule; RollingPinExt kinematics without the w[0] slide) */
  DSK_TOOL_GRIPPER2 = 7,       /* primitives.py:576 (Gripper with capsule jaws `h`, `r`) */
  DSK_TOOL_CYLINDER = 8,       /* primitives.py:302 (`h` = radial, `r` = axial half extent, as the reference names them) */
  DSK_TOOL_TORUS = 9,          /* primitives.py:337 (major radius cfg.tx in `h`, minor radius cfg.ty in `r`) */
  DSK_TOOL_CHOPSTICKS = 10     /* primitives.py:218 (two capsules `h`, `r` at -+gap/2 inside one tool frame; 8-float state, 7-D action) */
};

enum dsk_tool_param {
  DSK_PARAM_FRICTION = 0, /* prim.friction[None] = v */
  DSK_PARAM_SOFTNESS = 1, /* Primitives.set_softness, primitives.py:877-879 */
  DSK_PARAM_LOWER_X = 2, DSK_PARAM_LOWER_Y = 3, DSK_PARAM_LOWER_Z = 4, /* prim.xyz_limit[0] */
  DSK_PARAM_UPPER_X = 5, DSK_PARAM_UPPER_Y = 6, DSK_PARAM_UPPER_Z = 7  /* prim.xyz_limit[1] */
};

/* One rigid tool; the fields of Primitive.default_config and its subclasses. */
typedef struct dsk_tool_desc {
  int32_t type;            /* dsk_tool_type */
  int32_t action_dim;      /* cfg.action.dim : 0, 3, 6 or 7 */
  double action_scale[8];  /* cfg.action.scale */
  double friction;         /* cfg.friction */
  double softness;         /* GradModel softness (function.py:12), 666 by default */
  double lower_bound[3];   /* cfg.lower_bound -> xyz_limit[0] */
  double upper_bound[3];   /* cfg.upper_bound -> xyz_limit[1] */
  double size[3];          /* Box / Gripper / Knife.box half extents */
  double h, r;             /* Capsule / RollingPin(Ext) / Gripper2 jaws; Sphere: r = cfg.radius; Cylinder; Torus: (tx, ty) */
  double prism_h[2];       /* Knife.prism.h */
  double prot[4];          /* Knife.prism.prot */
  double minimal_gap, maximal_gap; /* Gripper, Gripper2 */
} dsk_tool_desc;

/* Scene + capacity.  Derived constants are passed as the python doubles the
 * reference computes in MPMSimulator.__init__ (mpm_simulator.py:19-39); the
 * engine rounds them to fp32 where Taichi does. */
typedef struct dsk_config {
  int32_t abi_version;       /* DSK_ABI_VERSION */
  int32_t device;            /* CUDA device ordinal */
  int32_t n_envs;            /* B */
  int32_t particle_capacity; /* max particles per env (cfg.n_particles) */
  int32_t n_grid;            /* multiple of 4 */
  int32_t substeps;          /* int(2e-3 // dt) */
  int32_t max_steps;         /* env-step horizon H: checkpoints 0..H are kept */
  int32_t step_slots;        /* substep-frame ring size in env steps (1 = pure checkpointing, H = full tape) */
  int32_t sort_particles;    /* 1: counting-sort particles by cell every env step */
  int32_t grid_tape_mib;     /* device memory budget (MiB, all step slots) for taping the active grid tiles of every substep so
                              * that the adjoint skips the p2g + grid_op recompute of mpm_simulator.py:330-333; 0 = off */
  double dt, dx, inv_dx, p_vol, p_mass;
  double mu, lam, yield_stress; /* initial per-particle fill, mpm_simulator.py:85-87 */
  double gravity[3];
  double ground_friction;
  double lower_bound;
  int32_t n_tools;
  int32_t n_pairs;
  dsk_tool_desc tools[DSK_MAX_TOOLS];
  int32_t pairs[DSK_MAX_PAIRS][2]; /* (moved tool i, obstacle box j), mpm_simulator.py:69-80 */
} dsk_config;

typedef struct dsk_engine dsk_engine;

const char* dsk_last_error(void);
int dsk_abi_version(void);
/* struct sizes, so that foreign-function bindings can verify their layout without a GPU */
int dsk_sizeof_config(void);
int dsk_sizeof_tool_desc(void);

/* MPMSimulator.__init__ + Primitives.__init__ + initialize (mpm_simulator.py:8-97, primitives.py:820-838) */
int dsk_create(const dsk_config* cfg, dsk_engine** out);
int dsk_destroy(dsk_engine* e);
/* cudaStream_t to enqueue on (0 = legacy default stream).  torch: torch.cuda.current_stream().cuda_stream */
int dsk_set_stream(dsk_engine* e, void* cuda_stream);
int dsk_synchronize(dsk_engine* e);
/* rand_num of mpm_simulator.py:89-97, [n_pairs,600,3] doubles (host) */
int dsk_set_rand_num(dsk_engine* e, const double* rand_num);

/* ---- particle / tool state at an env-step boundary ------------------------------------------- */
/* MPMSimulator.set_state / setframe (mpm_simulator.py:358-366,392-396): sets n_particles of env `env`.
 * x,v:[n,3]  F,C:[n,3,3] fp32, particle-major (the reference's layout). */
int dsk_set_particles(dsk_engine* e, int step, int env, int n, const float* x, const float* v, const float* F,
                      const float* C, int on_device);
/* MPMSimulator.get_state / readframe / get_x / get_v (mpm_simulator.py:348-356,380-390,419-438); any pointer may be NULL */
int dsk_get_particles(dsk_engine* e, int step, int env, float* x, float* v, float* F, float* C, int on_device);
int dsk_get_n_particles(dsk_engine* e, int env, int* n);
/* Primitive.set_state / get_state (primive_base.py:164-191, primitives.py:542-553): state8 = pos3, quat4 (w,x,y,z), gap */
int dsk_set_tool_state(dsk_engine* e, int step, int env, int tool, const float* state8);
int dsk_get_tool_state(dsk_engine* e, int step, int env, int tool, float* state8);
/* MPMSimulator.copyframe (mpm_simulator.py:368-378): particles + tool poses of all envs */
int dsk_copy_step(dsk_engine* e, int src_step, int dst_step);
/* sim.mu / lam / yield_stress fields (mpm_simulator.py:34-36); NULL leaves a field unchanged; [n] host floats */
int dsk_set_material(dsk_engine* e, int env, const float* mu, const float* lam, const float* yield_stress);
/* prim.friction[None], prim.softness[None], prim.xyz_limit (primive_base.py:34-35,45) */
int dsk_set_tool_param(dsk_engine* e, int tool, int which, double value);
int dsk_get_tool_param(dsk_engine* e, int tool, int which, double* value);
int dsk_set_gravity(dsk_engine* e, const double* g3);

/* ---- stepping --------------------------------------------------------------------------------- */
/* Primitives.set_action (primitives.py:863-867; clip to [-1,1], per-tool slices, set_velocity
 * primive_base.py:260-274).  actions: [n_envs, action_dim] fp32. */
int dsk_set_action(dsk_engine* e, int step, const float* actions, int on_device);
/* GradModel.forward_step without the set_action (function.py:174-175): `substeps` x MPMSimulator.substep
 * from checkpoint src_step into checkpoint dst_step, using the action stored for action_step.
 * Reference uses: (s, s+1, s) in gradient mode, (0, 0, 0) in copy mode (mpm_simulator.py:440-451). */
int dsk_forward_step(dsk_engine* e, int src_step, int dst_step, int action_step);
/* GradModel.backward_step (function.py:177-190): substep_grad over the step in reverse, set_velocity.grad.
 * Reads adjoint checkpoint step+1, accumulates into adjoint checkpoint `step` and the action gradient of `step`. */
int dsk_backward_step(dsk_engine* e, int step);
/* MPMSimulator.substep(f) / substep_grad(f) (mpm_simulator.py:307-345) for callers that drive single
 * substeps; f must follow the reference's access pattern (ascending within a step forward, descending backward). */
int dsk_substep(dsk_engine* e, int f);
int dsk_substep_grad(dsk_engine* e, int f);

/* ---- adjoints --------------------------------------------------------------------------------- */
/* GradModel.reset(clear_grad=True) (function.py:44-61) */
int dsk_zero_grad(dsk_engine* e);
/* GradModel._set_obs_grad (function.py:130-142): += into x.grad, v.grad [n_envs,cap,3] (rows >= n ignored) and
 * tool pose grads [n_envs,n_tools,8]; F/C adjoints are optional extras (NULL in the reference's use). */
int dsk_add_particle_grad(dsk_engine* e, int step, const float* gx, const float* gv, const float* gF,
                          const float* gC, int on_device);
int dsk_add_tool_grad(dsk_engine* e, int step, const float* gtool, int on_device);
/* x.grad[f] etc. at a step boundary ([n,3],[n,3],[n,3,3],[n,3,3] of one env; NULL skips) */
int dsk_get_particle_grad(dsk_engine* e, int step, int env, float* gx, float* gv, float* gF, float* gC,
                          int on_device);
int dsk_get_tool_grad(dsk_engine* e, int step, int env, int tool, float* g8);
/* GradModel.decay_kernel (function.py:66-77) */
int dsk_scale_grad(dsk_engine* e, int step, double alpha);
/* Primitive.get_action_grad(s, 1) for every tool, concatenated (primive_base.py:254-258,276-282): [n_envs, action_dim] */
int dsk_get_action_grad(dsk_engine* e, int step, float* out, int on_device);
/* the same for steps [step0, step0+nsteps) in one copy: [nsteps, n_envs, action_dim] (Primitives.get_grad(n), primitives.py:869-875) */
int dsk_get_action_grads(dsk_engine* e, int step0, int nsteps, float* out, int on_device);

/* ---- observations ----------------------------------------------------------------------------- */
/* GradModel._get_obs (function.py:90-102): xv [n_envs,cap,6] (x then v), tools [n_envs,n_tools,8] */
int dsk_get_obs(dsk_engine* e, int step, float* xv, float* tools, int on_device);
/* GradModel.compute_min_dist (+.grad) (function.py:79-88,154-159): [n_envs,cap,ncols]; ncols = sum(2 if gripper else 1) */
int dsk_min_dist_cols(dsk_engine* e, int* ncols);
int dsk_compute_min_dist(dsk_engine* e, int step, float* out, int on_device);
int dsk_compute_min_dist_grad(dsk_engine* e, int step, const float* gdist, int on_device);
/* MPMSimulator.compute_grid_m_kernel (+.grad) (mpm_simulator.py:456-471): [n_envs,n,n,n] */
int dsk_compute_grid_m(dsk_engine* e, int step, float* out, int on_device);
int dsk_compute_grid_m_grad(dsk_engine* e, int step, const float* gm, int on_device);

/* ---- introspection (parity tests, profiling) ---------------------------------------------------- */
/* Integer work of the path: base cell [n,3], sort key [n] (tile-major cell key) and the permutation in use
 * (sorted slot -> particle id) for env `env` at checkpoint `step`. */
int dsk_debug_cell_index(dsk_engine* e, int step, int env, int32_t* base, int32_t* key);
int dsk_debug_sort_order(dsk_engine* e, int env, int32_t* perm);
/* Dense [n,n,n,(3|1)] copies of the grids of the most recent substep / substep_grad of env `env`:
 * v_in, v_out [n,n,n,3]; m [n,n,n]; occupancy = tiles' nodes touched by a stencil (uint8). NULL skips. */
int dsk_debug_grid(dsk_engine* e, int env, float* v_in, float* v_out, float* m, uint8_t* occupied);
int dsk_debug_grid_grad(dsk_engine* e, int env, float* g_v_in, float* g_v_out, float* g_m);
/* particle / tool state at substep frame f (must still be in the slot ring) */
int dsk_debug_frame(dsk_engine* e, int f, int env, float* x, float* v, float* F, float* C);
int dsk_debug_tool_frame(dsk_engine* e, int f, int env, int tool, float* state8, int32_t* collision_idx);
int dsk_debug_tool_frame_grad(dsk_engine* e, int f, int env, int tool, float* g8);
/* device SVD probe: F [n,9] -> U,sig,V (host pointers) */
int dsk_debug_svd(dsk_engine* e, int n, const float* F, float* U, float* sig, float* V);
/* ---- measurement helpers (bench.py) ---------------------------------------------------------------------- */
/* Whole-step calls replay one captured CUDA graph per step slot (default on; DSK_NO_GRAPHS=1 disables). */
int dsk_set_graphs(dsk_engine* e, int on);
/* Profiling mode: whole-step calls launch eagerly and every kernel is bracketed by CUDA events on the engine's
 * stream; dsk_profile_report sums milliseconds / launches per kernel class (dsk_kernel_class_name). */
int dsk_profile_enable(dsk_engine* e, int on);
int dsk_kernel_class_count(void);
const char* dsk_kernel_class_name(int i);
int dsk_profile_report(dsk_engine* e, double* ms, int64_t* launches, int n, int reset);
int dsk_launch_counts(dsk_engine* e, int64_t* per_class, int n);
/* In-graph timeline: per-launch first/last %globaltimer stamps (ns), also inside replayed CUDA graphs.  Only the
 * profiling build (python -m diffskill_b200.build --timeline -> libdiffskill_mpm_tl.so, selected with
 * DSK_LIB=timeline) implements it; the product library returns an error.  enable drops the cached graphs so the
 * next capture assigns one record per launch; reset clears the stamps; read returns the number of records. */
int dsk_timeline_enable(dsk_engine* e, int on);
int dsk_timeline_reset(dsk_engine* e);
int dsk_timeline_read(dsk_engine* e, int* kid, unsigned long long* t0_ns, unsigned long long* t1_ns, int cap);
/* Deterministic stand-in for the reference's torch-side losses (taichi_env.py:246-275 needs geomloss):
 * loss[env] += weight * mean_p |x_p - target_p|^2 at checkpoint `step`; its gradient is added to the adjoint
 * checkpoint.  target: [n_envs, particle_capacity, 3]; dsk_loss_get: [n_envs]. */
int dsk_loss_reset(dsk_engine* e);
int dsk_loss_add_l2(dsk_engine* e, int step, const float* target, double weight, int on_device);
int dsk_loss_get(dsk_engine* e, float* out, int on_device);
/* Multi-step calls -- the batched fast path under the planner loops (Solver.solve_one_plan, plb/optimizer/solver.py:111-127;
 * plb/cut/solve_func.solve :74-105): the same work as the per-step calls above, issued by ONE host call each, so a whole
 * rollout is ~6 calls instead of ~5 per env step.  actions: [nsteps, n_envs, action_dim]; the loss is added at checkpoints
 * step0 .. step0+nsteps-1; backward runs steps step_hi, step_hi-1, ... (nsteps of them). */
int dsk_set_actions(dsk_engine* e, int step0, int nsteps, const float* actions, int on_device);
int dsk_forward_steps(dsk_engine* e, int step0, int nsteps);
int dsk_backward_steps(dsk_engine* e, int step_hi, int nsteps);
int dsk_loss_add_l2_steps(dsk_engine* e, int step0, int nsteps, const float* target, double weight, int on_device);
/* launches issued by this engine since creation (bench.py's gpu_launches) */
int dsk_launch_count(dsk_engine* e, int64_t* n);
/* bytes of device memory owned by the engine */
int dsk_memory_bytes(dsk_engine* e, int64_t* n);

#ifdef __cplusplus
}
#endif
#endif
