/* oracle/mpm_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * C interface of the CPU oracle: a literal, un-fused restatement of the
 * reference's differentiable MLS-MPM path (plb/engine/mpm_simulator.py,
 * plb/engine/primitive/{primive_base,primitives,utils}.py, plb/engine/function.py).
 *
 * PARITY UNPINNED: the reference's arithmetic lives in taichi==0.7.26
 * (environment.yml:21), which is neither vendored nor installable offline, and
 * the reference ships no golden vectors for this path.
 * The oracle is pinned only by (1) fp64 central finite differences of its own
 * adjoints, (2) invariants, (3) committed fixtures it generated itself.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.
 */
#ifndef MPM_ORACLE_H
#define MPM_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_TOOLS 8
#define ORC_MAX_PAIRS 8
#define ORC_NUM_COLLISION_POINTS 600 /* mpm_simulator.py:59 */

enum orc_tool_type {
  ORC_TOOL_CAPSULE = 0,        /* primitives.py:43  (base forward_kinematics) */
  ORC_TOOL_ROLLINGPIN_EXT = 1, /* primitives.py:120 */
  ORC_TOOL_BOX = 2,            /* primitives.py:359 */
  ORC_TOOL_GRIPPER = 3,        /* primitives.py:428 */
  ORC_TOOL_KNIFE = 4,          /* primitives.py:740 */
  ORC_TOOL_SPHERE = 5,         /* primitives.py:23  */
  ORC_TOOL_ROLLINGPIN = 6,     /* primitives.py:101 */
  ORC_TOOL_GRIPPER2 = 7,       /* primitives.py:576 (capsule jaws) */
  ORC_TOOL_CYLINDER = 8,       /* primitives.py:302 (h = radial, r = axial half extent) */
  ORC_TOOL_TORUS = 9,          /* primitives.py:337 (tx in h, ty in r) */
  ORC_TOOL_CHOPSTICKS = 10     /* primitives.py:218 (two capsules at +-gap/2 inside ONE tool frame) */
};

typedef struct orc_tool_cfg {
  int32_t type;
  int32_t action_dim;
  double action_scale[8];
  double friction;
  double softness;
  double lower_bound[3], upper_bound[3]; /* xyz_limit */
  double size[3];                        /* Box / Gripper / Knife.box half extents */
  double h, r;                           /* Capsule, RollingPin, Gripper2 jaws; Cylinder (h, r); Torus (tx, ty) */
  double radius;                         /* Sphere */
  double prism_h[2];                     /* Knife.prism.h */
  double prot[4];                        /* Knife.prism.prot */
  double minimal_gap, maximal_gap;       /* Gripper */
} orc_tool_cfg;

typedef struct orc_config {
  int32_t n_grid;
  int32_t substeps;
  int32_t max_frames;    /* tape length (reference: max_steps = 1024) */
  int32_t max_particles; /* capacity */
  double dt, dx, inv_dx, p_vol, p_mass;
  double mu, lam, yield_stress;
  double gravity[3];
  double ground_friction;
  double lower_bound;
  int32_t n_tools;
  orc_tool_cfg tools[ORC_MAX_TOOLS];
  int32_t n_pairs;
  int32_t pairs[ORC_MAX_PAIRS][2]; /* (i moved, j obstacle box) mpm_simulator.py:69-80 */
} orc_config;

void* orc_create(const orc_config* cfg, int use_f64);
void orc_destroy(void* h);
void orc_set_threads(int n);
/* 0: one-sided Jacobi (default); 1: Jacobi eigen-decomposition of A^T A + Givens QR (McAdams et al. construction): a second,
   independent algorithm to compare degenerate-subspace behaviour of backward_svd across (process-wide switch) */
void orc_set_svd_algorithm(int alg);
/* seed != 0: every grid sum of p2g / g2p.grad is perturbed by -ulps/0/+ulps units in the last place (hash of seed, frame, node): emulates the
 * run-to-run noise of the reference's unordered float atomics (mpm_simulator.py:224-225); 0 = exact (default) */
void orc_set_scatter_noise(int seed, double ulps);
/* amplitude > 0: log carries a pseudo-random absolute error in [-a, a], exp a relative one in [-a/2, a/2] (the reference
 * runs fast_math=True, taichi_env.py:20: hardware log/exp, |err(log)| <= 2^-21.4 = 3.7e-7); 0 = exact (default) */
void orc_set_fast_math_noise(double amplitude, int salt);
/* seed != 0: grid_op.grad adds every node's tool-pose adjoint terms to position.grad / rotation.grad one by one in the
 * simulation precision, in a seeded random node order -- what the reference's float atomics do (the terms are O(1/dt) and
 * cancel); 0 (default): order-independent sum in double, rounded once */
void orc_set_pose_adjoint_atomics(int seed);
int orc_is_f64(void* h);

void orc_initialize(void* h, int n_particles);            /* mpm_simulator.py:82-97 */
void orc_set_rand_num(void* h, const double* rand_num);   /* [pairs,600,3] */
void orc_set_material(void* h, const double* mu, const double* lam, const double* yield_stress);
void orc_set_tool_param(void* h, int tool, int which, double value); /* 0 friction 1 softness 2..4 lower 5..7 upper */
void orc_set_gravity(void* h, const double* g);

void orc_set_frame(void* h, int f, int n, const double* x, const double* v, const double* F, const double* C);
void orc_get_frame(void* h, int f, double* x, double* v, double* F, double* C);
void orc_set_tool_state(void* h, int f, int tool, const double* state8);
void orc_get_tool_state(void* h, int f, int tool, double* state8);
void orc_copyframe(void* h, int src, int dst);            /* mpm_simulator.py:368-378 */
int orc_n_particles(void* h);

void orc_set_action(void* h, int s, int n_substeps, const double* action); /* primitives.py:863-867 */
void orc_substep(void* h, int f);                         /* mpm_simulator.py:307-323 */
void orc_substep_grad(void* h, int f);                    /* mpm_simulator.py:325-345 */
void orc_set_velocity_grad(void* h, int s, int n_substeps); /* function.py:180-182 */
void orc_get_action_grad(void* h, int s, double* out);    /* primive_base.py:254-258 */

void orc_zero_grad(void* h);                              /* function.py:44-61 */
void orc_get_frame_grad(void* h, int f, double* gx, double* gv, double* gF, double* gC);
void orc_add_frame_grad(void* h, int f, const double* gx, const double* gv, const double* gF, const double* gC);
void orc_scale_frame_grad(void* h, int f, double alpha);  /* function.py:66-77 */
void orc_get_tool_grad(void* h, int f, int tool, double* g8);
void orc_add_tool_grad(void* h, int f, int tool, const double* g8);
void orc_get_tool_vel_grad(void* h, int f, int tool, double* g7); /* v3 w3 gap_vel */

void orc_get_grid(void* h, double* v_in, double* v_out, double* m);
void orc_get_grid_grad(void* h, double* g_v_in, double* g_v_out, double* g_m);
void orc_get_svd(void* h, double* F_tmp, double* U, double* sig, double* V);
void orc_cell_index(void* h, int f, int32_t* base, int32_t* key);
void orc_occupancy(void* h, int f, uint8_t* occ);
void orc_get_collision_idx(void* h, int f, int32_t* idx);

int orc_min_dist_cols(void* h);
void orc_compute_min_dist(void* h, int f, double* out);            /* function.py:79-88 */
void orc_compute_min_dist_grad(void* h, int f, const double* gin);
void orc_compute_grid_m(void* h, int f, double* out);              /* mpm_simulator.py:456-471 */
void orc_compute_grid_m_grad(void* h, int f, const double* gin);

/* single-function probes used by unit tests */
void orc_svd3(int use_f64, const double* F, double* U, double* sig, double* V);
double orc_tool_sdf(void* h, int tool, int f, const double* p);
void orc_tool_normal(void* h, int tool, int f, const double* p, double* n);
void orc_tool_collide(void* h, int tool, int f, const double* p, const double* v_in, double* v_out);
/* tape-AD adjoint of tool_sdf (what = 0, gout[1]), tool_normal (1, gout[3]) or tool_collide (2, gout[3]) at poses f, f+1:
 * out22 = [g(p) 3 | g(v_in) 3 | g(pose f) 8 | g(pose f+1) 8] */
void orc_tool_probe_grad(void* h, int tool, int f, int what, const double* p, const double* v_in, const double* gout,
                         double* out22);
/* forward_kinematics of one tool from (state8, vel7 = v3 w3 gap_vel); gnext8 != NULL: also [g(state) 8 | g(vel) 7] */
void orc_tool_probe_fk(void* h, int tool, const double* state8, const double* vel7, double* next8,
                       const double* gnext8, double* gout15);

#ifdef __cplusplus
}
#endif
#endif
