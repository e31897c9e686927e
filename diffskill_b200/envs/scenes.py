"""Scene parameters of the three DiffSkill environments.

These are the *inputs* of the hot path: the numeric values of
plb/envs/lift_spread.yml, plb/envs/gather_move.yml and
plb/cut/cut_rearrange.yml (gym ids registered at plb/envs/__init__.py:30-61),
restated as Python data so no reference file is needed at run time.  A user
with the reference checkout can equally pass the YAML path to
``diffskill_b200.scene.load_scene``.
"""

_WOOD = (0.7568, 0.6039, 0.4196)
_TILT = (0.707, 0.707, 0., 0.)          # note: not unit norm

LIFT_SPREAD = dict(
    SIMULATOR=dict(E=5000., n_particles=30000, yield_stress=200., ground_friction=1.5, gravity=(0, -20, 0), quality=1),
    SHAPES=[dict(shape='sphere', init_pos=(0.65, 0.08, 0.5), radius=0.05, color=100)],
    PRIMITIVES=[
        dict(shape='RollingPinExt', h=0.3, r=0.03, init_pos=(0.3, 0.25, 0.5), init_rot=_TILT, color=_WOOD,
             friction=0.9, action=dict(dim=6, scale=(0.7, 0.005, 0.005, 0.005, 0., 0.)), lower_bound=(0., 0.16, 0.)),
        dict(shape='Box', size=(0.1, 0.1, 0.02), init_pos=(0.65, 0.02, 0.5), init_rot=_TILT, color=_WOOD,
             friction=50., action=dict(dim=6, scale=(0.01, 0.01, 0., 0.0, 0., 0.05)), collision_group=(0, 0, 1)),
        dict(shape='Box', size=(0.2, 0.28, 0.07), init_pos=(0.3, 0.05, 0.5), init_rot=_TILT, color=(0.5, 0.5, 0.5),
             friction=5., action=dict(dim=0)),
    ],
    ENV=dict(cached_state_path='datasets/0202_liftspread', env_name='LiftSpread-v1'),
)

GATHER_MOVE = dict(
    SIMULATOR=dict(E=5000., n_particles=30000, yield_stress=200., ground_friction=1.5, gravity=(0, -20, 0), quality=1),
    SHAPES=[dict(shape='scatter', pos_min=(0.64, 0.02, 0.38), pos_max=(0.76, 0.035, 0.62), color=_WOOD, seed=0)],
    PRIMITIVES=[
        dict(shape='Gripper', size=(0.015, 0.09, 0.05), init_pos=(0.7, 0.06, 0.5), init_rot=(0.5, 0.5, -0.5, 0.5),
             init_gap=0.4, minimal_gap=0.05, color=_WOOD, friction=1.,
             action=dict(dim=7, scale=(0.015, 0.0, 0.015, 0.0, 0.0, 0.1, 0.03)), collision_group=(0, 0, 1)),
        dict(shape='Box', size=(0.07, 0.07, 0.02), init_pos=(0.7, 0.01, 0.5), init_rot=_TILT, color=_WOOD,
             friction=50., action=dict(dim=6, scale=(0.01, 0.01, 0., 0.0, 0., 0.05)), collision_group=(0, 0, 1)),
        dict(shape='Box', size=(0.2, 0.28, 0.04), init_pos=(0.33, 0.05, 0.5), init_rot=_TILT, color=(0.5, 0.5, 0.5),
             friction=5., action=dict(dim=0)),
    ],
    ENV=dict(cached_state_path='datasets/0202_gathermove', env_name='GatherMove-v1'),
)

CUT_REARRANGE = dict(
    SIMULATOR=dict(E=5000., n_particles=30000, quality_multiplier=1.25, yield_stress=150., ground_friction=0.5,
                   gravity=(0, -10, 0), quality=1, dtype='float32', lower_bound=1.),
    SHAPES=[dict(shape='box', init_pos=(0.5, 0.12, 0.5), width=(0.2, 0.08, 0.08), color=100, n_particles=5000)],
    PRIMITIVES=[
        dict(shape='Knife', h=(0.15, 0.15), size=(0.025, 0.2, 0.06), prot=(1.0, 0.0, 0.0, 0.58),
             init_pos=(0.5, 0.3, 0.5), color=_WOOD, friction=0., action=dict(dim=3, scale=(0.015, 0.015, 0.0))),
        dict(shape='Gripper', size=(0.015, 0.1, 0.06), init_pos=(0.5, 0.10, 0.5), init_gap=0.18, minimal_gap=0.08,
             maximal_gap=0.2, init_rot=(0.707, 0.0, 0.707, 0.0), color=_WOOD, friction=10.,
             action=dict(dim=7, scale=(0.015, 0.015, 0.015, 0., 0., 0., 0.015))),
    ],
    ENV=dict(cached_state_path='datasets/1215_cutrearrange', env_name='CutRearrange-v1'),
)

# PlasticineLab's Move-v1 (plb/envs/move.yml, first variant): two Sphere manipulators around a ball of dough.  Not one
# of the three DiffSkill envs -- registered to exercise the Sphere tool.
MOVE = dict(
    SIMULATOR=dict(E=5000., n_particles=10000, yield_stress=200.),
    SHAPES=[dict(shape='sphere', radius=0.2049069760770578 / 2,
                 init_pos=(0.6757143040494873, 0.5619162002773135, 0.7515980438048129), color=127 << 16)],
    PRIMITIVES=[
        dict(shape='Sphere', radius=0.03, init_pos=(0.5757143040494873, 0.5619162002773135, 0.7515980438048129),
             color=(0.7, 0.7, 0.7), friction=0.9, action=dict(dim=3, scale=(0.01, 0.01, 0.01))),
        dict(shape='Sphere', radius=0.03, init_pos=(0.7757143040494873, 0.5619162002773135, 0.7515980438048129),
             color=(0.7, 0.7, 0.7), friction=0.9, action=dict(dim=3, scale=(0.01, 0.01, 0.01))),
    ],
    ENV=dict(env_name='Move-v1'),
)

# Legacy PlasticineLab scenes that exercise the remaining tools: first variants of
# plb/envs/rollingpin.yml (RollingPin), torus.yml (Torus) and rope.yml (two Spheres + a static Cylinder pillar).
ROLLINGPIN = dict(
    SIMULATOR=dict(E=5000., n_particles=10000, yield_stress=50., ground_friction=1.5),
    SHAPES=[dict(shape='box', width=(0.3, 0.1, 0.3), init_pos=(0.5, 0.05, 0.5), color=100)],
    PRIMITIVES=[
        dict(shape='RollingPin', h=0.3, r=0.03, init_pos=(0.5, 0.123, 0.5), init_rot=_TILT, color=(0.8, 0.8, 0.8),
             friction=0.9, action=dict(dim=3, scale=(0.6666666666666667, 0.06666666666666668, 0.001))),
    ],
    ENV=dict(env_name='Rollingpin-v1'),
)
TORUS = dict(
    SIMULATOR=dict(yield_stress=50., ground_friction=100.),
    SHAPES=[dict(shape='box', width=(0.3, 0.1, 0.3), init_pos=(0.5, 0.05, 0.5), color=(((200 << 8) + 200) << 8))],
    PRIMITIVES=[
        dict(shape='Torus', tx=0.05, ty=0.03, init_pos=(0.5, 0.2, 0.5), init_rot=(0., 0., 0., 1.), friction=0.9,
             color=(0.8, 0.8, 0.8), lower_bound=(0., 0.05, 0.), action=dict(dim=3, scale=(0.004, 0.004, 0.004))),
    ],
    ENV=dict(env_name='Torus-v1'),
)
ROPE = dict(
    SIMULATOR=dict(yield_stress=50., ground_friction=0.3),
    SHAPES=[dict(shape='box', width=(0.6, 0.06, 0.06), init_pos=(0.5, 0.03, 0.73), color=(((0 << 8) + 150) << 8))],
    PRIMITIVES=[
        dict(shape='Sphere', radius=0.03, init_pos=(0.22, 0.015, 0.82), color=(0.8, 0.8, 0.8), friction=0.9,
             action=dict(dim=3, scale=(0.01, 0.01, 0.01))),
        dict(shape='Sphere', radius=0.03, init_pos=(0.78, 0.015, 0.82), color=(0.8, 0.8, 0.8), friction=0.9,
             action=dict(dim=3, scale=(0.01, 0.01, 0.01))),
        dict(shape='Cylinder', h=0.1, r=0.2, init_pos=(0.3919300650726247, 0., 0.4990770359432596),
             color=(0.3, 0.3, 0.3), friction=0.9),
    ],
    ENV=dict(env_name='Rope-v1'),
)
# No reference YAML instantiates Gripper2 (primitives.py:576-697, capsule jaws); this scene is the CutRearrange gripper
# pose and action scales with capsule jaws, so that the tool is exercised by the same parity tests.
GRIPPER2_SYNTHETIC = dict(
    SIMULATOR=dict(E=5000., n_particles=10000, yield_stress=150., ground_friction=0.5, gravity=(0, -10, 0)),
    SHAPES=[dict(shape='box', init_pos=(0.5, 0.06, 0.5), width=(0.12, 0.08, 0.08), color=100, n_particles=5000)],
    PRIMITIVES=[
        dict(shape='Gripper2', h=0.12, r=0.02, init_pos=(0.5, 0.08, 0.5), init_gap=0.18, minimal_gap=0.08,
             maximal_gap=0.2, init_rot=(0.707, 0.0, 0.707, 0.0), color=_WOOD, friction=10.,
             action=dict(dim=7, scale=(0.015, 0.015, 0.015, 0.02, 0.02, 0.02, 0.015))),
    ],
    ENV=dict(env_name='Gripper2-synthetic'),
)

# plb/envs/chopsticks.yml (PlasticineLab's Chopsticks-v1): a rope-like box and the two-stick tool
CHOPSTICKS = dict(
    SIMULATOR=dict(n_particles=10000, yield_stress=200., ground_friction=0., gravity=(0, -5, 0)),
    SHAPES=[dict(shape='box', width=(0.04, 0.04, 0.6), init_pos=(0.5, 0.02, 0.5), color=100)],
    PRIMITIVES=[
        dict(shape='Chopsticks', h=0.2, r=0.02, init_pos=(0.5, 0.15, 0.5), init_rot=(1., 0., 0., 0.), init_gap=0.06,
             color=(0.8, 0.8, 0.8), friction=10., action=dict(dim=7, scale=(0.02, 0.02, 0.02, 0.04, 0.04, 0.04, 0.02))),
    ],
    ENV=dict(env_name='Chopsticks-v1'),
)

SCENES = {
    'LiftSpread-v1': LIFT_SPREAD,
    'GatherMove-v1': GATHER_MOVE,
    'CutRearrange-v1': CUT_REARRANGE,
    'Move-v1': MOVE,
    'Rollingpin-v1': ROLLINGPIN,
    'Torus-v1': TORUS,
    'Rope-v1': ROPE,
    'Gripper2-synthetic': GRIPPER2_SYNTHETIC,
    'Chopsticks-v1': CHOPSTICKS,
}
