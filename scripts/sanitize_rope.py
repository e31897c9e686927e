"""Engine-only run (no oracle) of the Rope-v1 scene for compute-sanitizer: H env steps forward + backward."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
from helpers import small_dough, perturbed_state, tool_start
from diffskill_b200.engine import Engine
f32 = lambda a: np.asarray(a, np.float32)
name, H, n = (sys.argv[1] if len(sys.argv) > 1 else 'Rope-v1'), 2, 800
scene, cfg, x0 = small_dough(name, n)
v0, F0, C0 = perturbed_state(x0, 1)
eng = Engine(scene, n_envs=1, capacity=n, max_steps=H, step_slots=1)
eng.set_particles(0, 0, f32(x0), f32(v0), f32(F0), f32(C0))
for i, s in enumerate(tool_start(name, scene)):
    eng.set_tool_state(0, 0, i, f32(s))
acts = f32(np.random.RandomState(3).uniform(-1, 1, (H, scene.action_dim)) * 0.7)
for s in range(H):
    eng.set_action(s, acts[s][None]); eng.forward_step(s)
rng = np.random.RandomState(11)
eng.zero_grad(); eng.add_particle_grad(H, f32(rng.normal(size=(n, 3)))[None], f32(rng.normal(size=(n, 3)) * 0.01)[None])
for s in range(H - 1, -1, -1):
    eng.backward_step(s)
print('action grad', eng.get_action_grad(0)[0], 'launches', eng.launch_count())
