"""Short eager (graph-free) run of the bench workload for ncu: `H` env steps forward + backward."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from diffskill_b200.engine import Engine

wl = sys.argv[1] if len(sys.argv) > 1 else 'liftspread'
H = int(sys.argv[2]) if len(sys.argv) > 2 else 2
B = int(sys.argv[3]) if len(sys.argv) > 3 else 1
spec = bench.workload_spec(wl)
spec['horizon'] = H
scene, cfg, xs, targets, actions = bench.make_inputs(spec, 0, B)
cap = max(len(x) for x in xs)
eng = Engine(scene, n_envs=B, capacity=cap, max_steps=H, step_slots=H, grid_tape_mib=4096)
eng.set_graphs(False)
tgt = np.zeros((B, cap, 3), np.float32)
for b in range(B):
    eng.set_particles(0, b, xs[b]); tgt[b, :len(xs[b])] = targets[b]
# let the dough settle onto the tools first so contacts are active in the profiled steps
for it in range(int(os.environ.get('PROFILE_ITERS', '2'))):
    eng.zero_grad(); eng.loss_reset()
    for s in range(H):
        eng.set_action(s, actions[s]); eng.forward_step(s); eng.loss_add_l2(s + 1, tgt, 1.0 / H)
    for s in range(H - 1, -1, -1):
        eng.backward_step(s)
eng.synchronize()
print('done', eng.launch_count(), eng.get_action_grads(0, H).ravel()[:4])
