"""In-graph kernel timeline of one env step (forward and backward) of a bench workload.

Needs the profiling library: `python -m diffskill_b200.build --timeline`, run with DSK_LIB=timeline.
Prints, per kernel launch inside the replayed CUDA graph, start offset and duration (device %globaltimer),
then per-class totals and the critical-path gaps -- what CUDA events (eager only) and ncu (serialised, cold
cache) cannot show.
usage: DSK_LIB=timeline python scripts/timeline_step.py [workload] [envs] [--full]
"""
import os, sys
os.environ.setdefault('DSK_LIB', 'timeline')
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from diffskill_b200.engine import Engine

args = [a for a in sys.argv[1:] if not a.startswith('--')]
wl = args[0] if args else 'liftspread'
spec = bench.workload_spec(wl)
B = int(args[1]) if len(args) > 1 else spec.get('envs_per_gpu', 1)
H = 6
spec['horizon'] = H
scene, cfg, xs, targets, actions = bench.make_inputs(spec, 0, B)
cap = max(len(x) for x in xs)
eng = Engine(scene, n_envs=B, capacity=cap, max_steps=H, step_slots=H, grid_tape_mib=4096)
tgt = np.zeros((B, cap, 3), np.float32)
for b in range(B):
    eng.set_particles(0, b, xs[b]); tgt[b, :len(xs[b])] = targets[b]
eng.timeline_enable(True)


def iteration(probe=None):
    out = {}
    eng.zero_grad(); eng.loss_reset()
    for s in range(H):
        eng.set_action(s, actions[s])
        if probe == s: eng.timeline_reset()
        eng.forward_step(s)
        if probe == s: out['fwd'] = eng.timeline_read()
        eng.loss_add_l2(s + 1, tgt, 1.0 / H)
    for s in range(H - 1, -1, -1):
        if probe == s: eng.timeline_reset()
        eng.backward_step(s)
        if probe == s: out['bwd'] = eng.timeline_read()
    eng.synchronize()
    return out


for _ in range(3):
    iteration()
res = iteration(probe=H - 2)
full = '--full' in sys.argv
brief = '--brief' in sys.argv
for name in ('fwd', 'bwd'):
    recs = sorted(res[name], key=lambda r: r[1])
    if not recs:
        continue
    t_begin = min(r[1] for r in recs); t_end = max(r[2] for r in recs)
    tot = {}
    for k, t0, t1 in recs:
        d = tot.setdefault(k, [0, 0.0]); d[0] += 1; d[1] += (t1 - t0) / 1e3
    busy = sum(v[1] for v in tot.values())
    if brief:
        print(f'{wl} B={B} {name} span={(t_end - t_begin) / 1e3:.1f} ' + ' '.join(f'{k}={us / n:.2f}' for k, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:6]))
        continue
    print(f'== {wl} B={B} {name}: {len(recs)} launches, span {(t_end - t_begin) / 1e3:.1f} us')
    for k, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f'   {k:20s} n={n:3d}  total {us:8.1f} us  avg {us / n:6.2f} us  ({100 * us / (t_end - t_begin) * 1e3:5.1f}% of span)')
    print(f'   sum of kernel durations {busy:.1f} us (overlap or gaps: span - sum = {(t_end - t_begin) / 1e3 - busy:.1f} us)')
    if full:
        for k, t0, t1 in recs:
            print(f'     +{(t0 - t_begin) / 1e3:8.2f} us  {(t1 - t0) / 1e3:6.2f} us  {k}')
