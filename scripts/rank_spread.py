"""Attribution of the per-rank spread of the weak-scaling LiftSpread run (per_rank_ms 41.2 ... 50.2 at N=8).

Every rank simulates a different env (dough seed and action sequence = rank), so the max over ranks is the slowest of eight
contact histories.  This script replays the envs of ranks 0..7 one after the other on ONE GPU with the profiling library
(kernels stamp %globaltimer inside the replayed graphs) and sums the in-graph kernel durations per kernel class over the
whole H=50 iteration: the class whose total moves with the rank is the cause.
usage: DSK_LIB=timeline python scripts/rank_spread.py [workload] [ranks]
"""
import os, sys
os.environ.setdefault('DSK_LIB', 'timeline')
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from diffskill_b200.engine import Engine

wl = sys.argv[1] if len(sys.argv) > 1 else 'liftspread'
ranks = int(sys.argv[2]) if len(sys.argv) > 2 else 8
spec = bench.workload_spec(wl)
H, B = spec['horizon'], spec.get('envs_per_gpu') or 1
rows = []
for rank in range(ranks):
    scene, cfg, xs, targets, actions = bench.make_inputs(spec, rank, B)
    cap = max(len(x) for x in xs)
    eng = Engine(scene, n_envs=B, capacity=cap, max_steps=H, step_slots=H, grid_tape_mib=4096)
    tgt = np.zeros((B, cap, 3), np.float32)
    for b in range(B):
        eng.set_particles(0, b, xs[b]); tgt[b, :len(xs[b])] = targets[b]
    eng.timeline_enable(True)
    tot, span = {}, 0.0

    def iteration(probe):
        global span
        eng.zero_grad(); eng.loss_reset()
        for s in range(H):
            eng.set_action(s, actions[s])
            if probe: eng.timeline_reset()
            eng.forward_step(s)
            if probe: acc(eng.timeline_read())
            eng.loss_add_l2(s + 1, tgt, 1.0 / H)
        for s in range(H - 1, -1, -1):
            if probe: eng.timeline_reset()
            eng.backward_step(s)
            if probe: acc(eng.timeline_read())
        eng.synchronize()

    def acc(recs):
        global span
        if not recs:
            return
        span += (max(r[2] for r in recs) - min(r[1] for r in recs)) / 1e6
        for k, t0, t1 in recs:
            tot[k] = tot.get(k, 0.0) + (t1 - t0) / 1e6

    iteration(False); iteration(False)
    iteration(True)
    hits = 0
    for s in range(H):
        for j in range(1, scene.substeps + 1):
            for t in range(len(scene.tools)):
                pass
    rows.append((rank, span, dict(tot)))
    del eng
classes = sorted({k for _, _, t in rows for k in t}, key=lambda k: -max(t.get(k, 0) for _, _, t in rows))
print(f'# {wl}: sum over the {H} forward + {H} backward step graphs of one iteration, per rank (ms); span = sum of the graphs\' wall spans')
print('| rank | span | ' + ' | '.join(classes) + ' |')
print('|---|---|' + '---|' * len(classes))
for rank, span, t in rows:
    print(f'| {rank} | {span:.2f} | ' + ' | '.join(f'{t.get(k, 0):.2f}' for k in classes) + ' |')
lo = min(rows, key=lambda r: r[1]); hi = max(rows, key=lambda r: r[1])
print(f'\nslowest rank {hi[0]} ({hi[1]:.2f} ms) vs fastest rank {lo[0]} ({lo[1]:.2f} ms): difference per class (ms): ' +
      ', '.join(f'{k} {hi[2].get(k, 0) - lo[2].get(k, 0):+.2f}' for k in classes))
