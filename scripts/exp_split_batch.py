"""Experiment: the batch of a planner iteration split over several engines on concurrent CUDA streams.

The kernels of one env step form a dependent chain (~150 launches per step pair, each followed by a ~1.3 us graph edge and a
tail in which the last CTAs run alone); envs are independent, so k engines of B/k envs on k streams let the chains of different
sub-batches fill each other's gaps.  usage: python scripts/exp_split_batch.py [workload] [total envs] [k ...]
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
from diffskill_b200.engine import Engine

wl = sys.argv[1] if len(sys.argv) > 1 else 'gathermove'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
ks = [int(a) for a in sys.argv[3:]] or [1, 2, 4]
spec = bench.workload_spec(wl)
H = spec['horizon']
dev = torch.device('cuda', 0)
scene, cfg, xs, targets, actions = bench.make_inputs(spec, 0, B)
cap = max(len(x) for x in xs)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for k in ks:
    assert B % k == 0
    b = B // k
    streams = [torch.cuda.Stream(device=dev) for _ in range(k)]
    engs, tgts, acts = [], [], []
    for i in range(k):
        with torch.cuda.stream(streams[i]):
            e = Engine(scene, n_envs=b, capacity=cap, max_steps=H, step_slots=H, device=0, grid_tape_mib=8192 // k)
            e.set_stream(streams[i].cuda_stream)
            tgt = np.zeros((b, cap, 3), np.float32)
            for q in range(b):
                e.set_particles(0, q, xs[i * b + q]); tgt[q, :len(xs[i * b + q])] = targets[i * b + q]
            if spec['env'] == 'GatherMove-v1':
                from diffskill_b200.envs import generators as gen
                gen.settle(e)
            engs.append(e); tgts.append(torch.from_numpy(tgt).to(dev)); acts.append(torch.from_numpy(np.ascontiguousarray(actions[:, i * b:(i + 1) * b])).to(dev))
    torch.cuda.synchronize()

    def iteration():
        for i, e in enumerate(engs):
            e.zero_grad(); e.loss_reset()
            e.set_actions(0, acts[i])
            e.forward_steps(0, H)
            e.loss_add_l2_steps(1, H, tgts[i], 1.0 / H)
            e.backward_steps(H - 1, H)

    for _ in range(3):
        iteration()
    torch.cuda.synchronize()
    ms = []
    for _ in range(4):
        flush.zero_()
        torch.cuda.synchronize()
        a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for s in streams: s.wait_event(a)
        iteration()
        for s in streams: torch.cuda.current_stream().wait_stream(s)
        z.record()
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(z))
    loss = sum(float(e.loss_get().sum()) for e in engs)
    print(f'{wl} B={B} engines={k} x {b} envs: {np.mean(ms):.2f} ms per iteration (min {np.min(ms):.2f})  loss {loss:.6f}', flush=True)
    del engs
    torch.cuda.empty_cache()
