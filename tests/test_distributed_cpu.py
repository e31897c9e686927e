"""world_size-2 gloo test of the N>1 host path: env sharding + the planner's all-gather (NCCL on the GPU box)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from diffskill_b200.parallel import gather_planner_inputs, shard_envs


def _worker(rank, world, port, total_envs, H, A, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    lo, hi = shard_envs(total_envs, world, rank)
    env_ids = torch.arange(lo, hi, dtype=torch.float32)
    loss = env_ids * 10.0                                             # per-env loss [B_local]
    grads = env_ids[None, :, None] + torch.arange(H, dtype=torch.float32)[:, None, None] * 100 \
        + torch.arange(A, dtype=torch.float32)[None, None, :] * 0.01  # [H, B_local, A]
    all_loss, all_grads = gather_planner_inputs(loss, grads)
    q.put((rank, all_loss.numpy(), all_grads.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_reassembles_the_global_batch():
    world, total, H, A = 2, 8, 3, 5
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, H, A, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ids = np.arange(total, dtype=np.float32)
    exp_loss = ids * 10
    exp_grads = ids[None, :, None] + np.arange(H, dtype=np.float32)[:, None, None] * 100 + np.arange(A, dtype=np.float32)[None, None, :] * 0.01
    for rank, l, g in res:
        assert np.array_equal(l, exp_loss)
        assert g.shape == (H, total, A) and np.allclose(g, exp_grads)


def test_single_process_gather_is_identity():
    l, g = torch.ones(4), torch.ones(3, 4, 2)
    a, b = gather_planner_inputs(l, g)
    assert a is l and b is g
