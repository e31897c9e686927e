"""Env-batch sharding across GPUs (one process per GPU) and the planner's exchange step.

Environments are independent (start/goal pairs, action-sequence candidates), so the path shards with no data-path
collective: rank r owns a contiguous block of envs.  The only exchange is one all-gather per optimiser iteration of
the per-env losses ``[B_local]`` and action gradients ``[H, B_local, A]`` (KBs; latency-bound) -- NCCL over
NVLink/NVSwitch on the GPU box, gloo in the CPU tests.  Replaces the reference's one-process-per-GPU pipes
(plb/envs/mp_wrapper.py:46-167).
"""


def shard_envs(total_envs, world_size, rank):
    """Contiguous block [lo, hi) of env ids owned by `rank`; blocks differ by at most one env."""
    base, rem = divmod(total_envs, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_planner_inputs(loss, grads, group=None):
    """all-gather of per-env loss [B_local] and action grads [H, B_local, A] -> ([B], [H, B, A]) on every rank.
    Requires equal B_local on all ranks (the benchmark configs divide evenly)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return loss, grads
    w = dist.get_world_size(group)
    H, B, A = grads.shape
    all_loss = torch.empty(w * loss.numel(), dtype=loss.dtype, device=loss.device)
    all_grads = torch.empty(w * grads.numel(), dtype=grads.dtype, device=grads.device)
    dist.all_gather_into_tensor(all_loss, loss.contiguous().reshape(-1), group=group)
    dist.all_gather_into_tensor(all_grads, grads.contiguous().reshape(-1), group=group)
    return all_loss.reshape(w * B), all_grads.reshape(w, H, B, A).permute(1, 0, 2, 3).reshape(H, w * B, A)
