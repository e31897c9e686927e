"""Batched trajectory optimiser over the engine: the caller of the hot path.

Mirrors ``Solver.solve_one_plan`` (plb/optimizer/solver.py:97-152) and ``plb/cut/solve_func.solve``: Adam over an
action sequence ``[H, A]`` per env, loop = reset -> H x forward -> loss -> backward -> clamp to [-1,1] -> mask.
Differences: B envs are optimised at once inside one engine (the reference runs one process per env), the loss is
supplied as adjoints at the step boundaries (``loss_fn`` below) instead of a torch graph of 50 autograd nodes, and
across GPUs one ``gather_planner_inputs`` per iteration collects every rank's losses and gradients.
"""
import numpy as np

from .parallel import gather_planner_inputs


class BatchedSolver:
    def __init__(self, engine, horizon, lr=0.01, betas=(0.9, 0.999), eps=1e-8, action_mask=None):
        self.eng, self.H = engine, horizon
        self.lr, self.b1, self.b2, self.eps = lr, betas[0], betas[1], eps
        self.mask = None if action_mask is None else np.asarray(action_mask, np.float32).reshape(1, 1, -1)

    def l2_target_loss(self, targets, weight=None):
        """loss_fn: sum_t w * mean_p |x_t - target|^2 (deterministic stand-in for the EMD loss), engine-side."""
        w = 1.0 / self.H if weight is None else weight

        def fn(eng, step):
            eng.loss_add_l2(step, targets, w)
        fn.steps = lambda eng, step0, n: eng.loss_add_l2_steps(step0, n, targets, w)   # multi-step fast path
        return fn

    def rollout_grad(self, actions, loss_fn):
        """One iteration: returns (per-env loss [B], action gradient [H,B,A])."""
        eng = self.eng
        eng.zero_grad()
        eng.loss_reset()
        if hasattr(loss_fn, 'steps'):   # engine-side loss: the whole rollout in five host calls
            eng.set_actions(0, actions)
            eng.forward_steps(0, self.H)
            loss_fn.steps(eng, 1, self.H)
            eng.backward_steps(self.H - 1, self.H)
        else:
            for s in range(self.H):
                eng.set_action(s, actions[s])
                eng.forward_step(s)
                loss_fn(eng, s + 1)
            for s in range(self.H - 1, -1, -1):
                eng.backward_step(s)
        return eng.loss_get(), eng.get_action_grads(0, self.H)

    def solve(self, init_actions, loss_fn, max_iter=20, callback=None, distributed=False):
        """Adam on the action sequences; keeps the best-so-far plan per env (solver.py:135-141)."""
        a = np.array(init_actions, dtype=np.float32).reshape(self.H, self.eng.B, self.eng.A)
        m, v = np.zeros_like(a), np.zeros_like(a)
        best_loss = np.full(self.eng.B, np.inf, np.float32)
        best_a = a.copy()
        history = []
        for it in range(1, max_iter + 1):
            loss, g = self.rollout_grad(a, loss_fn)
            if distributed:
                import torch
                gl, gg = gather_planner_inputs(torch.from_numpy(loss), torch.from_numpy(g))
                history.append(float(gl.mean()))
            else:
                history.append(float(loss.mean()))
            if not np.isfinite(loss).all():          # "MEET NAN" (plb/cut/solve_func.py:124-129)
                break
            better = loss < best_loss
            best_loss = np.where(better, loss, best_loss)
            best_a[:, better] = a[:, better]
            if self.mask is not None:
                g = g * self.mask
            m = self.b1 * m + (1 - self.b1) * g
            v = self.b2 * v + (1 - self.b2) * g * g
            a = a - self.lr * (m / (1 - self.b1 ** it)) / (np.sqrt(v / (1 - self.b2 ** it)) + self.eps)
            a = np.clip(a, -1, 1)
            if self.mask is not None:
                a = a * self.mask
            if callback:
                callback(it, loss)
        return dict(best_action=best_a, best_loss=best_loss, history=history, last_action=a)
