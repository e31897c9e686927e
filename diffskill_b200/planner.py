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
        box = [targets]

        def fn(eng, step):
            eng.loss_add_l2(step, box[0], w)
        fn.steps = lambda eng, step0, n: eng.loss_add_l2_steps(step0, n, box[0], w)   # multi-step fast path

        def to_device(device):   # keep the targets resident next to the engine (no per-iteration staging copy)
            import torch
            if isinstance(box[0], np.ndarray):
                box[0] = torch.as_tensor(np.ascontiguousarray(box[0], dtype=np.float32), device=device)
        fn.to_device = to_device
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

    def rollout_grad_device(self, actions, loss_fn, loss_out, grad_out):
        """rollout_grad with everything resident on the device: `actions` [H,B,A], `loss_out` [B], `grad_out` [H,B,A] are
        torch tensors on the engine's GPU; no host synchronisation."""
        eng = self.eng
        eng.zero_grad()
        eng.loss_reset()
        eng.set_actions(0, actions)
        eng.forward_steps(0, self.H)
        if hasattr(loss_fn, 'steps'):
            loss_fn.steps(eng, 1, self.H)
        else:
            for s in range(1, self.H + 1):
                loss_fn(eng, s)
        eng.backward_steps(self.H - 1, self.H)
        eng.loss_get(loss_out)
        eng.get_action_grads(0, self.H, grad_out)

    def solve(self, init_actions, loss_fn, max_iter=20, callback=None, distributed=False, device=None):
        """Adam on the action sequences; keeps the best-so-far plan per env (solver.py:135-141).  The optimiser state, the
        clamp / mask and the best-so-far bookkeeping live in torch tensors on the engine's device (plb/optimizer/solver.py
        keeps them in torch too); one iteration makes no device->host copy unless a callback asks for the loss.  The NaN
        guard ("MEET NAN", plb/cut/solve_func.py:124-129) freezes the update on the device and is polled every 10th
        iteration."""
        import torch
        H, B, A = self.H, self.eng.B, self.eng.A
        if device is None:
            device = 'cuda' if torch.cuda.is_available() else 'cpu'
        on_gpu = str(device).startswith('cuda')
        if on_gpu and hasattr(loss_fn, 'to_device'):
            loss_fn.to_device(device)
        a = torch.as_tensor(np.array(init_actions, dtype=np.float32).reshape(H, B, A), device=device).contiguous()
        m, v = torch.zeros_like(a), torch.zeros_like(a)
        mask = None if self.mask is None else torch.as_tensor(self.mask, device=device)
        best_loss = torch.full((B,), float('inf'), device=device)
        best_a = a.clone()
        loss, g = torch.zeros(B, device=device), torch.zeros((H, B, A), device=device)
        history = []
        alive = torch.ones((), dtype=torch.bool, device=device)
        for it in range(1, max_iter + 1):
            if on_gpu:
                self.rollout_grad_device(a, loss_fn, loss, g)
            else:   # host engine (the CPU emulation of the test-suite): same arithmetic on CPU tensors
                l_, g_ = self.rollout_grad(a.numpy(), loss_fn)
                loss, g = torch.from_numpy(np.asarray(l_)), torch.from_numpy(np.asarray(g_))
            if distributed:
                gl, gg = gather_planner_inputs(loss, g)
                history.append(gl.mean())
            else:
                history.append(loss.mean())
            alive = alive & torch.isfinite(loss).all()
            better = (loss < best_loss) & alive
            best_loss = torch.where(better, loss, best_loss)
            best_a = torch.where(better.view(1, B, 1), a, best_a)
            gm = g if mask is None else g * mask
            m = self.b1 * m + (1 - self.b1) * gm
            v = self.b2 * v + (1 - self.b2) * gm * gm
            step = self.lr * (m / (1 - self.b1 ** it)) / (torch.sqrt(v / (1 - self.b2 ** it)) + self.eps)
            a_new = torch.clamp(a - step, -1, 1)
            if mask is not None:
                a_new = a_new * mask
            a = torch.where(alive, a_new, a).contiguous()
            if callback:
                callback(it, loss.detach().cpu().numpy())
            if it % 10 == 0 and not bool(alive):
                break
        hist = [float(h) for h in torch.stack(history).cpu()] if history else []
        if not bool(alive):   # history up to and including the first non-finite loss, as the reference's early break
            bad = next((i for i, h in enumerate(hist) if not np.isfinite(h)), len(hist) - 1)
            hist = hist[:bad + 1]
        return dict(best_action=best_a.cpu().numpy(), best_loss=best_loss.cpu().numpy(), history=hist,
                    last_action=a.cpu().numpy())


FUNCS = {}          # one GradModel per env, as plb/optimizer/solver.py:10,19-21


class Solver:
    """The reference's single-env trajectory optimiser with its own signatures (plb/optimizer/solver.py:13-165): torch Adam on
    an action sequence [H, A], each iteration = GradModel.reset -> H x GradModel.forward (autograd Function over the engine's
    forward_step / backward_step) -> loss_fn(idxes, observations, vel_loss_weight, loss_type=...) -> backward -> clamp to
    [-1, 1] -> mask, keeping the best-so-far plan.  `args` needs adam_loss_type, stop_action_n, vel_loss_weight, energy_weight,
    component_matching, enumerate_contact (debug plotting is not carried over).  BatchedSolver above is the fast path for many
    envs with an engine-side loss; this class is the drop-in for callers written against the reference."""

    def __init__(self, args, env, ouput_grid=(), device=None, **kwargs):
        import torch
        from .sim.function import GradModel
        self.args, self.env = args, env
        self.device = device or ('cuda' if torch.cuda.is_available() else 'cpu')
        self.env.update_loss_fn(self.args.adam_loss_type)
        if env not in FUNCS:
            FUNCS[env] = GradModel(env, output_grid=ouput_grid, **kwargs)
        self.func = FUNCS[env]
        self.buffer = []

    def solve(self, initial_actions, loss_fn, action_mask=None, lr=0.01, max_iter=200, verbose=True, scheduler=None):
        import torch
        from functools import partial
        if action_mask is not None:
            initial_actions = initial_actions * action_mask[None]
            action_mask = torch.as_tensor(np.asarray(action_mask)[None], dtype=torch.float32, device=self.device)
        self.initial_state = self.env.get_state()
        kw = dict(action_mask=action_mask, lr=lr, max_iter=max_iter, verbose=verbose, scheduler=scheduler)
        if getattr(self.args, 'component_matching', False):
            raise NotImplementedError("component_matching needs core.diffskill.utils.get_component_masks, which the reference "
                                      "tree does not contain (solver.py:3,30)")
        if getattr(self.args, 'enumerate_contact', False):
            # DBSCAN the initial dough; one optimisation per component with the contact loss restricted to it; keep the plan
            # with the largest relative improvement (solver.py:54-90)
            from sklearn.cluster import DBSCAN
            labels = DBSCAN(eps=0.01, min_samples=5).fit(self.initial_state['state'][0].reshape(-1, 3)).labels_
            infos, buffers = [], []
            for label in range(labels.max() + 1):
                self.env.set_state(**self.initial_state)
                mask = torch.as_tensor(labels == label, dtype=torch.bool, device=self.device)
                buffer, info = self.solve_one_plan(initial_actions, partial(loss_fn, state_mask=mask), **kw)
                buffers.append(buffer)
                infos.append(info)
            gain = np.array([(b[0]['loss'] - i['best_loss']) / b[0]['loss'] for b, i in zip(buffers, infos)])
            k = int(np.argmax(gain))
            self.buffer.append(buffers[k])
            return infos[k], buffers[k]
        buffer, info = self.solve_one_plan(initial_actions, loss_fn, **kw)
        self.buffer.append(buffer)
        return info, buffer

    def solve_one_plan(self, initial_actions, loss_fn, action_mask=None, lr=0.01, max_iter=200, verbose=True, scheduler=None):
        import torch
        action = torch.nn.Parameter(torch.as_tensor(np.array(initial_actions), dtype=torch.float32, device=self.device))
        optim = torch.optim.Adam([action], lr=lr)
        sched = None if scheduler is None else scheduler(optim)
        buffer, best_action, best_loss = [], initial_actions, np.inf
        loss, last, H = np.inf, initial_actions, action.shape[0]
        for iter_id in range(max_iter):
            optim.zero_grad()
            observations = self.func.reset(self.initial_state['state'], device=self.device)
            cached_obs = []
            for idx, a in enumerate(action):
                a = a.detach() if H - idx <= self.args.stop_action_n else a       # the last stop_action_n actions get no gradient
                observations = self.func.forward(idx, a, *observations)
                cached_obs.append(observations)
            loss = loss_fn(list(range(H)), cached_obs, self.args.vel_loss_weight, loss_type=self.args.adam_loss_type)
            assert self.args.energy_weight == 0.
            loss.backward()
            optim.step()
            if sched is not None:
                sched.step()
            with torch.no_grad():
                action.data.copy_(torch.clamp(action.data, -1, 1))
                if action_mask is not None:
                    action.data.copy_(action.data * action_mask)
                loss = loss.item()
                last = action.data.detach().cpu().numpy()
                if loss < best_loss:
                    best_loss, best_action = loss, last
            buffer.append({'action': last, 'loss': loss})
            if verbose:
                print(f"{iter_id}:  {loss}")
        self.env.set_state(**self.initial_state)
        return buffer, {'best_loss': best_loss, 'best_action': best_action, 'last_loss': loss, 'last_action': last}

    def eval(self, action, render_fn):
        self.env.simulator.cur = 0
        self.env.set_state(**self.initial_state)
        outs = []
        for a in action:
            self.env.step(a)
            outs.append(render_fn())
        self.env.set_state(**self.initial_state)
        return outs

    def save_plot_buffer(self, path, buffer=None):          # solver.py:181-193 (matplotlib only when asked for)
        import matplotlib.pyplot as plt
        for buf in (self.buffer if buffer is None else buffer):
            plt.plot(range(len(buf)), [b['loss'] for b in buf])
        plt.xlabel('Steps')
        plt.ylabel('Loss')
        plt.savefig(path)
        plt.close()

    def dump_buffer(self, path='/tmp/buffer.pkl'):
        import pickle
        with open(path, 'wb') as f:
            pickle.dump(self.buffer, f)


def solve(env, func, initial_actions, loss_fn, lr=0.01, max_iter=200, verbose=True, scheduler=None, action_dims=None,
          state=None, early_stop=None, compute_loss_in_end=False, device=None):
    """plb/cut/solve_func.solve (:32-166), the CutRearrange optimiser: resumable Adam on [H, A] actions with a per-step loss
    `loss_fn(idx, *observations)` (a tensor, or (tensor, dict of logged terms)), summed while stepping or after the rollout
    (`compute_loss_in_end`); after each update the actions are clamped to [-1, 1] and projected -- `action_dims` a tuple:
    only those columns stay non-zero; 'gripper': two 6-D tools move as one in y / z and keep their own x (:110-119) --
    NaN stops, a loss < -10000 is skipped as a bug guard, `early_stop` = patience in non-improving iterations.  Returns the
    plan and the optimiser state to pass back as `state=` for another `max_iter` iterations."""
    import torch
    dev = device or ('cuda' if torch.cuda.is_available() else 'cpu')
    if state is None:
        assert initial_actions is not None
        iter_id, optim_buffer = 0, []
        action = torch.nn.Parameter(torch.as_tensor(np.array(initial_actions), dtype=torch.float32, device=dev))
        optim = torch.optim.Adam([action], lr=lr)
        initial_state = env.get_state()
        scheduler = scheduler if scheduler is None else scheduler(optim)
        last_action, last_loss = best_action, best_loss = initial_actions, np.inf
    else:
        iter_id, optim_buffer, action, optim = state['iter_id'], state['optim_buffer'], state['action'], state['optim']
        initial_state, scheduler = state['initial_state'], state['scheduler']
        best_action, best_loss = state['best_action'], state['best_loss']
        last_action, last_loss = state['last_action'], state['last_loss']
    zero_masks = None
    if isinstance(action_dims, tuple):
        zero_masks = torch.ones(action.shape[-1], dtype=torch.bool, device=action.device)
        zero_masks[list(action_dims)] = False
    non_decrease_iters = 0
    for iter_id in range(iter_id, iter_id + max_iter):
        optim.zero_grad()
        loss, outputs = 0, []
        observations = func.reset(initial_state['state'], device=dev)

        def calc_loss(idx, observations):
            l = loss_fn(idx, *observations)
            if isinstance(l, tuple):
                outputs.append(l[1])
                return l[0]
            return l

        obs_array = []
        for idx, a in enumerate(action):
            observations = func.forward(idx, a, *observations)
            if compute_loss_in_end:
                obs_array.append(observations)
            else:
                loss = loss + calc_loss(idx, observations)
        for idx, observations in enumerate(obs_array):
            loss = loss + calc_loss(idx, observations)
        loss.backward()
        optim.step()
        if scheduler is not None:
            scheduler.step()
        with torch.no_grad():
            a = torch.clamp(action, -1, 1)
            if zero_masks is not None:
                a[:, zero_masks] = 0
            elif action_dims == 'gripper':
                a1, a2 = a[:, 0:3], a[:, 6:9]
                a = torch.zeros_like(a)
                a[:, [1, 2]] = a[:, [7, 8]] = (a1[:, 1:] + a2[:, 1:]) / 2
                a[:, 0], a[:, 6] = a1[:, 0], a2[:, 0]
            action.data[:] = a
            last_loss = loss.item()
            if np.isnan(last_loss):
                print("MEET NAN!!")
                break
            if last_loss < -10000:
                continue
            last_action = action.data.detach().cpu().numpy()
            if last_loss < best_loss:
                best_loss, best_action, non_decrease_iters = last_loss, last_action, 0
            else:
                non_decrease_iters += 1
        optim_buffer.append({'action': last_action, 'loss': last_loss})
        if verbose:
            word = f"{iter_id}: {last_loss:.4f}  {best_loss:.3f}"
            for k in (outputs[0] if outputs else ()):
                word += f', {k}: {sum(float(o[k]) for o in outputs):.3f}'
            print(word)
        if early_stop is not None and non_decrease_iters >= early_stop:
            break
    env.set_state(initial_state)
    return {'best_loss': best_loss, 'best_action': best_action, 'last_loss': last_loss, 'last_action': last_action,
            'iter_id': iter_id, 'optim_buffer': optim_buffer, 'action': action, 'optim': optim,
            'initial_state': initial_state, 'scheduler': scheduler}
