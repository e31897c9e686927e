"""``GradModel``: the torch.autograd bridge of the reference (plb/engine/function.py:11-241) over the CUDA engine.

One ``autograd.Function`` node per env step; the adjoint state between ``backward_step(s+1)`` and
``backward_step(s)`` lives inside the engine (adjoint checkpoints), exactly as it lives in Taichi's grad fields in
the reference -- the Function returns zeros for ``past_obs``.
"""
import numpy as np


class GradModel:
    def __init__(self, env, softness=666., init_sampler=None, output_grid=(), return_dist=False, env_index=0):
        self.env = env
        self.sim = env.simulator
        self.eng = self.sim.engine
        self.dim = self.sim.dim
        self.primitives = self.sim.primitives
        self.substeps = self.sim.substeps
        self.controllers = env.primitives
        self.softness = softness
        self._forward_func = None
        self.output_grid = output_grid
        self.return_dist = return_dist
        self.env_index = env_index
        self.init_state = self.env.get_state()['state']
        self.init_sampler = init_sampler if init_sampler is not None else (lambda: self.init_state)
        self.ncols = self.eng.ncols if return_dist else 0
        self.device = 'cuda'

    def reset(self, initial_states=None, device='cuda', clear_grad=True):
        self.device = device
        if initial_states is None:
            initial_states = self.init_sampler()
        if clear_grad:
            self.eng.zero_grad()                  # function.py:44-61
        self.env.set_state(initial_states, self.softness, False)
        return self.get_obs(0, self.device)

    def decay_kernel(self, f, alpha):             # function.py:66-77
        self.eng.scale_grad(self.sim._frame_to_step(f), float(alpha))

    def _bufs(self, device):
        """Device-resident staging tensors, allocated once per device: the per-step entry points below neither allocate
        nor synchronise (the reference's get_obs / set_obs_grad go through Taichi's to_torch / from_torch copies)."""
        import torch
        key = str(device)
        if getattr(self, '_buf_key', None) != key:
            eng = self.eng
            z = lambda *shape: torch.zeros(shape, device=device)
            self._buf = dict(xv=z(eng.B, eng.capacity, 6), c=z(eng.B, eng.K, 8), d=z(eng.B, eng.capacity, max(eng.ncols, 1)),
                             gx=z(eng.B, eng.capacity, 3), gv=z(eng.B, eng.capacity, 3), gt=z(eng.B, eng.K, 8),
                             gd=z(eng.B, eng.capacity, max(eng.ncols, 1)), ga=z(eng.B, max(eng.A, 1)))
            self._buf_key = key
            # gap.grad only exists for 8-dof tools (function.py:141-142)
            self._gap_mask = torch.tensor([1. if p.state_dim == 8 else 0. for p in self.primitives], device=device)
        return self._buf

    def get_obs(self, s, device):
        import torch
        eng, b = self.eng, self.env_index
        n = eng.n_particles(b)
        buf = self._bufs(device)
        xv, c = buf['xv'], buf['c']
        if xv.is_cuda:
            eng.get_obs(s, xv, c)
        else:
            a, bb = eng.get_obs(s)
            xv, c = torch.from_numpy(a), torch.from_numpy(bb)
        x = xv[b, :n]
        if self.return_dist:
            d = buf['d']
            if d.is_cuda:
                eng.compute_min_dist(s, d)
            else:
                d = torch.from_numpy(eng.compute_min_dist(s))
            x = torch.cat((x, d[b, :n]), 1)
        else:
            x = x.clone()
        outputs = x, c[b].clone()
        for _ in self.output_grid:
            self.sim.clear_and_compute_grid_m(s * self.substeps)
            outputs = outputs + (self.sim.grid_m.to_torch(device, env=b),)
        return outputs

    def set_obs_grad(self, s, obs_grad, manipulator_grad, *args):
        import torch
        eng, b = self.eng, self.env_index
        n = eng.n_particles(b)
        if len(self.output_grid) > 0:
            self.sim.grid_m.grad.from_torch(args[0], env=b)
            self.sim.compute_grid_m_kernel.grad(s * self.substeps)
        dev = obs_grad.device
        buf = self._bufs(dev)
        if self.return_dist:
            gd = buf['gd']
            gd[b, :n] = obs_grad[:, -eng.ncols:]
            eng.compute_min_dist_grad(s, gd if gd.is_cuda else gd.numpy())
            obs_grad = obs_grad[..., :-eng.ncols]
        obs_grad = obs_grad.reshape(-1, self.dim * 2)
        gx, gv, gt = buf['gx'], buf['gv'], buf['gt']     # rows of other envs / beyond n stay zero
        gx[b, :n], gv[b, :n] = obs_grad[:, :3], obs_grad[:, 3:6]
        gt[b] = manipulator_grad.reshape(eng.K, 8)
        gt[b, :, 7] *= self._gap_mask
        if gx.is_cuda:
            eng.add_particle_grad(s, gx, gv)
            eng.add_tool_grad(s, gt)
        else:
            eng.add_particle_grad(s, gx.numpy(), gv.numpy())
            eng.add_tool_grad(s, gt.numpy())

    def forward_step(self, s, a):
        a = a.reshape(-1).clamp(-1, 1)
        if a.is_cuda and self.eng.B == 1:
            self.eng.set_action(s, a.detach().float().contiguous().reshape(1, -1))
        else:
            self.sim._set_action(s, a.detach().cpu().numpy(), env=self.env_index if self.eng.B > 1 else None)
        self.eng.forward_step(s, s + 1, s)
        self.sim.cur = (s + 1) * self.substeps

    def backward_step(self, s):
        import torch
        self.eng.backward_step(s)
        if str(self.device).startswith('cuda'):
            ga = self._bufs(self.device)['ga']
            self.eng.get_action_grad(s, ga)                     # device -> device, no host round trip
            return ga[self.env_index, :self.eng.A].double()
        g = self.eng.get_action_grad(s)[self.env_index]
        return torch.tensor(g.astype(np.float64), device=self.device)

    def _make(self, gamma_lambda=None):
        import torch
        from torch.autograd import Function
        model = self

        class forward(Function):
            @staticmethod
            def forward(ctx, s, a, *past_obs):
                ctx.save_for_backward(torch.tensor([s]), *[torch.zeros_like(i) for i in past_obs])
                model.forward_step(s, a)
                return model.get_obs(s + 1, a.device)

            @staticmethod
            def backward(ctx, *obs_grad):
                tmp = ctx.saved_tensors
                s = tmp[0].item()
                model.set_obs_grad(s + 1, *obs_grad)
                if gamma_lambda is not None:
                    model.decay_kernel((s + 1) * model.substeps, gamma_lambda)
                actor_grad = model.backward_step(s)
                return (None, actor_grad.reshape(-1).to(obs_grad[0].dtype)) + tmp[1:]

        return forward.apply

    @property
    def forward(self):
        if self._forward_func is None:
            self._forward_func = self._make()
        return self._forward_func

    def make_func(self, γ, λ):
        return self._make(γ * λ)

    def render(self, mode='human', f=0, **kwargs):
        assert f == 0
        raise NotImplementedError("rendering is out of scope of the engine")
