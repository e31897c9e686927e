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

    def get_obs(self, s, device):
        import torch
        eng, b = self.eng, self.env_index
        n = eng.n_particles(b)
        xv = torch.zeros((eng.B, eng.capacity, 6), device=device)
        c = torch.zeros((eng.B, eng.K, 8), device=device)
        if xv.is_cuda:
            eng.get_obs(s, xv, c)
        else:
            a, bb = eng.get_obs(s)
            xv, c = torch.from_numpy(a), torch.from_numpy(bb)
        x = xv[b, :n]
        if self.return_dist:
            d = torch.zeros((eng.B, eng.capacity, eng.ncols), device=device)
            if d.is_cuda:
                eng.compute_min_dist(s, d)
            else:
                d = torch.from_numpy(eng.compute_min_dist(s))
            x = torch.cat((x, d[b, :n]), 1)
        outputs = x.clone(), c[b].clone()
        for _ in self.output_grid:
            self.sim.clear_and_compute_grid_m(s * self.substeps)
            outputs = outputs + (self.sim.grid_m.to_torch(device),)
        return outputs

    def set_obs_grad(self, s, obs_grad, manipulator_grad, *args):
        import torch
        eng, b = self.eng, self.env_index
        n = eng.n_particles(b)
        if len(self.output_grid) > 0:
            self.sim.grid_m.grad.from_torch(args[0])
            self.sim.compute_grid_m_kernel.grad(s * self.substeps)
        dev = obs_grad.device
        if self.return_dist:
            gd = torch.zeros((eng.B, eng.capacity, eng.ncols), device=dev)
            gd[b, :n] = obs_grad[:, -eng.ncols:]
            eng.compute_min_dist_grad(s, gd if gd.is_cuda else gd.numpy())
            obs_grad = obs_grad[..., :-eng.ncols]
        obs_grad = obs_grad.reshape(-1, self.dim * 2)
        gx = torch.zeros((eng.B, eng.capacity, 3), device=dev)
        gv = torch.zeros((eng.B, eng.capacity, 3), device=dev)
        gx[b, :n], gv[b, :n] = obs_grad[:, :3], obs_grad[:, 3:6]
        gt = torch.zeros((eng.B, eng.K, 8), device=dev)
        gt[b] = manipulator_grad.reshape(eng.K, 8)
        for i, p in enumerate(self.primitives):
            if p.state_dim != 8:
                gt[b, i, 7] = 0.                  # gap.grad only exists for 8-dof tools (function.py:141-142)
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
