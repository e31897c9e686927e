"""``MLP``: the closed-loop policy network of the reference (plb/engine/nn/mlp.py:12-183) over the CUDA engine.

The reference implements it as Taichi kernels writing into the tools' ``action_buffer`` so that ``ti.Tape`` differentiates
policy and simulator together (plb/optimizer/solver_nn.py:28-43).  Here the simulator's differentiable boundary is
``GradModel.forward`` (one torch autograd node per env step), so the policy is ordinary torch arithmetic between those nodes:
observation of step s -> layers -> clamp to [-1, 1] -> the action of step s.  Same constructor, same observation layout
(every ``obs_step``-th particle's position and velocity x ``velocity_weight``, then position + rotation of every tool,
mlp.py:63-87), same layer arithmetic (``W h + b``, relu / tanh on all but the last layer, :110-134), same flat parameter
layout for ``get_params / set_params / get_grad`` (W_0, b_0, W_1, b_1, ... [+ velocity_weight], :155-183).
"""
import numpy as np


class MLP:
    def __init__(self, simulator, primitives, hidden_dims, activation='relu', n_observed_particles=200, n_particles=None,
                 device=None, seed=0):
        import torch
        self.simulator, self.primitives = simulator, primitives
        for p in primitives:
            assert type(p).__name__ != 'Chopsticks', "Chopstick is not supported now.."      # mlp.py:28-29
        n_particle = int(n_particles if n_particles is not None else simulator.n_particles[None])
        assert n_particle > 0, "construct the MLP after the simulator holds particles (or pass n_particles)"
        self.n_observed_particles = n_observed_particles
        self.obs_step = max(1, n_particle // n_observed_particles)
        self.obs_num = n_particle // self.obs_step
        self.n_tools = len(primitives)
        inp_dim = self.obs_num * 6 + self.n_tools * 7     # the reference sizes it with primitives.state_dim but fills 7 per tool
        self.substeps = simulator.substeps
        self.dims = (inp_dim,) + tuple(hidden_dims) + (primitives.action_dim,)
        self.n_layer = len(self.dims) - 1
        self.activation = activation
        self.device = device or ('cuda' if torch.cuda.is_available() else 'cpu')
        g = torch.Generator().manual_seed(seed)
        self.W, self.b = [], []
        for i in range(self.n_layer):
            w = torch.randn((self.dims[i + 1], self.dims[i]), generator=g) * (1.0 / np.sqrt(self.dims[i]))
            self.W.append(w.to(self.device).requires_grad_(True))
            self.b.append(torch.zeros(self.dims[i + 1], device=self.device, requires_grad=True))
        self.velocity_weight = 1.0

    # ---- observation -> action (differentiable torch) ----------------------------------------------------------
    def observation(self, obs):
        """obs = (x [N, >= 6], c [K, 8]) as GradModel.get_obs returns it -> the network input (mlp.py:63-87)."""
        import torch
        x, c = obs[0], obs[1]
        idx = torch.arange(self.obs_num, device=x.device) * self.obs_step
        part = torch.cat([x[idx, :3], x[idx, 3:6] * self.velocity_weight], 1).reshape(-1)
        return torch.cat([part, c[:, :7].reshape(-1)]).to(torch.float32)

    def forward(self, obs):
        import torch
        h = self.observation(obs)
        for i in range(self.n_layer):
            h = self.W[i] @ h + self.b[i]
            if i != self.n_layer - 1:
                if self.activation == 'relu':
                    h = torch.relu(h)
                elif self.activation == 'tanh':
                    h = torch.tanh(h)
        return torch.clamp(h, -1.0, 1.0)                 # set_action: max(min(h, 1), -1), mlp.py:97

    __call__ = forward

    def rollout(self, func, horizon, step_loss, device=None):
        """The closed loop of solver_nn.py:28-43 over a GradModel: reset, then for every env step the policy acts on the
        current observation and the simulator advances; `step_loss(s, obs)` is added up.  Returns the (differentiable) loss."""
        obs = func.reset(device=device or self.device)
        loss = 0
        for s in range(horizon):
            obs = func.forward(s, self.forward(obs), *obs)
            loss = loss + step_loss(s, obs)
        return loss

    def set_action(self, s, n_substeps):
        """Open-loop use (mlp.py:143-153): act on the simulator's current state at env step s, without autograd."""
        import torch
        assert n_substeps == self.substeps
        f = s * self.substeps
        x = np.concatenate([self.simulator.get_x(f), self.simulator.get_v(f)], 1)
        c = np.stack([np.resize(np.asarray(p.get_state(f), np.float32), 8) for p in self.primitives])
        with torch.no_grad():
            a = self.forward((torch.as_tensor(x, dtype=torch.float32, device=self.device),
                              torch.as_tensor(c, dtype=torch.float32, device=self.device)))
        self.primitives.set_action(s, n_substeps, a.cpu().numpy().astype(np.float64))

    # ---- flat parameter interface (mlp.py:155-183) --------------------------------------------------------------
    def parameters(self):
        return [t for pair in zip(self.W, self.b) for t in pair]

    def zero_grad(self):
        for t in self.parameters():
            t.grad = None

    def get_grad(self):
        return np.concatenate([(t.grad if t.grad is not None else t * 0).detach().cpu().numpy().reshape(-1) for t in self.parameters()])

    def get_params(self):
        return np.concatenate([t.detach().cpu().numpy().reshape(-1) for t in self.parameters()])

    def set_params(self, param):
        import torch
        param = np.asarray(param, dtype=np.float32)
        with torch.no_grad():
            for i in range(self.n_layer):
                n = self.dims[i + 1] * self.dims[i]
                self.W[i].copy_(torch.as_tensor(param[:n].reshape(self.dims[i + 1], self.dims[i])))
                param = param[n:]
                n = self.dims[i + 1]
                self.b[i].copy_(torch.as_tensor(param[:n]))
                param = param[n:]
        if len(param) == 1:
            self.velocity_weight = float(param[-1])
        else:
            self.velocity_weight = 1.0
            assert len(param) == 0
