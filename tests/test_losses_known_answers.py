"""The torch-side losses that feed the adjoint (taichi_env.py:23-26, 246-275).  `geomloss` / `pykeops`
are not installable offline, so `sinkhorn_emd` stands in for `SamplesLoss('sinkhorn', p=1, blur=0.001)`.  These are
known-answer tests for it: with p = 1 and blur -> 0 the debiased Sinkhorn divergence between two uniform clouds of equal size
converges to the earth mover's distance, which scipy's Hungarian solver gives exactly; its gradient with respect to the points
is the (unit-direction) transport displacement of the optimal matching.  Chamfer is checked against a brute-force loop."""
import numpy as np
import pytest
import torch
from scipy.optimize import linear_sum_assignment

from diffskill_b200.sim.taichi_env import chamfer_loss, sinkhorn_emd


def _exact_emd(x, y):
    C = np.linalg.norm(x[:, None] - y[None], axis=-1)
    r, c = linear_sum_assignment(C)
    return C[r, c].mean(), c


@pytest.mark.parametrize('seed,n', [(0, 60), (1, 200), (2, 500)])
def test_sinkhorn_stand_in_converges_to_the_exact_emd(seed, n):
    rng = np.random.RandomState(seed)
    x = rng.uniform(0.3, 0.7, (n, 3))                      # a dough-sized cloud in the unit box
    y = x[rng.permutation(n)] * np.array([1.3, 0.6, 1.3]) + np.array([-0.08, 0.01, 0.02]) + rng.normal(size=(n, 3)) * 0.01
    exact, match = _exact_emd(x, y)
    xt = torch.tensor(x, dtype=torch.float64, requires_grad=True)
    val = sinkhorn_emd(xt, torch.tensor(y, dtype=torch.float64), blur=0.001, p=1)
    val.backward()
    print(f'n={n}: Sinkhorn divergence {float(val):.6f}  exact EMD {exact:.6f}')
    assert abs(float(val) - exact) < 0.02 * exact + 1e-3   # entropic bias at blur = 1e-3 on a 0.1-sized displacement
    # gradient: d EMD / d x_i = (x_i - y_match(i)) / |x_i - y_match(i)| / n ; compare directions where the displacement is
    # well above the blur
    d = x - y[match]
    far = np.linalg.norm(d, axis=1) > 0.03
    g = xt.grad.numpy()[far]
    ref = d[far] / np.linalg.norm(d[far], axis=1, keepdims=True) / n
    cos = (g * ref).sum(1) / (np.linalg.norm(g, axis=1) * np.linalg.norm(ref, axis=1) + 1e-30)
    assert np.median(cos) > 0.9 and np.isfinite(xt.grad.numpy()).all()
    assert abs(np.linalg.norm(g, axis=1).mean() * n - 1.0) < 0.15   # unit transport direction / n: no double counting
    # and the autograd gradient is the derivative of the value it returns (central differences on one point)
    i = int(np.argmax(far))
    fd = np.zeros(3)
    for k in range(3):
        e = np.zeros_like(x)
        e[i, k] = 1e-5
        fd[k] = (float(sinkhorn_emd(torch.tensor(x + e), torch.tensor(y))) - float(sinkhorn_emd(torch.tensor(x - e), torch.tensor(y)))) / 2e-5
    assert np.abs(xt.grad.numpy()[i] - fd).max() < 0.15 * np.abs(fd).max()


def test_sinkhorn_is_zero_on_identical_clouds_and_symmetric():
    rng = np.random.RandomState(3)
    x = torch.tensor(rng.uniform(0.3, 0.7, (150, 3)))
    y = torch.tensor(rng.uniform(0.3, 0.7, (150, 3)))
    assert abs(float(sinkhorn_emd(x, x.clone()))) < 1e-6
    assert float(sinkhorn_emd(x, y)) == pytest.approx(float(sinkhorn_emd(y, x)), rel=2e-2)   # the alternating updates end on different half-steps
    assert float(sinkhorn_emd(x, y)) > 0


@pytest.mark.parametrize('bidirectional', [False, True])
def test_chamfer_matches_brute_force(bidirectional):
    rng = np.random.RandomState(4)
    a, b = rng.normal(size=(40, 3)), rng.normal(size=(55, 3))
    want = sum(min(((p - q) ** 2).sum() for q in b) for p in a)
    if bidirectional:
        want += sum(min(((p - q) ** 2).sum() for p in a) for q in b)
    got = chamfer_loss(bidirectional)(torch.tensor(a)[None], torch.tensor(b)[None])
    assert float(got) == pytest.approx(want, rel=1e-10)
