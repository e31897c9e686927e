"""The warp-aggregated scatters of the particle kernels on an emulated warp (no GPU).

Round 2 added the transposed shared-memory scatters (`warp_scatter27_ts`, `warp_scatter27_ts_affine`, `warp_scatter9_ts`: every
lane stores its contributions to a [node][lane] tile, lanes 0..26 add up the row segment of every run of equal cell keys);
they run here on the same patterns (modes 127 / 227 / 109).

`warp_scatter27` / `warp_scatter9` (diffskill_b200/csrc/kernels_common.cuh) choose between three regimes per warp from
the pattern of cell keys: a recursive-halving butterfly per group of lanes that share a cell, one segmented shuffle-down
reduction over all runs when there are more than two such groups, and per-lane reductions for strays.  `tests/host_check`
runs them under g++ on 32 lock-step host threads (simt_shim.h: ballots, match.any, shuffles, vector reductions) and this
file compares the resulting grid -- and the active-tile list P2G builds on the way -- with a plain per-particle
accumulation, for key patterns that drive every regime, inactive lanes included.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
HC_DIR = os.path.join(HERE, 'host_check')
CSRC = os.path.join(os.path.dirname(HERE), 'diffskill_b200', 'csrc')
N_GRID = 16
INV_DX = np.float32(N_GRID)


@pytest.fixture(scope='module')
def lib():
    so = os.path.join(HC_DIR, 'libhost_scatter.so')
    srcs = [os.path.join(HC_DIR, f) for f in ('host_scatter.cpp', 'simt_shim.h', 'cuda_shim.h')] + \
           [os.path.join(CSRC, f) for f in ('kernels_common.cuh', 'particle_math.cuh', 'mpm_math.cuh', 'tools.cuh', 'svd3.cuh')]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(['g++', '-O1', '-std=c++17', '-ffp-contract=off', '-fPIC', '-shared', '-pthread', '-w', '-U_FORTIFY_SOURCE', '-D_FORTIFY_SOURCE=0', '-o', so,
                               srcs[0]])
    return C.CDLL(so)


def _stencil(x):
    """bspline1 of kernels_common.cuh in numpy fp32: base cell (truncation, clamped) and the three weights per axis."""
    xg = (x.astype(np.float32) * INV_DX).astype(np.float32)
    b = np.clip((xg - np.float32(0.5)).astype(np.int32), 0, N_GRID - 3)      # astype(int) truncates toward zero
    fx = (xg - b.astype(np.float32)).astype(np.float32)
    a, c, d = np.float32(1.5) - fx, fx - np.float32(1.0), fx - np.float32(0.5)
    w = np.stack([np.float32(0.5) * (a * a), np.float32(0.75) - c * c, np.float32(0.5) * (d * d)], axis=-1).astype(np.float32)
    return b, w


def _reference(x, active, a):
    grid = np.zeros((N_GRID, N_GRID, N_GRID, 4))
    tiles = np.zeros((N_GRID // 4,) * 3, np.int32)
    b, w = _stencil(x)
    for p in range(len(x)):
        if not active[p]:
            continue
        for i in range(3):
            for j in range(3):
                for l in range(3):
                    ww = np.float32(np.float32(w[p, 0, i] * w[p, 1, j]) * w[p, 2, l])
                    v = a[p, 0, :3] + i * a[p, 1, :3] + j * a[p, 2, :3] + l * a[p, 3, :3]
                    X, Y, Z = b[p, 0] + i, b[p, 1] + j, b[p, 2] + l
                    grid[X, Y, Z, :3] += np.float64(ww) * v.astype(np.float32)
                    grid[X, Y, Z, 3] += np.float64(ww) * a[p, 0, 3]
                    tiles[X // 4, Y // 4, Z // 4] = 1
    return grid.reshape(-1, 4), tiles.reshape(-1)


def _cells_to_positions(cells, rng):
    """A particle somewhere inside each requested base cell (base = int(x * inv_dx - 0.5))."""
    frac = rng.uniform(0.05, 0.95, (len(cells), 3))
    return ((np.asarray(cells, np.float64) + 0.5 + frac) / N_GRID).astype(np.float32)


def _patterns(rng):
    A, B, Cc, D, E = (5, 6, 7), (5, 6, 8), (9, 3, 2), (0, 0, 0), (N_GRID - 3,) * 3
    yield 'one cell, full warp (butterfly)', [A] * 32, np.ones(32, bool)
    yield 'two cells 16/16 (two butterflies)', [A] * 16 + [B] * 16, np.ones(32, bool)
    yield 'four cells 8/8/8/8 (segmented reduction)', [A] * 8 + [B] * 8 + [Cc] * 8 + [D] * 8, np.ones(32, bool)
    yield 'five runs with strays between them (segmented + strays)', \
        [A] * 7 + [E] + [B] * 6 + [Cc] * 2 + [D] * 9 + [A] * 5 + [E] * 2, np.ones(32, bool)
    yield 'one big group, non-adjacent members, strays (match.any)', \
        [A] * 10 + [B] * 2 + [A] * 9 + [Cc] + [A] * 8 + [D] * 2, np.ones(32, bool)
    yield 'two groups + groups of three (per-lane reductions)', [A] * 12 + [B] * 3 + [Cc] * 14 + [D] * 3, np.ones(32, bool)
    act = np.ones(32, bool)
    act[27:] = False
    yield 'inactive tail (last warp of an env)', [A] * 20 + [B] * 12, act
    act = np.ones(32, bool)
    act[[3, 4, 17, 30]] = False
    yield 'inactive lanes inside runs', [A] * 9 + [B] * 9 + [Cc] * 9 + [D] * 5, act
    yield 'all lanes inactive', [A] * 32, np.zeros(32, bool)
    yield 'every lane its own cell (unsorted particles)', [tuple(rng.randint(0, N_GRID - 2, 3)) for _ in range(32)], np.ones(32, bool)
    cells = [tuple(c) for c in rng.randint(0, N_GRID - 2, (6, 3))]
    yield 'three warps, random runs of random cells', [cells[i] for i in np.sort(rng.randint(0, 6, 96))], rng.rand(96) > 0.1


@pytest.mark.parametrize('mode', [27, 9, 127, 227, 109])
def test_warp_scatter_equals_per_particle_accumulation(mode, lib):
    rng = np.random.RandomState(3)
    FP, IP = C.POINTER(C.c_float), C.POINTER(C.c_int)
    for name, cells, active in _patterns(rng):
        x = _cells_to_positions(cells, rng)
        b, _ = _stencil(x)
        assert np.array_equal(b, np.clip(np.asarray(cells), 0, N_GRID - 3)), name    # the pattern is what it says
        n = len(x)
        a = rng.normal(size=(n, 4, 4)).astype(np.float32)
        a[:, 0, 3] = rng.uniform(0.5, 1.5, n)                                          # "mass"
        grid = np.zeros((N_GRID ** 3, 4), np.float32)
        tiles = np.zeros((N_GRID // 4) ** 3, np.int32)
        act_i = np.ascontiguousarray(active, dtype=np.int32)
        lib.hc_scatter(N_GRID, C.c_float(float(INV_DX)), n, x.ctypes.data_as(FP), act_i.ctypes.data_as(IP), a.ctypes.data_as(FP),
                       mode, grid.ctypes.data_as(FP), tiles.ctypes.data_as(IP))
        ref, ref_tiles = _reference(x, active, a)
        scale = max(np.abs(ref).max(), 1e-30)
        assert np.abs(grid - ref).max() <= 2e-6 * scale, (mode, name, np.abs(grid - ref).max() / scale)
        assert np.array_equal(grid[:, 3] != 0, ref[:, 3] != 0), (mode, name, 'occupancy')
        assert np.array_equal(tiles, ref_tiles), (mode, name, 'every touched tile is in the active list exactly once')
