"""The analytic residual/Jacobian arithmetic the CUDA kernels run (csrc/mcba_math.cuh), compiled for the host
by tests/host_math, against the oracle's autodiff on random blocks: every stage, both camera models, the
refine_camera / refine_board flags and the small-angle branch of AngleAxisRotatePoint. This checks the math, not
the kernels (those are checked on the GPU in test_gpu_parity.py)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import mcba

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def mathlib():
    d = os.path.join(HERE, "host_math")
    subprocess.check_call(["make", "-s", "-C", d])
    L = C.CDLL(os.path.join(d, "libmathtest.so"))
    dp = mcba.c_f64p
    L.mathtest_obs.restype = C.c_double
    L.mathtest_obs.argtypes = [C.c_int] * 3 + [dp] * 6 + [C.c_double] + [dp] * 3
    L.mathtest_chol6.argtypes = [dp, dp]
    return L


def test_analytic_jacobian_matches_autodiff(mathlib, oracle):
    rng = np.random.default_rng(0)
    p = lambda a: a.ctypes.data_as(mcba.c_f64p)
    worst, n_checked, seen = 0.0, 0, set()
    for _ in range(3000):
        stage, model, flags = int(rng.integers(1, 6)), int(rng.integers(0, 2)), int(rng.integers(0, 4))
        small = rng.random() < 0.1
        cam = np.concatenate([rng.normal(size=3) * (1e-9 if small else 0.5), rng.normal(size=3) * 0.3])
        pose = np.concatenate([rng.normal(size=3) * (1e-10 if rng.random() < 0.1 else 0.8), rng.normal(size=3) * 0.3 + [0, 0, 2.5]])
        board = np.concatenate([rng.normal(size=3) * (0 if rng.random() < 0.1 else 0.5), rng.normal(size=3) * 0.3])
        intr = np.array([1400 + rng.normal() * 50, 1380 + rng.normal() * 50, 960 + rng.normal() * 10, 540 + rng.normal() * 10,
                         *(rng.normal(size=5) * [0.1, 0.05, 1e-3, 1e-3, 0.01])])
        if model == 1:
            intr[:2] = 600 + rng.normal(size=2) * 10
            intr[8] = 0
        xyz = np.array([rng.uniform(0, 0.8), rng.uniform(0, 0.6), 0.0 if rng.random() < 0.7 else rng.normal() * 0.1])
        proj, _ = oracle.functor(stage, model, flags, cam, pose, board, intr, xyz, np.zeros(2))
        if np.abs((proj - intr[2:4]) / intr[:2]).max() > 3:
            continue
        uv = proj + rng.normal(size=2) * (0.4 if rng.random() < 0.5 else 30.0)    # inliers and Huber outliers
        r_o, J_o = oracle.functor(stage, model, flags, cam, pose, board, intr, xyz, uv)
        rho = oracle.huber(1.0, float((r_o ** 2).sum()))
        w = np.sqrt(rho[1])
        K = mcba.STAGE_K[stage]
        r, J, co = np.zeros(2), np.zeros((2, K)), C.c_double()
        hr = mathlib.mathtest_obs(stage, model, flags, p(cam), p(pose), p(board), p(intr), p(xyz), p(uv), 1.0, p(r), p(J), C.byref(co))
        sc = max(1.0, np.abs(proj).max())
        errs = [np.abs(r - w * r_o).max() / sc, np.abs(J - w * J_o).max() / max(1.0, np.abs(J_o).max()),
                abs(hr - 0.5 * rho[0]) / max(1.0, rho[0], sc), abs(co.value - 0.5 * rho[0]) / max(1.0, rho[0], sc)]
        worst = max(worst, *errs)
        n_checked += 1
        seen.add((stage, model, flags & (1 if stage >= 3 else 0), small))
    assert n_checked > 2000 and worst <= 1e-11, worst


def test_chol6(mathlib):
    rng = np.random.default_rng(1)
    for _ in range(50):
        M = rng.normal(size=(6, 6)); A = M @ M.T + 0.1 * np.eye(6); b = rng.normal(size=6)
        Ac, bc = A.copy(), b.copy()
        assert mathlib.mathtest_chol6(Ac.ctypes.data_as(mcba.c_f64p), bc.ctypes.data_as(mcba.c_f64p)) == 1
        assert np.abs(bc - np.linalg.solve(A, b)).max() <= 1e-10 * np.abs(bc).max()
    A = -np.eye(6)
    assert mathlib.mathtest_chol6(A.ctypes.data_as(mcba.c_f64p), np.zeros(6).ctypes.data_as(mcba.c_f64p)) == 0
