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


def test_copy_free_context_gives_the_same_jacobian(mathlib):
    """Round-2 fused kernels (csrc/mcba_obs_fused.cuh): the context is the raw prep records (identity record for a block
    that is not refined) + 63 three-term products (ctx_raw_operands), instead of 144 entries built one by one. Same
    residuals, Jacobians and costs as the entry-by-entry context - which is the one checked against the oracle's autodiff
    above - for every stage, model and reference-camera / reference-board flag, small angles included."""
    dp = mcba.c_f64p
    mathlib.mathtest_obs_raw.restype = C.c_double
    mathlib.mathtest_obs_raw.argtypes = [C.c_int] * 3 + [dp] * 6 + [C.c_double] + [dp] * 2
    rng = np.random.default_rng(5)
    p = lambda a: a.ctypes.data_as(dp)
    seen = set()
    for _ in range(2000):
        stage, model, flags = int(rng.integers(1, 6)), int(rng.integers(0, 2)), int(rng.integers(0, 4))
        small = rng.random() < 0.15
        cam = np.concatenate([rng.normal(size=3) * (0.0 if small else 0.5), rng.normal(size=3) * 0.3])
        pose = np.concatenate([rng.normal(size=3) * (1e-10 if rng.random() < 0.1 else 0.8), rng.normal(size=3) * 0.3 + [0, 0, 2.5]])
        board = np.concatenate([rng.normal(size=3) * (0 if rng.random() < 0.1 else 0.5), rng.normal(size=3) * 0.3])
        intr = np.array([1400 + rng.normal() * 50, 1380 + rng.normal() * 50, 960, 540, *(rng.normal(size=5) * [0.1, 0.05, 1e-3, 1e-3, 0.01])])
        if model == 1:
            intr[:2] = 600; intr[8] = 0
        xyz = np.array([rng.uniform(0, 0.8), rng.uniform(0, 0.6), 0.0 if rng.random() < 0.7 else rng.normal() * 0.1])
        uv = np.array([960.0, 540.0]) + rng.normal(size=2) * 300
        K = mcba.STAGE_K[stage]
        r0, J0, r1, J1, co = np.zeros(2), np.zeros((2, K)), np.zeros(2), np.zeros((2, K)), C.c_double()
        h0 = mathlib.mathtest_obs(stage, model, flags, p(cam), p(pose), p(board), p(intr), p(xyz), p(uv), 1.0, p(r0), p(J0), C.byref(co))
        h1 = mathlib.mathtest_obs_raw(stage, model, flags, p(cam), p(pose), p(board), p(intr), p(xyz), p(uv), 1.0, p(r1), p(J1))
        if not np.isfinite(J0).all():
            continue                                   # point behind the camera etc.: nothing to compare
        sc = max(1.0, np.abs(J0).max())
        assert np.abs(J1 - J0).max() <= 1e-13 * sc and np.abs(r1 - r0).max() <= 1e-13 * max(1.0, np.abs(r0).max())
        assert abs(h1 - h0) <= 1e-13 * max(1.0, abs(h0))
        seen.add((stage, model, flags))
    assert len(seen) == 5 * 2 * 4


@pytest.mark.parametrize("stage", [1, 2, 3, 4, 5])
def test_group_pair_cover_writes_every_record_entry_exactly_once(mathlib, stage):
    """The DMMA products of the fused kernels cover the upper triangle of J~^T J~ by pairs of 4-column groups (stage 5:
    the 7 lines of the Fano plane, 7 products instead of 10 aligned tiles). Through the accumulator layout of
    mma.m8n8k4 and the lane -> record-index map the kernels keep in shared memory (k1_store_index), every entry of the
    compact record must be written by exactly one (product, lane, accumulator entry) and nothing else may be written."""
    mathlib.mathtest_cover.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    mathlib.mathtest_cidx.argtypes = [C.c_int] * 3
    wf = {1: 9, 2: 6, 3: 6, 4: 12, 5: 21}[stage]
    nc = wf + 7
    ncomp = wf * (wf + 1) // 2 + wf + 6 * wf + 27
    count = (C.c_int * ncomp)()
    groups = (C.c_int * 64)()
    n_prod = C.c_int()
    assert mathlib.mathtest_cover(stage, count, C.byref(n_prod), groups) == ncomp
    assert list(count) == [1] * ncomp
    # the record holds every pair (a <= b) of the nc columns except residual x residual: the map is onto and one to one
    idx = sorted(mathlib.mathtest_cidx(stage, a, b) for a in range(nc) for b in range(a, nc) if not (a == b == nc - 1))
    assert idx == list(range(ncomp))
    # the products of the cover: every unordered pair of groups (and every group with itself) appears in some product
    ng = (nc + 3) // 4
    assert n_prod.value == {4: 3, 5: 5, 7: 7}[ng]
    covered = set()
    for r in range(n_prod.value):
        r0, r1, c0, c1 = groups[4 * r:4 * r + 4]
        covered |= {tuple(sorted(x)) for x in ((r0, c0), (r0, c1), (r1, c0), (r1, c1))}
    assert covered == {(a, b) for a in range(ng) for b in range(a, ng)}
    if ng == 7:                                         # Fano plane: 7 products x 4 blocks = the 28 blocks, none twice
        assert n_prod.value * 4 == len(covered)
