"""The pose-initialisation arithmetic the CUDA kernels run (csrc/mcba_pnp_math.cuh), compiled for the host by
tests/host_math, against OpenCV (the third-party library the reference calls for this step: cv::solvePnP with
SOLVEPNP_P3P / SOLVEPNP_ITERATIVE, cv::undistortPoints, cv::fisheye::undistortPoints, cv::Rodrigues;
/root/reference/McCalib/src/geometrytools.cpp:233-326, 709-751). cv2 is the checker here, never the product."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import mcba

cv2 = pytest.importorskip("cv2")
HERE = os.path.dirname(os.path.abspath(__file__))
dp = mcba.c_f64p


@pytest.fixture(scope="module")
def lib():
    d = os.path.join(HERE, "host_math")
    subprocess.check_call(["make", "-s", "-C", d])
    L = C.CDLL(os.path.join(d, "libmathtest.so"))
    L.pnptest_undistort.argtypes = [C.c_int, dp, C.c_double, C.c_double, dp]
    L.pnptest_p3p.argtypes = [dp, dp, dp, dp]
    L.pnptest_hypothesis.argtypes = [dp, dp, dp, dp, dp, dp]
    L.pnptest_rvec.argtypes = [dp, dp]
    L.pnptest_sample4.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.POINTER(C.c_int)]
    L.pnptest_replay.argtypes = [C.POINTER(C.c_int), C.c_int, C.c_int, C.c_int, C.c_double, C.POINTER(C.c_int)]
    L.pnptest_residual_jac.argtypes = [dp, dp, dp, C.c_double, C.c_double, dp, dp]
    return L


def P(a):
    return a.ctypes.data_as(dp)


def random_pose(rng, near_pi=False):
    ax = rng.normal(size=3)
    ax /= np.linalg.norm(ax)
    ang = (np.pi - abs(rng.normal()) * 1e-3) if near_pi else rng.uniform(0.0, 3.0)
    return ax * ang, np.array([rng.normal() * 0.2, rng.normal() * 0.2, rng.uniform(0.8, 3.0)])


def test_rvec_matches_rodrigues(lib):
    rng = np.random.default_rng(0)
    worst = 0.0
    for k in range(2000):
        rv, _ = random_pose(rng, near_pi=(k % 4 == 0))
        if k % 50 == 0:
            rv = rv * 1e-9
        R, _ = cv2.Rodrigues(rv)
        out = np.zeros(3)
        lib.pnptest_rvec(P(np.ascontiguousarray(R.reshape(-1))), P(out))
        R2, _ = cv2.Rodrigues(out)
        worst = max(worst, np.abs(R2 - R).max())
    assert worst < 1e-9


def test_undistort_inverts_the_projection_model(lib):
    """(a, b) -> pixel through cv2.projectPoints / cv2.fisheye.projectPoints, back through undistort_pixel; and the
    same points through cv2's own undistortPoints (looser: OpenCV stops the Brown fixed point after 5 sweeps)."""
    rng = np.random.default_rng(1)
    K = np.array([[1400.0, 0, 960], [0, 1380, 540], [0, 0, 1]])
    for model in (0, 1):
        d = np.array([0.1, -0.05, 1e-3, -1e-3, 0.01]) if model == 0 else np.array([0.02, -0.01, 0.003, -0.001])
        intr = np.array([K[0, 0], K[1, 1], K[0, 2], K[1, 2], *d, *([0.0] if model else [])])
        ab = rng.uniform(-0.6, 0.6, size=(500, 2))
        X = np.concatenate([ab, np.ones((500, 1))], axis=1)
        if model == 0:
            uv, _ = cv2.projectPoints(X, np.zeros(3), np.zeros(3), K, d)
        else:
            uv, _ = cv2.fisheye.projectPoints(X.reshape(1, -1, 3), np.zeros(3), np.zeros(3), K, d)
        uv = uv.reshape(-1, 2)
        got = np.zeros((500, 2))
        for i in range(500):
            assert lib.pnptest_undistort(model, P(intr), uv[i, 0], uv[i, 1], P(got[i]))
        assert np.abs(got - ab).max() < 1e-11
        if model == 0:
            cvu = cv2.undistortPoints(uv.reshape(-1, 1, 2), K, d).reshape(-1, 2)
        else:
            cvu = cv2.fisheye.undistortPoints(uv.reshape(-1, 1, 2), K, d).reshape(-1, 2)
        assert np.abs(got - cvu).max() < 2e-4


def test_p3p_contains_the_true_pose_and_matches_cv2(lib):
    """Random triangles and poses: every returned (R, t) is a rotation that maps the points onto their rays, the true
    pose is among them, and the solution set agrees with cv2.solveP3P."""
    rng = np.random.default_rng(2)
    n_sol_hist = np.zeros(5, int)
    for trial in range(1500):
        rv, tv = random_pose(rng)
        R, _ = cv2.Rodrigues(rv)
        planar = trial % 2 == 0
        Pw = rng.uniform(-0.3, 0.3, size=(3, 3))
        if planar:
            Pw[:, 2] = 0.0
        Xc = Pw @ R.T + tv
        if np.any(Xc[:, 2] < 0.2):
            continue
        f = Xc / np.linalg.norm(Xc, axis=1, keepdims=True)
        Rs, ts = np.zeros((4, 9)), np.zeros((4, 3))
        n = lib.pnptest_p3p(P(np.ascontiguousarray(f)), P(np.ascontiguousarray(Pw)), P(Rs), P(ts))
        assert 1 <= n <= 4
        n_sol_hist[n] += 1
        found = False
        for s in range(n):
            Rm = Rs[s].reshape(3, 3)
            assert np.abs(Rm @ Rm.T - np.eye(3)).max() < 1e-7 and abs(np.linalg.det(Rm) - 1) < 1e-7
            Y = Pw @ Rm.T + ts[s]
            Yn = Y / np.linalg.norm(Y, axis=1, keepdims=True)
            assert np.abs(Yn - f).max() < 1e-7
            if np.abs(Rm - R).max() < 1e-6 and np.abs(ts[s] - tv).max() < 1e-6:
                found = True
        assert found, (trial, n)
        # cv2: same solution set (float32 inputs -> loose tolerance), on a subset to keep the test short
        if trial % 10 == 0:
            uvn = (Xc[:, :2] / Xc[:, 2:3])
            nc, rvs, tvs = cv2.solveP3P(Pw.astype(np.float32), uvn.astype(np.float32), np.eye(3), None, flags=cv2.SOLVEPNP_P3P)
            for r_cv, t_cv in zip(rvs, tvs):
                Rc, _ = cv2.Rodrigues(r_cv)
                Yc = Pw @ Rc.T + t_cv.ravel()
                if np.abs(Yc[:, :2] / Yc[:, 2:3] - uvn).max() > 1e-4:
                    continue     # an inaccurate OpenCV root: nothing to compare
                d = min(np.abs(Rs[s].reshape(3, 3) - Rc).max() + np.abs(ts[s] - t_cv.ravel()).max() for s in range(n))
                assert d < 5e-3, (trial, d)
    assert n_sol_hist[2:].sum() > 100      # multi-solution configurations were exercised


def test_hypothesis_matches_cv2_solvepnp_p3p(lib):
    """The four-point hypothesis (three points + disambiguation) against cv2.solvePnP(..., SOLVEPNP_P3P) on board corners
    seen through a distorted camera."""
    rng = np.random.default_rng(3)
    K = np.array([[1400.0, 0, 960], [0, 1380, 540], [0, 0, 1]])
    d = np.array([0.1, -0.05, 1e-3, -1e-3, 0.01])
    intr = np.array([K[0, 0], K[1, 1], K[0, 2], K[1, 2], *d])
    grid = np.array([[x * 0.04, y * 0.04, 0.0] for y in range(6) for x in range(8)])
    n_cmp = 0
    for trial in range(400):
        rv, tv = random_pose(rng)
        idx = rng.choice(len(grid), 4, replace=False)
        X = grid[idx]
        if np.linalg.norm(np.cross(X[1] - X[0], X[2] - X[0])) < 1e-4:
            continue
        uv, _ = cv2.projectPoints(X, rv, tv, K, d)
        uv = uv.reshape(-1, 2)
        Rm, _ = cv2.Rodrigues(rv)
        if np.any((X @ Rm.T + tv)[:, 2] < 0.3) or np.abs(uv - [960, 540]).max() > 900:
            continue
        ab = np.zeros((4, 2))
        for i in range(4):
            assert lib.pnptest_undistort(0, P(intr), uv[i, 0], uv[i, 1], P(ab[i]))
        r_out, t_out = np.zeros(3), np.zeros(3)
        ok = lib.pnptest_hypothesis(P(ab), P(np.ascontiguousarray(X)), P(intr), P(np.ascontiguousarray(uv[3])), P(r_out), P(t_out))
        assert ok
        R_out, _ = cv2.Rodrigues(r_out)
        assert np.abs(R_out - Rm).max() < 1e-6 and np.abs(t_out - tv).max() < 1e-6   # exact data: the true pose wins
        okc, r_cv, t_cv = cv2.solvePnP(X.astype(np.float32), uv.astype(np.float32), K, d, flags=cv2.SOLVEPNP_P3P)
        if okc:
            R_cv, _ = cv2.Rodrigues(r_cv)
            if np.abs(R_cv - Rm).max() < 1e-3:        # OpenCV itself found the true pose from float32 inputs
                assert np.abs(R_cv - R_out).max() < 1e-3 and np.abs(t_cv.ravel() - t_out).max() < 2e-3
                n_cmp += 1
    assert n_cmp > 200


def test_sample4_distinct_and_uniform(lib):
    idx = (C.c_int * 4)()
    counts = np.zeros(16)
    for h in range(4000):
        lib.pnptest_sample4(12345, 7, h, 16, idx)
        v = list(idx)
        assert len(set(v)) == 4 and min(v) >= 0 and max(v) < 16
        counts[v] += 1
    assert counts.min() > 800 and counts.max() < 1200      # 1000 expected per index
    lib.pnptest_sample4(1, 0, 0, 4, idx)
    assert sorted(idx) == [0, 1, 2, 3]


def replay_reference(counts, total, it, p):
    """ransacP3P's loop (geometrytools.cpp:243-300) restated on a list of inlier counts."""
    N, trial, countit, best, best_h, h = it, 0, 0, 0, -1, 0
    while N > trial and countit < it and h < len(counts):
        nb = counts[h]
        trial += 1
        if nb > best:
            best, best_h = nb, h
            frac = best / float(total)
            pno = 1 - frac ** 3
            if pno == 0:
                pno = 0.00001
            if pno > 1 - 0.00001:
                pno = 1 - 0.00001
            N = int(np.floor(np.log(1 - p) / np.log(pno) + 0.5))
            trial = 0
        countit += 1
        h += 1
    return best_h, h


def test_stopping_rule(lib):
    rng = np.random.default_rng(4)
    for _ in range(500):
        total = int(rng.integers(4, 60))
        it = int(rng.integers(1, 80))
        counts = rng.integers(0, total + 1, size=100).astype(np.int32)
        if rng.random() < 0.3:
            counts = np.minimum(counts, total // 3)
        used = C.c_int(0)
        got = lib.pnptest_replay(counts.ctypes.data_as(C.POINTER(C.c_int)), 100, total, it, 0.99, C.byref(used))
        assert (got, used.value) == replay_reference(list(counts), total, it, 0.99)


def test_pose_jacobian_matches_cv2(lib):
    rng = np.random.default_rng(5)
    K = np.array([[1400.0, 0, 960], [0, 1380, 540], [0, 0, 1]])
    d = np.array([0.1, -0.05, 1e-3, -1e-3, 0.01])
    intr = np.array([K[0, 0], K[1, 1], K[0, 2], K[1, 2], *d])
    worst = 0.0
    for _ in range(300):
        rv, tv = random_pose(rng)
        X = rng.uniform(-0.3, 0.3, size=3)
        uv, jac = cv2.projectPoints(X.reshape(1, 3), rv, tv, K, d)
        r, J = np.zeros(2), np.zeros(12)
        pose = np.concatenate([rv, tv])
        if not lib.pnptest_residual_jac(P(pose), P(intr), P(X), 100.0, 200.0, P(r), P(J)):
            continue
        assert np.abs(r - (uv.reshape(2) - [100.0, 200.0])).max() < 1e-8
        worst = max(worst, np.abs(J.reshape(2, 6) - jac[:, :6]).max() / max(1.0, np.abs(jac[:, :6]).max()))
    assert worst < 1e-8
