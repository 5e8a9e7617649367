"""Pins the CPU oracle's functors (oracle/oracle.cpp, restating OptimizationCeres.h) against independent
evaluators: the committed mpmath known-answer vectors, OpenCV's projectPoints (values + analytic Jacobians)
and the published closed forms of ceres::HuberLoss / AngleAxisRotatePoint."""
import os

import cv2
import numpy as np
import pytest

import mcba
from util import rel

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def kat():
    return np.load(os.path.join(GOLD, "functor_kat.npz"))


def test_functor_kat_mpmath(oracle, kat):
    """Residual and autodiff Jacobian of all five functors vs the 50-digit mpmath evaluation. Tolerance 1e-11 relative
    to the largest Jacobian entry (fp64 forward-mode AD of a 1e3-pixel projection)."""
    n = len(kat["stage"])
    seen = set()
    for i in range(n):
        st, md, fl = int(kat["stage"][i]), int(kat["model"][i]), int(kat["flags"][i])
        r, J = oracle.functor(st, md, fl, kat["cam"][i], kat["pose"][i], kat["board"][i], kat["intr"][i], kat["xyz"][i], kat["uv"][i])
        K = mcba.STAGE_K[st]
        assert np.abs(r - kat["r"][i]).max() <= 1e-9, (i, st, md, fl)
        assert rel(J, kat["J"][i][:, :K]) <= 1e-11, (i, st, md, fl)
        seen.add((st, md))
    assert len(seen) == 10   # every stage x camera model is covered


def test_reference_flags_zero_columns(oracle, kat):
    """refine_camera=false / refine_board=false give structurally zero Jacobian blocks (OptimizationCeres.h:401-427),
    and the Kannala intrinsics column 8 is unused."""
    i = 0
    for st, (cam_off, board_off, intr_off) in {4: (0, 12, None), 5: (0, 12, 18)}.items():
        for md in (0, 1):
            r, J = oracle.functor(st, md, 3, kat["cam"][i], kat["pose"][i], kat["board"][i], kat["intr"][i], kat["xyz"][i], kat["uv"][i])
            assert np.all(J[:, cam_off:cam_off + 6] == 0) and np.all(J[:, board_off:board_off + 6] == 0)
            assert np.any(J[:, 6:12] != 0)
            if intr_off is not None and md == 1:
                assert np.all(J[:, intr_off + 8] == 0)


def test_stage1_brown_matches_opencv(oracle):
    """ReprojectionError (OptimizationCeres.h:11-117), Brown: cv2.projectPoints values and its analytic Jacobian
    (columns rvec, tvec, f, c, k1 k2 p1 p2 k3 = the functor's parameter order)."""
    rng = np.random.default_rng(3)
    for _ in range(50):
        pose = np.concatenate([rng.normal(size=3) * 0.6, rng.normal(size=3) * 0.2 + [0, 0, 2.0]])
        intr = np.array([1400.0, 1390.0, 960.0, 540.0, -0.12, 0.06, 5e-4, -3e-4, -0.01]) * (1 + 0.05 * rng.normal(size=9))
        xyz = np.array([rng.uniform(0, 0.8), rng.uniform(0, 0.6), 0.0])
        Kmat = np.array([[intr[0], 0, intr[2]], [0, intr[1], intr[3]], [0, 0, 1.0]])
        img, jac = cv2.projectPoints(xyz.reshape(1, 1, 3), pose[:3], pose[3:], Kmat, intr[4:9])
        r, J = oracle.functor(1, 0, 0, None, pose, None, intr, xyz, np.zeros(2))
        assert np.abs(r - img.reshape(2)).max() <= 1e-9
        assert rel(J, jac[:, :15]) <= 1e-9


def test_stage1_kannala_matches_opencv_fisheye(oracle):
    """Fisheye branch (OptimizationCeres.h:63-101) vs cv2.fisheye.projectPoints; its Jacobian columns are
    f(2) c(2) k(4) rvec(3) tvec(3) alpha(1)."""
    rng = np.random.default_rng(4)
    for _ in range(50):
        pose = np.concatenate([rng.normal(size=3) * 0.6, rng.normal(size=3) * 0.2 + [0, 0, 1.2]])
        intr = np.array([600.0, 598.0, 960.0, 540.0, -0.02, 3e-3, -1e-3, 2e-4, 0.0]) * (1 + 0.05 * rng.normal(size=9))
        xyz = np.array([rng.uniform(0, 0.8), rng.uniform(0, 0.6), 0.0])
        Kmat = np.array([[intr[0], 0, intr[2]], [0, intr[1], intr[3]], [0, 0, 1.0]])
        img, jac = cv2.fisheye.projectPoints(xyz.reshape(1, 1, 3), pose[:3].reshape(3, 1), pose[3:].reshape(3, 1), Kmat, intr[4:8].reshape(4, 1))
        r, J = oracle.functor(1, 1, 0, None, pose, None, intr, xyz, np.zeros(2))
        assert np.abs(r - img.reshape(2)).max() <= 1e-9
        ours = np.concatenate([J[:, 6:14], J[:, 0:6]], 1)     # -> f c k rvec tvec
        assert rel(ours, jac[:, :14]) <= 1e-8


def test_angle_axis_rotate(oracle):
    """ceres::AngleAxisRotatePoint: agrees with cv2.Rodrigues away from 0 and switches to p + w x p (derivative -[p]x)
    when theta^2 <= DBL_EPSILON."""
    rng = np.random.default_rng(5)
    for _ in range(20):
        w, p = rng.normal(size=3), rng.normal(size=3)
        out, d_aa, d_pt = oracle.angle_axis_rotate(w, p)
        R, jacR = cv2.Rodrigues(w)
        assert np.abs(out - R @ p).max() <= 1e-14
        assert np.abs(d_pt - R).max() <= 1e-14
        # cv2: jacobian is 3x9, d R(row-major) / d w
        dR = jacR.reshape(3, 3, 3)
        assert np.abs(d_aa - np.einsum("kij,j->ik", dR, p)).max() <= 1e-12
    w = np.array([1e-9, -2e-9, 3e-9]); p = np.array([0.3, -0.2, 0.9])
    out, d_aa, d_pt = oracle.angle_axis_rotate(w, p)
    assert np.abs(out - (p + np.cross(w, p))).max() == 0
    px = np.array([[0, -p[2], p[1]], [p[2], 0, -p[0]], [-p[1], p[0], 0]])
    assert np.abs(d_aa + px).max() == 0


def test_huber(oracle):
    """ceres::HuberLoss(a): rho(s) = s for s <= a^2, 2 a sqrt(s) - a^2 beyond; rho' = a/sqrt(s); rho'' = -rho'/(2 s)."""
    assert np.allclose(oracle.huber(1.0, 0.25), [0.25, 1.0, 0.0])
    assert np.allclose(oracle.huber(1.0, 1.0), [1.0, 1.0, 0.0])
    rho = oracle.huber(1.0, 16.0)
    assert np.allclose(rho, [7.0, 0.25, -0.25 / 32.0])
    rho = oracle.huber(2.0, 100.0)
    assert np.allclose(rho, [36.0, 0.2, -0.2 / 200.0])
