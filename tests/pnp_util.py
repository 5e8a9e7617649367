"""Shared helpers of the pose-initialisation tests: seeded problems, the comparison rule, the host harness."""
import ctypes as C
import os
import subprocess

import numpy as np

import mcba
from mcba import synth
from mcba.pnp import PnpProblem, mcba_pnp_options

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden", "pnp_golden.npz")
# (config, frames, threshold px, intrinsics perturbation): thresholds below the 20 px outliers and near the 0.3 px noise
# floor exercise multi-draw RANSAC; c2 has Kannala cameras.
CASES = {"c2_t10": ("c2", 40, 10.0, 0.0), "c2_t1": ("c2", 40, 1.0, 0.0), "c1_t2": ("c1", 100, 2.0, 0.0),
         "c3_t3": ("c3", 4, 3.0, 0.0), "c2_t3_pert": ("c2", 30, 3.0, 0.01)}
SEED = 42


def make_case(name):
    cfg, nf, thr, pert = CASES[name]
    s = synth.make_scene(cfg, n_frames=nf)
    o = synth.generate_observations(s)
    prob, init, gt = synth.make_problem(s, o, 1)
    intr = s.intr_gt.copy()
    if pert:
        rng = np.random.default_rng(7)
        intr[:, :2] *= 1 + pert * rng.standard_normal((s.n_cam, 2))
    P = PnpProblem(prob, intr)
    opt = mcba_pnp_options(thr, 1000, 0.99, 1, 20, SEED)
    return P, opt, gt.pose


def rot(rv):
    th = np.linalg.norm(rv)
    if th < 1e-12:
        return np.eye(3)
    k = rv / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K


def pose_diff(a, b):
    """max |R_a - R_b| and relative translation difference."""
    return max(np.abs(rot(a[:3]) - rot(b[:3])).max(), np.abs(a[3:] - b[3:]).max() / max(1.0, np.abs(b[3:]).max()))


def compare(got, ref, P, tol_refined=1e-5, tol_unrefined=2e-4):
    """got/ref: dicts with best_h, n_hyp, n_inliers, inlier_mask, pose, pose0. The integer outputs (selected draw,
    draws consumed, inlier set) must be identical; poses agree within the tolerances (OpenCV works from float32 points
    and stops its LM at 20 iterations / FLT_EPSILON)."""
    assert np.array_equal(got["best_h"], ref["best_h"])
    assert np.array_equal(got["n_hyp"], ref["n_hyp"])
    assert np.array_equal(got["inlier_mask"], ref["inlier_mask"])
    assert np.array_equal(got["n_inliers"], ref["n_inliers"])
    w0 = w1 = 0.0
    for b in range(P.n_bobs):
        if ref["best_h"][b] < 0:
            continue
        w0 = max(w0, pose_diff(got["pose0"][b], ref["pose0"][b]))
        w1 = max(w1, pose_diff(got["pose"][b], ref["pose"][b]))
    assert w0 < tol_unrefined and w1 < tol_refined, (w0, w1)
    return w0, w1


def host_lib():
    d = os.path.join(HERE, "host_math")
    subprocess.check_call(["make", "-s", "-C", d])
    L = C.CDLL(os.path.join(d, "libmathtest.so"))
    dp = mcba.c_f64p
    L.pnptest_ransac.argtypes = [C.c_int, mcba.c_f32p, mcba.c_f32p, mcba.c_f32p, C.c_int, dp, C.c_double, C.c_int,
                                 C.c_double, C.c_int, C.c_int, C.c_uint64, C.c_uint64, dp, dp, mcba.c_u8p,
                                 C.POINTER(C.c_int), C.POINTER(C.c_int)]
    return L


def host_ransac_all(L, P, opt):
    """The sequential host restatement of the kernel (tests/host_math) over every observation run."""
    nb = P.n_bobs
    res = dict(pose=np.zeros((nb, 6)), pose0=np.zeros((nb, 6)), n_inliers=np.zeros(nb, np.int32),
               inlier_mask=np.zeros(P.n_obs, np.uint8), best_h=np.zeros(nb, np.int32), n_hyp=np.zeros(nb, np.int32))
    dp = mcba.c_f64p
    for b in range(nb):
        o0, o1 = int(P.bobs_ptr[b]), int(P.bobs_ptr[b + 1])
        n, cam = o1 - o0, int(P.bobs_cam[b])
        xyz = np.ascontiguousarray(P.pts_xyz[P.pt_idx[o0:o1]])
        u, v = np.ascontiguousarray(P.u[o0:o1]), np.ascontiguousarray(P.v[o0:o1])
        mask = np.zeros(max(n, 1), np.uint8)
        bh, nh = C.c_int(), C.c_int()
        cnt = L.pnptest_ransac(n, u.ctypes.data_as(mcba.c_f32p), v.ctypes.data_as(mcba.c_f32p),
                               xyz.ctypes.data_as(mcba.c_f32p), int(P.cam_model[cam]),
                               P.intrinsics[cam].ctypes.data_as(dp), opt.threshold_px, opt.max_iterations,
                               opt.confidence, opt.refine, opt.refine_max_iterations, opt.seed, b,
                               res["pose"][b].ctypes.data_as(dp), res["pose0"][b].ctypes.data_as(dp),
                               mask.ctypes.data_as(mcba.c_u8p), C.byref(bh), C.byref(nh))
        res["n_inliers"][b], res["best_h"][b], res["n_hyp"][b] = cnt, bh.value, nh.value
        res["inlier_mask"][o0:o1] = mask[:n]
    return res


def golden_case(name):
    g = np.load(GOLDEN)
    return {k: g[f"{name}/{k}"] for k in ("pose", "pose0", "n_inliers", "inlier_mask", "best_h", "n_hyp")}
