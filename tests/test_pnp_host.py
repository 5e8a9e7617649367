"""Pose initialisation, CPU side: (1) the OpenCV-backed checker (oracle/oracle_pnp.py) reproduces the committed golden
outputs; (2) the sequential host restatement of the CUDA kernel's algorithm (tests/host_math, same header arithmetic
as csrc/mcba_pnp.cu) agrees with it: identical selected draw, draws consumed and inlier sets, poses within tolerance;
(3) edge cases; (4) the C-ABI refuses to run without a GPU instead of falling back."""
import numpy as np
import pytest

import mcba
import pnp_util
from mcba import pnp


@pytest.fixture(scope="module")
def hostlib():
    return pnp_util.host_lib()


@pytest.mark.parametrize("name", ["c2_t1", "c1_t2"])
def test_checker_reproduces_golden(name):
    pytest.importorskip("cv2")
    import oracle_pnp
    P, opt, _ = pnp_util.make_case(name)
    ref = oracle_pnp.ransac_all(P, opt)
    pnp_util.compare(ref, pnp_util.golden_case(name), P, tol_refined=1e-9, tol_unrefined=1e-9)


@pytest.mark.parametrize("name", list(pnp_util.CASES))
def test_host_restatement_matches_golden(hostlib, name):
    P, opt, gt_pose = pnp_util.make_case(name)
    got = pnp_util.host_ransac_all(hostlib, P, opt)
    pnp_util.compare(got, pnp_util.golden_case(name), P)
    if "pert" not in name:      # exact intrinsics: the refined poses are the ground truth up to the pixel noise
        err = max(pnp_util.pose_diff(got["pose"][b], gt_pose[b]) for b in range(P.n_bobs))
        assert err < 0.08


def test_edge_cases(hostlib):
    P, opt, _ = pnp_util.make_case("c1_t2")
    # a run with fewer than four corners has no pose; an impossible threshold leaves every run without inliers
    sub = pnp.PnpProblem.__new__(pnp.PnpProblem)
    sub.__dict__.update(P.__dict__)
    sub.bobs_ptr = np.array([0, 3, 3 + 16], np.int64)
    sub.bobs_cam = P.bobs_cam[:2].copy()
    sub.u, sub.v, sub.pt_idx = P.u[:19].copy(), P.v[:19].copy(), np.concatenate([P.pt_idx[:3], P.pt_idx[16:32]]).astype(np.int32)
    sub.u[3:], sub.v[3:] = P.u[16:32], P.v[16:32]
    got = pnp_util.host_ransac_all(hostlib, sub, opt)
    assert got["best_h"][0] == -1 and got["n_inliers"][0] == 0 and not got["pose"][0].any()
    assert got["best_h"][1] >= 0 and got["n_inliers"][1] >= 4
    tiny = pnp.mcba_pnp_options(1e-9, 50, 0.99, 1, 20, 1)
    got = pnp_util.host_ransac_all(hostlib, P, tiny)
    assert (got["n_hyp"] == 50).all()            # never satisfied: all draws consumed
    assert (got["n_inliers"] <= 4).all()


def test_abi_has_no_cpu_fallback():
    import os
    import torch
    if torch.cuda.is_available() or not os.path.exists(mcba.LIB_PATH):
        pytest.skip("needs a CPU-only box with the library built")
    P, opt, _ = pnp_util.make_case("c1_t2")
    with pytest.raises(mcba.McbaError):
        pnp.ransac(P, opt)


def _run_one(hostlib, xyz, uv, intr, model=0, thr=2.0, it=50, seed=3):
    import ctypes as C
    n = len(xyz)
    xyz = np.ascontiguousarray(xyz, np.float32)
    u, v = np.ascontiguousarray(uv[:, 0], np.float32), np.ascontiguousarray(uv[:, 1], np.float32)
    pose, pose0, mask = np.zeros(6), np.zeros(6), np.zeros(max(n, 1), np.uint8)
    bh, nh = C.c_int(), C.c_int()
    dp = mcba.c_f64p
    cnt = hostlib.pnptest_ransac(n, u.ctypes.data_as(mcba.c_f32p), v.ctypes.data_as(mcba.c_f32p), xyz.ctypes.data_as(mcba.c_f32p),
                                 model, np.ascontiguousarray(intr, np.float64).ctypes.data_as(dp), thr, it, 0.99, 1, 20, seed, 0,
                                 pose.ctypes.data_as(dp), pose0.ctypes.data_as(dp), mask.ctypes.data_as(mcba.c_u8p),
                                 C.byref(bh), C.byref(nh))
    return cnt, bh.value, nh.value, pose, mask[:n]


def test_degenerate_and_hostile_inputs(hostlib):
    """Collinear corners, non-finite pixels, points behind the camera, exactly four corners: never a crash, never a
    non-finite pose, and 'no pose' (best = -1, zero inliers) where no hypothesis can exist."""
    from mcba import synth
    intr = synth.BROWN_GT
    rng = np.random.default_rng(0)
    rv, t = np.array([0.2, -0.1, 0.05]), np.array([0.1, -0.05, 2.0])
    R = pnp_util.rot(rv)
    grid = np.array([[x * 0.1, y * 0.1, 0.0] for y in range(4) for x in range(4)])
    uv = synth.project(intr, np.int32(0), grid @ R.T + t)
    # exact data: all inliers, pose recovered
    cnt, bh, nh, pose, mask = _run_one(hostlib, grid, uv, intr)
    assert cnt == 16 and bh >= 0 and pnp_util.pose_diff(pose, np.concatenate([rv, t])) < 1e-5
    # four corners only
    cnt, bh, nh, pose, mask = _run_one(hostlib, grid[[0, 3, 12, 15]], uv[[0, 3, 12, 15]], intr)
    assert cnt == 4 and np.isfinite(pose).all() and pnp_util.pose_diff(pose, np.concatenate([rv, t])) < 1e-5
    # every corner on one line: P3P is degenerate for every draw
    line = np.array([[x * 0.05, 0.0, 0.0] for x in range(10)])
    uvl = synth.project(intr, np.int32(0), line @ R.T + t)
    cnt, bh, nh, pose, mask = _run_one(hostlib, line, uvl, intr)
    assert (cnt, bh) == (0, -1) and nh == 50 and not pose.any() and not mask.any()
    # non-finite and absurd pixels among good ones: they are never inliers, the pose is still found
    bad = uv.copy()
    bad[1] = [np.nan, 5.0]
    bad[7] = [np.inf, -np.inf]
    bad[9] = [1e30, 1e30]
    cnt, bh, nh, pose, mask = _run_one(hostlib, grid, bad, intr)
    assert np.isfinite(pose).all() and cnt == 13 and not mask[[1, 7, 9]].any()
    assert pnp_util.pose_diff(pose, np.concatenate([rv, t])) < 1e-5
    # everything non-finite: no pose
    cnt, bh, nh, pose, mask = _run_one(hostlib, grid, np.full_like(uv, np.nan), intr)
    assert (cnt, bh) == (0, -1) and not pose.any()
    # pure noise: whatever is selected is finite
    for k in range(20):
        cnt, bh, nh, pose, mask = _run_one(hostlib, rng.normal(size=(12, 3)), rng.uniform(0, 1900, size=(12, 2)), intr, seed=k)
        assert np.isfinite(pose).all() and cnt == int(mask.sum())
