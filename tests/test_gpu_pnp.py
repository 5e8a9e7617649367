"""Pose initialisation on the GPU (mcba_pnp_ransac through the C ABI) against the committed OpenCV golden outputs and,
when cv2 is importable, against the OpenCV-backed checker run on the same inputs."""
import numpy as np
import pytest

import pnp_util
from mcba import pnp

pytestmark = pytest.mark.gpu


def as_dict(r):
    return dict(pose=r.pose, pose0=r.unrefined_pose, n_inliers=r.n_inliers, inlier_mask=r.inlier_mask,
                best_h=r.best_hypothesis, n_hyp=r.n_hypotheses)


@pytest.mark.parametrize("name", list(pnp_util.CASES))
def test_gpu_matches_opencv_golden(name):
    P, opt, gt_pose = pnp_util.make_case(name)
    r = pnp.ransac(P, opt)
    assert r.kernel_launches == 4 and r.kernel_ms > 0
    pnp_util.compare(as_dict(r), pnp_util.golden_case(name), P)


def test_gpu_matches_checker_live():
    pytest.importorskip("cv2")
    import oracle_pnp
    P, opt, _ = pnp_util.make_case("c2_t1")
    opt.seed = 777                                  # a seed the golden file does not hold
    r = pnp.ransac(P, opt)
    pnp_util.compare(as_dict(r), oracle_pnp.ransac_all(P, opt), P)


def test_gpu_reproducible_and_seed_dependent():
    P, opt, _ = pnp_util.make_case("c2_t1")
    a, b = pnp.ransac(P, opt), pnp.ransac(P, opt)
    assert np.array_equal(a.pose, b.pose) and np.array_equal(a.inlier_mask, b.inlier_mask)
    opt.seed = 43
    c = pnp.ransac(P, opt)
    assert not np.array_equal(a.best_hypothesis, c.best_hypothesis)


def test_gpu_matches_host_restatement_bitwise_on_integers():
    L = pnp_util.host_lib()
    P, opt, _ = pnp_util.make_case("c3_t3")
    r = pnp.ransac(P, opt)
    h = pnp_util.host_ransac_all(L, P, opt)
    pnp_util.compare(as_dict(r), h, P, tol_refined=1e-8, tol_unrefined=1e-10)


def test_gpu_edge_cases():
    P, opt, _ = pnp_util.make_case("c1_t2")
    empty = pnp.PnpProblem.__new__(pnp.PnpProblem)
    empty.__dict__.update(P.__dict__)
    empty.bobs_ptr, empty.bobs_cam = np.zeros(1, np.int64), np.zeros(0, np.int32)
    empty.u, empty.v, empty.pt_idx = np.zeros(0, np.float32), np.zeros(0, np.float32), np.zeros(0, np.int32)
    r = pnp.ransac(empty, opt)
    assert r.pose.shape == (0, 6)
    short = pnp.PnpProblem.__new__(pnp.PnpProblem)
    short.__dict__.update(P.__dict__)
    short.bobs_ptr, short.bobs_cam = np.array([0, 3, 19], np.int64), P.bobs_cam[:2].copy()
    short.u, short.v, short.pt_idx = P.u[:19].copy(), P.v[:19].copy(), P.pt_idx[:19].copy()
    r = pnp.ransac(short, opt)
    assert r.best_hypothesis[0] == -1 and r.n_inliers[0] == 0 and not r.pose[0].any()
    bad = pnp.PnpProblem.__new__(pnp.PnpProblem)
    bad.__dict__.update(P.__dict__)
    bad.pt_idx = P.pt_idx.copy()
    bad.pt_idx[5] = 10 ** 6
    with pytest.raises(Exception):
        pnp.ransac(bad, opt)


def test_gpu_degenerate_and_hostile_inputs():
    """Collinear corners, non-finite pixels, pure noise: same verdicts as the host restatement, never a non-finite pose."""
    from mcba import synth
    L = pnp_util.host_lib()
    intr = synth.BROWN_GT
    rng = np.random.default_rng(0)
    rv, t = np.array([0.2, -0.1, 0.05]), np.array([0.1, -0.05, 2.0])
    R = pnp_util.rot(rv)
    grid = np.array([[x * 0.1, y * 0.1, 0.0] for y in range(4) for x in range(4)])
    line = np.array([[x * 0.05, 0.0, 0.0] for x in range(10)])
    uv = synth.project(intr, np.int32(0), grid @ R.T + t)
    bad = uv.copy()
    bad[1], bad[7], bad[9] = [np.nan, 5.0], [np.inf, -np.inf], [1e30, 1e30]
    runs = [(grid, uv), (grid[[0, 3, 12, 15]], uv[[0, 3, 12, 15]]), (line, synth.project(intr, np.int32(0), line @ R.T + t)),
            (grid, bad), (grid, np.full_like(uv, np.nan))]
    runs += [(rng.normal(size=(12, 3)), rng.uniform(0, 1900, size=(12, 2))) for _ in range(20)]
    P = pnp.PnpProblem.__new__(pnp.PnpProblem)
    P.pts_xyz = np.ascontiguousarray(np.concatenate([r[0] for r in runs]), np.float32)
    P.u = np.ascontiguousarray(np.concatenate([r[1][:, 0] for r in runs]), np.float32)
    P.v = np.ascontiguousarray(np.concatenate([r[1][:, 1] for r in runs]), np.float32)
    P.pt_idx = np.arange(len(P.u), dtype=np.int32)
    P.bobs_ptr = np.concatenate([[0], np.cumsum([len(r[0]) for r in runs])]).astype(np.int64)
    P.bobs_cam = np.zeros(len(runs), np.int32)
    P.cam_model = np.zeros(1, np.int32)
    P.intrinsics = np.ascontiguousarray(intr.reshape(1, 9))
    opt = pnp.default_options(threshold_px=2.0, max_iterations=50, seed=3)
    r = pnp.ransac(P, opt)
    assert np.isfinite(r.pose).all()
    h = pnp_util.host_ransac_all(L, P, opt)
    assert np.array_equal(r.best_hypothesis, h["best_h"]) and np.array_equal(r.n_hypotheses, h["n_hyp"])
    assert np.array_equal(r.inlier_mask, h["inlier_mask"])
    assert r.n_inliers[0] == 16 and r.n_inliers[1] == 4 and r.best_hypothesis[2] == -1 and r.n_inliers[3] == 13
    assert r.best_hypothesis[4] == -1 and not r.pose[4].any() and not r.pose[2].any()
    for b in (0, 1, 3):
        assert pnp_util.pose_diff(r.pose[b], np.concatenate([rv, t])) < 1e-5
