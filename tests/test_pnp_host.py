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
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    P, opt, _ = pnp_util.make_case("c1_t2")
    with pytest.raises(mcba.McbaError):
        pnp.ransac(P, opt)
