"""GPU parity tests: libmcba (CUDA, through the C ABI) against the CPU oracle on identical inputs.

Tolerances: residuals/Jacobians/costs 1e-11 relative (fp64, different but equivalent operation order);
reduced system 1e-9 relative; LM traces: identical accept/reject sequence, iteration counts and
termination; final RMS within 1e-6 px and parameters within 1e-6 relative (BASELINE.json north_star).
"""
import numpy as np
import pytest

import mcba
from mcba import synth
from util import dense_jacobian, layout, rel, trace_tuple

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c1():
    return synth.config_problem("c1", 5)


@pytest.fixture(scope="module")
def c2():
    return synth.config_problem("c2", 5)


def stage_problem(cfg, stage):
    s, o = cfg[3], cfg[4]
    return synth.make_problem(s, o, stage)


@pytest.mark.parametrize("cfgname", ["c1", "c2"])
@pytest.mark.parametrize("stage", [1, 2, 3, 4, 5])
def test_eval_matches_oracle(cfgname, stage, oracle, c1, c2):
    cfg = c1 if cfgname == "c1" else c2
    prob, init, gt = stage_problem(cfg, stage)
    h = mcba.Handle(prob)
    for p in (init, gt):
        cost, res, jac = h.eval(p)
        rc, ocost, ores, ojac, _ = oracle.eval(prob, p)
        assert rc == 0
        assert abs(cost - ocost) <= 1e-12 * abs(ocost)
        assert np.abs(res - ores).max() <= 1e-9 * max(1.0, np.abs(ores).max())
        assert rel(jac, ojac) <= 1e-11
    err = h.reprojection_errors(gt)
    _, _, ores, _, _ = oracle.eval(prob, gt, mcba.default_options(huber_a=0.0))
    assert np.abs(err - np.sqrt((ores ** 2).sum(1))).max() <= 1e-9
    h.close()


@pytest.mark.parametrize("stage", [1, 2, 3, 4, 5])
def test_normal_equations_and_schur(stage, oracle, c1):
    """F^T F, F^T r, reduced system and its solution against a dense numpy computation from the oracle's J."""
    prob, init, gt = stage_problem(c1, stage)
    h = mcba.Handle(prob)
    opt = mcba.default_options(max_num_iterations=1)
    h.solve_begin(init.copy(), opt)
    L = layout(prob)
    n_e, n_f = L["n_e"], L["n_f"]
    FF = h.debug_fetch("FF").reshape(n_f, n_f)          # blocks at the initial point (iteration 0)
    gf = h.debug_fetch("gf")
    Hoo = h.debug_fetch("Hoo").reshape(-1, 6, 6)
    scale = np.concatenate([h.debug_fetch("scale_e"), h.debug_fetch("scale_f")])
    h.solve_step()      # records iteration 0, computes the first step (then re-evaluates at the accepted point)
    _, _, ores, ojac, _ = oracle.eval(prob, init)
    J = dense_jacobian(prob, ojac)
    r = ores.reshape(-1)
    Hd, g = J.T @ J, J.T @ r
    assert rel(FF, Hd[n_e:, n_e:]) <= 1e-11
    assert rel(gf, g[n_e:]) <= 1e-11
    for o in range(0, prob.n_pose, 17):
        assert rel(Hoo[o], Hd[o * 6:o * 6 + 6, o * 6:o * 6 + 6]) <= 1e-11
    s = 1.0 / (1.0 + np.sqrt(np.diag(Hd)))
    assert rel(scale, s) <= 1e-12
    Hs = Hd * s[:, None] * s[None, :]
    diag = np.clip(np.diag(Hs), 1e-6, 1e32)
    A = Hs + np.diag(diag / 1e4)
    gs = g * s
    z = np.linalg.solve(A, gs)
    Aee, Aef, Aff = A[:n_e, :n_e], A[:n_e, n_e:], A[n_e:, n_e:]
    S = Aff - Aef.T @ np.linalg.solve(Aee, Aef)
    rhs = gs[n_e:] - Aef.T @ np.linalg.solve(Aee, gs[:n_e])
    # the GPU factorises S in place (lower triangle); compare through the solution and the rhs
    assert rel(h.debug_fetch("rhs"), rhs) <= 1e-9
    assert rel(h.debug_fetch("zf"), z[n_e:]) <= 1e-8
    Lg = np.tril(h.debug_fetch("S").reshape(n_f, n_f))
    assert rel(Lg @ Lg.T, S) <= 1e-9
    h.close()


def _solve_both(prob, init, oracle, **kw):
    opt = mcba.default_options(**kw)
    pg, po = init.copy(), init.copy()
    h = mcba.Handle(prob)
    sg, tg = h.solve(pg, opt)
    h.close()
    rc, so, to = oracle.solve(prob, po, opt)
    assert rc == 0
    return (pg, sg, tg), (po, so, to)


def _assert_same_solution(prob, g, o):
    (pg, sg, tg), (po, so, to) = g, o
    assert sg.termination == so.termination
    assert (sg.num_successful_steps, sg.num_unsuccessful_steps, sg.num_iterations) == \
           (so.num_successful_steps, so.num_unsuccessful_steps, so.num_iterations)
    assert trace_tuple(tg) == trace_tuple(to)
    for a, b in zip(tg, to):
        assert abs(a.cost - b.cost) <= 1e-9 * abs(b.cost)
        assert abs(a.trust_region_radius - b.trust_region_radius) <= 1e-6 * b.trust_region_radius
    assert abs(sg.rms_px - so.rms_px) <= 1e-6
    for name in ("cam_pose", "pose", "board_pose", "intrinsics"):
        a, b = getattr(pg, name), getattr(po, name)
        if a.size:
            assert np.abs(a - b).max() <= 1e-6 * max(1.0, np.abs(b).max()), name


@pytest.mark.parametrize("stage", [1, 2, 3, 4, 5])
def test_solve_c1_matches_oracle(stage, oracle, c1):
    prob, init, gt = stage_problem(c1, stage)
    g, o = _solve_both(prob, init, oracle)
    _assert_same_solution(prob, g, o)


@pytest.mark.parametrize("stage", [4, 5])
def test_solve_c2_hybrid_matches_oracle(stage, oracle, c2):
    prob, init, gt = stage_problem(c2, stage)
    g, o = _solve_both(prob, init, oracle)
    _assert_same_solution(prob, g, o)


def test_materialized_path_matches_fused(c1):
    prob, init, gt = stage_problem(c1, 5)
    out = []
    for mat in (0, 1):
        p = init.copy()
        h = mcba.Handle(prob)
        s, t = h.solve(p, mcba.default_options(materialize_jacobian=mat))
        h.close()
        out.append((p, s, t))
    (p0, s0, t0), (p1, s1, t1) = out
    assert trace_tuple(t0) == trace_tuple(t1)
    assert abs(s0.final_cost - s1.final_cost) <= 1e-10 * s0.final_cost
    assert np.abs(p0.intrinsics - p1.intrinsics).max() <= 1e-8 * np.abs(p0.intrinsics).max()


def test_stage4_then_stage5_on_one_handle(oracle, c1):
    """calibrate.cpp:60-64: refineAllCameraGroupAndObjects then ...AndIntrinsic on the same observations."""
    s, o = c1[3], c1[4]
    prob4, init, gt = synth.make_problem(s, o, 4)
    pg, po = init.copy(), init.copy()
    pg.intrinsics[:] = s.intr_init
    po.intrinsics[:] = s.intr_init
    h = mcba.Handle(prob4)
    s4, _ = h.solve(pg)
    h.set_stage(5)
    s5, t5 = h.solve(pg)
    h.close()
    rc, o4, _ = oracle.solve(prob4, po)
    rc, o5, u5 = oracle.solve(prob4.with_stage(5), po)
    assert (s5.num_successful_steps, s5.num_unsuccessful_steps) == (o5.num_successful_steps, o5.num_unsuccessful_steps)
    assert abs(s5.rms_px - o5.rms_px) <= 1e-6
    assert np.abs(pg.intrinsics - po.intrinsics).max() <= 1e-6 * np.abs(po.intrinsics).max()


def _camera_subproblem(prob, params, c):
    """Stage-1 problem of one camera cut out of a batched problem (what the reference solves per camera)."""
    sel = np.nonzero(prob.bobs_cam == c)[0]
    b0, b1 = sel[0], sel[-1] + 1
    o0, o1 = prob.bobs_ptr[b0], prob.bobs_ptr[b1]
    sub = mcba.Problem(1, prob.u[o0:o1], prob.v[o0:o1], prob.pt_idx[o0:o1], prob.bobs_ptr[b0:b1 + 1] - o0,
                       np.zeros(b1 - b0), np.arange(b1 - b0), np.zeros(b1 - b0), prob.pts_xyz, prob.cam_model[c:c + 1],
                       b1 - b0, 0)
    p = mcba.Params(np.zeros((1, 6)), params.pose[b0:b1], np.zeros((0, 6)), params.intrinsics[c:c + 1])
    return sub, p, (b0, b1)


def test_batched_intrinsics_matches_per_camera_oracle(oracle):
    """Config 5 shape: independent per-camera intrinsic refinements solved in one batch; every camera must follow the
    same trajectory as its own single-problem oracle solve (Camera::refineIntrinsicCalibration per camera)."""
    prob, init, gt = synth.batched_intrinsics_problem(n_cameras=7, n_frames=40, kannala_every=3)
    pg = init.copy()
    h = mcba.Handle(prob)
    summ = h.solve_batched(pg)
    h.close()
    assert len(summ) == 7
    for c in range(7):
        sub, p, (b0, b1) = _camera_subproblem(prob, init, c)
        rc, so, to = oracle.solve(sub, p)
        assert rc == 0
        s = summ[c]
        assert (s.termination, s.num_successful_steps, s.num_unsuccessful_steps, s.num_iterations) == \
               (so.termination, so.num_successful_steps, so.num_unsuccessful_steps, so.num_iterations), c
        assert abs(s.final_cost - so.final_cost) <= 1e-9 * so.final_cost
        assert abs(s.rms_px - so.rms_px) <= 1e-6
        assert np.abs(pg.intrinsics[c] - p.intrinsics[0]).max() <= 1e-6 * np.abs(p.intrinsics[0]).max()
        assert np.abs(pg.pose[b0:b1] - p.pose).max() <= 1e-6 * max(1.0, np.abs(p.pose).max())


def test_batched_problems_terminate_independently(oracle):
    """One camera starts at its optimum-ish (ground truth), the others far away: iteration counts differ per problem."""
    prob, init, gt = synth.batched_intrinsics_problem(n_cameras=4, n_frames=30)
    p0 = init.copy()
    p0.intrinsics[0] = gt.intrinsics[0]
    sel = prob.bobs_cam == 0
    p0.pose[sel] = gt.pose[sel]
    h = mcba.Handle(prob)
    summ = h.solve_batched(p0)
    h.close()
    its = [s.num_iterations for s in summ]
    assert its[0] < max(its[1:])
    assert all(s.termination == 0 for s in summ)


@pytest.mark.parametrize("cfgname,frames,stage", [("c3", 40, 5), ("c3", 40, 4), ("c4", 30, 5)])
def test_solve_large_reduced_systems(cfgname, frames, stage, oracle):
    """16-camera ring (63-corner boards: three warp passes per board observation, n_f = 270 = 4 panels + 14) and the
    64-camera dome (n_f = 1020 = 15 panels + 60): multi-panel cooperative Cholesky and the split-K Schur SYRK."""
    s = synth.make_scene(cfgname, n_frames=frames)
    o = synth.generate_observations(s)
    prob, init, gt = synth.make_problem(s, o, stage)
    g, oo = _solve_both(prob, init, oracle)
    _assert_same_solution(prob, g, oo)


def test_ragged_and_empty_inputs(oracle):
    """Poses without any observation, single-corner board observations and a camera that is never observed."""
    s = synth.make_scene("c2", n_frames=30)
    o = synth.generate_observations(s)
    prob, init, gt = synth.make_problem(s, o, 5)
    counts = np.diff(prob.bobs_ptr)
    keep_bobs = np.ones(prob.n_bobs, bool)
    keep_bobs[prob.bobs_pose == 7] = False          # frame 7 loses all its observations -> empty e-block
    keep_bobs[prob.bobs_cam == 3] &= (np.arange(prob.n_bobs) % 2 == 0)[prob.bobs_cam == 3]
    new_counts = counts.copy()
    new_counts[prob.bobs_pose == 11] = 1            # frame 11: one corner per board observation
    obs_keep = np.zeros(prob.n_obs, bool)
    for b in range(prob.n_bobs):
        if keep_bobs[b]:
            obs_keep[prob.bobs_ptr[b]:prob.bobs_ptr[b] + new_counts[b]] = True
    p2 = mcba.Problem(5, prob.u[obs_keep], prob.v[obs_keep], prob.pt_idx[obs_keep],
                      np.concatenate([[0], np.cumsum(new_counts[keep_bobs])]), prob.bobs_cam[keep_bobs],
                      prob.bobs_pose[keep_bobs], prob.bobs_board[keep_bobs], prob.pts_xyz, prob.cam_model, prob.n_pose,
                      prob.n_board, prob.cam_is_ref, prob.board_is_ref)
    g, oo = _solve_both(p2, init, oracle)
    _assert_same_solution(p2, g, oo)
    assert np.array_equal(g[0].pose[7], init.pose[7])      # untouched: zero gradient, zero step


def test_invalid_inputs_are_rejected():
    s = synth.make_scene("c1", n_frames=5)
    o = synth.generate_observations(s)
    prob, init, gt = synth.make_problem(s, o, 4)
    bad = prob.with_stage(4)
    bad.pt_idx = prob.pt_idx.copy(); bad.pt_idx[3] = 10 ** 6
    with pytest.raises(mcba.McbaError):
        mcba.Handle(bad)
    bad = prob.with_stage(4)
    bad.bobs_pose = prob.bobs_pose[::-1].copy()
    with pytest.raises(mcba.McbaError):
        mcba.Handle(bad)
    with pytest.raises(mcba.McbaError):
        mcba.Handle(prob.with_stage(9))
    bad = prob.with_stage(4)
    bad.cam_model = prob.cam_model.copy(); bad.cam_model[1] = 7          # neither Brown nor Kannala
    with pytest.raises(mcba.McbaError):
        mcba.Handle(bad)
    h = mcba.Handle(prob)
    with pytest.raises(mcba.McbaError):
        h.solve_batched(init.copy())
    h.close()


def test_non_finite_start_fails_like_ceres():
    """NaN in the initial point: Ceres reports 'Initial residual and Jacobian evaluation failed' (FAILURE)."""
    s = synth.make_scene("c1", n_frames=5)
    o = synth.generate_observations(s)
    prob, init, gt = synth.make_problem(s, o, 5)
    p = init.copy(); p.intrinsics[0, 0] = np.nan
    h = mcba.Handle(prob)
    with pytest.raises(mcba.McbaError) as e:
        h.solve(p)
    assert "rc=-4" in str(e.value)
    h.close()


def test_solve_is_bit_reproducible(c2):
    """No floating-point atomics anywhere: two solves from the same start give bit-identical parameters and traces."""
    prob, init, gt = stage_problem(c2, 5)
    outs = []
    for _ in range(2):
        p = init.copy()
        h = mcba.Handle(prob)
        s, t = h.solve(p)
        h.close()
        outs.append((p, [(r.cost, r.trust_region_radius, r.gradient_max_norm) for r in t]))
    assert outs[0][1] == outs[1][1]
    for name in ("cam_pose", "pose", "board_pose", "intrinsics"):
        assert np.array_equal(getattr(outs[0][0], name), getattr(outs[1][0], name))


def test_empty_problem(oracle):
    """A problem whose parameter blocks have no residuals: cost 0, zero gradient -> CONVERGENCE at iteration 0."""
    s = synth.make_scene("c1", n_frames=4)
    o = synth.generate_observations(s)
    prob, init, gt = synth.make_problem(s, o, 4)
    e = mcba.Problem(4, [], [], [], [0], [], [], [], prob.pts_xyz, prob.cam_model, prob.n_pose, prob.n_board,
                     prob.cam_is_ref, prob.board_is_ref)
    p = init.copy()
    h = mcba.Handle(e)
    summ, tr = h.solve(p)
    h.close()
    po = init.copy()
    rc, so, to = oracle.solve(e, po)
    assert (summ.termination, summ.num_iterations) == (so.termination, so.num_iterations) == (0, 1)
    assert summ.final_cost == 0.0 and np.array_equal(p.pose, init.pose)


def test_active_blocks_with_exactly_zero_rotation(oracle, c2):
    """The reference value-initialises a pose that was never set (std::map::operator[], CameraGroup.cpp:445-451): an
    ACTIVE camera / board / object block with omega == 0 exactly takes AngleAxisRotatePoint's small-angle branch
    (R = I + [w]x, dR/dw = -[p]x). Device prep_block (mcba_math.cuh) against the oracle's dual numbers, then a solve."""
    prob, init, gt = stage_problem(c2, 5)
    p = init.copy()
    p.cam_pose[1, :3] = 0.0          # active (non-reference) camera, zero rotation, non-zero translation
    p.board_pose[2, :3] = 0.0        # active board
    p.pose[5, :3] = 0.0              # an e-block
    p.cam_pose[2, :] = 0.0           # a fully zero active pose (what operator[] creates)
    h = mcba.Handle(prob)
    cost, res, jac = h.eval(p)
    rc, ocost, ores, ojac, _ = oracle.eval(prob, p)
    assert rc == 0 and np.isfinite(jac).all()
    assert abs(cost - ocost) <= 1e-12 * abs(ocost)
    assert np.abs(res - ores).max() <= 1e-9 * max(1.0, np.abs(ores).max())
    assert rel(jac, ojac) <= 1e-11
    h.close()
    g, o = _solve_both(prob, p, oracle)
    _assert_same_solution(prob, g, o)


def test_kannala_point_exactly_on_the_optical_axis(oracle):
    """r = 0: the functor's `r > 1e-8 ? theta_d / r : 1` guard (OptimizationCeres.h:472-477) selects the constant, so
    the NaN partials of sqrt(0) in Ceres' Jets never reach the residual: the Jacobian is finite (d/dk = 0, the pose and
    focal columns those of xd = a). Same on the device and in the oracle."""
    pts = np.array([[0, 0, 0], [0.1, 0, 0], [0, 0.1, 0], [0.1, 0.1, 0], [0.05, 0.02, 0]], np.float32)
    intr = np.array([[600.0, 600, 960, 540, -0.02, 3e-3, -1e-3, 2e-4, 0]])
    pose = np.array([[0.0, 0, 0, 0, 0, 2.0], [0.1, -0.2, 0.05, 0, 0, 1.5]])      # frame 0: corner 0 projects to r = 0 exactly
    uv = np.array([[961.0, 541.0]] * 10, np.float32)
    prob = mcba.Problem(1, uv[:, 0], uv[:, 1], np.tile(np.arange(5), 2), [0, 5, 10], [0, 0], [0, 1], [0, 0], pts, [1], 2, 0)
    p = mcba.Params(np.zeros((1, 6)), pose, np.zeros((0, 6)), intr)
    h = mcba.Handle(prob)
    cost, res, jac = h.eval(p)
    h.close()
    rc, ocost, ores, ojac, _ = oracle.eval(prob, p)
    assert rc == 0 and np.isfinite(jac).all() and np.isfinite(ojac).all()
    assert np.array_equal(jac[0, :, 10:14], np.zeros((2, 4)))        # d/d(k1..k4) of the on-axis corner
    assert np.abs(res - ores).max() <= 1e-10 and rel(jac, ojac) <= 1e-11


def test_failure_returns_the_start_point(oracle):
    """ceres::Solve leaves the ORIGINAL parameter values in place when the summary is not usable (FAILURE)."""
    import lm_cases
    prob, p0, opt, _ = lm_cases.make_case("invalid_failure_c2_s5")
    pg = p0.copy()
    h = mcba.Handle(prob)
    sg, _ = h.solve(pg, opt)
    h.close()
    assert sg.termination == 2
    for name in ("cam_pose", "pose", "board_pose", "intrinsics"):
        assert np.array_equal(getattr(pg, name), getattr(p0, name)), name


def test_full_config3_against_the_committed_oracle_fixture():
    """BASELINE configs[2] in full (16-camera ring, 5 boards, 5 000 frames, 10 188 613 observations) on ONE B200: stage 4
    then stage 5 (calibrate.cpp:60-64) against tests/golden/full_c3.json (the CPU oracle, 160 s on 8 threads):
    identical accept / reject sequence and termination, costs within 1e-9, RMS within 1e-6 px, f-blocks within 1e-6."""
    import json, os
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "full_c3.json")))
    prob, init, gt, s, o = synth.config_problem("c3", 4)
    assert prob.n_obs == gold["observations"]
    p = init.copy()
    p.intrinsics[:] = s.intr_init
    h = mcba.Handle(prob)
    for st in gold["stages"]:
        if st["stage"] == 5:
            h.set_stage(5)
        summ, tr = h.solve(p)
        assert [(r.iteration, r.step_is_valid, r.step_is_successful) for r in tr] == \
               [(r["iteration"], r["valid"], r["successful"]) for r in st["trace"]]
        assert summ.termination == st["termination"]
        for a, b in zip(tr, st["trace"]):
            assert abs(a.cost - b["cost"]) <= 1e-9 * b["cost"]
            assert abs(a.trust_region_radius - b["radius"]) <= 1e-6 * b["radius"]
        assert abs(summ.rms_px - st["rms_px"]) <= 1e-6
        for name in ("cam_pose", "board_pose", "intrinsics"):
            a, b = getattr(p, name), np.array(st[name])
            assert np.abs(a - b).max() <= 1e-6 * np.abs(b).max(), (st["stage"], name)
    h.close()


def test_registered_observations_are_shared_between_pose_init_and_bundle_adjustment():
    """mcba_obs_register: one device copy of the corners serves mcba_pnp_ransac and mcba_create; results unchanged."""
    from mcba import pnp
    s = synth.make_scene("c2", n_frames=40)
    o = synth.generate_observations(s)
    prob, init, gt = synth.make_problem(s, o, 5)
    p1, _, _ = synth.make_problem(s, o, 1)
    p1.u, p1.v, p1.pt_idx = prob.u, prob.v, prob.pt_idx           # the same host arrays, as in the reference's BoardObs
    P = pnp.PnpProblem(p1, s.intr_gt)
    opt = pnp.default_options(threshold_px=10.0, max_iterations=200, seed=3)
    r0 = pnp.ransac(P, opt)
    pa = init.copy(); h = mcba.Handle(prob); sa, ta = h.solve(pa); h.close()
    mcba.obs_register(prob)
    r1 = pnp.ransac(P, opt)
    pb = init.copy(); h = mcba.Handle(prob); sb, tb = h.solve(pb)
    with pytest.raises(mcba.McbaError):
        mcba.obs_release(prob)                                     # a live handle still reads the shared copy
    h.close()
    mcba.obs_release(prob)
    assert np.array_equal(r0.pose, r1.pose) and np.array_equal(r0.inlier_mask, r1.inlier_mask)
    assert trace_tuple(ta) == trace_tuple(tb) and np.array_equal(pa.intrinsics, pb.intrinsics)


def _ragged_problem(stage, seed=3):
    """c2 scene with random corner counts per board observation (0 .. all 16, odd and even) and an odd number of
    board observations: the tail handling of the fused observation kernels (half-empty passes, the zeroed partner row
    of an odd corner count, the unpaired last board observation of k_observations_pair)."""
    s = synth.make_scene("c2", n_frames=41)
    o = synth.generate_observations(s)
    prob, init, gt = synth.make_problem(s, o, stage)
    rng = np.random.default_rng(seed)
    counts = np.diff(prob.bobs_ptr)
    keep_bobs = np.ones(prob.n_bobs, bool)
    if prob.n_bobs % 2 == 0:
        keep_bobs[-1] = False
    new_counts = np.minimum(counts, rng.integers(1, counts.max() + 1, prob.n_bobs))
    new_counts[[5, 40, 41, 77, 78, 79]] = 0                      # empty board observations, also a whole empty pair / warp turn
    obs_keep = np.zeros(prob.n_obs, bool)
    for b in range(prob.n_bobs):
        if keep_bobs[b]:
            obs_keep[prob.bobs_ptr[b]:prob.bobs_ptr[b] + new_counts[b]] = True
    p2 = mcba.Problem(stage, prob.u[obs_keep], prob.v[obs_keep], prob.pt_idx[obs_keep],
                      np.concatenate([[0], np.cumsum(new_counts[keep_bobs])]), prob.bobs_cam[keep_bobs],
                      prob.bobs_pose[keep_bobs], prob.bobs_board[keep_bobs], prob.pts_xyz, prob.cam_model, prob.n_pose,
                      prob.n_board, prob.cam_is_ref, prob.board_is_ref)
    return p2, init


def _record_layout(stage):
    """(first index of the e-block part, length) of a compact per-board-observation record (Cols<STAGE>, mcba_math.cuh)."""
    wf = {1: 9, 2: 6, 3: 6, 4: 12, 5: 21}[stage]
    ci_fe = wf * (wf + 1) // 2 + wf
    return ci_fe, ci_fe + 6 * wf + 27


@pytest.mark.parametrize("stage", [1, 2, 3, 4, 5])
def test_fused_observation_kernel_variants_write_identical_records(stage, oracle):
    """The fused observation kernels against the first version (variant 1): k_observations_fused (0), k_observations_pair
    (3), the streaming experiment (2) and the 16-warp experiment (4). Same per-corner arithmetic and the same row order
    of the DMMA sums: the per-board-observation records and costs are identical bit for bit - on ragged corner counts,
    empty board observations and an odd count too. With the in-kernel pair accumulation on (32 + variant; an
    experiment, off by default) the e-block part of every record is still identical and the pair records
    (sums over the board observations of a (camera, board) pair, taken in another order) agree to rounding. Variant 1
    is the one whose normal equations are compared with the oracle's (test_normal_equations_and_schur runs the default)."""
    lib = mcba.load_lib()
    prob, init = _ragged_problem(stage)
    assert prob.n_bobs % 2 == 1 and len(set(np.diff(prob.bobs_ptr) % 2)) == 2
    ci_fe, ncomp = _record_layout(stage)
    ref = None
    try:
        for variant in (1, 0, 3, 2, 4, 32 + 0, 32 + 3):
            lib.mcba_debug_set_k1(variant)
            h = mcba.Handle(prob)
            h.solve_begin(init.copy(), mcba.default_options())
            rec = h.debug_fetch("brec").reshape(prob.n_bobs, ncomp).copy()
            pair = h.debug_fetch("pair_rec").copy()
            cost = h.eval(init.copy(), want_jac=False)[0]
            h.solve_end(init.copy())
            h.close()
            assert np.isfinite(pair).all() and np.isfinite(rec[:, ci_fe:]).all()
            if ref is None:
                assert np.isfinite(rec).all()
                ref = (rec, cost, pair)
                rc, co, _, _, _ = oracle.eval(prob, init.copy())
                assert rc == 0 and abs(cost - co) <= 1e-11 * abs(co)
                continue
            scale = np.abs(ref[0]).max()
            # variant 4 (24-corner passes) compiles obs_jacobian with other multiply-add contractions: rounding level
            tol = 1e-12 if variant == 4 else 1e-15
            accumulating = variant >= 32
            part = slice(ci_fe, None) if accumulating else slice(None)
            assert np.abs(rec[:, part] - ref[0][:, part]).max() <= tol * scale, f"variant {variant}"
            assert np.abs(pair - ref[2]).max() <= (1e-13 if accumulating or variant == 4 else 1e-15) * np.abs(ref[2]).max(), f"variant {variant}"
            assert abs(cost - ref[1]) <= 1e-14 * abs(ref[1])
    finally:
        lib.mcba_debug_set_k1(-1)


def test_problem_file_replay_and_resume(tmp_path, c2):
    """mcba_problem_save / mcba_problem_load (checkpoint / resume): a solve replayed from the file is
    bit-identical to the solve from the arrays in memory, and a solve resumed from the parameters saved after three
    iterations converges to the same minimum (to the solver's own tolerance)."""
    prob, init = stage_problem(c2, 5)[:2]
    path = str(tmp_path / "c2.mcba")
    mcba.save_problem(path, prob, init)
    p2, x2 = mcba.load_problem(path)
    ha, hb = mcba.Handle(prob), mcba.Handle(p2)
    pa, pb = init.copy(), x2.copy()
    sa, ta = ha.solve(pa)
    sb, tb = hb.solve(pb)
    assert trace_tuple(ta) == trace_tuple(tb) and sa.final_cost == sb.final_cost
    for k in ("cam_pose", "pose", "board_pose", "intrinsics"):
        assert np.array_equal(getattr(pa, k), getattr(pb, k)), k
    # checkpoint after 3 iterations, resume from the file
    pc = init.copy()
    sc, _ = ha.solve(pc, mcba.default_options(max_num_iterations=3))
    assert sc.termination == 1                                     # NO_CONVERGENCE: stopped by the iteration limit
    ck = str(tmp_path / "c2_it3.mcba")
    mcba.save_problem(ck, prob, pc)
    p3, x3 = mcba.load_problem(ck)
    hc = mcba.Handle(p3)
    sd, _ = hc.solve(x3)
    # a resumed solve restarts the trust region, so it stops (function tolerance 1e-6) at another point of the same basin
    assert sd.termination == 0 and abs(sd.rms_px - sa.rms_px) <= 1e-4 and sd.final_cost <= sc.final_cost
    assert rel(x3.intrinsics, pa.intrinsics) <= 1e-3
    for h in (ha, hb, hc):
        h.close()
