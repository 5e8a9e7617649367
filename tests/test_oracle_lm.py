"""Pins the oracle's Levenberg-Marquardt restatement: the committed traces, the minimum found by an independent
algorithm (scipy trust-region-reflective with the same Huber objective), and Ceres' bookkeeping rules."""
import json
import os

import numpy as np
import pytest
from scipy.optimize import least_squares

import mcba
from mcba import synth
from util import dense_jacobian, layout, params_to_x

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("cfg,stage", [("c1", 4), ("c1", 5), ("c2", 4), ("c2", 5)])
def test_trace_fixture(cfg, stage, oracle):
    doc = json.load(open(os.path.join(GOLD, f"trace_{cfg}_stage{stage}.json")))
    prob4, init, gt, s, o = synth.config_problem(cfg, 4)
    p = init.copy()
    if stage == 5:   # the fixture chain is stage 4 then stage 5, like apps/calibrate/src/calibrate.cpp:60-64
        rc, _, _ = oracle.solve(prob4, p)
        p.intrinsics[:] = s.intr_init
    rc, summ, tr = oracle.solve(prob4.with_stage(stage), p)
    assert rc == 0 and prob4.n_obs == doc["n_obs"]
    assert summ.termination == doc["termination"]
    assert (summ.num_successful_steps, summ.num_unsuccessful_steps) == (doc["num_successful_steps"], doc["num_unsuccessful_steps"])
    assert len(tr) == len(doc["rows"])
    for r, d in zip(tr, doc["rows"]):
        assert (r.iteration, r.step_is_valid, r.step_is_successful) == (d["iteration"], d["valid"], d["successful"])
        assert abs(r.cost - d["cost"]) <= 1e-9 * d["cost"]
        assert abs(r.trust_region_radius - d["radius"]) <= 1e-6 * d["radius"]
    assert abs(summ.rms_px - doc["rms_px"]) <= 1e-8
    assert np.abs(p.intrinsics - np.array(doc["intrinsics"])).max() <= 1e-7 * np.abs(p.intrinsics).max()


@pytest.mark.parametrize("stage", [1, 2, 3])
def test_trace_fixture_stages_1_to_3(stage, oracle):
    doc = json.load(open(os.path.join(GOLD, f"trace_c2_stage{stage}.json")))
    prob, init, gt, s, o = synth.config_problem("c2", stage, n_frames=doc["n_frames"])
    p = init.copy()
    rc, summ, tr = oracle.solve(prob, p)
    assert rc == 0 and prob.n_obs == doc["n_obs"]
    assert (summ.termination, summ.num_successful_steps, summ.num_unsuccessful_steps) == \
           (doc["termination"], doc["num_successful_steps"], doc["num_unsuccessful_steps"])
    assert len(tr) == len(doc["rows"])
    for r, d in zip(tr, doc["rows"]):
        assert (r.iteration, r.step_is_valid, r.step_is_successful) == (d["iteration"], d["valid"], d["successful"])
        assert abs(r.cost - d["cost"]) <= 1e-9 * d["cost"]
        assert abs(r.trust_region_radius - d["radius"]) <= 1e-6 * d["radius"]
    assert abs(summ.rms_px - doc["rms_px"]) <= 1e-8
    assert np.abs(p.pose[:10] - np.array(doc["pose_first10"])).max() <= 1e-7


def test_trace_bookkeeping(oracle):
    """Ceres rules: iteration 0 is recorded as a successful step with radius = initial radius; after a successful
    step with ratio rho the radius grows by 1/max(1/3, 1-(2 rho-1)^3); cost is monotone over successful steps."""
    prob, init, gt, s, o = synth.config_problem("c1", 5)
    rc, summ, tr = oracle.solve(prob, init.copy())
    assert tr[0].iteration == 0 and tr[0].step_is_successful == 1 and tr[0].trust_region_radius == 1e4
    assert summ.num_successful_steps + summ.num_unsuccessful_steps == len(tr)
    for a, b in zip(tr[:-1], tr[1:]):
        if b.step_is_successful:
            assert b.cost < a.cost
            grow = 1.0 / max(1.0 / 3.0, 1.0 - (2.0 * b.relative_decrease - 1.0) ** 3)
            assert abs(b.trust_region_radius - min(1e16, a.trust_region_radius * grow)) <= 1e-9 * b.trust_region_radius
    assert summ.termination == 0


def test_max_iterations_and_no_convergence(oracle):
    prob, init, gt, s, o = synth.config_problem("c1", 5)
    rc, summ, tr = oracle.solve(prob, init.copy(), mcba.default_options(max_num_iterations=2))
    assert summ.termination == 1 and len(tr) == 3 and tr[-1].iteration == 2


def test_rejected_steps_shrink_radius(oracle):
    """A tiny min_relative_decrease can't be violated, so force rejections with an absurd one: every step is
    rejected, radius halves, then quarters ... (StepRejected: radius /= decrease_factor; decrease_factor *= 2)."""
    prob, init, gt, s, o = synth.config_problem("c1", 4)
    rc, summ, tr = oracle.solve(prob, init.copy(), mcba.default_options(max_num_iterations=4, min_relative_decrease=1e9))
    radii = [r.trust_region_radius for r in tr]
    assert [r.step_is_successful for r in tr] == [1, 0, 0, 0, 0]
    assert np.allclose(radii, [1e4, 5e3, 1.25e3, 156.25, 9.765625])


def test_minimum_matches_scipy_huber(oracle):
    """Independent minimiser, same objective 1/2 sum rho(|r_i|^2): scipy's loss applies rho per scalar residual, so
    the per-observation Huber on the 2-vector is reproduced by handing scipy the pre-robustified residuals."""
    s = synth.make_scene("c1", n_frames=40)
    o = synth.generate_observations(s)
    prob, init, gt = synth.make_problem(s, o, 4)
    p = init.copy()
    rc, summ, tr = oracle.solve(prob, p, mcba.default_options(function_tolerance=1e-14, parameter_tolerance=1e-14))
    x_star = params_to_x(prob, p)
    L = layout(prob)

    def unpack(x):
        q = init.copy()
        q.pose[:] = x[:L["n_e"]].reshape(-1, 6)
        f = x[L["n_e"]:]
        q.cam_pose[:] = f[:prob.n_cam * 6].reshape(-1, 6)
        q.board_pose[:] = f[prob.n_cam * 6:].reshape(-1, 6)
        return q

    def fun_jac(x):     # |r_i| per observation; scipy's loss='huber' (f_scale=1) applies the same rho to |r_i|^2
        q = unpack(x)
        _, cost, res, jac, _ = oracle.eval(prob, q, mcba.default_options(huber_a=0.0))
        f = np.sqrt((res ** 2).sum(1))
        J = dense_jacobian(prob, jac)                       # (2N, n_x), rows 2i, 2i+1
        Jf = (res[:, 0:1] * J[0::2] + res[:, 1:2] * J[1::2]) / np.maximum(f, 1e-300)[:, None]
        return f, Jf

    free = np.ones(L["n_e"] + L["n_f"], bool)
    free[L["n_e"]:L["n_e"] + 6] = False                      # reference camera
    free[L["n_e"] + prob.n_cam * 6:L["n_e"] + prob.n_cam * 6 + 6] = False   # reference board
    x0 = params_to_x(prob, init)

    def full(z):
        x = x0.copy(); x[free] = z
        return x

    z0 = x_star[free] + 1e-4 * np.random.default_rng(0).normal(size=int(free.sum()))
    sol = least_squares(lambda z: fun_jac(full(z))[0], z0, x_scale="jac", jac=lambda z: fun_jac(full(z))[1][:, free],
                        method="trf", loss="huber", f_scale=1.0, xtol=1e-15, ftol=1e-15, gtol=1e-13, max_nfev=30)
    # scipy's TRF crawls on this objective; what it must not do is find a lower minimum than the oracle's
    assert sol.cost >= summ.final_cost * (1 - 1e-10)
    assert sol.cost - summ.final_cost <= 1e-3 * summ.final_cost
    # first-order optimality of the oracle's point under the independent objective: g = sum rho' J^T r
    def grad(x):
        _, _, res, jac, _ = oracle.eval(prob, unpack(x), mcba.default_options(huber_a=0.0))
        sq = (res ** 2).sum(1)
        w = np.where(sq <= 1, 1.0, 1.0 / np.sqrt(np.maximum(sq, 1e-300)))
        J = dense_jacobian(prob, jac)
        return ((w[:, None] * res).reshape(-1) @ J)[free]
    g0, g1 = np.abs(grad(x0)).max(), np.abs(grad(x_star)).max()
    assert g1 <= 1e-7 * g0


@pytest.mark.parametrize("stage", [1, 2, 3, 4, 5])
def test_first_lm_step_matches_dense_numpy(stage, oracle):
    """LevenbergMarquardtStrategy::ComputeStep restated densely in numpy (Jacobi scaling 1/(1+|J_j|), diagonal clamp,
    D = sqrt(diag/radius), one dense solve of (Js^T Js + D^2) y = Js^T r, delta = -s o y): the oracle's Schur elimination
    + Cholesky must produce the same first step (step norm, model cost change through rho, cost at the new point)."""
    prob, init, gt, s, o = synth.config_problem("c1", stage, n_frames=30)
    p = init.copy()
    if stage in (1, 5):
        p.intrinsics[:] = s.intr_init
    opt = mcba.default_options(max_num_iterations=1)
    p0 = p.copy()
    rc, summ, tr = oracle.solve(prob, p, opt)
    assert rc == 0 and len(tr) == 2 and tr[1].step_is_valid == 1
    _, cost0, res, jac, _ = oracle.eval(prob, p0, opt)
    J = dense_jacobian(prob, jac)
    r = res.reshape(-1)
    sc = 1.0 / (1.0 + np.sqrt((J * J).sum(0)))
    Js = J * sc
    diag = np.clip((Js * Js).sum(0), opt.min_lm_diagonal, opt.max_lm_diagonal)
    H = Js.T @ Js + np.diag(diag / opt.initial_trust_region_radius)
    y = np.linalg.solve(H, Js.T @ r)
    delta = -sc * y
    assert tr[1].step_norm == 0.0   # Ceres only fills step_norm in ParameterToleranceReached, i.e. after a first success
    m = Js @ (-y)
    mcc = -m @ (r + 0.5 * m)
    x1 = params_to_x(prob, p0) + delta
    assert tr[1].step_is_successful == 1
    assert np.abs(params_to_x(prob, p) - x1).max() <= 1e-7 * max(1.0, np.abs(x1).max())
    assert abs((cost0 - tr[1].cost) / mcc - tr[1].relative_decrease) <= 1e-6 * abs(tr[1].relative_decrease)
