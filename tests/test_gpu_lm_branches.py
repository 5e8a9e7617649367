"""-m gpu parity on the non-accept branches of ceres::Solve: rejected steps, invalid steps
(recovering and 5 in a row -> FAILURE), NO_CONVERGENCE at max_num_iterations, parameter-tolerance, gradient-tolerance
and min-trust-region-radius exits - for the host-driven single-problem driver (mcba_solve) and the device-side
batched driver (mcba_solve_batched / k_b_decide). The 2-rank driver runs a subset in tests/mgpu_worker.py.

Bar: identical (iteration, valid, successful) sequences and terminations, radii within 1e-6, accepted costs within
1e-9, final parameters within 1e-6 - against the oracle run live on the same inputs AND against the committed golden
traces (tests/golden/lm_branch_traces.json, generator beside it)."""
import numpy as np
import pytest

import lm_cases
import mcba
from test_gpu_parity import _camera_subproblem

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", list(lm_cases.CASES))
def test_single_problem_driver_follows_the_oracle_through_every_branch(name, oracle):
    prob, p0, opt, _ = lm_cases.make_case(name)
    pg, po = p0.copy(), p0.copy()
    h = mcba.Handle(prob)
    sg, tg = h.solve(pg, opt)
    h.close()
    rc, so, to = oracle.solve(prob, po, opt)
    assert rc == 0
    lm_cases.check_expectation(name, sg, tg)            # the GPU run itself contains the rejections / invalid steps
    lm_cases.assert_same_run(name, (pg, sg, tg), (po, so, to))
    g = lm_cases.golden()["single"][name]
    assert lm_cases.sequence(tg) == g["sequence"] and sg.termination == g["termination"]
    for r, w in zip(tg, g["rows"]):
        assert abs(r.trust_region_radius - w["radius"]) <= 1e-6 * w["radius"]
    assert np.abs(pg.intrinsics - np.array(g["intrinsics"])).max() <= 1e-6 * np.abs(pg.intrinsics).max()


def test_rejected_step_leaves_the_current_point_and_tiles_untouched():
    """After a rejected step the driver re-solves with the cached blocks: parameters, F^T F and the e-block rows must be
    bit-identical to what they were before the step (mcba_api.cu: alternate tile set, no swap on rejection)."""
    prob, p0, opt, _ = lm_cases.make_case("reject_c1_s5")
    h = mcba.Handle(prob)
    h.solve_begin(p0.copy(), opt)
    seen_reject = False
    for _ in range(12):
        ff0, wt0 = h.debug_fetch("FF"), h.debug_fetch("Wt")
        if h.solve_step() <= 0:
            break
        ff1, wt1 = h.debug_fetch("FF"), h.debug_fetch("Wt")
        same = np.array_equal(ff0, ff1) and np.array_equal(wt0, wt1)
        seen_reject = seen_reject or same
    assert seen_reject
    p = p0.copy()
    h.solve_end(p)
    h.close()


@pytest.mark.parametrize("name", list(lm_cases.BATCHED_CASES))
def test_batched_device_driver_follows_each_cameras_oracle(name, oracle):
    prob, p0, opt = lm_cases.make_batched_case(name)
    pg = p0.copy()
    h = mcba.Handle(prob)
    summ = h.solve_batched(pg, opt)
    h.close()
    g = lm_cases.golden()["batched"][name]
    for c in range(prob.n_cam):
        sub, p, (b0, b1) = _camera_subproblem(prob, p0, c)
        rc, so, to = oracle.solve(sub, p, opt)
        assert rc == 0
        s = summ[c]
        assert (s.termination, s.num_successful_steps, s.num_unsuccessful_steps, s.num_iterations) == \
               (so.termination, so.num_successful_steps, so.num_unsuccessful_steps, so.num_iterations), (name, c)
        assert (s.termination, s.num_iterations) == (g[c]["termination"], g[c]["num_iterations"])
        assert abs(s.final_cost - so.final_cost) <= (1e-9 if so.termination == 0 else 1e-6) * so.final_cost, (name, c)
        assert np.abs(pg.intrinsics[c] - p.intrinsics[0]).max() <= 1e-6 * np.abs(p.intrinsics[0]).max(), (name, c)
        assert np.abs(pg.pose[b0:b1] - p.pose).max() <= 1e-6 * max(1.0, np.abs(p.pose).max()), (name, c)
