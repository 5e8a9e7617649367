"""CPU side of the LM branch cases (tests/lm_cases.py): the oracle takes the intended branch on each case, reproduces
the committed golden traces, and a second, independent dense restatement of ceres::Solve (test_oracle_lm.py) follows
the same accept / reject / invalid sequence - so the -m gpu parity tests on these cases are not vacuous."""
import numpy as np
import pytest

import lm_cases
from test_gpu_parity import _camera_subproblem
from test_oracle_lm import dense_trust_region_solve
from util import params_to_x


@pytest.mark.parametrize("name", list(lm_cases.CASES))
def test_oracle_takes_the_intended_branch_and_matches_golden(name, oracle):
    prob, p, opt, _ = lm_cases.make_case(name)
    rc, summ, tr = oracle.solve(prob, p, opt)
    assert rc == 0
    lm_cases.check_expectation(name, summ, tr)
    g = lm_cases.golden()["single"][name]
    assert lm_cases.sequence(tr) == g["sequence"]
    assert (summ.termination, summ.num_iterations) == (g["termination"], g["num_iterations"])
    for r, w in zip(tr, g["rows"]):
        assert abs(r.cost - w["cost"]) <= 1e-9 * abs(w["cost"])
        assert abs(r.trust_region_radius - w["radius"]) <= 1e-12 * w["radius"]
    assert np.abs(p.intrinsics - np.array(g["intrinsics"])).max() <= 1e-9 * np.abs(p.intrinsics).max()


@pytest.mark.parametrize("name", ["reject_c1_s4", "invalid_recover_c2_s5", "invalid_failure_c1_s4", "min_radius_c1_s5",
                                  "no_convergence_c1_s5", "parameter_tolerance_c1_s5", "gradient_tolerance_at_start_c1_s5",
                                  "small_radius_c1_s5"])
def test_independent_dense_minimizer_follows_the_same_branches(name, oracle):
    prob, p, opt, _ = lm_cases.make_case(name)
    po = p.copy()
    rc, summ, tr = oracle.solve(prob, po, opt)
    rows, term, x = dense_trust_region_solve(oracle, prob, p, opt)
    assert term == summ.termination
    assert "".join("A" if r["successful"] else ("r" if r["valid"] else "i") for r in rows) == lm_cases.sequence(tr)
    for a, b in zip(rows, tr):
        assert abs(a["radius"] - b.trust_region_radius) <= 1e-6 * b.trust_region_radius
    assert np.abs(x - params_to_x(prob, po)).max() <= 1e-6 * max(1.0, np.abs(x).max())


@pytest.mark.parametrize("name", list(lm_cases.BATCHED_CASES))
def test_batched_cases_per_camera_oracle_matches_golden(name, oracle):
    prob, p0, opt = lm_cases.make_batched_case(name)
    g = lm_cases.golden()["batched"][name]
    seqs = []
    for c in range(prob.n_cam):
        sub, p, _ = _camera_subproblem(prob, p0, c)
        rc, summ, tr = oracle.solve(sub, p, opt)
        assert rc == 0
        assert (lm_cases.sequence(tr), summ.termination) == (g[c]["sequence"], g[c]["termination"])
        seqs.append(lm_cases.sequence(tr))
    if name == "batched_reject":
        assert max(s.count("r") for s in seqs) >= 5
    if name == "batched_invalid_failure":
        assert sorted(set(x["termination"] for x in g)) == [0, 2]
