"""The product's trust-region bookkeeping (csrc/mcba_lm.cuh: the __host__ __device__ functions the host driver of the
single problem and the device-side driver of the batched problems both call), compiled for the host by tests/host_math:
 - against a Python restatement of ceres-solver 2.2.0's TrustRegionMinimizer / LevenbergMarquardtStrategy control flow
    on random step sequences that reach every branch: invalid steps (and FAILURE after
   max_num_consecutive_invalid_steps), rejected steps with the doubling decrease factor, accepted steps with the
   1/max(1/3, 1-(2 rho-1)^3) radius growth, parameter / function tolerance only after a first successful step,
   gradient tolerance, minimum radius, iteration limit;
 - against the CPU oracle's own traces: replaying the per-iteration quantities the oracle recorded through the
   product's rules reproduces its accept / reject sequence, radii and counters."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import mcba
from mcba import synth

HERE = os.path.dirname(os.path.abspath(__file__))
DBL_MAX = np.finfo(np.float64).max


@pytest.fixture(scope="module")
def lib():
    d = os.path.join(HERE, "host_math")
    subprocess.check_call(["make", "-s", "-C", d])
    L = C.CDLL(os.path.join(d, "libmathtest.so"))
    L.lmtest_replay.argtypes = [C.POINTER(mcba.mcba_options), C.c_double, C.c_double, C.c_double, C.c_int, mcba.c_f64p,
                                C.POINTER(mcba.mcba_trace_row), C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                C.POINTER(C.c_int), mcba.c_f64p]
    return L


def run_product(lib, opt, cost0, grad0, xnorm0, steps):
    steps = np.ascontiguousarray(steps, np.float64).reshape(-1, 7)
    rows = (mcba.mcba_trace_row * 4096)()
    term, ns, nu, fc = C.c_int(), C.c_int(), C.c_int(), C.c_double()
    n = lib.lmtest_replay(C.byref(opt), cost0, grad0, xnorm0, len(steps), steps.ctypes.data_as(mcba.c_f64p), rows, 4096,
                          C.byref(term), C.byref(ns), C.byref(nu), C.byref(fc))
    out = [dict(iteration=r.iteration, valid=r.step_is_valid, successful=r.step_is_successful, cost=r.cost,
                cost_change=r.cost_change, grad=r.gradient_max_norm, step_norm=r.step_norm, rel=r.relative_decrease,
                radius=r.trust_region_radius) for r in rows[:n]]
    return out, term.value, ns.value, nu.value, fc.value


def run_ceres_restatement(opt, cost0, grad0, xnorm0, steps):
    """trust_region_minimizer.cc / levenberg_marquardt_strategy.cc control flow (monotonic steps, no inner iterations)."""
    radius, decrease_factor = opt.initial_trust_region_radius, 2.0
    x_cost, x_norm = cost0, xnorm0
    it = dict(iteration=0, valid=1, successful=1, cost=cost0, cost_change=0.0, grad=grad0, step_norm=0.0, rel=0.0, radius=0.0)
    rows, n_succ, n_unsucc, n_invalid, at_least_one = [], 0, 0, 0, False
    k = 0
    while True:
        # FinalizeIterationAndCheckIfMinimizerCanContinue
        if it["successful"]:
            n_succ += 1
        else:
            n_unsucc += 1
        it["radius"] = radius
        rows.append(dict(it))
        if it["iteration"] >= opt.max_num_iterations:
            return rows, 1, n_succ, n_unsucc, x_cost
        if it["successful"] and it["grad"] <= opt.gradient_tolerance:
            return rows, 0, n_succ, n_unsucc, x_cost
        if radius <= opt.min_trust_region_radius:
            return rows, 0, n_succ, n_unsucc, x_cost
        prev_grad = it["grad"]
        it = dict(iteration=it["iteration"] + 1, valid=0, successful=0, cost=0.0, cost_change=0.0, grad=0.0, step_norm=0.0, rel=0.0, radius=0.0)
        if k >= len(steps):
            return rows, -1, n_succ, n_unsucc, x_cost
        solver_ok, mcc, step_norm, cand_norm, cand_cost, acc_cost, acc_grad = steps[k]
        k += 1
        it["valid"] = int(bool(solver_ok) and mcc > 0.0)
        if not it["valid"]:                                   # HandleInvalidStep
            n_invalid += 1
            if n_invalid >= opt.max_num_consecutive_invalid_steps:
                return rows, 2, n_succ, n_unsucc, x_cost
            radius /= decrease_factor
            decrease_factor *= 2.0
            it["cost"], it["grad"] = x_cost, prev_grad
            continue
        n_invalid = 0
        if at_least_one:
            it["step_norm"] = step_norm                       # ParameterToleranceReached
            if step_norm <= opt.parameter_tolerance * (x_norm + opt.parameter_tolerance):
                return rows, 0, n_succ, n_unsucc, x_cost
            it["cost_change"] = x_cost - cand_cost            # FunctionToleranceReached
            if abs(it["cost_change"]) <= opt.function_tolerance * x_cost:
                return rows, 0, n_succ, n_unsucc, x_cost
        rel = -DBL_MAX if cand_cost >= DBL_MAX else (x_cost - cand_cost) / mcc
        it["rel"] = rel
        if rel > opt.min_relative_decrease:                   # HandleSuccessfulStep + StepAccepted
            at_least_one = True
            x_norm, x_cost = cand_norm, acc_cost
            it["successful"], it["cost"], it["grad"] = 1, acc_cost, acc_grad
            t = 2.0 * rel - 1.0      # Ceres writes pow(t, 3): correctly rounded, at most 1 ulp from t * t * t (which is what
            radius = min(opt.max_trust_region_radius, radius / max(1.0 / 3.0, 1.0 - t * t * t))   # host and device can share)
            decrease_factor = 2.0
        else:                                                 # StepRejected
            it["successful"], it["cost"], it["grad"] = 0, cand_cost, prev_grad
            radius /= decrease_factor
            decrease_factor *= 2.0


def same(a, b):
    ra, ta, sa, ua, ca = a
    rb, tb, sb, ub, cb = b
    assert (ta, sa, ua) == (tb, sb, ub) and len(ra) == len(rb)
    assert ca == cb
    for x, y in zip(ra, rb):
        assert (x["iteration"], x["valid"], x["successful"]) == (y["iteration"], y["valid"], y["successful"])
        for key in ("cost", "grad", "radius"):
            assert x[key] == y[key], (key, x, y)        # same arithmetic, same order: bit-identical


def test_rules_match_the_ceres_restatement_on_random_sequences(lib):
    rng = np.random.default_rng(0)
    seen = set()
    for trial in range(600):
        opt = mcba.default_options(max_num_iterations=int(rng.choice([3, 8, 50])))
        if trial % 7 == 0:
            opt.min_trust_region_radius = 1e-2
        if trial % 11 == 0:
            opt.gradient_tolerance = 1.0
        cost, steps = float(10 ** rng.uniform(0, 6)), []
        c = cost
        for _ in range(60):
            kind = rng.choice(["good", "good", "good", "bad", "invalid", "nosolve", "tiny", "flat", "inf"])
            mcc = float(c * 10 ** rng.uniform(-6, -0.3))
            if kind == "good":
                cand = c - mcc * rng.uniform(0.05, 1.5)
            elif kind == "bad":
                cand = c + mcc * rng.uniform(0.0, 3.0)
            elif kind == "inf":
                cand = DBL_MAX
            elif kind == "flat":
                cand = c * (1 - 1e-9)
            else:
                cand = c - 0.5 * mcc
            if kind == "invalid":
                mcc = -mcc
            step_norm = 1e-12 if kind == "tiny" else float(10 ** rng.uniform(-4, 1))
            acc_grad = float(10 ** rng.uniform(-2, 6))
            steps.append([0.0 if kind == "nosolve" else 1.0, mcc, step_norm, 100.0 + rng.normal(), cand, cand, acc_grad])
            if kind == "good":
                c = cand
        got = run_product(lib, opt, cost, 1e3, 100.0, steps)
        want = run_ceres_restatement(opt, cost, 1e3, 100.0, steps)
        same(got, want)
        seen.add((want[1], len(want[0])))
    assert {t for t, _ in seen} >= {0, 1, 2}          # CONVERGENCE, NO_CONVERGENCE and FAILURE all occurred


def test_specific_branches(lib):
    opt = mcba.default_options(max_num_iterations=100)
    good = lambda c, frac=0.5: [1.0, c * 0.1, 1.0, 100.0, c - c * 0.1 * frac, c - c * 0.1 * frac, 10.0]
    # five consecutive invalid steps -> FAILURE; the radius was divided by 2, 4, 8, 16 on the way
    rows, term, ns, nu, _ = run_product(lib, opt, 100.0, 1.0, 10.0, [[1.0, -1.0, 1.0, 10.0, 50.0, 50.0, 1.0]] * 5)
    assert term == 2 and [r["radius"] for r in rows] == [1e4, 5e3, 1.25e3, 156.25, 9.765625]
    # tolerances are not tested before the first successful step: a tiny first step is accepted on its merits
    tiny = [1.0, 10.0, 1e-30, 10.0, 95.0, 95.0, 1.0]
    rows, term, ns, nu, _ = run_product(lib, opt, 100.0, 1.0, 10.0, [tiny, tiny])
    assert rows[1]["successful"] == 1 and term == 0 and len(rows) == 2      # second tiny step: parameter tolerance
    # function tolerance after a success
    rows, term, ns, nu, fc = run_product(lib, opt, 100.0, 1.0, 10.0, [good(100.0), [1.0, 1e-3, 1.0, 10.0, 95.0 - 1e-6, 0, 0]])
    assert term == 0 and len(rows) == 2 and fc == 95.0
    # gradient tolerance after a successful step
    s = good(100.0)
    s[6] = 1e-11
    rows, term, ns, nu, _ = run_product(lib, opt, 100.0, 1.0, 10.0, [s, good(95.0)])
    assert term == 0 and len(rows) == 2
    # rho = 1 triples the radius, capped at max_trust_region_radius
    opt2 = mcba.default_options(max_num_iterations=100, max_trust_region_radius=2.5e4)
    rows, term, ns, nu, _ = run_product(lib, opt2, 100.0, 1.0, 10.0, [good(100.0, 1.0), good(90.0, 1.0)])
    assert [r["radius"] for r in rows] == [1e4, 2.5e4, 2.5e4]


@pytest.mark.parametrize("cfg,stage", [("c1", 5), ("c2", 4)])
def test_replaying_oracle_traces_through_the_product_rules(lib, oracle, cfg, stage):
    """The oracle's recorded iterations carry cost, cost_change, relative_decrease, step_norm and gradient norm: enough to
    rebuild what the product's driver would have been handed at each iteration."""
    prob, init, gt, s, o = synth.config_problem(cfg, stage, n_frames=60 if cfg == "c2" else None)
    p = init.copy()
    if stage == 5:
        p.intrinsics[:] = s.intr_init
    opt = mcba.default_options(initial_trust_region_radius=1e2 if cfg == "c1" else 1e4)
    rc, summ, tr = oracle.solve(prob, p, opt)
    assert rc == 0
    steps, x_cost = [], tr[0].cost
    for r in tr[1:]:
        if not r.step_is_valid:
            steps.append([1.0, -1.0, 0.0, 0.0, x_cost, 0.0, 0.0])
            continue
        cand = x_cost - r.cost_change if r.cost_change != 0.0 else r.cost
        mcc = (x_cost - cand) / r.relative_decrease
        steps.append([1.0, mcc, r.step_norm if r.step_norm > 0 else 1.0, 1e3, cand, r.cost, r.gradient_max_norm])
        if r.step_is_successful:
            x_cost = r.cost
    rows, term, ns, nu, _ = run_product(lib, opt, tr[0].cost, tr[0].gradient_max_norm, 1e3, steps)
    assert len(rows) == len(tr) and term == -1          # the terminating iteration itself is never recorded by Ceres
    for a, b in zip(rows, tr):
        assert (a["iteration"], a["valid"], a["successful"]) == (b.iteration, b.step_is_valid, b.step_is_successful)
        assert abs(a["radius"] - b.trust_region_radius) <= 1e-9 * b.trust_region_radius
        assert abs(a["cost"] - b.cost) <= 1e-12 * abs(b.cost)
    assert (ns, nu) == (summ.num_successful_steps, summ.num_unsuccessful_steps)
