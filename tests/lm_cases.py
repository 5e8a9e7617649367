"""Named start points / solver options that drive ceres::Solve's trust-region loop through every branch other than
"step accepted": rejected steps, invalid steps (recovering and 5-in-a-row -> FAILURE), NO_CONVERGENCE at
max_num_iterations, and the parameter-tolerance / gradient-tolerance / min-trust-region-radius exits.

Reference defaults these options override: /root/reference/McCalib/src/CameraGroup.cpp:463-468, Camera.cpp:296-301,
Object3D.cpp:218-223 (only max_num_iterations is set there; everything else is the Ceres 2.2.0 default).

Used by the CPU suite (the oracle really takes the intended branch; committed golden traces) and by the -m gpu parity
tests of mcba_solve, mcba_solve_batched and the 2-rank worker, so that "identical accept/reject sequence" is checked
on sequences that do contain rejections.
"""
import json
import os

import numpy as np

import mcba
from mcba import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lm_branch_traces.json")

# name -> (config, frames, stage, far, options, expectation)
# `far` moves the initial e-block poses `far` times further from the ground truth than the usual start; with a strict
# min_relative_decrease most steps of such a start are rejected (radius halved, quartered, ...).
CASES = {
    "reject_c1_s1": ("c1", 40, 1, 60, dict(min_relative_decrease=0.9, max_num_iterations=40), dict(min_unsuccessful=5, min_successful=5)),
    "reject_c1_s4": ("c1", 40, 4, 45, dict(min_relative_decrease=0.9, max_num_iterations=40), dict(min_unsuccessful=5, min_successful=5)),
    "reject_c1_s5": ("c1", 40, 5, 60, dict(min_relative_decrease=0.9, max_num_iterations=40), dict(min_unsuccessful=5, min_successful=5)),
    "reject_c2_s1": ("c2", 60, 1, 60, dict(min_relative_decrease=0.9, max_num_iterations=40), dict(min_unsuccessful=5, min_successful=5)),
    "reject_c2_s4": ("c2", 60, 4, 40, dict(min_relative_decrease=0.9, max_num_iterations=40), dict(min_unsuccessful=5, min_successful=5)),
    "reject_c2_s5": ("c2", 60, 5, 60, dict(min_relative_decrease=0.9, max_num_iterations=40), dict(min_unsuccessful=5, min_successful=5)),
    # invalid steps that recover: with min_lm_diagonal = 1e-20 and a trust-region radius of 1e306 the damping
    # diag / radius of the structurally zero Jacobian columns (reference camera / board, Kannala intr[8]) underflows to
    # exactly 0 -> singular reduced system -> invalid step, radius / 2, / 4, / 8, / 16 -> 1e-20 / 9.8e302 is a positive
    # denormal again and the step is valid. Zero / non-zero is decided by IEEE division alone, not by conditioning.
    "invalid_recover_c2_s5": ("c2", 60, 5, 1, dict(min_lm_diagonal=1e-20, initial_trust_region_radius=1e306, max_trust_region_radius=1e307),
                              dict(termination=0, min_invalid=6, min_successful=4)),
    "small_radius_c1_s5": ("c1", 40, 5, 1, dict(initial_trust_region_radius=1e-1, max_num_iterations=40), dict(termination=0)),
    "no_convergence_c1_s5": ("c1", 40, 5, 1, dict(max_num_iterations=3), dict(termination=1, iterations=4)),
    "no_convergence_c2_s5": ("c2", 60, 5, 1, dict(max_num_iterations=3), dict(termination=1, iterations=4)),
    "parameter_tolerance_c1_s5": ("c1", 40, 5, 1, dict(parameter_tolerance=1e-4, function_tolerance=1e-12), dict(termination=0)),
    "parameter_tolerance_c2_s5": ("c2", 60, 5, 1, dict(parameter_tolerance=1e-4, function_tolerance=1e-12), dict(termination=0)),
    "gradient_tolerance_at_start_c1_s5": ("c1", 40, 5, 1, dict(gradient_tolerance=1e12), dict(termination=0, iterations=1)),
    "min_radius_c1_s5": ("c1", 40, 5, 60, dict(min_relative_decrease=0.9, min_trust_region_radius=1e2, max_num_iterations=40), dict(termination=0, min_unsuccessful=3)),
    "min_radius_c2_s5": ("c2", 60, 5, 60, dict(min_relative_decrease=0.9, min_trust_region_radius=1e2, max_num_iterations=40), dict(termination=0, min_unsuccessful=3)),
    # min_lm_diagonal = 0: the zero Jacobian columns of the reference camera / reference board (and Kannala intr[8])
    # make the damped reduced system singular -> five invalid steps in a row -> FAILURE
    "invalid_failure_c1_s4": ("c1", 40, 4, 1, dict(min_lm_diagonal=0.0), dict(termination=2, min_invalid=4)),
    "invalid_failure_c2_s5": ("c2", 60, 5, 1, dict(min_lm_diagonal=0.0), dict(termination=2, min_invalid=4)),
}

_cache = {}


def make_case(name):
    """-> (Problem, start Params, mcba_options, expectation dict)"""
    cfg, frames, stage, far, okw, expect = CASES[name]
    key = (cfg, frames, stage)
    if key not in _cache:
        _cache[key] = synth.config_problem(cfg, stage, n_frames=frames)
    prob, init, gt, s, o = _cache[key]
    p = init.copy()
    if stage in (1, 5):
        p.intrinsics[:] = s.intr_init
    p.pose[:] = gt.pose + far * (p.pose - gt.pose)
    return prob, p, mcba.default_options(**okw), expect


def sequence(trace):
    """'A' accepted, 'r' rejected (valid, unsuccessful), 'i' invalid — one letter per recorded iteration."""
    return "".join("A" if r.step_is_successful else ("r" if r.step_is_valid else "i") for r in trace)


def check_expectation(name, summ, trace):
    e = CASES[name][5]
    seq = sequence(trace)
    if "termination" in e:
        assert summ.termination == e["termination"], (name, summ.termination, seq)
    if "iterations" in e:
        assert summ.num_iterations == e["iterations"], (name, summ.num_iterations)
    assert seq.count("r") + seq.count("i") >= e.get("min_unsuccessful", 0), (name, seq)
    assert seq.count("i") >= e.get("min_invalid", 0), (name, seq)
    assert seq.count("A") >= e.get("min_successful", 0), (name, seq)


def golden():
    with open(GOLD) as f:
        return json.load(f)


def assert_same_run(name, got, want, final_tol=1e-6):
    """got / want: (Params, summary, trace). Identical (iteration, valid, successful) sequence and termination, radii
    within 1e-6 relative, costs within 1e-6 relative (the far starts sit at costs of 1e6..1e13 where the step is
    ill-conditioned: one LM step amplifies the 1e-12 evaluation differences to ~1e-8 of the cost; a rejected step's
    recorded cost is that of a wild candidate - up to 1e20 - and is compared at 1e-2), final cost within 1e-9 (1e-6 for the far starts),
    final parameters within `final_tol` relative."""
    (pg, sg, tg), (po, so, to) = got, want
    assert sequence(tg) == sequence(to), (name, sequence(tg), sequence(to))
    assert [r.iteration for r in tg] == [r.iteration for r in to]
    assert sg.termination == so.termination, name
    assert (sg.num_successful_steps, sg.num_unsuccessful_steps, sg.num_iterations) == \
           (so.num_successful_steps, so.num_unsuccessful_steps, so.num_iterations), name
    for a, b in zip(tg, to):
        assert abs(a.trust_region_radius - b.trust_region_radius) <= 1e-6 * b.trust_region_radius, (name, a.iteration)
        tol = 1e-6 if b.step_is_successful else 1e-2
        assert abs(a.cost - b.cost) <= tol * abs(b.cost), (name, a.iteration, a.cost, b.cost)
    far = CASES[name][3] if name in CASES else 1     # far starts: see above (observed: a constant 1e-8 relative offset)
    assert abs(sg.final_cost - so.final_cost) <= (1e-9 if far == 1 else 1e-6) * abs(so.final_cost), name
    for attr in ("cam_pose", "pose", "board_pose", "intrinsics"):
        a, b = getattr(pg, attr), getattr(po, attr)
        if a.size:
            assert np.abs(a - b).max() <= final_tol * max(1.0, np.abs(b).max()), (name, attr, np.abs(a - b).max())


# ---- batched (config 5 shape) branch cases: name -> (n_cameras, n_frames, kannala_every, far, options)
BATCHED_CASES = {
    "batched_reject": (6, 40, 3, 30, dict(min_relative_decrease=0.9, max_num_iterations=40)),
    "batched_no_convergence": (5, 30, 0, 1, dict(max_num_iterations=3)),
    # Kannala cameras (every 2nd) have the structurally zero intr[8] column: singular with min_lm_diagonal = 0 -> FAILURE;
    # the Brown cameras of the same batch converge
    "batched_invalid_failure": (6, 30, 2, 1, dict(min_lm_diagonal=0.0)),
    "batched_min_radius": (4, 30, 0, 30, dict(min_relative_decrease=0.9, min_trust_region_radius=1e2, max_num_iterations=40)),
}


def make_batched_case(name):
    n_cam, n_frames, kan, far, okw = BATCHED_CASES[name]
    prob, init, gt = synth.batched_intrinsics_problem(n_cameras=n_cam, n_frames=n_frames, kannala_every=kan)
    p = init.copy()
    p.pose[:] = gt.pose + far * (p.pose - gt.pose)
    return prob, p, mcba.default_options(**okw)
