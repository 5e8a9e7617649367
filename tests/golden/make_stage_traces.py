"""Generates tests/golden/trace_c2_stage{1,2,3}.json: Levenberg-Marquardt traces of the CPU oracle for the three
refinement stages that make_golden.py's stage 4/5 chain does not cover (config 2, 120 frames, each stage from its own
perturbed initial values). Regression pins for the oracle and, through tests/test_gpu_parity.py, for the CUDA path."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(REPO, "mc-calib_b200", "python"))
sys.path.insert(0, os.path.join(REPO, "oracle"))
import mcba  # noqa: E402
from mcba import synth  # noqa: E402
import oracle_py  # noqa: E402

for stage in (1, 2, 3):
    prob, init, gt, s, o = synth.config_problem("c2", stage, n_frames=120)
    p = init.copy()
    rc, summ, tr = oracle_py.solve(prob, p, mcba.default_options())
    assert rc == 0
    doc = dict(config="c2", stage=stage, n_frames=120, n_obs=int(prob.n_obs), termination=int(summ.termination),
               num_successful_steps=int(summ.num_successful_steps), num_unsuccessful_steps=int(summ.num_unsuccessful_steps),
               initial_cost=summ.initial_cost, final_cost=summ.final_cost, rms_px=summ.rms_px,
               rows=[dict(iteration=r.iteration, valid=r.step_is_valid, successful=r.step_is_successful, cost=r.cost,
                          radius=r.trust_region_radius, gradient_max_norm=r.gradient_max_norm) for r in tr],
               intrinsics=p.intrinsics.tolist(), cam_pose=p.cam_pose.tolist(), board_pose=p.board_pose.tolist(),
               pose_first10=p.pose[:10].tolist())
    json.dump(doc, open(os.path.join(HERE, f"trace_c2_stage{stage}.json"), "w"), indent=1)
    print(f"stage {stage}: {prob.n_obs} obs, {len(tr)} recorded iterations, termination {summ.termination}, rms {summ.rms_px:.6f}")
