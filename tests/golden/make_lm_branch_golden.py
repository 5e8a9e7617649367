"""Generates tests/golden/lm_branch_traces.json: the CPU oracle's LM traces, terminations and final f-blocks on the
branch cases of tests/lm_cases.py (rejected / invalid steps, FAILURE, NO_CONVERGENCE, tolerance exits) and the
per-camera outcomes of the batched cases. Run here (CPU container): python tests/golden/make_lm_branch_golden.py"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
for d in ("mc-calib_b200/python", "oracle", "tests"):
    sys.path.insert(0, os.path.join(REPO, d))
import lm_cases  # noqa: E402
import oracle_py  # noqa: E402
from test_gpu_parity import _camera_subproblem  # noqa: E402  (pure numpy helper)


def summary_doc(summ, tr):
    return dict(termination=int(summ.termination), num_successful_steps=int(summ.num_successful_steps),
                num_unsuccessful_steps=int(summ.num_unsuccessful_steps), num_iterations=int(summ.num_iterations),
                final_cost=summ.final_cost, rms_px=summ.rms_px, sequence=lm_cases.sequence(tr),
                rows=[dict(iteration=r.iteration, valid=r.step_is_valid, successful=r.step_is_successful, cost=r.cost,
                           radius=r.trust_region_radius) for r in tr])


def main():
    doc = dict(generator="tests/golden/make_lm_branch_golden.py", single={}, batched={})
    for name in lm_cases.CASES:
        prob, p, opt, _ = lm_cases.make_case(name)
        rc, summ, tr = oracle_py.solve(prob, p, opt)
        assert rc == 0
        lm_cases.check_expectation(name, summ, tr)
        d = summary_doc(summ, tr)
        d.update(cam_pose=p.cam_pose.tolist(), board_pose=p.board_pose.tolist(), intrinsics=p.intrinsics.tolist())
        doc["single"][name] = d
        print(name, d["sequence"], d["termination"])
    for name in lm_cases.BATCHED_CASES:
        prob, p0, opt = lm_cases.make_batched_case(name)
        cams = []
        for c in range(prob.n_cam):
            sub, p, _ = _camera_subproblem(prob, p0, c)
            rc, summ, tr = oracle_py.solve(sub, p, opt)
            assert rc == 0
            d = summary_doc(summ, tr)
            d.update(intrinsics=p.intrinsics[0].tolist())
            cams.append(d)
        doc["batched"][name] = cams
        print(name, [(c["sequence"], c["termination"]) for c in cams])
    with open(lm_cases.GOLD, "w") as f:
        json.dump(doc, f, indent=0)


if __name__ == "__main__":
    main()
