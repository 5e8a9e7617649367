"""tests/golden/c5_sample.json: CPU-oracle outcomes of a few cameras of BASELINE config 5 (4 096 independent
per-camera intrinsic refinements x 300 frames, Camera::refineIntrinsicCalibration, Camera.cpp:268-305), each solved
as its own single problem - what bench.py's config-5 line and the -m gpu tests compare the batched launch with.
The generator is keyed by (seed, camera), so camera c of the full batch is reproduced without building the batch.
    python tests/golden/make_c5_golden.py"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
for d in ("mc-calib_b200/python", "oracle"):
    sys.path.insert(0, os.path.join(REPO, d))
import mcba  # noqa: E402
from mcba import synth  # noqa: E402
import oracle_py  # noqa: E402

CAMERAS = [0, 3, 1365, 2731, 4095]      # 3 and 2731 are Kannala cameras (kannala_every = 4)


def main():
    doc = dict(generator="tests/golden/make_c5_golden.py", n_cameras=4096, n_frames=300, kannala_every=4, cameras={})
    for c in CAMERAS:
        # kannala_every tests c % 4 == 3 on the absolute camera index, and the RNG is keyed by it
        prob, init, gt = synth.batched_intrinsics_problem(n_cameras=4096, n_frames=300, kannala_every=4, cam_lo=c, cam_hi=c + 1)
        sub = mcba.Problem(1, prob.u, prob.v, prob.pt_idx, prob.bobs_ptr, prob.bobs_cam, prob.bobs_pose, prob.bobs_board,
                           prob.pts_xyz, prob.cam_model, prob.n_pose, 0)
        p = init.copy()
        rc, summ, tr = oracle_py.solve(sub, p)
        assert rc == 0
        doc["cameras"][str(c)] = dict(model=int(prob.cam_model[0]), observations=int(prob.n_obs), termination=int(summ.termination),
                                      num_successful_steps=int(summ.num_successful_steps),
                                      num_unsuccessful_steps=int(summ.num_unsuccessful_steps),
                                      num_iterations=int(summ.num_iterations), final_cost=summ.final_cost, rms_px=summ.rms_px,
                                      intrinsics=p.intrinsics[0].tolist())
        print(c, prob.cam_model[0], summ.num_iterations, summ.rms_px)
    with open(os.path.join(HERE, "c5_sample.json"), "w") as f:
        json.dump(doc, f, indent=0)


if __name__ == "__main__":
    main()
