"""Generates tests/golden/pnp_golden.npz: outputs of the OpenCV-backed checker (oracle/oracle_pnp.py, cv2 4.13:
cv::solvePnP P3P / ITERATIVE, cv::projectPoints, cv::fisheye::undistortPoints called with the reference's conventions)
on the seeded cases of tests/pnp_util.py. Inputs are regenerated from the seeds by the tests; only outputs are stored."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
for p in ("mc-calib_b200/python", "oracle", "tests"):
    sys.path.insert(0, os.path.join(REPO, p))
import oracle_pnp  # noqa: E402
import pnp_util  # noqa: E402

out = {}
for name in pnp_util.CASES:
    P, opt, _ = pnp_util.make_case(name)
    ref = oracle_pnp.ransac_all(P, opt)
    for k, v in ref.items():
        out[f"{name}/{k}"] = v
    print(name, P.n_bobs, "runs,", int(ref["n_hyp"].sum()), "draws,", int((ref["best_h"] < 0).sum()), "without a pose")
np.savez_compressed(pnp_util.GOLDEN, **out)
