#!/usr/bin/env python
"""Side-by-side of the 8-GPU north-star run and the CPU oracle on the same 50 M-observation problem."""
import json
import sys

import numpy as np

g, o = json.load(open(sys.argv[1])), json.load(open(sys.argv[2]))
out = dict(observations=[g["observations"], o["observations"]], gpus=g["world"], cpu_threads=o.get("threads"), stages=[])
for sg, so in zip(g["stages"], o["stages"]):
    tg = [(r["iteration"], r["valid"], r["successful"]) for r in sg["trace"]]
    to = [(r["iteration"], r["valid"], r["successful"]) for r in so["trace"]]
    out["stages"].append(dict(
        stage=sg["stage"], same_accept_reject_sequence=tg == to,
        iterations=[sg["summary"]["num_iterations"], so["summary"]["num_iterations"]],
        termination=[sg["summary"]["termination"], so["summary"]["termination"]],
        final_cost=[sg["summary"]["final_cost"], so["summary"]["final_cost"]],
        rms_px=[sg["summary"]["rms_px"], so["summary"]["rms_px"]],
        rms_abs_diff=abs(sg["summary"]["rms_px"] - so["summary"]["rms_px"]),
        max_rel_cost_diff_per_iteration=max(abs(a["cost"] - b["cost"]) / b["cost"] for a, b in zip(sg["trace"], so["trace"])),
        wall_s=[sg["wall_s"], so["wall_s"]], speedup=so["wall_s"] / sg["wall_s"]))
for k in ("cam_pose", "board_pose", "intrinsics"):
    a, b = np.array(g[k]), np.array(o[k])
    out[k + "_max_abs_diff"] = float(np.abs(a - b).max())
    out[k + "_max_rel_diff"] = float(np.abs(a - b).max() / np.abs(b).max())
print(json.dumps(out, indent=1))
