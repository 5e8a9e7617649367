#!/usr/bin/env python
"""Full solves (default Ceres options, stage 4 then stage 5) of BASELINE configs 1-3 on one GPU or with the CPU
oracle; writes traces and refined f-blocks for an offline comparison (tools/northstar_compare.py works on the output)."""
import argparse
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "mc-calib_b200", "python"))
sys.path.insert(0, os.path.join(REPO, "oracle"))
sys.path.insert(0, os.path.join(REPO, "tools"))
from northstar_c4 import rows  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", default="mcba", choices=["mcba", "oracle"])
    ap.add_argument("--config", required=True)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--out", required=True)
    args = ap.parse_args()
    import mcba
    from mcba import synth
    t0 = time.time()
    prob, init, gt, s, o = synth.config_problem(args.config, 4)
    p = init.copy()
    p.intrinsics[:] = s.intr_init
    doc = dict(impl=args.impl, config=args.config, world=1, observations=int(prob.n_obs), gen_s=time.time() - t0, stages=[])
    if args.impl == "mcba":
        t0 = time.perf_counter()
        h = mcba.Handle(prob)
        doc["create_s"] = time.perf_counter() - t0
        for stage in (4, 5):
            if stage == 5:
                h.set_stage(5)
            h.solve(p.copy())            # warm-up (module load), discarded
            h.reset_kernel_stats()
            t0 = time.perf_counter()
            summ, tr = h.solve(p)
            doc["stages"].append(dict(stage=stage, wall_s=time.perf_counter() - t0, summary=summ.as_dict(), trace=rows(tr),
                                      kernel_ms={k: v[1] for k, v in h.kernel_stats().items() if v[0]}))
        h.close()
    else:
        import oracle_py
        threads = args.threads or oracle_py.max_threads()
        doc["threads"] = threads
        for stage in (4, 5):
            t0 = time.perf_counter()
            rc, summ, tr = oracle_py.solve(prob.with_stage(stage), p, n_threads=threads)
            doc["stages"].append(dict(stage=stage, wall_s=time.perf_counter() - t0, summary=summ.as_dict(), trace=rows(tr)))
    doc["cam_pose"], doc["board_pose"], doc["intrinsics"] = p.cam_pose.tolist(), p.board_pose.tolist(), p.intrinsics.tolist()
    json.dump(doc, open(args.out, "w"))
    print(args.config, args.impl, [(st["stage"], round(st["wall_s"], 4), st["summary"]["num_iterations"]) for st in doc["stages"]])


if __name__ == "__main__":
    main()
