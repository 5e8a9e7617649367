#!/usr/bin/env python
"""BASELINE config 5: batched independent per-camera intrinsic refinement (Camera::refineIntrinsicCalibration for
every camera at once). Prints one JSON line: whole-batch solve time on one GPU and the CPU oracle solving a sample
of the same cameras one after the other (the reference's loop, McCalib.cpp:713-716)."""
import argparse
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "mc-calib_b200", "python"))
sys.path.insert(0, os.path.join(REPO, "oracle"))
sys.path.insert(0, os.path.join(REPO, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cameras", type=int, default=4096)
    ap.add_argument("--frames", type=int, default=300)
    ap.add_argument("--cpu-cameras", type=int, default=16)
    args = ap.parse_args()
    import mcba
    from mcba import synth
    import oracle_py
    t0 = time.time()
    prob, init, gt = synth.batched_intrinsics_problem(n_cameras=args.cameras, n_frames=args.frames, kannala_every=4)
    t_gen = time.time() - t0
    h = mcba.Handle(prob)
    p = init.copy()
    h.solve_batched(p)            # warm-up (module load, allocator)
    p = init.copy()
    h.reset_kernel_stats()
    t0 = time.perf_counter()
    summ = h.solve_batched(p)
    t_solve = time.perf_counter() - t0
    stats = h.kernel_stats()
    h.close()
    its = np.array([s.num_iterations for s in summ])
    ok = int(sum(s.termination == 0 for s in summ))
    # CPU: the same cameras, one problem after the other, single thread (reference configuration)
    from test_gpu_parity import _camera_subproblem
    t_cpu, same = 0.0, 0
    for c in range(min(args.cpu_cameras, args.cameras)):
        sub, pc, _ = _camera_subproblem(prob, init, c)
        t1 = time.perf_counter()
        rc, so, _ = oracle_py.solve(sub, pc)
        t_cpu += time.perf_counter() - t1
        same += int((so.num_successful_steps, so.num_unsuccessful_steps, so.termination) ==
                    (summ[c].num_successful_steps, summ[c].num_unsuccessful_steps, summ[c].termination)
                    and np.abs(pc.intrinsics[0] - p.intrinsics[c]).max() <= 1e-6 * np.abs(pc.intrinsics[0]).max())
    n_cpu = min(args.cpu_cameras, args.cameras)
    print(json.dumps({
        "workload": f"config 5: {args.cameras} independent cameras x {args.frames} frames x 16 corners, stage 1, batched",
        "observations": int(prob.n_obs), "gpu_solve_s": t_solve, "problems_per_s": args.cameras / t_solve,
        "iterations_min_mean_max": [int(its.min()), float(its.mean()), int(its.max())], "converged": ok,
        "obs_iterations_per_s_M": float(prob.n_obs / args.cameras * its.sum() / t_solve / 1e6),
        "cpu_oracle_s_per_problem_1thread": t_cpu / n_cpu, "cpu_sample_problems": n_cpu, "cpu_sample_matching_gpu": same,
        "speedup_vs_1thread_cpu_loop": (t_cpu / n_cpu * args.cameras) / t_solve,
        "kernel_ms": {k: v[1] for k, v in stats.items() if v[0]}, "gen_s": t_gen}))


if __name__ == "__main__":
    main()
