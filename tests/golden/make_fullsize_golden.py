"""Full-size golden fixtures: the CPU oracle's stage 4 + stage 5 solves (reference order,
/root/reference/apps/calibrate/src/calibrate.cpp:60-64) of BASELINE.json's full-size configs.

  full_c3.json   config 3: 16-camera ring, 5 boards, 5 000 frames, 10 188 613 observations
  full_c4.json   config 4: 64-camera dome, 10 boards, 20 000 frames, 49 256 002 observations

Each holds, per stage, the summary (termination, step counts, costs, RMS), the LM trace (iteration, valid,
successful, cost, radius, gradient max norm) and the refined f-blocks (camera poses, board poses, intrinsics) —
what the -m gpu tests and bench.py's `parity` object compare the CUDA path with at full size, where the oracle
itself takes minutes (c3: ~2.5 min, c4: ~20 min and 25 GB on 8 threads).

    python tests/golden/make_fullsize_golden.py --config c3 [--threads 8]

The oracle is test infrastructure (oracle/oracle.h); the reference itself cannot be built in this image.
"""
import argparse
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(REPO, "mc-calib_b200", "python"))
sys.path.insert(0, os.path.join(REPO, "oracle"))


def rows(tr):
    return [dict(iteration=r.iteration, valid=r.step_is_valid, successful=r.step_is_successful, cost=r.cost,
                 radius=r.trust_region_radius, gradient_max_norm=r.gradient_max_norm) for r in tr]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", required=True, choices=["c1", "c2", "c3", "c4"])
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--frames", type=int, default=None, help="default: the config's full frame count")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import mcba
    from mcba import synth
    import oracle_py
    threads = args.threads or oracle_py.max_threads()
    t0 = time.time()
    prob, init, gt, s, o = synth.config_problem(args.config, 4, n_frames=args.frames)
    p = init.copy()
    p.intrinsics[:] = s.intr_init          # stage 4 runs with the initial intrinsics, stage 5 refines them
    doc = dict(config=args.config, frames=int(s.n_frames), observations=int(prob.n_obs), threads=threads,
               generator="tests/golden/make_fullsize_golden.py", stages=[])
    print(f"{args.config}: {prob.n_obs} observations generated in {time.time() - t0:.1f} s", flush=True)
    for stage in (4, 5):
        t0 = time.perf_counter()
        rc, summ, tr = oracle_py.solve(prob.with_stage(stage), p, mcba.default_options(), n_threads=threads)
        assert rc == 0
        doc["stages"].append(dict(stage=stage, wall_s=time.perf_counter() - t0, termination=int(summ.termination),
                                  num_successful_steps=int(summ.num_successful_steps),
                                  num_unsuccessful_steps=int(summ.num_unsuccessful_steps),
                                  num_iterations=int(summ.num_iterations), initial_cost=summ.initial_cost,
                                  final_cost=summ.final_cost, rms_px=summ.rms_px, trace=rows(tr),
                                  cam_pose=p.cam_pose.tolist(), board_pose=p.board_pose.tolist(),
                                  intrinsics=p.intrinsics.tolist()))
        print(f"  stage {stage}: {summ.num_iterations} iterations, rms {summ.rms_px:.9f}, {time.perf_counter() - t0:.1f} s", flush=True)
    out = args.out or os.path.join(HERE, f"full_{args.config}.json")
    with open(out, "w") as f:
        json.dump(doc, f, indent=0)
    print("wrote", out)


if __name__ == "__main__":
    main()
