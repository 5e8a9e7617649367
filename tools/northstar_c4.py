#!/usr/bin/env python
"""North-star run (BASELINE.json): a FULL Levenberg-Marquardt bundle adjustment (default Ceres options, to
convergence) over BASELINE config 4 - 64-camera dome, 10 boards, 20,000 frames, ~50 M observations - stage 4 then
stage 5 like apps/calibrate/src/calibrate.cpp:60-64.

  torchrun --nproc-per-node 8 tools/northstar_c4.py --impl mcba --out gpurun_out/northstar_gpu.json
  python tools/northstar_c4.py --impl oracle --threads 8 --out profiles/northstar_oracle.json     (CPU, test oracle)

Both write the per-iteration trace, termination, final cost / RMS and the refined camera, board and intrinsics blocks;
tools/northstar_compare.py puts them side by side."""
import argparse
import ctypes
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "mc-calib_b200", "python"))
sys.path.insert(0, os.path.join(REPO, "oracle"))


def rows(tr):
    return [dict(iteration=r.iteration, valid=r.step_is_valid, successful=r.step_is_successful, cost=r.cost,
                 cost_change=r.cost_change, gradient_max_norm=r.gradient_max_norm, step_norm=r.step_norm,
                 relative_decrease=r.relative_decrease, radius=r.trust_region_radius) for r in tr]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", default="mcba", choices=["mcba", "oracle"])
    ap.add_argument("--frames", type=int, default=20000)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--out", required=True)
    ap.add_argument("--warm", action="store_true", help="solve every stage once untimed first (kernel modules, NCCL channels)")
    args = ap.parse_args()
    import mcba
    from mcba import synth
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    s = synth.make_scene("c4", n_frames=args.frames)
    per = args.frames // world
    t0 = time.time()
    o = synth.generate_observations(s, rank * per, (rank + 1) * per if rank < world - 1 else args.frames)
    t_gen = time.time() - t0
    prob, init, gt = synth.make_problem(s, o, 4)
    p = init.copy()
    p.intrinsics[:] = s.intr_init
    doc = dict(impl=args.impl, frames=args.frames, world=world, gen_s=t_gen, warm=bool(args.warm), stages=[])
    if args.impl == "mcba":
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        if world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        h = mcba.Handle(prob, device=local)
        if world > 1:
            uid = torch.zeros(128, dtype=torch.uint8)
            if rank == 0:
                buf = ctypes.create_string_buffer(128)
                assert h.lib.mcba_nccl_unique_id(buf) == 0
                uid = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
            uid = uid.cuda()
            dist.broadcast(uid, 0)
            h.comm_init(rank, world, uid.cpu().numpy().tobytes())
        n_obs = prob.n_obs
        if world > 1:
            t = torch.tensor([float(n_obs)], dtype=torch.float64, device="cuda")
            dist.all_reduce(t)
            n_obs = int(t.item())
        for stage in (4, 5):
            if stage == 5:
                h.set_stage(5)
            if args.warm:
                h.solve(p.copy())
            h.reset_kernel_stats()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            summ, tr = h.solve(p)
            torch.cuda.synchronize()
            wall = time.perf_counter() - t0
            doc["stages"].append(dict(stage=stage, wall_s=wall, summary=summ.as_dict(), trace=rows(tr),
                                      kernel_ms={k: v[1] for k, v in h.kernel_stats().items() if v[0]}))
        h.close()
        doc["observations"] = n_obs
    else:
        import oracle_py
        threads = args.threads or oracle_py.max_threads()
        doc["threads"] = threads
        doc["observations"] = int(prob.n_obs)
        for stage in (4, 5):
            t0 = time.perf_counter()
            rc, summ, tr = oracle_py.solve(prob.with_stage(stage), p, n_threads=threads)
            wall = time.perf_counter() - t0
            assert rc == 0
            doc["stages"].append(dict(stage=stage, wall_s=wall, summary=summ.as_dict(), trace=rows(tr)))
            print(f"oracle stage {stage}: {wall:.1f} s, {summ.num_iterations} iterations, cost {summ.final_cost:.6f}", flush=True)
    doc["cam_pose"] = p.cam_pose.tolist()
    doc["board_pose"] = p.board_pose.tolist()
    doc["intrinsics"] = p.intrinsics.tolist()
    doc["pose_first_rank_head"] = p.pose[:8].tolist()
    if rank == 0:
        with open(args.out, "w") as f:
            json.dump(doc, f)
        print("wrote", args.out, [(st["stage"], st["wall_s"], st["summary"]["num_iterations"]) for st in doc["stages"]])
    if args.impl == "mcba" and world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
