#!/usr/bin/env python
"""Pose initialisation throughput: every board observation of a config-4 shard (64-camera dome,
2500 frames, ~129 k board observations of 48 corners) through mcba_pnp_ransac in one call, beside the OpenCV loop the
reference runs (oracle/oracle_pnp.py = the reference's ransacP3P on cv2, single thread as the reference) on a bounded
sample. Prints one JSON line.

  python tools/bench_pnp.py [--frames 2500] [--threshold 10] [--reps 5] [--cpu-runs 400]
"""
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
    ap.add_argument("--frames", type=int, default=2500)
    ap.add_argument("--threshold", type=float, default=10.0)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--cpu-runs", type=int, default=400)
    args = ap.parse_args()
    import mcba
    from mcba import synth, pnp
    s = synth.make_scene("c4", n_frames=args.frames)
    o = synth.generate_observations(s)
    prob, init, gt = synth.make_problem(s, o, 1)
    P = pnp.PnpProblem(prob, s.intr_gt)
    try:      # pinned host arrays, like bench.py's e2e leg
        import torch
        P._pinned = []
        for name in ("u", "v", "pt_idx", "bobs_ptr", "bobs_cam"):
            t = torch.from_numpy(getattr(P, name)).pin_memory()
            P._pinned.append(t)
            setattr(P, name, t.numpy())
    except Exception:
        pass
    opt = pnp.default_options(threshold_px=args.threshold, max_iterations=1000, seed=1)
    pnp.ransac(P, opt)                       # warm-up: context, module load
    k_ms, t_ms = [], []
    for _ in range(args.reps):
        r = pnp.ransac(P, opt)
        k_ms.append(r.kernel_ms)
        t_ms.append(r.total_ms)
    k, t = float(np.median(k_ms)), float(np.median(t_ms))
    import pnp_util
    err = [pnp_util.pose_diff(r.pose[b], gt.pose[b]) for b in range(0, P.n_bobs, max(1, P.n_bobs // 2000))]
    line = {
        "metric": "board-observation poses per second (P3P RANSAC + pose refinement)", "unit": "bobs/s",
        "value": P.n_bobs / (k * 1e-3), "e2e": {"value": P.n_bobs / (t * 1e-3), "unit": "bobs/s",
                                                "includes": "H2D of the observations (pinned host arrays), 4 kernels, D2H of poses/masks, device allocation"},
        "kernel_ms": k, "total_ms": t, "n_bobs": int(P.n_bobs), "n_obs": int(P.n_obs),
        "draws_total": int(r.n_hypotheses.sum()), "draws_evaluated_lanes": int(32 * np.ceil(r.n_hypotheses / 32).sum()),
        "without_pose": int((r.best_hypothesis < 0).sum()), "mean_inlier_fraction": float(r.n_inliers.sum() / P.n_obs),
        "max_pose_error_vs_gt_sample": float(max(err)),
        "config": {"workload": f"config 4 shard: 64-camera dome, 10 boards x 48 corners, {args.frames} frames, threshold {args.threshold} px, 1000 max draws"},
    }
    if args.cpu_runs > 0:
        try:
            import oracle_pnp
            sub = pnp.PnpProblem.__new__(pnp.PnpProblem)
            sub.__dict__.update(P.__dict__)
            nb = min(args.cpu_runs, P.n_bobs)
            sub.bobs_ptr, sub.bobs_cam = P.bobs_ptr[:nb + 1].copy(), P.bobs_cam[:nb].copy()
            ne = int(sub.bobs_ptr[-1])
            sub.u, sub.v, sub.pt_idx = P.u[:ne], P.v[:ne], P.pt_idx[:ne]
            oracle_pnp.ransac_all(sub, opt)
            t0 = time.perf_counter()
            oracle_pnp.ransac_all(sub, opt)
            sec = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": nb / sec, "unit": "bobs/s", "cores": 1, "kind": "reference-library",
                                    "sample": f"first {nb} board observations, the reference's ransacP3P loop on OpenCV {__import__('cv2').__version__} through cv2 (python call overhead included)"}
        except ImportError:
            line["cpu_baseline"] = None
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
