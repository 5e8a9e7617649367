#!/usr/bin/env python
"""Fused observation kernel (K1) alone: time per launch and record parity of the variants (mcba_debug_set_k1) on the
bench shard (config 4 shape, stage 5 and stage 4) and on the small configurations of every stage.

  python tools/bench_k1.py [--frames 2500] [--variants 1,0] [--steps 10]
Prints one JSON object: per stage and variant the K1 time per launch (CUDA events around the kernel family, direct
launches), the whole-iteration time, and the largest difference of the per-board-observation records against variant 1
relative to the largest entry of the record row."""
import argparse, json, os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "mc-calib_b200", "python"))
sys.path.insert(0, REPO)
os.environ["MCBA_GRAPH"] = "0"      # per-kernel events
import mcba
from mcba import synth
from bench import bench_options

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=2500)
ap.add_argument("--variants", default="1,0,3")
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--stages", default="5,4")
ap.add_argument("--small", action="store_true", help="also compare the records on configs c1 / c2, all five stages")
args = ap.parse_args()
variants = [int(v) for v in args.variants.split(",")]
lib = mcba.load_lib()
out = {"frames": args.frames, "runs": []}


WF = {1: 9, 2: 6, 3: 6, 4: 12, 5: 21}


def records(h, prob, p, opt):
    """e-block part of every record (the f part is not written when the kernel accumulates pair records itself) and the
    pair records after the assembly."""
    h.solve_begin(p.copy(), opt)
    wf = WF[prob.stage]
    ci_fe = wf * (wf + 1) // 2 + wf
    rec = h.debug_fetch("brec").reshape(prob_nb(prob), -1)[:, ci_fe:].copy()
    return rec, h.debug_fetch("pair_rec").copy()


def run(prob, init, label, steps):
    opt = bench_options(mcba)
    ref = None
    for v in variants:
        lib.mcba_debug_set_k1(v)
        h = mcba.Handle(prob)
        p = init.copy()
        rec, pair = records(h, prob, p, opt)
        row = {"case": label, "variant": v, "observations": int(prob.n_obs)}
        if steps:
            for _ in range(3):
                assert h.solve_step() == 1
            h.reset_kernel_stats()
            h.timer_start()
            for _ in range(steps):
                assert h.solve_step() == 1
            ms = h.timer_stop()
            st = h.kernel_stats()
            row["ms_per_iteration"] = ms / steps
            row["k1_ms_per_launch"] = st["resid_jac_fused"][1] / max(1, st["resid_jac_fused"][0])
            row["kernel_ms_per_iteration"] = {k: x[1] / steps for k, x in st.items() if x[0]}
        summ = h.solve_end(p)
        h.close()
        if ref is None:
            ref = (rec, pair)
        else:
            a = ref[0]; b = rec
            scale = np.abs(a).max(axis=1, keepdims=True) + 1e-300
            row["max_rel_diff_vs_first"] = float((np.abs(a - b) / scale).max())
            row["pair_rec_max_rel_diff_vs_first"] = float(np.abs(pair - ref[1]).max() / (np.abs(ref[1]).max() + 1e-300))
            row["n_nan"] = int(np.isnan(b).sum() + np.isnan(pair).sum())
        out["runs"].append(row)
        print(json.dumps(row), file=sys.stderr, flush=True)
    lib.mcba_debug_set_k1(-1)


def prob_nb(prob):
    return int(prob.n_bobs)


if args.small:
    for cfg in ("c1", "c2"):
        for stage in (1, 2, 3, 4, 5):
            prob, init, gt, s, o = synth.config_problem(cfg, stage)
            run(prob, init, f"{cfg} stage {stage}", 0)
s = synth.make_scene("c4", n_frames=args.frames)
o = synth.generate_observations(s, 0, args.frames, workers=8)
for stage in [int(x) for x in args.stages.split(",")]:
    prob, init, gt = synth.make_problem(s, o, stage)
    run(prob, init, f"c4 {args.frames} frames stage {stage}", args.steps)
print(json.dumps(out))
