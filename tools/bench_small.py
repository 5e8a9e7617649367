#!/usr/bin/env python
"""ms per LM iteration on the small configurations (launch / synchronisation bound): config 1 (2 cameras, 200 frames,
5 744 observations) and config 2 (4 cameras, 1 000 frames, 141 k observations), stage 5, with the per-iteration launch
sequences replayed as CUDA graphs (default) and launched kernel by kernel (MCBA_GRAPH=0)."""
import json, os, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "mc-calib_b200", "python"))
import mcba
from mcba import synth
sys.path.insert(0, REPO)
from bench import bench_options
out = {}
for cfg in ("c1", "c2"):
    prob, init, gt, s, o = synth.config_problem(cfg, 5)
    row = {"observations": int(prob.n_obs)}
    for mode in ("graph", "direct"):
        os.environ["MCBA_GRAPH"] = "1" if mode == "graph" else "0"
        h = mcba.Handle(prob)
        p = init.copy(); p.intrinsics[:] = s.intr_init
        summ, tr = h.solve(p.copy())                       # real solve: trace for the parity of the two modes
        opt = bench_options(mcba)
        best = 1e9
        for rep in range(3):
            q = p.copy()
            h.solve_begin(q, opt)
            for _ in range(5): h.solve_step()
            h.timer_start(); t0 = time.perf_counter()
            for _ in range(50): assert h.solve_step() == 1
            ms_dev = h.timer_stop(); ms_wall = (time.perf_counter() - t0) * 1e3
            h.solve_end(q)
            best = min(best, ms_wall / 50)
        row[mode] = {"ms_per_iteration": best, "ms_per_iteration_device": ms_dev / 50, "solve_iterations": summ.num_iterations,
                     "rms_px": summ.rms_px, "sequence": "".join("A" if r.step_is_successful else "r" for r in tr)}
        h.close()
    out[cfg] = row
print(json.dumps(out))
