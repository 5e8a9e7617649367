#!/usr/bin/env python
"""Kernel time of the two reduced-system Cholesky kernels at a few sizes (mcba_debug_cholesky)."""
import json, os, sys
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "mc-calib_b200", "python")); sys.path.insert(0, os.path.join(REPO, "tests"))
import mcba
from mcba import synth
from test_gpu_cholesky import spd
s = synth.make_scene("c1", n_frames=4); o = synth.generate_observations(s)
prob, init, gt = synth.make_problem(s, o, 4)
h = mcba.Handle(prob)
out = {}
for n in (36, 78, 270, 512, 1020, 1400):
    A, b = spd(n, n)
    row = {}
    for v in (0, 1):
        handle = h.debug_cholesky(A, b, variant=v, iters=3)
        z, L, ms, fail = h.debug_cholesky(A, b, variant=v, iters=20)
        row["coop_ms" if v == 0 else "tiles_ms"] = ms
        row["err%d" % v] = float(np.abs(z - np.linalg.solve(A, b)).max())
    out[n] = row
print(json.dumps(out))
h.close()
