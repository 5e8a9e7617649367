#!/usr/bin/env python
"""Timeline of one k_chol_tiles launch from its %globaltimer stamps (MCBA_CHOL_PROF=1)."""
import json, os, sys
import numpy as np
os.environ["MCBA_CHOL_PROF"] = "1"
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "mc-calib_b200", "python")); sys.path.insert(0, os.path.join(REPO, "tests"))
import mcba
from mcba import synth
from test_gpu_cholesky import spd
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1020
s = synth.make_scene("c1", n_frames=4); o = synth.generate_observations(s)
prob, init, gt = synth.make_problem(s, o, 4)
h = mcba.Handle(prob)
A, b = spd(n, n)
h.debug_cholesky(A, b, variant=1, iters=3)
z, L, ms, fail = h.debug_cholesky(A, b, variant=1, iters=1)
t = h.debug_fetch("chol_prof").reshape(-1, 16)
NT = (n + 63) // 64
tid = lambda i, j: j * (NT + 1) - j * (j - 1) // 2 + (i - j)
t0 = t[t[:, 0] > 0, 0].min()
rel = lambda x: (x - t0) / 1e3 if x > 0 else float("nan")
print(f"n {n} kernel {ms*1e3:.1f} us; per column j: diag [updates_done potrf_done published] | subdiag (j+1,j) [upd_done Ljj_seen loaded trsm_done published] | z_j [w_ready z_published]")
for j in range(NT):
    d = t[tid(j, j)]
    line = f"{j:2d} diag {rel(d[1]):7.1f} {rel(d[2]):7.1f} {rel(d[3]):7.1f}"
    if j + 1 < NT:
        e = t[tid(j + 1, j)]
        line += f" | sub {rel(e[1]):7.1f} {rel(e[4]):7.1f} {rel(e[5]):7.1f} {rel(e[2]):7.1f} {rel(e[3]):7.1f} | p {rel(e[6]):7.1f} {rel(e[7]):7.1f}"
    line += f" | z {rel(d[6]):7.1f} {rel(d[7]):7.1f}"
    print(line)
r = t[tid(NT, 0)]
print("rhs tile col 0:", [round(rel(x), 1) for x in r[:6]])
d = t[tid(3, 3)]
print("diag tile 3 potrf cycles: b1 %d  wait1 %d  b2 %d  wait2 %d" % (d[4] // 100000, d[4] % 100000, d[5] // 100000, d[5] % 100000))
print("  b1 split: diag-update %d  loads %d  pivots %d  rsqrt+stores %d" % (d[8], d[9], d[10], d[11]))
print("first stamp spread (us):", (t[t[:, 0] > 0, 0].max() - t0) / 1e3)
h.close()
