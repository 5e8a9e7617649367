#!/usr/bin/env python
"""Small pass over every kernel family for compute-sanitizer (memcheck / racecheck): all five stages, the n_f = 1020
reduced system (multi-panel Cholesky, split-K SYRK), the materialised path, mcba_eval, reprojection errors, batched."""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "mc-calib_b200", "python"))
import mcba
from mcba import synth

opt = mcba.default_options(max_num_iterations=3)
s = synth.make_scene("c2", n_frames=12)
o = synth.generate_observations(s)
for stage in (1, 2, 3, 4, 5):
    prob, init, gt = synth.make_problem(s, o, stage)
    h = mcba.Handle(prob)
    summ, tr = h.solve(init.copy(), opt)
    h.eval(init)
    h.reprojection_errors(init)
    if stage == 5:
        h.solve(init.copy(), mcba.default_options(max_num_iterations=2, materialize_jacobian=1))
    h.close()
    print("stage", stage, "ok", summ.num_iterations, flush=True)
s = synth.make_scene("c4", n_frames=6)
o = synth.generate_observations(s)
prob, init, gt = synth.make_problem(s, o, 5)
h = mcba.Handle(prob)
summ, tr = h.solve(init.copy(), opt)
h.close()
print("c4 n_f=1020 ok", summ.num_iterations, flush=True)
s = synth.make_scene("c3", n_frames=6)
o = synth.generate_observations(s)
prob, init, gt = synth.make_problem(s, o, 4)
h = mcba.Handle(prob)
summ, tr = h.solve(init.copy(), opt)
h.set_stage(5)
summ, tr = h.solve(init.copy(), opt)
h.close()
print("c3 ok", summ.num_iterations, flush=True)
prob, init, gt = synth.batched_intrinsics_problem(n_cameras=5, n_frames=12, kannala_every=2)
h = mcba.Handle(prob)
summ = h.solve_batched(init.copy(), opt)
h.close()
print("batched ok", [x.num_iterations for x in summ], flush=True)
