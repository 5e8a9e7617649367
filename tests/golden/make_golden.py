"""Generates the committed golden fixtures (run here, on the CPU container):

  functor_kat.npz   known-answer tests of the five reprojection functors: random (stage, model, flags,
                    parameter blocks, point, pixel) -> residual and Jacobian from an mpmath 50-digit
                    evaluator (values by the exact Rodrigues formula, Jacobian by central differences with
                    h = 1e-22). Independent of the oracle and of libmcba.
  trace_*.json      Levenberg-Marquardt traces (cost, radius, accept flags, termination) of the CPU oracle on
                    configs 1 and 2, stages 4 and 5: the regression pin for the LM driver on both sides.

The reference (MC-Calib + ceres 2.2.0) cannot be built in this image, so no fixture comes from the
reference binary itself; see oracle/oracle.h.
"""
import json
import os
import sys

import mpmath as mp
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(REPO, "mc-calib_b200", "python"))
sys.path.insert(0, os.path.join(REPO, "oracle"))
import mcba  # noqa: E402
from mcba import synth  # noqa: E402
import oracle_py  # noqa: E402

mp.mp.dps = 50
BLOCKS = {1: ["pose", "intr"], 2: ["pose", "board"], 3: ["cam", "pose"], 4: ["cam", "pose", "board"],
          5: ["cam", "pose", "board", "intr"]}


def rot(w, p):
    th2 = w[0] ** 2 + w[1] ** 2 + w[2] ** 2
    if th2 == 0:
        return list(p)
    th = mp.sqrt(th2)
    k = [x / th for x in w]
    c, s = mp.cos(th), mp.sin(th)
    kxp = [k[1] * p[2] - k[2] * p[1], k[2] * p[0] - k[0] * p[2], k[0] * p[1] - k[1] * p[0]]
    kdp = k[0] * p[0] + k[1] * p[1] + k[2] * p[2]
    return [p[i] * c + kxp[i] * s + k[i] * kdp * (1 - c) for i in range(3)]


def f_mp(stage, model, flags, blocks, xyz, uv):
    """blocks: dict name -> list of mpf"""
    names = BLOCKS[stage]
    p = [mp.mpf(x) for x in xyz]
    if "board" in names and not (flags & 2):
        b = blocks["board"]
        p = rot(b[:3], p)
        p = [p[i] + b[3 + i] for i in range(3)]
    o = blocks["pose"]
    p = rot(o[:3], p)
    p = [p[i] + o[3 + i] for i in range(3)]
    if "cam" in names and not (flags & 1):
        c = blocks["cam"]
        p = rot(c[:3], p)
        p = [p[i] + c[3 + i] for i in range(3)]
    a, b = p[0] / p[2], p[1] / p[2]
    it = blocks["intr"]
    r2 = a * a + b * b
    if model == 0:
        k1, k2, p1, p2, k3 = it[4], it[5], it[6], it[7], it[8]
        rc = 1 + k1 * r2 + k2 * r2 ** 2 + k3 * r2 ** 3
        xd = a * rc + 2 * p1 * a * b + p2 * (r2 + 2 * a * a)
        yd = b * rc + p1 * (r2 + 2 * b * b) + 2 * p2 * a * b
    else:
        r = mp.sqrt(r2)
        th = mp.atan(r)
        thd = th + it[4] * th ** 3 + it[5] * th ** 5 + it[6] * th ** 7 + it[7] * th ** 9
        cd = thd / r if r > mp.mpf("1e-8") else mp.mpf(1)
        xd, yd = a * cd, b * cd
    return [it[0] * xd + it[2] - mp.mpf(uv[0]), it[1] * yd + it[3] - mp.mpf(uv[1])]


def kat_case(rng):
    stage = int(rng.integers(1, 6)); model = int(rng.integers(0, 2)); flags = int(rng.integers(0, 4))
    cam = np.concatenate([rng.normal(size=3) * 0.4, rng.normal(size=3) * 0.3])
    pose = np.concatenate([rng.normal(size=3) * 0.7, rng.normal(size=3) * 0.2 + [0, 0, 2.5]])
    board = np.concatenate([rng.normal(size=3) * 0.4, rng.normal(size=3) * 0.2])
    intr = np.array([1400 + rng.normal() * 50, 1380 + rng.normal() * 50, 960 + rng.normal() * 10, 540 + rng.normal() * 10,
                     *(rng.normal(size=5) * [0.1, 0.05, 1e-3, 1e-3, 0.01])])
    if model == 1:
        intr[:2] = 600 + rng.normal(size=2) * 10
        intr[4:8] = rng.normal(size=4) * [0.02, 3e-3, 1e-3, 2e-4]
        intr[8] = 0
    xyz = np.array([rng.uniform(0, 0.8), rng.uniform(0, 0.6), 0.0 if rng.random() < 0.7 else rng.normal() * 0.1])
    uv = np.array([rng.uniform(0, 1920), rng.uniform(0, 1080)])
    return stage, model, flags, cam, pose, board, intr, xyz, uv


def make_kat(n=160, seed=20261019):
    rng = np.random.default_rng(seed)
    rows = []
    while len(rows) < n:
        stage, model, flags, cam, pose, board, intr, xyz, uv = kat_case(rng)
        blocks = {"cam": [mp.mpf(x) for x in cam], "pose": [mp.mpf(x) for x in pose],
                  "board": [mp.mpf(x) for x in board], "intr": [mp.mpf(x) for x in intr]}
        r = f_mp(stage, model, flags, blocks, xyz, uv)
        proj = [r[0] + mp.mpf(uv[0]), r[1] + mp.mpf(uv[1])]
        ab = max(abs((proj[0] - blocks["intr"][2]) / blocks["intr"][0]), abs((proj[1] - blocks["intr"][3]) / blocks["intr"][1]))
        if ab > 2.0:   # keep well-conditioned geometry
            continue
        K = mcba.STAGE_K[stage]
        J = np.zeros((2, 27))
        h = mp.mpf("1e-22")
        col = 0
        for name in BLOCKS[stage]:
            w = 9 if name == "intr" else 6
            for k in range(w):
                bp = {n_: list(v) for n_, v in blocks.items()}
                bm = {n_: list(v) for n_, v in blocks.items()}
                bp[name][k] += h
                bm[name][k] -= h
                fp, fm = f_mp(stage, model, flags, bp, xyz, uv), f_mp(stage, model, flags, bm, xyz, uv)
                J[0, col], J[1, col] = float((fp[0] - fm[0]) / (2 * h)), float((fp[1] - fm[1]) / (2 * h))
                col += 1
        assert col == K
        rows.append(dict(stage=stage, model=model, flags=flags, cam=cam, pose=pose, board=board, intr=intr, xyz=xyz,
                         uv=uv, r=np.array([float(r[0]), float(r[1])]), J=J))
    out = {k: np.array([row[k] for row in rows]) for k in rows[0]}
    np.savez(os.path.join(HERE, "functor_kat.npz"), **out)
    print("functor_kat.npz:", n, "cases")


def make_traces():
    for cfg in ("c1", "c2"):
        prob4, init, gt, s, o = synth.config_problem(cfg, 4)
        p = init.copy()
        for stage in (4, 5):
            prob = prob4.with_stage(stage)
            if stage == 5:
                p.intrinsics[:] = s.intr_init
            rc, summ, tr = oracle_py.solve(prob, p, mcba.default_options())
            assert rc == 0
            doc = dict(config=cfg, stage=stage, n_obs=int(prob.n_obs), termination=int(summ.termination),
                       num_successful_steps=int(summ.num_successful_steps),
                       num_unsuccessful_steps=int(summ.num_unsuccessful_steps),
                       initial_cost=summ.initial_cost, final_cost=summ.final_cost, rms_px=summ.rms_px,
                       rows=[dict(iteration=r.iteration, valid=r.step_is_valid, successful=r.step_is_successful, cost=r.cost,
                                  radius=r.trust_region_radius, gradient_max_norm=r.gradient_max_norm) for r in tr],
                       intrinsics=p.intrinsics.tolist(), cam_pose=p.cam_pose.tolist(), board_pose=p.board_pose.tolist())
            with open(os.path.join(HERE, f"trace_{cfg}_stage{stage}.json"), "w") as f:
                json.dump(doc, f, indent=1)
            print(f"trace_{cfg}_stage{stage}.json", summ.num_successful_steps, summ.num_unsuccessful_steps, summ.final_cost)


if __name__ == "__main__":
    make_kat()
    make_traces()
