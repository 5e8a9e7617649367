"""Shared helpers for the parity tests (numpy only)."""
import numpy as np

import mcba

STAGE_BLOCKS = {  # reference argument order of each stage's residual block: (name, width)
    1: [("pose", 6), ("intr", 9)],
    2: [("pose", 6), ("board", 6)],
    3: [("cam", 6), ("pose", 6)],
    4: [("cam", 6), ("pose", 6), ("board", 6)],
    5: [("cam", 6), ("pose", 6), ("board", 6), ("intr", 9)],
}


def layout(problem):
    """State-vector layout used by libmcba and the oracle: e-blocks, then per camera [pose][intr], then boards."""
    names = [n for n, _ in STAGE_BLOCKS[problem.stage]]
    wc = (6 if "cam" in names else 0) + (9 if "intr" in names else 0)
    wb = 6 if "board" in names else 0
    n_e = problem.n_pose * 6
    n_f = problem.n_cam * wc + problem.n_board * wb
    return dict(names=names, wc=wc, wb=wb, n_e=n_e, n_f=n_f)


def dense_jacobian(problem, jac):
    """Scatter per-observation blocks [n_obs,2,K] (reference order) into a dense (2 n_obs) x n_x matrix."""
    L = layout(problem)
    n_x = L["n_e"] + L["n_f"]
    J = np.zeros((2 * problem.n_obs, n_x))
    counts = np.diff(problem.bobs_ptr)
    cam = np.repeat(problem.bobs_cam, counts)
    pose = np.repeat(problem.bobs_pose, counts)
    board = np.repeat(problem.bobs_board, counts)
    rows = np.arange(problem.n_obs)
    off = 0
    for name, w in STAGE_BLOCKS[problem.stage]:
        if name == "pose":
            base = pose * 6
        elif name == "cam":
            base = L["n_e"] + cam * L["wc"]
        elif name == "intr":
            base = L["n_e"] + cam * L["wc"] + (6 if "cam" in L["names"] else 0)
        else:
            base = L["n_e"] + problem.n_cam * L["wc"] + board * L["wb"]
        for k in range(w):
            for r in range(2):
                J[2 * rows + r, base + k] = jac[:, r, off + k]
        off += w
    return J


def params_to_x(problem, p):
    L = layout(problem)
    x = [p.pose.ravel()]
    f = []
    for c in range(problem.n_cam):
        if "cam" in L["names"]:
            f.append(p.cam_pose[c])
        if "intr" in L["names"]:
            f.append(p.intrinsics[c])
    if "board" in L["names"]:
        f.append(p.board_pose.ravel())
    return np.concatenate(x + [np.ravel(a) for a in f])


def trace_tuple(rows):
    return [(r.iteration, r.step_is_valid, r.step_is_successful) for r in rows]


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return np.abs(a - b).max() / max(1e-300, np.abs(b).max())
