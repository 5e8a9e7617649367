"""Seeded synthetic rigs/boards/observations for BASELINE.json's configs.

Produces the same structures the reference's refine* stages walk (frames -> camera-group
observation -> object observations -> corners; /root/reference/McCalib/src/CameraGroup.cpp:371-461)
already flattened into the mcba_problem layout. Observations are generated per frame chunk from an
RNG keyed by (seed, chunk), so any contiguous frame shard reproduces exactly the frames of the full
problem: rank r of a multi-GPU run generates only its own shard.
"""
import numpy as np
from scipy.spatial.transform import Rotation as Rot

from . import Problem, Params

IMG_W, IMG_H = 1920.0, 1080.0
BROWN_GT = np.array([1400.0, 1400.0, 960.0, 540.0, -0.12, 0.06, 5e-4, -3e-4, -0.01])
KANNALA_GT = np.array([600.0, 600.0, 960.0, 540.0, -0.02, 3e-3, -1e-3, 2e-4, 0.0])
CHUNK = 250  # frames per RNG chunk

CONFIGS = {
    # name: rig, n_cam, kannala cams, n_board, (nx, ny) corners, square, n_frames, keep prob
    "c1": dict(rig="stereo", n_cam=2, kannala=[], n_board=1, grid=(4, 4), square=0.192, n_frames=200, keep=1.0),
    "c2": dict(rig="square", n_cam=4, kannala=[2, 3], n_board=3, grid=(4, 4), square=0.192, n_frames=1000, keep=1.0),
    "c3": dict(rig="ring", n_cam=16, kannala=[], n_board=5, grid=(9, 7), square=0.1, n_frames=5000, keep=1.0, facing=0.03),
    "c4": dict(rig="dome", n_cam=64, kannala=[], n_board=10, grid=(8, 6), square=0.08, n_frames=20000, keep=0.26),
}


def project(intr, model, X):
    """numpy restatement of the camera models (independent of the oracle; used to synthesise pixels).
    Brown: OptimizationCeres.h:436-443; Kannala: :460-477. intr [...,9], X [...,3]."""
    a, b = X[..., 0] / X[..., 2], X[..., 1] / X[..., 2]
    r2 = a * a + b * b
    fx, fy, u0, v0 = intr[..., 0], intr[..., 1], intr[..., 2], intr[..., 3]
    k1, k2, p1, p2, k3 = intr[..., 4], intr[..., 5], intr[..., 6], intr[..., 7], intr[..., 8]
    rc = 1 + k1 * r2 + k2 * r2 ** 2 + k3 * r2 ** 3
    xb = a * rc + 2 * p1 * a * b + p2 * (r2 + 2 * a * a)
    yb = b * rc + p1 * (r2 + 2 * b * b) + 2 * p2 * a * b
    r = np.sqrt(r2)
    th = np.arctan(r)
    thd = th + intr[..., 4] * th ** 3 + intr[..., 5] * th ** 5 + intr[..., 6] * th ** 7 + intr[..., 7] * th ** 9
    cd = np.where(r > 1e-8, thd / np.maximum(r, 1e-300), 1.0)
    xd = np.where(model == 0, xb, a * cd)
    yd = np.where(model == 0, yb, b * cd)
    return np.stack([fx * xd + u0, fy * yd + v0], -1)


def compose(rv2, t2, rv1, t1):
    """(R2,t2) o (R1,t1): x -> R2 (R1 x + t1) + t2, as rotation vectors."""
    R2, R1 = Rot.from_rotvec(rv2), Rot.from_rotvec(rv1)
    return (R2 * R1).as_rotvec(), R2.apply(t1) + t2


def look_at(center, target, up=np.array([0.0, 0.0, 1.0])):
    """world->camera pose (rotvec, t) of a camera at `center` looking at `target` (OpenCV axes: z forward, y down)."""
    z = target - center
    z = z / np.linalg.norm(z)
    x = np.cross(z, up)
    if np.linalg.norm(x) < 1e-6:
        x = np.cross(z, np.array([0.0, 1.0, 0.0]))
    x = x / np.linalg.norm(x)
    y = np.cross(z, x)
    R = np.stack([x, y, z], 0)  # rows = camera axes in world
    return Rot.from_matrix(R).as_rotvec(), -R @ center


class Scene:
    pass


def make_scene(cfg, seed=None, n_frames=None):
    """Ground truth rig, boards and initial (perturbed) parameters; no observations yet."""
    if isinstance(cfg, str):
        name = cfg
        cfg = dict(CONFIGS[cfg])
        cfg["name"] = name
    s = Scene()
    s.cfg = cfg
    s.seed = int(seed if seed is not None else 1000 + sum(map(ord, cfg.get("name", "x"))))
    s.n_frames = int(n_frames if n_frames is not None else cfg["n_frames"])
    rng = np.random.default_rng([s.seed, 0xC0FFEE])
    n_cam, n_board = cfg["n_cam"], cfg["n_board"]
    nx, ny = cfg["grid"]
    sq = cfg["square"]
    s.n_cam, s.n_board, s.P = n_cam, n_board, nx * ny
    # board corners (x*sq, y*sq, 0) in float, like Board.cpp:66-72
    gx, gy = np.meshgrid(np.arange(nx), np.arange(ny))
    corners = np.stack([gx.ravel() * sq, gy.ravel() * sq, np.zeros(nx * ny)], -1).astype(np.float32)
    s.board_pts = np.tile(corners[None], (n_board, 1, 1))  # [n_board, P, 3]
    cx, cy = (nx - 1) * sq / 2, (ny - 1) * sq / 2
    s.cam_model = np.zeros(n_cam, np.int32)
    s.cam_model[cfg["kannala"]] = 1
    s.intr_gt = np.where(s.cam_model[:, None] == 0, BROWN_GT, KANNALA_GT).astype(np.float64)
    s.intr_gt = s.intr_gt * (1 + 0.01 * rng.standard_normal((n_cam, 9)) * np.array([1, 1, 0.3, 0.3, 1, 1, 1, 1, 1]))
    s.intr_gt[s.cam_model == 1, 8] = 0.0
    # board poses in the object: board 0 is the reference (identity). Visible side is -z (towards a viewer on -z).
    brv, bt = np.zeros((n_board, 3)), np.zeros((n_board, 3))
    rig = cfg["rig"]
    if rig in ("stereo", "square"):
        for b in range(1, n_board):
            side = 1 if b % 2 else -1
            k = (b + 1) // 2
            brv[b] = [0.0, side * 0.35 * k, 0.0]
            bt[b] = [side * k * (2 * cx + 0.15), 0.02 * b, 0.05 * k]
    else:
        # boards on the faces of a prism around the object's z axis, visible side facing outward
        rad = max(0.25, 1.2 * cx * n_board / np.pi / 2)
        poses = []
        for b in range(n_board):
            ang = 2 * np.pi * b / n_board
            out = np.array([np.cos(ang), np.sin(ang), 0.3 * np.sin(3 * ang)])
            out /= np.linalg.norm(out)
            # board frame: -z = out
            zb = -out
            xb = np.cross(np.array([0, 0, 1.0]), zb)
            xb /= np.linalg.norm(xb)
            yb = np.cross(zb, xb)
            R = np.stack([xb, yb, zb], 1)  # columns = board axes in object frame
            c = rad * out
            t = c - R @ np.array([cx, cy, 0.0])
            poses.append((R, t))
        R0, t0 = poses[0]
        for b in range(n_board):  # re-express relative to board 0 so that board 0 is the identity
            R, t = poses[b]
            Rr = R0.T @ R
            brv[b] = Rot.from_matrix(Rr).as_rotvec()
            bt[b] = R0.T @ (t - t0)
    s.board_gt = np.concatenate([brv, bt], 1)
    # camera poses world->cam
    cams = []
    if rig == "stereo":
        for c in range(n_cam):
            cams.append((np.array([0.0, 0.02 * c, 0.0]), -Rot.from_rotvec([0.0, 0.02 * c, 0.0]).apply([0.3 * c, 0, 0])))
    elif rig == "square":
        offs = [(0, 0), (0.5, 0), (0, 0.5), (0.5, 0.5)]
        for c in range(n_cam):
            rv = np.array([0.03 * (c % 2), -0.03 * (c // 2), 0.01 * c])
            cams.append((rv, -Rot.from_rotvec(rv).apply([offs[c % 4][0], offs[c % 4][1], 0.02 * c])))
    elif rig == "ring":
        for c in range(n_cam):
            ang = 2 * np.pi * c / n_cam
            cams.append(look_at(np.array([3 * np.cos(ang), 3 * np.sin(ang), 0.4 * np.sin(2 * ang)]), np.zeros(3)))
    elif rig == "dome":
        ga = np.pi * (3 - np.sqrt(5))
        for c in range(n_cam):
            z = 0.05 + 0.9 * (c + 0.5) / n_cam
            r = np.sqrt(1 - z * z)
            cams.append(look_at(4 * np.array([r * np.cos(ga * c), r * np.sin(ga * c), z]), np.array([0, 0, 0.6])))
    else:
        raise ValueError(rig)
    s.cam_world = np.array([np.concatenate(c) for c in cams])  # world->cam
    # relative to camera 0 (group frame = camera 0 frame): T_c o T_0^-1
    R0 = Rot.from_rotvec(s.cam_world[0, :3])
    inv0 = (R0.inv().as_rotvec(), -R0.inv().apply(s.cam_world[0, 3:]))
    rel = [compose(s.cam_world[c, :3], s.cam_world[c, 3:], *inv0) for c in range(n_cam)]
    s.cam_gt = np.array([np.concatenate(r) for r in rel])
    s.cam_gt[0] = 0.0
    # initial values: GT perturbed (rot 0.01 rad, trans 1 %, f 1 %, pp 5 px, distortion -> 0)
    s.cam_init = s.cam_gt + np.concatenate([0.01 * rng.standard_normal((n_cam, 3)),
                                            0.01 * np.abs(s.cam_gt[:, 3:]).max() * rng.standard_normal((n_cam, 3))], 1)
    s.cam_init[0] = 0.0
    s.board_init = s.board_gt + np.concatenate([0.01 * rng.standard_normal((n_board, 3)),
                                                0.01 * (np.abs(s.board_gt[:, 3:]).max() + 0.1) * rng.standard_normal((n_board, 3))], 1)
    s.board_init[0] = 0.0
    s.intr_init = s.intr_gt.copy()
    s.intr_init[:, :2] *= 1 + 0.01 * rng.standard_normal((n_cam, 2))
    s.intr_init[:, 2:4] += 5 * rng.standard_normal((n_cam, 2))
    s.intr_init[:, 4:] = 0.0
    return s


def _frame_poses(s, rng, n):
    """GT object pose (group frame = camera-0 frame) for n frames."""
    rig = s.cfg["rig"]
    if rig in ("stereo", "square"):
        # object in front of the cameras, visible (-z) side towards them, tilt <= 45 deg
        tilt = Rot.from_rotvec(rng.uniform(-1, 1, (n, 3)) * np.array([0.6, 0.6, 0.8]))
        rv = tilt.as_rotvec()
        nx, ny = s.cfg["grid"]
        ctr = np.array([(nx - 1) * s.cfg["square"] / 2, (ny - 1) * s.cfg["square"] / 2, 0])
        pos = np.stack([rng.uniform(-0.5, 0.8, n), rng.uniform(-0.4, 0.7, n), rng.uniform(1.2, 2.6, n)], 1)
        if 1 in s.cam_model:
            pos[:, 2] = rng.uniform(0.8, 1.8, n)
        t = pos - tilt.apply(ctr)
        return rv, t
    # ring/dome: object near the world centre, any orientation; convert world->group (camera 0)
    if rig == "ring":   # prism axis near the ring's axis: spin about z, tilt <= ~30 deg
        Rw = Rot.from_rotvec(rng.uniform(-0.5, 0.5, (n, 3)) * np.array([1, 1, 0])) * \
            Rot.from_rotvec(np.stack([np.zeros(n), np.zeros(n), rng.uniform(-np.pi, np.pi, n)], 1))
    else:
        Rw = Rot.random(n, random_state=rng.integers(1 << 31))
    tw = rng.uniform(-0.3, 0.3, (n, 3)) + (np.array([0, 0, 0.6]) if rig == "dome" else 0)
    R0 = Rot.from_rotvec(s.cam_world[0, :3])
    return (R0 * Rw).as_rotvec(), R0.apply(tw) + s.cam_world[0, 3:]


def _generate_span(args):
    s, lo, hi, noise, outlier_frac, outlier_sigma = args
    return generate_observations(s, lo, hi, noise, outlier_frac, outlier_sigma)


def generate_observations(s, frame_lo=0, frame_hi=None, noise=0.3, outlier_frac=0.01, outlier_sigma=20.0, workers=1):
    """Observations of frames [frame_lo, frame_hi) -> dict of flat arrays (bobs sorted by frame, cam, board).
    workers > 1: the RNG chunks are independent, so spans of chunks are generated in forked worker processes and
    concatenated - bit-identical to the serial result."""
    frame_hi = s.n_frames if frame_hi is None else frame_hi
    c_lo, c_hi = frame_lo // CHUNK, (frame_hi + CHUNK - 1) // CHUNK
    if workers > 1 and c_hi - c_lo >= 4:
        import multiprocessing as mp
        nspan = min(workers * 2, c_hi - c_lo)
        cuts = [c_lo + (c_hi - c_lo) * i // nspan for i in range(nspan + 1)]
        spans = [(s, max(frame_lo, cuts[i] * CHUNK), min(frame_hi, cuts[i + 1] * CHUNK), noise, outlier_frac, outlier_sigma)
                 for i in range(nspan) if cuts[i + 1] > cuts[i]]
        with mp.get_context("fork").Pool(workers) as pool:
            parts = pool.map(_generate_span, spans)
        o = Scene()
        o.frame_lo, o.frame_hi = frame_lo, frame_hi
        for k in ("bobs_frame", "bobs_cam", "bobs_board", "u", "v", "pt_idx", "obj_gt", "obj_init"):
            setattr(o, k, np.concatenate([getattr(p, k) for p in parts]))
        cnt = np.concatenate([np.diff(p.bobs_ptr) for p in parts])
        o.bobs_ptr = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
        return o
    n_cam, n_board, P = s.n_cam, s.n_board, s.P
    keep = s.cfg["keep"]
    out = dict(frame=[], cam=[], board=[], cnt=[], u=[], v=[], pt=[], obj_rv=[], obj_t=[], obj_init=[])
    Rc = Rot.from_rotvec(s.cam_gt[:, :3]).as_matrix()  # [n_cam,3,3]
    Rb = Rot.from_rotvec(s.board_gt[:, :3]).as_matrix()
    Xb = np.einsum("bij,bpj->bpi", Rb, s.board_pts.astype(np.float64)) + s.board_gt[:, None, 3:]  # object frame
    for chunk in range(frame_lo // CHUNK, (frame_hi + CHUNK - 1) // CHUNK):
        f0, f1 = chunk * CHUNK, min((chunk + 1) * CHUNK, s.n_frames)
        rng = np.random.default_rng([s.seed, 7, chunk])
        n = f1 - f0
        rv, t = _frame_poses(s, rng, n)
        init = np.concatenate([rv + 0.01 * rng.standard_normal((n, 3)), t * (1 + 0.01 * rng.standard_normal((n, 3)))], 1)
        Ro = Rot.from_rotvec(rv).as_matrix()
        Xg = np.einsum("fij,bpj->fbpi", Ro, Xb) + t[:, None, None, :]            # [n, B, P, 3] group frame
        Xc = np.einsum("cij,fbpj->fcbpi", Rc, Xg) + s.cam_gt[None, :, None, None, 3:]  # [n, C, B, P, 3]
        uv = project(s.intr_gt[None, :, None, None, :], s.cam_model[None, :, None, None], Xc)
        # visible side of a board is its -z: normal in camera frame
        nb = np.einsum("cij,fjk,bk->fcbi", Rc, Ro, -Rb[:, :, 2])  # [n, C, B, 3]
        ctr = Xc.mean(3)
        facing = (nb * ctr).sum(-1) < -s.cfg.get("facing", 0.25) * np.linalg.norm(ctr, axis=-1)
        zmin = np.where(s.cam_model == 1, 0.05, 0.2)[None, :, None, None]
        a2 = (Xc[..., 0] ** 2 + Xc[..., 1] ** 2) / np.maximum(Xc[..., 2], 1e-9) ** 2
        fov_ok = a2 < np.where(s.cam_model == 1, 9.0, 1.2)[None, :, None, None]
        inimg = (Xc[..., 2] > zmin) & fov_ok & (uv[..., 0] >= 0) & (uv[..., 0] < IMG_W) & (uv[..., 1] >= 0) & (uv[..., 1] < IMG_H)
        inimg &= facing[..., None]
        bvis = inimg.sum(-1) >= 0.5 * P   # min_perc_pts (McCalib.cpp:320-322)
        bvis &= rng.random(bvis.shape) < keep
        noise_uv = noise * rng.standard_normal(uv.shape)
        is_out = rng.random(uv.shape[:-1]) < outlier_frac
        noise_uv += is_out[..., None] * outlier_sigma * rng.standard_normal(uv.shape)
        uvn = (uv + noise_uv).astype(np.float32)
        sel = inimg & bvis[..., None]
        fi, ci, bi = np.nonzero(bvis)
        # keep only frames in the requested range
        m = (fi + f0 >= frame_lo) & (fi + f0 < frame_hi)
        cnt = sel.sum(-1)[bvis]
        fsel, csel, bsel, psel = np.nonzero(sel)
        mo = (fsel + f0 >= frame_lo) & (fsel + f0 < frame_hi)
        out["frame"].append((fi + f0)[m]); out["cam"].append(ci[m]); out["board"].append(bi[m]); out["cnt"].append(cnt[m])
        out["u"].append(uvn[..., 0][sel][mo]); out["v"].append(uvn[..., 1][sel][mo])
        out["pt"].append((bsel * P + psel)[mo])
        fr = np.arange(f0, f1)
        mf = (fr >= frame_lo) & (fr < frame_hi)
        out["obj_rv"].append(rv[mf]); out["obj_t"].append(t[mf]); out["obj_init"].append(init[mf])
    cat = lambda k, dt: np.concatenate(out[k]).astype(dt) if out[k] else np.zeros(0, dt)
    o = Scene()
    o.frame_lo, o.frame_hi = frame_lo, frame_hi
    o.bobs_frame, o.bobs_cam, o.bobs_board = cat("frame", np.int32), cat("cam", np.int32), cat("board", np.int32)
    cnt = cat("cnt", np.int64)
    o.bobs_ptr = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
    o.u, o.v, o.pt_idx = cat("u", np.float32), cat("v", np.float32), cat("pt", np.int32)
    o.obj_gt = np.concatenate([np.concatenate(out["obj_rv"]), np.concatenate(out["obj_t"])], 1)
    o.obj_init = np.concatenate(out["obj_init"])
    return o


def make_problem(s, o, stage):
    """Flatten (scene, observations) into (Problem, initial Params, ground-truth Params) for a stage."""
    n_cam, n_board = s.n_cam, s.n_board
    cam_is_ref = np.zeros(n_cam, np.uint8); cam_is_ref[0] = 1
    board_is_ref = np.zeros(n_board, np.uint8); board_is_ref[0] = 1
    pts = s.board_pts.reshape(-1, 3)
    n_frames = o.frame_hi - o.frame_lo
    pose_slot = o.bobs_frame - o.frame_lo
    if stage in (4, 5):
        prob = Problem(stage, o.u, o.v, o.pt_idx, o.bobs_ptr, o.bobs_cam, pose_slot, o.bobs_board, pts,
                       s.cam_model, n_frames, n_board, cam_is_ref, board_is_ref)
        init = Params(s.cam_init, o.obj_init, s.board_init, s.intr_init if stage == 5 else s.intr_gt)
        gt = Params(s.cam_gt, o.obj_gt, s.board_gt, s.intr_gt)
        return prob, init, gt
    if stage == 3:
        # object-frame points, rounded to float like Object3D::pts_3d_ (geometrytools.cpp:346-352)
        Rb = Rot.from_rotvec(s.board_gt[:, :3]).as_matrix()
        Xo = (np.einsum("bij,bpj->bpi", Rb, s.board_pts.astype(np.float64)) + s.board_gt[:, None, 3:]).astype(np.float32)
        prob = Problem(3, o.u, o.v, o.pt_idx, o.bobs_ptr, o.bobs_cam, pose_slot, o.bobs_board, Xo.reshape(-1, 3),
                       s.cam_model, n_frames, 0, cam_is_ref, None)
        return prob, Params(s.cam_init, o.obj_init, np.zeros((0, 6)), s.intr_gt), Params(s.cam_gt, o.obj_gt, np.zeros((0, 6)), s.intr_gt)
    if stage == 2:
        # e-block = Object3DObs::pose_ per (frame, camera): T_cam o T_obj
        key = pose_slot.astype(np.int64) * n_cam + o.bobs_cam
        uniq, slot = np.unique(key, return_inverse=True)
        fr, cm = (uniq // n_cam).astype(int), (uniq % n_cam).astype(int)
        def comp(cam, obj):
            rv, t = compose(cam[cm, :3], cam[cm, 3:], obj[fr, :3], obj[fr, 3:])
            return np.concatenate([rv, t], 1)
        prob = Problem(2, o.u, o.v, o.pt_idx, o.bobs_ptr, o.bobs_cam, slot, o.bobs_board, pts, s.cam_model,
                       len(uniq), n_board, None, board_is_ref)
        return (prob, Params(np.zeros((n_cam, 6)), comp(s.cam_init, o.obj_init), s.board_init, s.intr_gt),
                Params(np.zeros((n_cam, 6)), comp(s.cam_gt, o.obj_gt), s.board_gt, s.intr_gt))
    if stage == 1:
        # e-block = BoardObs::pose_ per board observation: T_cam o T_obj o T_board; intrinsics free
        nb = len(o.bobs_cam)
        def comp(cam, obj, brd):
            rv, t = compose(obj[pose_slot, :3], obj[pose_slot, 3:], brd[o.bobs_board, :3], brd[o.bobs_board, 3:])
            rv, t = compose(cam[o.bobs_cam, :3], cam[o.bobs_cam, 3:], rv, t)
            return np.concatenate([rv, t], 1)
        prob = Problem(1, o.u, o.v, o.pt_idx, o.bobs_ptr, o.bobs_cam, np.arange(nb), o.bobs_board, pts, s.cam_model,
                       nb, 0, None, None)
        return (prob, Params(np.zeros((n_cam, 6)), comp(s.cam_init, o.obj_init, s.board_init), np.zeros((0, 6)), s.intr_init),
                Params(np.zeros((n_cam, 6)), comp(s.cam_gt, o.obj_gt, s.board_gt), np.zeros((0, 6)), s.intr_gt))
    raise ValueError(stage)


def config_problem(name, stage, n_frames=None, frame_lo=0, frame_hi=None, seed=None):
    s = make_scene(name, seed=seed, n_frames=n_frames)
    o = generate_observations(s, frame_lo, frame_hi)
    return make_problem(s, o, stage) + (s, o)


def batched_intrinsics_problem(n_cameras=4096, n_frames=300, seed=1005, cam_lo=0, cam_hi=None, kannala_every=0):
    """Config 5: independent per-camera intrinsic refinement (Camera::refineIntrinsicCalibration,
    Camera.cpp:268-305): n_cameras problems x n_frames board observations x 16 corners, batched."""
    cam_hi = n_cameras if cam_hi is None else cam_hi
    nx = ny = 4; sq = 0.192; P = 16
    gx, gy = np.meshgrid(np.arange(nx), np.arange(ny))
    corners = np.stack([gx.ravel() * sq, gy.ravel() * sq, np.zeros(P)], -1).astype(np.float32)
    ctr = np.array([1.5 * sq, 1.5 * sq, 0.0])
    us, vs, pts_i, poses_gt, poses_init, intr_gt, intr_init, models = [], [], [], [], [], [], [], []
    bobs_cam, cnts = [], []
    for c in range(cam_lo, cam_hi):
        rng = np.random.default_rng([seed, c])
        model = 1 if (kannala_every and c % kannala_every == kannala_every - 1) else 0
        gt = (KANNALA_GT if model else BROWN_GT) * (1 + 0.02 * rng.standard_normal(9) * np.array([1, 1, 0.3, 0.3, 1, 1, 1, 1, 1]))
        if model:
            gt[8] = 0.0
        tilt = Rot.from_rotvec(rng.uniform(-1, 1, (n_frames, 3)) * np.array([0.6, 0.6, 0.8]))
        pos = np.stack([rng.uniform(-0.5, 0.5, n_frames), rng.uniform(-0.3, 0.3, n_frames),
                        rng.uniform(0.8, 1.6, n_frames) if model else rng.uniform(1.0, 2.4, n_frames)], 1)
        t = pos - tilt.apply(ctr)
        X = np.einsum("fij,pj->fpi", tilt.as_matrix(), corners.astype(np.float64)) + t[:, None, :]
        uv = project(gt[None, None, :], np.int32(model), X)
        uv = uv + 0.3 * rng.standard_normal(uv.shape)
        is_out = rng.random(uv.shape[:-1]) < 0.01
        uv += is_out[..., None] * 20.0 * rng.standard_normal(uv.shape)
        ok = (uv[..., 0] >= 0) & (uv[..., 0] < IMG_W) & (uv[..., 1] >= 0) & (uv[..., 1] < IMG_H)
        keep = ok.sum(-1) >= 8
        sel = ok & keep[:, None]
        fsel, psel = np.nonzero(sel)
        us.append(uv[..., 0][sel].astype(np.float32)); vs.append(uv[..., 1][sel].astype(np.float32)); pts_i.append(psel)
        cnts.append(sel.sum(-1)[keep]); bobs_cam.append(np.full(keep.sum(), c - cam_lo))
        g = np.concatenate([tilt.as_rotvec(), t], 1)[keep]
        poses_gt.append(g)
        poses_init.append(g + np.concatenate([0.01 * rng.standard_normal((len(g), 3)), 0.01 * g[:, 3:] * rng.standard_normal((len(g), 3))], 1))
        ini = gt.copy(); ini[:2] *= 1 + 0.01 * rng.standard_normal(2); ini[2:4] += 5 * rng.standard_normal(2); ini[4:] = 0
        intr_gt.append(gt); intr_init.append(ini); models.append(model)
    bobs_cam = np.concatenate(bobs_cam)
    nb = len(bobs_cam)
    ptr = np.concatenate([[0], np.cumsum(np.concatenate(cnts))])
    ncam = cam_hi - cam_lo
    prob = Problem(1, np.concatenate(us), np.concatenate(vs), np.concatenate(pts_i), ptr, bobs_cam, np.arange(nb),
                   np.zeros(nb), corners, np.array(models), nb, 0, None, None, cam_problem=np.arange(ncam), n_problems=ncam)
    init = Params(np.zeros((ncam, 6)), np.concatenate(poses_init), np.zeros((0, 6)), np.array(intr_init))
    gt = Params(np.zeros((ncam, 6)), np.concatenate(poses_gt), np.zeros((0, 6)), np.array(intr_gt))
    return prob, init, gt


def shard_problem(prob, params, rank, world):
    """Frame shard `rank` of `world` of a stage 3-5 problem: a contiguous range of e-blocks with all
    their board observations; camera / board / intrinsics blocks are replicated. -> (Problem, Params, (lo, hi))."""
    lo, hi = (prob.n_pose * rank) // world, (prob.n_pose * (rank + 1)) // world
    b0, b1 = np.searchsorted(prob.bobs_pose, [lo, hi])
    o0, o1 = prob.bobs_ptr[b0], prob.bobs_ptr[b1]
    sub = Problem(prob.stage, prob.u[o0:o1], prob.v[o0:o1], prob.pt_idx[o0:o1], prob.bobs_ptr[b0:b1 + 1] - o0,
                  prob.bobs_cam[b0:b1], prob.bobs_pose[b0:b1] - lo, prob.bobs_board[b0:b1], prob.pts_xyz, prob.cam_model,
                  hi - lo, prob.n_board, prob.cam_is_ref, prob.board_is_ref)
    return sub, Params(params.cam_pose, params.pose[lo:hi], params.board_pose, params.intrinsics), (lo, hi)
