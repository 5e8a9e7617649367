"""The McCalib-shaped C++ facade (mc-calib_b200/host) end to end: a scene is written to a text file, the C++ driver
builds the Camera/Board/Object3D/Frame/CameraGroup graph, calls the Calibration::refineAll* entry points, and the
refined blocks are compared with solving the same stage directly through the C ABI from Python."""
import os
import subprocess

import numpy as np
import pytest

import mcba
from mcba import synth

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(REPO, "mc-calib_b200", "host")


def write_scene(path, s, o, n_iter=1000, fix_intrinsic=0):
    with open(path, "w") as f:
        nb = len(o.bobs_cam)
        f.write(f"{s.n_cam} {s.n_board} {o.frame_hi - o.frame_lo} {nb} {n_iter} {fix_intrinsic}\n")
        for c in range(s.n_cam):
            f.write(f"{int(s.cam_model[c])} " + " ".join(repr(float(x)) for x in s.intr_init[c]) + " " +
                    " ".join(repr(float(x)) for x in s.cam_init[c]) + "\n")
        for b in range(s.n_board):
            pts = s.board_pts[b]
            f.write(f"{len(pts)} " + " ".join(repr(float(x)) for x in pts.reshape(-1)) + " " +
                    " ".join(repr(float(x)) for x in s.board_init[b]) + "\n")
        for fr in range(o.frame_hi - o.frame_lo):
            f.write(" ".join(repr(float(x)) for x in o.obj_init[fr]) + "\n")
        for k in range(nb):
            i0, i1 = o.bobs_ptr[k], o.bobs_ptr[k + 1]
            ids = o.pt_idx[i0:i1] - o.bobs_board[k] * s.P
            f.write(f"{o.bobs_frame[k] - o.frame_lo} {o.bobs_cam[k]} {o.bobs_board[k]} {i1 - i0} " +
                    " ".join(f"{int(a)} {float(u)!r} {float(v)!r}" for a, u, v in zip(ids, o.u[i0:i1], o.v[i0:i1])) + "\n")


def run_facade(tmp_path, s, o, flow):
    subprocess.check_call(["make", "-s", "-C", HOST])
    scene, out = str(tmp_path / "scene.txt"), str(tmp_path / f"out_{flow}.txt")
    write_scene(scene, s, o)
    r = subprocess.run([os.path.join(HOST, "facade_main"), scene, flow, out], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    rows = [np.array(l.split(), float) for l in open(out)]
    avg = float(rows[0][0])
    nc, nbd, nf = s.n_cam, s.n_board, o.frame_hi - o.frame_lo
    cams = np.array(rows[1:1 + nc])
    boards = np.array(rows[1 + nc:1 + nc + nbd])
    objs = np.array(rows[1 + nc + nbd:1 + nc + nbd + nf])
    rest = rows[1 + nc + nbd + nf:]
    nbo = len(o.bobs_cam)
    return dict(avg=avg, intr=cams[:, :9], cam_pose=cams[:, 9:], board=boards, obj=objs,
                bobs_pose=np.array(rest[:nbo]), oo_pose=np.array(rest[nbo:]), cameras_yml=out + ".cameras.yml")


@pytest.fixture(scope="module")
def scene():
    s = synth.make_scene("c2", n_frames=60)
    o = synth.generate_observations(s)
    return s, o


def test_final_bundle_adjustment_flow(tmp_path, scene):
    """refineAllCameraGroupAndObjects then ...AndIntrinsic (calibrate.cpp:60-64) == stage 4 then stage 5 through the ABI."""
    s, o = scene
    got = run_facade(tmp_path, s, o, "final")
    prob, init, gt = synth.make_problem(s, o, 4)
    p = init.copy()
    p.intrinsics[:] = s.intr_init
    h = mcba.Handle(prob)
    h.solve(p)
    h.set_stage(5)
    h.solve(p)
    err = h.reprojection_errors(p)
    h.close()
    assert np.abs(got["intr"] - p.intrinsics).max() <= 1e-9 * np.abs(p.intrinsics).max()
    assert np.abs(got["cam_pose"] - p.cam_pose).max() <= 1e-9
    assert np.abs(got["board"] - p.board_pose).max() <= 1e-9
    assert np.abs(got["obj"] - p.pose).max() <= 1e-9
    # mean over object observations (frame, camera) of the mean corner error (McCalib.cpp:2168-2245); the facade goes
    # through float object points like the reference, so only ~1e-4 px agreement is expected
    key = o.bobs_frame.astype(np.int64) * s.n_cam + o.bobs_cam
    per_obs_key = np.repeat(key, np.diff(o.bobs_ptr))
    means = [err[per_obs_key == k].mean() for k in np.unique(per_obs_key)]
    assert abs(got["avg"] - np.mean(means)) <= 1e-3
    # the output files are OpenCV-FileStorage YAML like the reference's savers: read them back with OpenCV itself
    import cv2
    base = got["cameras_yml"][:-len(".cameras.yml")]
    fs = cv2.FileStorage(base + ".cameras.yml", cv2.FILE_STORAGE_READ)
    assert int(fs.getNode("nb_camera").real()) == s.n_cam
    for c in range(s.n_cam):
        node = fs.getNode(f"camera_{c}")
        K = node.getNode("camera_matrix").mat()
        assert np.allclose([K[0, 0], K[1, 1], K[0, 2], K[1, 2]], p.intrinsics[c, :4], rtol=1e-8)
        assert int(node.getNode("distortion_type").real()) == int(s.cam_model[c])
        nd = 5 if s.cam_model[c] == 0 else 4
        assert np.allclose(node.getNode("distortion_vector").mat().ravel(), p.intrinsics[c, 4:4 + nd], rtol=1e-7, atol=1e-12)
        T = node.getNode("camera_pose_matrix").mat()       # inverse of relative_camera_pose_ (McCalib.cpp:398-399)
        R = synth.Rot.from_rotvec(p.cam_pose[c, :3]).as_matrix()
        Tf = np.eye(4); Tf[:3, :3] = R; Tf[:3, 3] = p.cam_pose[c, 3:]
        assert np.allclose(T @ Tf, np.eye(4), atol=1e-9)
    fs.release()
    fs = cv2.FileStorage(base + ".objects.yml", cv2.FILE_STORAGE_READ)
    pts = fs.getNode("object_0").getNode("points").mat()
    assert pts.shape == (5, s.n_board * s.P) and set(np.unique(pts[3]).astype(int)) == set(range(s.n_board))
    fs.release()
    fs = cv2.FileStorage(base + ".objects_pose.yml", cv2.FILE_STORAGE_READ)
    poses = fs.getNode("object_0").getNode("poses").mat()
    assert poses.shape[0] == 6 and poses.shape[1] == len(np.unique(key))
    fs.release()
    # reprojection_error_data.yml, reproErrorAllCamGroup, the average: the reference's values (its float roundings)
    import ref_errors
    res = dict(got); res["out"] = base
    n_exact, n_all = ref_errors.check_error_files(base, s, o, res, o.frame_hi - o.frame_lo)
    assert n_all == len(o.u)


def test_intrinsics_flow(tmp_path, scene, oracle):
    """refineIntrinsicAndPoseAllCam: per camera stage 1 (Camera.cpp:268-305)."""
    s, o = scene
    got = run_facade(tmp_path, s, o, "intrinsics")
    prob, init, gt = synth.make_problem(s, o, 1)
    for c in range(s.n_cam):
        sel = np.nonzero(prob.bobs_cam == c)[0]
        counts = np.diff(prob.bobs_ptr)[sel]
        obs = np.concatenate([np.arange(prob.bobs_ptr[b], prob.bobs_ptr[b + 1]) for b in sel])
        sub = mcba.Problem(1, prob.u[obs], prob.v[obs], prob.pt_idx[obs], np.concatenate([[0], np.cumsum(counts)]),
                           np.zeros(len(sel)), np.arange(len(sel)), np.zeros(len(sel)), prob.pts_xyz, prob.cam_model[c:c + 1],
                           len(sel), 0)
        p = mcba.Params(np.zeros((1, 6)), init.pose[sel], np.zeros((0, 6)), init.intrinsics[c:c + 1])
        rc, so, _ = oracle.solve(sub, p)
        assert rc == 0
        assert np.abs(got["intr"][c] - p.intrinsics[0]).max() <= 1e-6 * np.abs(p.intrinsics[0]).max()
        assert np.abs(got["bobs_pose"][sel] - p.pose).max() <= 1e-6


def test_objects_and_group_flows(tmp_path, scene, oracle):
    s, o = scene
    got = run_facade(tmp_path, s, o, "objects")          # refineAllObject3D, stage 2
    prob, init, gt = synth.make_problem(s, o, 2)
    p = init.copy()
    p.intrinsics[:] = s.intr_init          # the facade's cameras carry the initial intrinsics (constants in stage 2)
    rc, so, _ = oracle.solve(prob, p)
    assert np.abs(got["board"] - p.board_pose).max() <= 1e-6
    assert np.abs(got["oo_pose"] - p.pose).max() <= 1e-6
    got = run_facade(tmp_path, s, o, "group")            # refineAllCameraGroup, stage 3 (object-frame float points)
    prob, init, gt = synth.make_problem(s, o, 3)
    # the facade's object points come from the *initial* board poses (updateObjectPts), as in the reference
    Rb = synth.Rot.from_rotvec(s.board_init[:, :3]).as_matrix()
    Xo = (np.einsum("bij,bpj->bpi", Rb, s.board_pts.astype(np.float64)) + s.board_init[:, None, 3:]).astype(np.float32)
    prob.pts_xyz = np.ascontiguousarray(Xo.reshape(-1, 3))
    p = init.copy()
    p.intrinsics[:] = s.intr_init
    rc, so, _ = oracle.solve(prob, p)
    assert np.abs(got["cam_pose"] - p.cam_pose).max() <= 1e-4
    assert np.abs(got["obj"] - p.pose).max() <= 1e-4


def run_init(tmp_path, s, o, flow, thr=10.0, seed=5):
    subprocess.check_call(["make", "-s", "-C", HOST])
    scene, out = str(tmp_path / "scene.txt"), str(tmp_path / f"out_{flow}.txt")
    write_scene(scene, s, o)
    r = subprocess.run([os.path.join(HOST, "facade_main"), scene, flow, out, repr(thr), str(seed)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    rows = [np.array(l.split(), float) for l in open(out)]
    init_rows = [np.array(l.split(), float) for l in open(out + ".init.txt")]
    nc, nbd, nf, nbo = s.n_cam, s.n_board, o.frame_hi - o.frame_lo, len(o.bobs_cam)
    rest = rows[1 + nc + nbd + nf:]
    cams = np.array(rows[1:1 + nc])
    return dict(avg=float(rows[0][0]), intr=cams[:, :9], cam_pose=cams[:, 9:], board=np.array(rows[1 + nc:1 + nc + nbd]),
                obj=np.array(rows[1 + nc + nbd:1 + nc + nbd + nf]), bobs_pose=np.array(rest[:nbo]), oo_pose=np.array(rest[nbo:]),
                valid=np.array([r[0] for r in init_rows[:nbo]], int), kept=np.array([r[1] for r in init_rows[:nbo]], int),
                oo_npts=np.array([r[0] for r in init_rows[nbo:]], int), oo_group=np.array([r[1:] for r in init_rows[nbo:]]))


def test_pose_initialisation_flow(tmp_path, scene):
    """No pose comes from the scene file: Calibration::estimatePoseAllBoards / estimatePoseAllObjects (one batched GPU
    launch each) and computeAllObjPoseInCameraGroup, against mcba_pnp_ransac called from Python on the same packing and
    a numpy restatement of the group-pose step."""
    from mcba import pnp
    s, o = scene
    got = run_init(tmp_path, s, o, "init")
    prob, init, gt = synth.make_problem(s, o, 1)
    P = pnp.PnpProblem(prob, s.intr_init)                  # the facade's cameras carry the initial intrinsics
    r = pnp.ransac(P, pnp.default_options(threshold_px=10.0, max_iterations=1000, seed=5))
    assert np.array_equal(got["bobs_pose"], r.pose)        # same draws, same arithmetic: identical
    assert np.array_equal(got["valid"], (r.n_inliers >= 4).astype(int))
    assert np.array_equal(got["kept"], r.n_inliers)
    assert (r.n_inliers < np.diff(prob.bobs_ptr)).any()    # some 20 px outliers were dropped
    # object observations: (frame, camera) runs of the kept corners, object-frame float points from the initial board poses
    key = o.bobs_frame.astype(np.int64) * s.n_cam + o.bobs_cam
    runs = [np.nonzero(key == k)[0] for k in np.unique(key)]
    assert np.array_equal(got["oo_npts"], [r.n_inliers[b].sum() for b in runs])
    Rb = synth.Rot.from_rotvec(s.board_init[:, :3]).as_matrix()
    Xo = (np.einsum("bij,bpj->bpi", Rb, s.board_pts.astype(np.float64)) + s.board_init[:, None, 3:]).astype(np.float32).reshape(-1, 3)
    keep = r.inlier_mask.astype(bool)
    u, v, pt, ptr, cam = [], [], [], [0], []
    for b in runs:
        for bb in b:
            sl = np.arange(prob.bobs_ptr[bb], prob.bobs_ptr[bb + 1])
            sl = sl[keep[sl]]
            u.append(prob.u[sl]); v.append(prob.v[sl]); pt.append(prob.pt_idx[sl])
        ptr.append(ptr[-1] + int(r.n_inliers[b].sum()))
        cam.append(prob.bobs_cam[b[0]])
    Po = pnp.PnpProblem.__new__(pnp.PnpProblem)
    Po.u, Po.v, Po.pt_idx = np.concatenate(u), np.concatenate(v), np.concatenate(pt).astype(np.int32)
    Po.bobs_ptr, Po.bobs_cam = np.array(ptr, np.int64), np.array(cam, np.int32)
    Po.pts_xyz, Po.cam_model, Po.intrinsics = Xo, prob.cam_model, P.intrinsics
    ro = pnp.ransac(Po, pnp.default_options(threshold_px=10.0, max_iterations=1000, seed=6))
    assert np.abs(got["oo_pose"] - ro.pose).max() < 1e-6   # float object points come from two rotation-matrix routines
    # group pose per frame: the reference camera's observation if there is one, else the quaternion average
    oo_frame = np.array([o.bobs_frame[b[0]] for b in runs]) - o.frame_lo
    oo_cam = np.array(cam)
    for f in range(o.frame_hi - o.frame_lo):
        idx = np.nonzero(oo_frame == f)[0]
        if len(idx) == 0:
            continue
        ing = []
        for j in idx:
            Rc = synth.Rot.from_rotvec(s.cam_init[oo_cam[j], :3])
            ing.append(np.concatenate(synth.compose(Rc.inv().as_rotvec(), -Rc.inv().apply(s.cam_init[oo_cam[j], 3:]),
                                                    got["oo_pose"][j, :3], got["oo_pose"][j, 3:])))
        ing = np.array(ing)
        if (oo_cam[idx] == 0).any():
            want = ing[np.nonzero(oo_cam[idx] == 0)[0][-1]]
        else:
            want = np.concatenate([synth.Rot.from_rotvec(ing[:, :3]).mean().as_rotvec(), ing[:, 3:].mean(0)])
        assert np.abs(synth.Rot.from_rotvec(got["obj"][f, :3]).as_matrix() - synth.Rot.from_rotvec(want[:3]).as_matrix()).max() < 1e-8
        assert np.abs(got["obj"][f, 3:] - want[3:]).max() < 1e-8
        assert np.allclose(got["oo_group"][idx], got["obj"][f], rtol=0, atol=0)


def test_calibration_from_observations_only(tmp_path, scene):
    """Pose initialisation followed by the final bundle adjustment (no per-frame pose given) lands on the same
    calibration as the run that starts from the perturbed ground-truth poses."""
    s, o = scene
    a = run_init(tmp_path, s, o, "init_final")
    b = run_facade(tmp_path, s, o, "final")
    # the RANSAC drops the corners farther than 10 px (most of the synthetic 20 px outliers), so the mean error is lower
    assert 0.2 < a["avg"] <= b["avg"] + 0.01
    assert np.abs(a["intr"][:, :4] - b["intr"][:, :4]).max() <= 2e-3 * np.abs(b["intr"][:, :4]).max()
    assert np.abs(a["cam_pose"] - b["cam_pose"]).max() <= 5e-3
    assert np.abs(a["intr"][:, :2] - s.intr_gt[:, :2]).max() <= 0.01 * s.intr_gt[:, :2].max()


def test_workflow_head_from_files(tmp_path):
    """calibrate.cpp's first steps from files only: Calibration(config) -> boardExtraction (keypoint YAML written by
    OpenCV) -> initIntrinsic (camera-parameter YAML written by OpenCV, batched P3P RANSAC, batched stage-1 LM). The
    refined intrinsics (read back with OpenCV) are checked against a Python run of the same two GPU calls and the
    ground truth."""
    import cv2
    from mcba import pnp
    import test_host_io as hio
    s = synth.make_scene("c2", n_frames=80)
    o = synth.generate_observations(s)
    subprocess.check_call(["make", "-s", "-C", HOST])
    cfg, kp, cp, out, kp_out = (str(tmp_path / n) for n in ("config.yml", "kp.yml", "cams.yml", "intr.yml", "kp_out.yml"))
    hio.write_keypoints_with_opencv(kp, s, o)
    hio.write_camera_params_with_opencv(cp, s, s.intr_init)
    nx, ny = s.cfg["grid"]
    hio.write_config(cfg, n_cam=s.n_cam, n_board=s.n_board, nx=nx + 1, ny=ny + 1, square=s.cfg["square"],
                     dist_pc="[" + ", ".join(str(int(m)) for m in s.cam_model) + "]", cam_params=cp, keypoints=kp, thr=10, iters=200)
    r = subprocess.run([os.path.join(HOST, "calibrate_main"), cfg, out, kp_out], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("valid board observations")][0].split()
    n_valid, n_kept, mean_err = int(line[3]), int(line[6]), float(line[-2])
    fs = cv2.FileStorage(out, cv2.FILE_STORAGE_READ)
    got = np.zeros((s.n_cam, 9))
    for c in range(s.n_cam):
        n = fs.getNode(f"camera_{c}")
        K, D = n.getNode("camera_matrix").mat(), n.getNode("distortion_vector").mat().ravel()
        got[c, :4] = [K[0, 0], K[1, 1], K[0, 2], K[1, 2]]
        got[c, 4:4 + len(D)] = D
        assert int(n.getNode("distortion_type").real()) == int(s.cam_model[c])
    fs.release()
    # the same two GPU calls from Python, on the facade's packing order (camera-major)
    prob, init, gt = synth.make_problem(s, o, 1)
    order = np.concatenate([np.nonzero(prob.bobs_cam == c)[0] for c in range(s.n_cam)])
    obs = np.concatenate([np.arange(prob.bobs_ptr[b], prob.bobs_ptr[b + 1]) for b in order])
    counts = np.diff(prob.bobs_ptr)[order]
    intr0 = s.intr_init.copy()
    intr0[s.cam_model == 1, 8] = 0.0
    sub = mcba.Problem(1, prob.u[obs], prob.v[obs], prob.pt_idx[obs], np.concatenate([[0], np.cumsum(counts)]), prob.bobs_cam[order],
                       np.arange(len(order)), np.zeros(len(order)), prob.pts_xyz, prob.cam_model, len(order), 0)
    # estimatePoseAllBoards walks board_observations_ in insertion order = camera-major file order
    r1 = pnp.ransac(pnp.PnpProblem(sub, intr0), pnp.default_options(threshold_px=10.0, max_iterations=200, seed=0))
    assert n_valid == int((r1.n_inliers >= 4).sum()) and n_kept == int(r1.inlier_mask.sum())
    keep = r1.inlier_mask.astype(bool)
    cnt2 = np.array([keep[a:b].sum() for a, b in zip(sub.bobs_ptr[:-1], sub.bobs_ptr[1:])])
    sub2 = mcba.Problem(1, sub.u[keep], sub.v[keep], sub.pt_idx[keep], np.concatenate([[0], np.cumsum(cnt2)]), sub.bobs_cam,
                        np.arange(len(order)), np.zeros(len(order)), prob.pts_xyz, prob.cam_model, len(order), 0,
                        cam_problem=np.arange(s.n_cam), n_problems=s.n_cam)
    p = mcba.Params(np.zeros((s.n_cam, 6)), r1.pose, np.zeros((0, 6)), intr0)
    h = mcba.Handle(sub2)
    h.solve_batched(p, mcba.default_options(max_num_iterations=200))
    err = None
    h.close()
    assert np.abs(got - p.intrinsics).max() <= 1e-9 * np.abs(p.intrinsics).max()
    assert np.abs(got[:, :2] - s.intr_gt[:, :2]).max() <= 0.01 * s.intr_gt[:, :2].max()
    assert 0.2 < mean_err < 0.6
    # the re-saved keypoints no longer hold the corners the RANSAC rejected
    kept_file = hio.read_keypoints_with_opencv(kp_out, s.n_cam)
    assert sum(len(x) for c in kept_file.values() for x in c["ids"]) == n_kept
