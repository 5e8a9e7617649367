"""Host logic of the McCalib-shaped facade on a CPU-only box: the drivers of mc-calib_b200/host are linked against a
CPU stand-in of the C ABI (tests/stub_abi: the CPU oracle behind mcba_solve / mcba_solve_batched /
mcba_reprojection_errors, the sequential host restatement of the PnP kernels behind mcba_pnp_ransac) and run end to end.
What is checked is the facade itself — walking order, packing into the flat ABI arrays, scattering results back, the
pose-initialisation flows, the file formats — against the same checker called directly on problems packed in Python.
The product library is not involved (it has no CPU path); tests/test_gpu_facade.py runs the same flows on the GPU."""
import os
import subprocess

import numpy as np
import pytest

import mcba
import pnp_util
from mcba import pnp, synth
from test_gpu_facade import write_scene

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STUB = os.path.join(REPO, "tests", "stub_abi")


@pytest.fixture(scope="module")
def bins(oracle):
    subprocess.check_call(["make", "-s", "-C", STUB])
    return os.path.join(STUB, "facade_main_cpu"), os.path.join(STUB, "calibrate_main_cpu")


@pytest.fixture(scope="module")
def scene():
    s = synth.make_scene("c2", n_frames=24)
    o = synth.generate_observations(s)
    return s, o


def run(binary, tmp_path, s, o, flow, extra=()):
    scene_f, out = str(tmp_path / "scene.txt"), str(tmp_path / f"out_{flow}.txt")
    write_scene(scene_f, s, o)
    r = subprocess.run([binary, scene_f, flow, out, *extra], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    rows = [np.array(l.split(), float) for l in open(out)]
    nc, nbd, nf, nbo = s.n_cam, s.n_board, o.frame_hi - o.frame_lo, len(o.bobs_cam)
    cams = np.array(rows[1:1 + nc])
    rest = rows[1 + nc + nbd + nf:]
    res = dict(avg=float(rows[0][0]), intr=cams[:, :9], cam_pose=cams[:, 9:], board=np.array(rows[1 + nc:1 + nc + nbd]),
               obj=np.array(rows[1 + nc + nbd:1 + nc + nbd + nf]), bobs_pose=np.array(rest[:nbo]), oo_pose=np.array(rest[nbo:]), out=out)
    if os.path.exists(out + ".init.txt"):
        ir = [np.array(l.split(), float) for l in open(out + ".init.txt")]
        res.update(valid=np.array([r_[0] for r_ in ir[:nbo]], int), kept=np.array([r_[1] for r_ in ir[:nbo]], int))
    return res


def test_final_flow_packs_and_scatters_like_the_direct_call(bins, tmp_path, scene, oracle):
    """refineAllCameraGroupAndObjects then ...AndIntrinsic == stage 4 then stage 5 on the problem packed in Python."""
    s, o = scene
    got = run(bins[0], tmp_path, s, o, "final")
    prob, init, gt = synth.make_problem(s, o, 4)
    p = init.copy()
    p.intrinsics[:] = s.intr_init
    assert oracle.solve(prob, p)[0] == 0
    assert oracle.solve(prob.with_stage(5), p)[0] == 0
    assert np.abs(got["intr"] - p.intrinsics).max() <= 1e-9 * np.abs(p.intrinsics).max()
    assert np.abs(got["cam_pose"] - p.cam_pose).max() <= 1e-9
    assert np.abs(got["board"] - p.board_pose).max() <= 1e-9
    assert np.abs(got["obj"] - p.pose).max() <= 1e-9
    import cv2       # the savers' files read back with OpenCV
    fs = cv2.FileStorage(got["out"] + ".cameras.yml", cv2.FILE_STORAGE_READ)
    K = fs.getNode("camera_1").getNode("camera_matrix").mat()
    assert np.allclose([K[0, 0], K[1, 1], K[0, 2], K[1, 2]], p.intrinsics[1, :4], rtol=1e-8)
    fs.release()


def test_error_outputs_hold_the_reference_values(bins, tmp_path, scene):
    """reprojection_error_data.yml, reproErrorAllCamGroup and computeAvgReprojectionError against the reference's own
    arithmetic (OpenCV projection of float points, float stores: tests/ref_errors.py), key by key with cv::FileStorage;
    one frame has no observation at all and must still be listed (McCalib.cpp:2262-2270 walks the group's frames)."""
    import copy
    import ref_errors
    s, o = scene
    o2 = copy.copy(o)
    keep = o.bobs_frame != 5
    cnt = np.diff(o.bobs_ptr)
    obs_keep = np.repeat(keep, cnt)
    o2.bobs_frame, o2.bobs_cam, o2.bobs_board = o.bobs_frame[keep], o.bobs_cam[keep], o.bobs_board[keep]
    o2.bobs_ptr = np.concatenate([[0], np.cumsum(cnt[keep])]).astype(np.int64)
    o2.u, o2.v, o2.pt_idx = o.u[obs_keep], o.v[obs_keep], o.pt_idx[obs_keep]
    got = run(bins[0], tmp_path, s, o2, "final")
    n_exact, n_all = ref_errors.check_error_files(got["out"], s, o2, got, o.frame_hi - o.frame_lo)
    assert n_all == len(o2.u)


def test_intrinsics_flow_batches_the_cameras(bins, tmp_path, scene, oracle):
    """refineIntrinsicAndPoseAllCam hands all cameras to mcba_solve_batched: per camera identical to a single solve."""
    s, o = scene
    got = run(bins[0], tmp_path, s, o, "intrinsics")
    prob, init, gt = synth.make_problem(s, o, 1)
    for c in range(s.n_cam):
        sel = np.nonzero(prob.bobs_cam == c)[0]
        counts = np.diff(prob.bobs_ptr)[sel]
        obs = np.concatenate([np.arange(prob.bobs_ptr[b], prob.bobs_ptr[b + 1]) for b in sel])
        sub = mcba.Problem(1, prob.u[obs], prob.v[obs], prob.pt_idx[obs], np.concatenate([[0], np.cumsum(counts)]),
                           np.zeros(len(sel)), np.arange(len(sel)), np.zeros(len(sel)), prob.pts_xyz, prob.cam_model[c:c + 1],
                           len(sel), 0)
        p = mcba.Params(np.zeros((1, 6)), init.pose[sel], np.zeros((0, 6)), init.intrinsics[c:c + 1])
        assert oracle.solve(sub, p)[0] == 0
        assert np.abs(got["intr"][c] - p.intrinsics[0]).max() <= 1e-9 * np.abs(p.intrinsics[0]).max()
        assert np.abs(got["bobs_pose"][sel] - p.pose).max() <= 1e-9


def test_objects_flow(bins, tmp_path, scene, oracle):
    s, o = scene
    got = run(bins[0], tmp_path, s, o, "objects")
    prob, init, gt = synth.make_problem(s, o, 2)
    p = init.copy()
    p.intrinsics[:] = s.intr_init
    assert oracle.solve(prob, p)[0] == 0
    assert np.abs(got["board"] - p.board_pose).max() <= 1e-9
    assert np.abs(got["oo_pose"] - p.pose).max() <= 1e-9


def test_pose_initialisation_flow(bins, tmp_path, scene):
    """estimatePoseAllBoards on the facade's packing == the same checker on the Python packing (draws keyed by the
    observation index, so any difference in order or content shows); validity and kept corners follow BoardObs.cpp:143-161."""
    s, o = scene
    got = run(bins[0], tmp_path, s, o, "init", extra=("3.0", "5"))
    prob, init, gt = synth.make_problem(s, o, 1)
    P = pnp.PnpProblem(prob, s.intr_init)
    opt = pnp.mcba_pnp_options(3.0, 1000, 0.99, 1, 20, 5)
    want = pnp_util.host_ransac_all(pnp_util.host_lib(), P, opt)
    assert np.array_equal(got["bobs_pose"], want["pose"])
    assert np.array_equal(got["kept"], want["n_inliers"]) and np.array_equal(got["valid"], (want["n_inliers"] >= 4).astype(int))
    assert (want["n_inliers"] < np.diff(prob.bobs_ptr)).any()
    # every frame got a group pose close to the ground truth (reference camera's observation, or the average)
    err = [pnp_util.pose_diff(got["obj"][f], o.obj_gt[f]) for f in range(len(got["obj"])) if got["obj"][f].any()]
    assert len(err) > 0.9 * len(got["obj"]) and np.median(err) < 0.05


def test_workflow_head_from_files(bins, tmp_path, oracle):
    """Calibration(config) -> boardExtraction -> initIntrinsic from a config file, a keypoint file and a camera-parameter
    file written by OpenCV; refined focal lengths land near the ground truth."""
    import cv2
    import test_host_io as hio
    s = synth.make_scene("c1", n_frames=40)
    o = synth.generate_observations(s)
    cfg, kp, cp, out = (str(tmp_path / n) for n in ("config.yml", "kp.yml", "cams.yml", "intr.yml"))
    hio.write_keypoints_with_opencv(kp, s, o)
    hio.write_camera_params_with_opencv(cp, s, s.intr_init)
    nx, ny = s.cfg["grid"]
    hio.write_config(cfg, n_cam=s.n_cam, n_board=s.n_board, nx=nx + 1, ny=ny + 1, square=s.cfg["square"], cam_params=cp,
                     keypoints=kp, thr=10, iters=100)
    r = subprocess.run([bins[1], cfg, out], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    fs = cv2.FileStorage(out, cv2.FILE_STORAGE_READ)
    for c in range(s.n_cam):
        K = fs.getNode(f"camera_{c}").getNode("camera_matrix").mat()
        assert abs(K[0, 0] - s.intr_gt[c, 0]) <= 0.01 * s.intr_gt[c, 0] and abs(K[1, 1] - s.intr_gt[c, 1]) <= 0.01 * s.intr_gt[c, 1]
    fs.release()
    line = [l for l in r.stdout.splitlines() if l.startswith("valid board observations")][0].split()
    assert int(line[3]) == len(o.bobs_cam) and 0.2 < float(line[-2]) < 0.6


def test_facade_dumps_the_packed_problems(bins, tmp_path, scene, monkeypatch):
    """MCBA_DUMP_PROBLEM_DIR: solveStages checkpoints every packed problem with the parameters the solve starts from
    (mcba_problem_save). The stage-4 dump of the final flow holds, board observation for board observation, the problem
    packed in Python - the packer's slot assignment and corner tables checked without going through a solve - and the
    stage-5 dump the same arrays with the parameters stage 4 ended at."""
    s, o = scene
    d = tmp_path / "dumps"
    d.mkdir()
    monkeypatch.setenv("MCBA_DUMP_PROBLEM_DIR", str(d))
    got = run(bins[0], tmp_path, s, o, "final")
    files = sorted(os.listdir(d))
    assert files == ["problem_0_stage4.mcba", "problem_1_stage5.mcba"]
    p4, x4 = mcba.load_problem(str(d / files[0]))
    p5, x5 = mcba.load_problem(str(d / files[1]))
    prob, init, gt = synth.make_problem(s, o, 4)
    assert (p4.stage, p5.stage) == (4, 5) and p4.n_board == prob.n_board and p4.n_cam == prob.n_cam
    # the facade numbers cameras / boards in the order its walk first meets them (Packed::camSlot / boardSlot): recover
    # the relabelling from the parameter blocks (initial intrinsics and board poses are distinct per camera / board)
    cam_of = [int(np.argmin(np.abs(s.intr_init - x4.intrinsics[f]).max(axis=1))) for f in range(p4.n_cam)]
    board_of = [int(np.argmin(np.abs(init.board_pose - x4.board_pose[f]).max(axis=1))) for f in range(p4.n_board)]
    assert sorted(cam_of) == list(range(s.n_cam)) and sorted(board_of) == list(range(s.n_board))
    assert np.abs(x4.intrinsics - s.intr_init[cam_of]).max() == 0.0 and np.abs(x4.board_pose - init.board_pose[board_of]).max() <= 1e-12

    def canon(p, cam_of=None, board_of=None):   # board observations keyed by (e-block, camera, board)
        out = {}
        rank = np.unique(p.bobs_pose, return_inverse=True)[1]   # e-block slots: only observed (frame, object) pairs get one
        for b in range(p.n_bobs):
            sl = slice(p.bobs_ptr[b], p.bobs_ptr[b + 1])
            c, bd = int(p.bobs_cam[b]), int(p.bobs_board[b])
            key = (int(rank[b]), cam_of[c] if cam_of else c, board_of[bd] if board_of else bd)
            assert key not in out
            out[key] = (p.u[sl].tobytes(), p.v[sl].tobytes(), p.pts_xyz[p.pt_idx[sl]].tobytes())
        return out
    ref = canon(prob)
    assert canon(p4, cam_of, board_of) == ref and canon(p5, cam_of, board_of) == ref and p4.n_obs == prob.n_obs
    assert np.all(np.diff(p4.bobs_pose) >= 0) and np.array_equal(p4.cam_model, prob.cam_model[cam_of])
    for k in ("u", "v", "pt_idx", "bobs_ptr", "bobs_cam", "bobs_pose", "bobs_board", "cam_model"):
        assert np.array_equal(getattr(p4, k), getattr(p5, k)), k
    assert np.array_equal(p4.cam_is_ref, prob.cam_is_ref[cam_of]) and np.array_equal(p4.board_is_ref, prob.board_is_ref[board_of])
    seen = np.unique(prob.bobs_pose)
    assert p4.n_pose == len(seen)
    assert np.abs(x4.pose - init.pose[seen]).max() <= 1e-12 and np.abs(x4.cam_pose - init.cam_pose[cam_of]).max() <= 1e-12
    assert np.abs(x5.pose - x4.pose).max() > 0.0                                     # stage 4 moved the e-blocks
    assert np.abs(x5.intrinsics - x4.intrinsics).max() == 0.0                        # ... and not the intrinsics
