"""The reference's error OUTPUTS recomputed with its own arithmetic, for the facade tests: OpenCV (cv2, in this image)
does the projection exactly as MC-Calib calls it, numpy float32 reproduces the cv::Point3f / cv::Point2f / float stores.

  Object3D::updateObjectPts          board points -> object frame, transform3DPts (geometrytools.cpp:328-352), float
  transform3DPts (object pose)       float again
  projectPointsWithDistortion        cv::projectPoints / cv::fisheye::projectPoints on float points -> float pixels
  computeDistanceBetweenPoints       float difference, double square/sum/sqrt, float store (McCalib.cpp:2150-2160)
"""
import cv2
import numpy as np
from scipy.spatial.transform import Rotation as Rot


def transform3d_f32(pts_f32, pose):
    R = cv2.Rodrigues(np.asarray(pose[:3], np.float64))[0]
    p = pts_f32.astype(np.float64)
    out = np.empty_like(p)
    for i in range(3):      # tx + r11 x + r12 y + r13 z, left to right in double, then the float store
        out[:, i] = ((pose[3 + i] + R[i, 0] * p[:, 0]) + R[i, 1] * p[:, 1]) + R[i, 2] * p[:, 2]
    return out.astype(np.float32)


def project_f32(pts_f32, cam_pose, intr, model):
    K = np.array([[intr[0], 0, intr[2]], [0, intr[1], intr[3]], [0, 0, 1]], np.float64)
    rv, tv = np.asarray(cam_pose[:3], np.float64).reshape(3, 1), np.asarray(cam_pose[3:], np.float64).reshape(3, 1)
    if model == 0:
        uv, _ = cv2.projectPoints(pts_f32.reshape(-1, 1, 3), rv, tv, K, np.asarray(intr[4:9], np.float64))
    else:
        uv = cv2.fisheye.projectPoints(pts_f32.reshape(-1, 1, 3), rv, tv, K, np.asarray(intr[4:8], np.float64))[0]
    return uv.reshape(-1, 2).astype(np.float32)


def object_observation_errors(s, o, cam_pose, board_pose, obj_pose, intr):
    """-> {(frame, cam): float32 error list}, corners in the order the facade appends them (board after board)."""
    obj_pts = np.concatenate([transform3d_f32(s.board_pts[b].astype(np.float32), board_pose[b]) for b in range(s.n_board)])
    out = {}
    for k in range(len(o.bobs_cam)):
        fr, cam = int(o.bobs_frame[k]), int(o.bobs_cam[k])
        i0, i1 = o.bobs_ptr[k], o.bobs_ptr[k + 1]
        X = transform3d_f32(obj_pts[o.pt_idx[i0:i1]], obj_pose[fr - o.frame_lo])
        uv = project_f32(X, cam_pose[cam], intr[cam], int(s.cam_model[cam]))
        du = (o.u[i0:i1].astype(np.float32) - uv[:, 0]).astype(np.float64)
        dv = (o.v[i0:i1].astype(np.float32) - uv[:, 1]).astype(np.float64)
        e = np.sqrt(du * du + dv * dv).astype(np.float32)
        out.setdefault((fr, cam), []).append(e)
    return {k: np.concatenate(v) for k, v in out.items()}


def check_error_files(base, s, o, res, n_frames):
    """reprojection_error_data.yml key by key, the per-observation means and the average against the recomputation."""
    want = object_observation_errors(s, o, res["cam_pose"], res["board"], res["obj"], res["intr"])
    fs = cv2.FileStorage(base + ".reprojection_error.yml", cv2.FILE_STORAGE_READ)
    assert int(fs.getNode("nb_camera_group").real()) == 1
    grp = fs.getNode("camera_group_0")
    frames = grp.getNode("frame_list").mat().ravel().astype(int)
    assert list(frames) == list(range(n_frames))                 # every frame of the group, observed or not
    n_exact, n_all, worst = 0, 0, 0.0
    for fr in frames:
        node = grp.getNode(f"frame_{fr}")
        cams_want = sorted(c for (f_, c) in want if f_ == fr + o.frame_lo)
        cl = node.getNode("camera_list").mat()
        cams = [] if cl is None or cl.size == 0 else list(cl.ravel().astype(int))
        assert cams == cams_want, (fr, cams, cams_want)
        for c in cams:
            cn = node.getNode(f"camera_{c}")
            e = cn.getNode("error_list").mat().ravel()
            w = want[(fr + o.frame_lo, c)]
            assert int(cn.getNode("nb_pts").real()) == len(w) == len(e)
            assert e.dtype == np.float32
            n_exact += int((e == w).sum()); n_all += len(w); worst = max(worst, float(np.abs(e - w).max()))
    fs.release()
    # identical floats, up to a last-bit flip where OpenCV's and this library's double projections straddle a float
    # rounding boundary (they differ by ~1e-13 px)
    assert worst <= 2e-4 and n_exact >= 0.995 * n_all, (worst, n_exact, n_all)
    rows = [l.split() for l in open(base + ".repro_obs.txt")]
    assert len(rows) == len(want)
    for r in rows:
        w = want[(int(r[1]) + o.frame_lo, int(r[2]))]
        assert int(r[5]) == len(w) and abs(float(r[4]) - float(w.sum(dtype=np.float32) / np.float32(len(w)))) <= 2e-5
    avg = np.mean([float(w.astype(np.float64).mean()) for w in want.values()])
    assert abs(res["avg"] - avg) <= 1e-6
    # calibrated_cameras_data.yml: the reference's key set and order (McCalib.cpp:385-400)
    text = open(base + ".cameras.yml").read()
    blk = text[text.index("camera_0:"):text.index("camera_1:")]
    keys = [l.strip().split(":")[0] for l in blk.splitlines()[1:] if l.startswith("   ") and not l.startswith("    ")]
    assert keys == ["camera_matrix", "distortion_vector", "distortion_type", "camera_group", "img_width", "img_height", "camera_pose_matrix"], keys
    return n_exact, n_all
