"""Host pieces either side of the refinement path, CPU only:
 - detected_keypoints_data.yml / camera-parameter YAML in cv::FileStorage's dialect: files written by OpenCV itself
   (cv2.FileStorage, the library the reference uses) are read by the C++ facade, and the facade's writer is read back by
   OpenCV (McCalib.cpp:172-213, :493-548, :662-688);
 - getAverageRotation (geometrytools.cpp:832-893) against scipy's Markley mean and numpy medians;
 - CameraGroup::computeObjPoseInCameraGroup + CameraGroupObs::computeObjectsPose (CameraGroup.cpp:153-183,
   CameraGroupObs.cpp:42-108) against a numpy restatement."""
import os
import subprocess

import numpy as np
import pytest
from scipy.spatial.transform import Rotation as Rot

from mcba import synth

cv2 = pytest.importorskip("cv2")
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(REPO, "mc-calib_b200", "host")


@pytest.fixture(scope="module")
def io_main():
    if not os.path.exists(os.path.join(REPO, "mc-calib_b200", "csrc", "libmcba.so")):
        pytest.skip("libmcba.so not built")
    subprocess.check_call(["make", "-s", "-C", HOST])
    return os.path.join(HOST, "io_main")


def write_keypoints_with_opencv(path, s, o, frame_path=lambda c, f: f"/data/Cam_{c:03d}/{f:06d}.png"):
    """What Calibration::saveDetectedKeypoints emits, through OpenCV's own emitter."""
    fs = cv2.FileStorage(path, cv2.FILE_STORAGE_WRITE)
    fs.write("nb_camera", int(s.n_cam))
    for c in range(s.n_cam):
        sel = np.nonzero(o.bobs_cam == c)[0]
        fs.startWriteStruct(f"camera_{c}", cv2.FileNode_MAP)
        fs.write("img_width", 1920)
        fs.write("img_height", 1080)
        fs.startWriteStruct("frame_idxs", cv2.FileNode_SEQ | cv2.FileNode_FLOW)
        for b in sel:
            fs.write("", int(o.bobs_frame[b]))
        fs.endWriteStruct()
        fs.startWriteStruct("frame_paths", cv2.FileNode_SEQ)
        for b in sel:
            fs.write("", frame_path(c, int(o.bobs_frame[b])))
        fs.endWriteStruct()
        fs.startWriteStruct("board_idxs", cv2.FileNode_SEQ | cv2.FileNode_FLOW)
        for b in sel:
            fs.write("", int(o.bobs_board[b]))
        fs.endWriteStruct()
        fs.startWriteStruct("pts_2d", cv2.FileNode_SEQ)
        for b in sel:
            fs.startWriteStruct("", cv2.FileNode_SEQ | cv2.FileNode_FLOW)
            for i in range(o.bobs_ptr[b], o.bobs_ptr[b + 1]):
                fs.write("", float(o.u[i]))
                fs.write("", float(o.v[i]))
            fs.endWriteStruct()
        fs.endWriteStruct()
        fs.startWriteStruct("charuco_idxs", cv2.FileNode_SEQ)
        for b in sel:
            fs.startWriteStruct("", cv2.FileNode_SEQ | cv2.FileNode_FLOW)
            for i in range(o.bobs_ptr[b], o.bobs_ptr[b + 1]):
                fs.write("", int(o.pt_idx[i] - o.bobs_board[b] * s.P))
            fs.endWriteStruct()
        fs.endWriteStruct()
        fs.endWriteStruct()
    fs.release()


def write_camera_params_with_opencv(path, s, intr):
    fs = cv2.FileStorage(path, cv2.FILE_STORAGE_WRITE)
    fs.write("nb_camera", int(s.n_cam))
    for c in range(s.n_cam):
        fs.startWriteStruct(f"camera_{c}", cv2.FileNode_MAP)
        K = np.array([[intr[c, 0], 0, intr[c, 2]], [0, intr[c, 1], intr[c, 3]], [0, 0, 1.0]])
        fs.write("camera_matrix", K)
        nd = 5 if s.cam_model[c] == 0 else 4
        fs.write("distortion_vector", intr[c, 4:4 + nd].reshape(1, nd))
        fs.write("distortion_type", int(s.cam_model[c]))
        fs.endWriteStruct()
    fs.release()


def write_boards(path, s):
    nx, ny = s.cfg["grid"]
    with open(path, "w") as f:
        f.write(f"{s.n_board}\n" + "".join(f"{nx} {ny} {s.cfg['square']!r}\n" for _ in range(s.n_board)))


def read_keypoints_with_opencv(path, n_cam):
    fs = cv2.FileStorage(path, cv2.FILE_STORAGE_READ)
    assert int(fs.getNode("nb_camera").real()) == n_cam
    out = {}
    for c in range(n_cam):
        n = fs.getNode(f"camera_{c}")
        seq = lambda node: [node.at(i) for i in range(node.size())]
        out[c] = dict(
            w=int(n.getNode("img_width").real()), h=int(n.getNode("img_height").real()),
            frames=[int(x.real()) for x in seq(n.getNode("frame_idxs"))],
            paths=[x.string() for x in seq(n.getNode("frame_paths"))],
            boards=[int(x.real()) for x in seq(n.getNode("board_idxs"))],
            pts=[np.array([y.real() for y in seq(x)], np.float32) for x in seq(n.getNode("pts_2d"))],
            ids=[[int(y.real()) for y in seq(x)] for x in seq(n.getNode("charuco_idxs"))])
    fs.release()
    return out


def test_keypoint_and_intrinsics_files_round_trip_with_opencv(io_main, tmp_path):
    s = synth.make_scene("c2", n_frames=12)
    o = synth.generate_observations(s)
    kin, kout, cp, dump, boards = (str(tmp_path / n) for n in ("kp_in.yml", "kp_out.yml", "cams.yml", "dump.txt", "boards.txt"))
    write_keypoints_with_opencv(kin, s, o, frame_path=lambda c, f: f"/data/My \"Rig\"/Cam_{c:03d}\\{f:06d} #1.png")
    write_camera_params_with_opencv(cp, s, s.intr_init)
    write_boards(boards, s)
    r = subprocess.run([io_main, "keypoints", boards, kin, cp, kout, dump], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    rows = [l.split() for l in open(dump)]
    n_cam, n_frames, n_bo = map(int, rows[0])
    assert (n_cam, n_bo) == (s.n_cam, len(o.bobs_cam)) and n_frames == len(np.unique(o.bobs_frame))
    for c in range(s.n_cam):                          # intrinsics file -> Camera::intrinsics_
        row = rows[1 + c]
        assert [int(x) for x in row[:4]] == [c, 1920, 1080, int(s.cam_model[c])]
        assert int(row[4]) == int((o.bobs_cam == c).sum())
        want = s.intr_init[c].copy()
        if s.cam_model[c] == 1:
            want[8] = 0.0
        assert np.array_equal(np.array(row[5:], float), want)
    # board observations arrive camera by camera, in file order, bit-exact floats
    order = np.concatenate([np.nonzero(o.bobs_cam == c)[0] for c in range(s.n_cam)])
    for k, b in enumerate(order):
        row = rows[1 + s.n_cam + k]
        npts = int(o.bobs_ptr[b + 1] - o.bobs_ptr[b])
        assert [int(x) for x in row[:5]] == [int(o.bobs_cam[b]), int(o.bobs_frame[b]), int(o.bobs_board[b]), npts, 1]
        vals = np.array(row[5:], float).reshape(npts, 3)
        sl = slice(o.bobs_ptr[b], o.bobs_ptr[b + 1])
        assert np.array_equal(vals[:, 0].astype(int), o.pt_idx[sl] - o.bobs_board[b] * s.P)
        assert np.array_equal(vals[:, 1].astype(np.float32), o.u[sl]) and np.array_equal(vals[:, 2].astype(np.float32), o.v[sl])
    # the facade's writer, read by OpenCV: identical content (frame-major order inside a camera, like the reference)
    a, b_ = read_keypoints_with_opencv(kin, s.n_cam), read_keypoints_with_opencv(kout, s.n_cam)
    for c in range(s.n_cam):
        assert (a[c]["w"], a[c]["h"]) == (b_[c]["w"], b_[c]["h"])
        key = lambda d: sorted(zip(d["frames"], d["boards"], d["paths"], [tuple(p) for p in d["pts"]], [tuple(i) for i in d["ids"]]))
        assert key(a[c]) == key(b_[c])
        assert b_[c]["frames"] == sorted(b_[c]["frames"])


def test_keypoint_loader_rejects_inconsistent_file(io_main, tmp_path):
    bad = tmp_path / "bad.yml"
    bad.write_text('%YAML:1.0\n---\nnb_camera: 1\ncamera_0:\n   img_width: 10\n   img_height: 10\n   frame_idxs: [ 0, 1 ]\n'
                   '   frame_paths:\n      - "a"\n      - "b"\n   board_idxs: [ 0 ]\n   pts_2d:\n      - [ 1., 2. ]\n   charuco_idxs:\n      - [ 0 ]\n')
    boards = tmp_path / "boards.txt"
    boards.write_text("1\n4 4 0.1\n")
    r = subprocess.run([io_main, "keypoints", str(boards), str(bad), "None", str(tmp_path / "o.yml"), str(tmp_path / "d.txt")],
                       capture_output=True, text=True)
    assert r.returncode == 1 and "inconsistent" in r.stderr


def test_average_rotation(io_main, tmp_path):
    rng = np.random.default_rng(0)
    cases = []
    for k in range(40):
        quat = k % 2
        n = int(rng.integers(1, 9))
        base = Rot.random(random_state=int(rng.integers(1 << 30)))
        spread = 0.3 if k % 4 < 2 else 1e-3
        rv = (base * Rot.from_rotvec(spread * rng.standard_normal((n, 3)))).as_rotvec()
        cases.append((quat, rv))
    fin, fout = tmp_path / "in.txt", tmp_path / "out.txt"
    with open(fin, "w") as f:
        f.write(f"{len(cases)}\n")
        for quat, rv in cases:
            f.write(f"{quat} {len(rv)}\n" + "".join(" ".join(repr(float(x)) for x in r) + "\n" for r in rv))
    subprocess.check_call([io_main, "avgrot", str(fin), str(fout)])
    got = np.loadtxt(fout).reshape(-1, 3)
    for (quat, rv), g in zip(cases, got):
        if quat:
            want = Rot.from_rotvec(rv).mean()          # Markley et al.: largest eigenvector of sum q q^T
            assert np.abs(Rot.from_rotvec(g).as_matrix() - want.as_matrix()).max() < 1e-9
        else:
            assert np.allclose(g, np.median(rv, axis=0), atol=1e-15)


def test_group_pose_initialisation(io_main, tmp_path):
    rng = np.random.default_rng(1)
    cases = []
    for k in range(30):
        n_cam = int(rng.integers(2, 6))
        ref = 0
        cams = np.concatenate([Rot.random(n_cam, random_state=int(rng.integers(1 << 30))).as_rotvec() * 0.3, rng.normal(size=(n_cam, 3))], 1)
        cams[ref] = 0
        n_obj = int(rng.integers(1, 3))
        obj_gt = np.concatenate([Rot.random(n_obj, random_state=int(rng.integers(1 << 30))).as_rotvec(), rng.normal(size=(n_obj, 3)) + [0, 0, 3]], 1)
        obs = []
        for c in rng.permutation(n_cam)[:int(rng.integers(1, n_cam + 1))]:
            if k % 3 == 0 and c == ref:
                continue                              # no observation from the reference camera: the average is used
            for ob in range(n_obj):
                if rng.random() < 0.8:
                    rv, t = synth.compose(cams[c, :3], cams[c, 3:], obj_gt[ob, :3], obj_gt[ob, 3:])
                    rv = (Rot.from_rotvec(rv) * Rot.from_rotvec(0.01 * rng.standard_normal(3))).as_rotvec()
                    obs.append((int(c), ob, np.concatenate([rv, t + 0.01 * rng.standard_normal(3)])))
        if obs:
            cases.append((ref, cams, obs, k % 2))
    fin, fout = tmp_path / "in.txt", tmp_path / "out.txt"
    with open(fin, "w") as f:
        f.write(f"{len(cases)}\n")
        for ref, cams, obs, quat in cases:
            f.write(f"{ref} {len(cams)} {len(obs)} {quat}\n")
            for c in cams:
                f.write(" ".join(repr(float(x)) for x in c) + "\n")
            for c, ob, p in obs:
                f.write(f"{c} {ob} " + " ".join(repr(float(x)) for x in p) + "\n")
    subprocess.check_call([io_main, "grouppose", str(fin), str(fout)])
    lines = [l.split() for l in open(fout)]
    li = 0
    for ref, cams, obs, quat in cases:
        head = lines[li]; li += 1
        n_obj = int(head[0])
        got = {int(head[1 + 7 * j]): np.array(head[2 + 7 * j:8 + 7 * j], float) for j in range(n_obj)}
        group_pose = np.array([lines[li + j] for j in range(len(obs))], float); li += len(obs)
        # numpy restatement
        ing = []
        for c, ob, p in obs:
            Rc = Rot.from_rotvec(cams[c, :3])
            inv = (Rc.inv().as_rotvec(), -Rc.inv().apply(cams[c, 3:]))
            ing.append(np.concatenate(synth.compose(inv[0], inv[1], p[:3], p[3:])))
        ing = np.array(ing)
        for ob in sorted({ob for _, ob, _ in obs}):
            idx = [j for j, (_, o2, _) in enumerate(obs) if o2 == ob]
            refs = [j for j in idx if obs[j][0] == ref]
            if refs:
                want = ing[refs[-1]]
            else:
                r = Rot.from_rotvec(ing[idx, :3]).mean().as_rotvec() if quat else np.median(ing[idx, :3], axis=0)
                want = np.concatenate([r, ing[idx, 3:].mean(0)])
            assert np.abs(Rot.from_rotvec(got[ob][:3]).as_matrix() - Rot.from_rotvec(want[:3]).as_matrix()).max() < 1e-9
            assert np.abs(got[ob][3:] - want[3:]).max() < 1e-12
            for j in idx:                                 # every observation of the object carries the group pose
                assert np.allclose(group_pose[j], got[ob], atol=0, rtol=0)


CONFIG_TEMPLATE = """%YAML:1.0

######################################## Boards Parameters ###################################################
number_x_square: {nx}         # number of squares in the X direction
number_y_square: {ny}         # number of squares the Y direction
resolution_x: 500          # horizontal resolution in pixel
resolution_y: 500          # vertical resolution in pixel
length_square: 0.04        # parameters on the marker (can be kept as it is)
length_marker: 0.03        # parameters on the marker (can be kept as it is)
number_board: {n_board}            # number of boards used for calibration
boards_index: {boards_index}           # leave it empty [] if the board index are ranging from zero to number_board
square_size: {square}         # size of each square of the board in cm/mm/whatever you want

############# Boards Parameters for different board size (leave empty if all boards have the same size) #################
number_x_square_per_board: {nx_pb}
number_y_square_per_board: {ny_pb}
square_size_per_board: {sq_pb}

######################################## Camera Parameters ###################################################
distortion_model: {dist}             # 0:Brown (perspective) // 1: Kannala (fisheye)
distortion_per_camera : {dist_pc}         # specify the model per camera
number_camera: {n_cam}           # number of cameras in the rig to calibrate
refine_corner: 1           # activate or deactivate the corner refinement
min_perc_pts: 0.5           # min percentage of points visible to assume a good detection

cam_params_path: "{cam_params}"   # file with cameras intrinsics to initialize the intrinsic, write "None" if no initialization available
fix_intrinsic: {fix} #if 1 then the intrinsic parameters will not be estimated nor refined (initial value needed)

######################################## Images Parameters ###################################################
root_path: "../data/Images"
cam_prefix: "Cam_"
keypoints_path: "{keypoints}" # "None" #

######################################## Optimization Parameters #############################################
quaternion_averaging: 1    # use Quaternion Averaging or median for average rotation
ransac_threshold: {thr}       # RANSAC threshold in pixel (keep it high just to remove strong outliers)
number_iterations: {iters}    # Max number of iterations for the non linear refinement

######################################## Output Parameters ###################################################
save_path: "{save}"
save_detection: 0
save_reprojection: 0
camera_params_file_name: "" # "name.yml"
"""


def write_config(path, **kw):
    d = dict(nx=5, ny=5, n_board=1, boards_index="[]", square=0.192, nx_pb="[]", ny_pb="[]", sq_pb="[]", dist=0, dist_pc="[]",
             n_cam=2, cam_params="None", fix=0, keypoints="None", thr=10, iters=1000, save="/tmp/out")
    d.update(kw)
    with open(path, "w") as f:
        f.write(CONFIG_TEMPLATE.format(**d))


def test_config_constructor(io_main, tmp_path):
    """Calibration::Calibration(config) (McCalib.cpp:30-135, Board.cpp:22-72) on config files in the reference's layout
    (a '#' comment on almost every line, 'key : value' with a space before the colon, empty lists)."""
    cfg, out = str(tmp_path / "c.yml"), str(tmp_path / "o.txt")
    write_config(cfg, n_cam=3, n_board=2, nx=5, ny=4, square=0.192, dist=0, dist_pc="[0, 1, 0]", thr=7.5, iters=250, fix=1,
                 cam_params="/x/cams.yml", keypoints="/x/kp.yml")
    subprocess.check_call([io_main, "config", cfg, out])
    rows = open(out).read().split("\n")
    assert rows[0].split() == ["3", "2", "250", "7.5", "1", "1"]
    assert rows[1:4] == ["/x/cams.yml", "/x/kp.yml", "/tmp/out"]
    assert rows[4].split() == ["0", "1", "0"]
    for b in range(2):
        v = np.array(rows[5 + b].split(), float)
        assert int(v[0]) == 4 * 3                          # (nx - 1)(ny - 1) corners
        pts = v[1:].reshape(-1, 3)
        want = np.array([[x * np.float32(0.192), y * np.float32(0.192), 0] for y in range(3) for x in range(4)], np.float32)
        assert np.allclose(pts, want, rtol=1e-7)
    # boards of different sizes selected through boards_index
    write_config(cfg, n_cam=1, n_board=2, boards_index="[2, 0]", nx_pb="[5, 9, 7]", ny_pb="[5, 9, 4]", sq_pb="[0.1, 0.2, 0.3]")
    subprocess.check_call([io_main, "config", cfg, out])
    rows = open(out).read().split("\n")
    assert rows[0].split()[:2] == ["1", "2"]
    v0, v1 = (np.array(rows[5 + b].split(), float) for b in range(2))
    assert int(v0[0]) == 6 * 3 and int(v1[0]) == 4 * 4
    assert np.isclose(v0[1:].reshape(-1, 3)[1, 0], 0.3) and np.isclose(v1[1:].reshape(-1, 3)[1, 0], 0.1)


def test_yaml_reader_on_random_opencv_files(io_main, tmp_path):
    """Random nested documents emitted by cv::FileStorage (block and flow sequences and maps, strings that need
    quoting, negative / exponent numbers, long flow sequences that wrap, matrices) parse to the tree OpenCV reads back."""
    import json
    rng = np.random.default_rng(0)
    words = ["a", "Cam_001/000012.png", "with space", "x:y", "#hash", "1e5", "", "-", "[b]", "q\"uote", "tab\tin"]

    def gen(fs, depth, name):
        """writes a random value under `name`, returns the expected canonical tree"""
        kind = rng.integers(0, 6 if depth < 3 else 3)
        if kind == 0:
            v = int(rng.integers(-10**6, 10**6))
            fs.write(name, v)
            return str(v)
        if kind == 1:
            v = float(rng.choice([rng.normal() * 10.0 ** rng.integers(-12, 12), float(rng.integers(-5, 5)), 0.1, -2.5e-7]))
            fs.write(name, v)
            return ("f", v)
        if kind == 2:
            v = str(rng.choice(words))
            fs.write(name, v)
            return ("s", v)
        if kind == 3:
            n = int(rng.integers(0, 40))
            flow = bool(rng.integers(0, 2))
            fs.startWriteStruct(name, cv2.FileNode_SEQ | (cv2.FileNode_FLOW if flow else 0))
            out = []
            for _ in range(n):
                if flow:
                    v = float(np.float32(rng.normal() * 1000))
                    fs.write("", v)
                    out.append(("f", v))
                else:
                    out.append(gen(fs, depth + 1, ""))
            fs.endWriteStruct()
            return out
        if kind == 4:
            n = int(rng.integers(1, 6))
            fs.startWriteStruct(name, cv2.FileNode_MAP)
            out = {}
            for k in range(n):
                key = f"key_{depth}_{k}"
                out[key] = gen(fs, depth + 1, key)
            fs.endWriteStruct()
            return out
        r, c = int(rng.integers(1, 5)), int(rng.integers(1, 7))
        m = rng.normal(size=(r, c)) * 100
        fs.write(name, m)
        return {"rows": str(r), "cols": str(c), "dt": ("s", "d"), "data": [("f", float(x)) for x in m.ravel()]}

    def check(got, want, path):
        if isinstance(want, dict):
            assert isinstance(got, dict) and list(got.keys()) == list(want.keys()), path
            for k in want:
                check(got[k], want[k], path + "/" + k)
        elif isinstance(want, list):
            assert isinstance(got, list) and len(got) == len(want), (path, got, want)
            for i, (g, w) in enumerate(zip(got, want)):
                check(g, w, f"{path}[{i}]")
        elif isinstance(want, tuple) and want[0] == "f":
            assert float(got) == pytest.approx(want[1], rel=1e-15, abs=0), (path, got, want)
        elif isinstance(want, tuple):
            assert got == want[1], (path, got, want)
        else:
            assert got == want, (path, got, want)

    for trial in range(80):
        f, out = str(tmp_path / f"r{trial}.yml"), str(tmp_path / f"r{trial}.json")
        fs = cv2.FileStorage(f, cv2.FILE_STORAGE_WRITE)
        want = {}
        for k in range(int(rng.integers(1, 8))):
            want[f"top_{k}"] = gen(fs, 0, f"top_{k}")
        fs.release()
        r = subprocess.run([io_main, "yamldump", f, out], capture_output=True, text=True)
        assert r.returncode == 0, (r.stderr, open(f).read())
        check(json.loads(open(out).read(), strict=False), want, f"trial{trial}")


def test_keypoint_file_with_an_empty_camera(io_main, tmp_path):
    """A camera without any detection: OpenCV writes its block sequences as '[]' on a line of their own."""
    s = synth.make_scene("c2", n_frames=6)
    o = synth.generate_observations(s)
    # drop every observation of camera 1
    keep = np.nonzero(o.bobs_cam != 1)[0]
    o2 = synth.Scene()
    o2.frame_lo, o2.frame_hi = o.frame_lo, o.frame_hi
    sel = np.concatenate([np.arange(o.bobs_ptr[b], o.bobs_ptr[b + 1]) for b in keep])
    cnt = np.diff(o.bobs_ptr)[keep]
    o2.bobs_frame, o2.bobs_cam, o2.bobs_board = o.bobs_frame[keep], o.bobs_cam[keep], o.bobs_board[keep]
    o2.bobs_ptr = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
    o2.u, o2.v, o2.pt_idx = o.u[sel], o.v[sel], o.pt_idx[sel]
    kin, kout, dump, boards = (str(tmp_path / n) for n in ("kp.yml", "kp_out.yml", "dump.txt", "boards.txt"))
    write_keypoints_with_opencv(kin, s, o2)
    assert "[]" in open(kin).read()
    write_boards(boards, s)
    r = subprocess.run([io_main, "keypoints", boards, kin, "None", kout, dump], capture_output=True, text=True)
    # the reference asserts frame_idxs.size() != 0 for every camera (McCalib.cpp:201); the loader reports it instead of aborting
    assert r.returncode == 1 and "camera_1" in r.stderr


def test_reference_geometrytools_known_answers(io_main, tmp_path):
    """The four cases of the reference's own tests/test_geometrytools.cpp (:9-101), same inputs, same expected values,
    same 1 % BOOST_CHECK_CLOSE tolerance; the round-trip sweep additionally against cv2.Rodrigues."""
    fin, fout = tmp_path / "in.txt", tmp_path / "out.txt"
    lines = ["m 1 0 0 0 1 0 0 0 1",
             "m -0.9999 -0.1998 -0.3996 0.1998 0.6000 -0.8000 0.3996 -0.8000 -0.5999",
             "q 0.2809946 0.8377387 -0.4188693 -0.2092473"]
    sweep = [(ax, ay, az) for ax in range(-20, 21, 4) for ay in range(-45, 91, 5) for az in range(-270, -179, 6)]
    lines += [f"b {ax} {ay} {az}" for ax, ay, az in sweep]
    fin.write_text("\n".join(lines) + "\n")
    subprocess.check_call([io_main, "quat", str(fin), str(fout)])
    rows = [np.array(l.split(), float) for l in open(fout)]

    def close(a, b, pct=1.0):            # BOOST_CHECK_CLOSE: relative difference in percent w.r.t. both values
        a, b = np.asarray(a, float), np.asarray(b, float)
        return np.all((a == b) | (np.abs(a - b) <= pct / 100.0 * np.minimum(np.abs(a), np.abs(b))))

    assert close(rows[0], [0, 0, 0, 1])                                          # CheckRotationMatrixToQuaternionConversion1
    assert close(rows[1][1:], [0.8944, -0.4472, -0.2234]) and abs(rows[1][0]) < 1e-3   # ...Conversion2 (x = 0 in the reference)
    assert close(rows[2], [-0.7545152, 0.2955056, -0.5859892, 0.6460947, 0.4911810, -0.5842113,
                           0.1151891, -0.8194008, -0.5615281])                  # CheckQuaternionToRotationMatrixConversion
    for (ax, ay, az), r in zip(sweep, rows[3:]):                                 # CheckRotationMatrixToQuaternionAndBackConversion
        before, after = r[:9], r[9:]
        assert np.abs(after - before).max() < 1e-12
        Rcv, _ = cv2.Rodrigues(np.array([ax, ay, az], float))
        assert np.abs(before - Rcv.ravel()).max() < 1e-12


def test_pose_helpers_against_opencv(io_main, tmp_path):
    """rotVecToMat / matToRotVec / composePose / invertPose (the facade's stand-ins for cv::Rodrigues, RVecT2Proj,
    Proj2RT, invertRvecT) against cv2.Rodrigues and 4x4 matrix algebra, including the reference's own ProjToVecAllZeros
    case (tests/simple_unit_tests_example.cpp:10-18: the identity gives an all-zero vector) and rotations near pi."""
    rng = np.random.default_rng(0)
    lines, checks = ["v 1 0 0 0 1 0 0 0 1"], [("zero",)]
    for k in range(300):
        ax = rng.normal(size=3); ax /= np.linalg.norm(ax)
        ang = [rng.uniform(0, np.pi), np.pi - abs(rng.normal()) * 1e-6, abs(rng.normal()) * 1e-9][k % 3]
        rv = ax * ang
        R, _ = cv2.Rodrigues(rv)
        lines.append("r " + " ".join(repr(float(x)) for x in rv)); checks.append(("r", R))
        lines.append("v " + " ".join(repr(float(x)) for x in R.ravel())); checks.append(("v", R))
        a = np.concatenate([rv, rng.normal(size=3)])
        b = np.concatenate([Rot.random(random_state=int(rng.integers(1 << 30))).as_rotvec(), rng.normal(size=3)])
        lines.append("c " + " ".join(repr(float(x)) for x in np.concatenate([a, b]))); checks.append(("c", a, b))
        lines.append("i " + " ".join(repr(float(x)) for x in a)); checks.append(("i", a))
    fin, fout = tmp_path / "in.txt", tmp_path / "out.txt"
    fin.write_text("\n".join(lines) + "\n")
    subprocess.check_call([io_main, "pose", str(fin), str(fout)])
    rows = [np.array(l.split(), float) for l in open(fout)]

    def T(p):
        M = np.eye(4); M[:3, :3] = cv2.Rodrigues(np.asarray(p[:3], float))[0]; M[:3, 3] = p[3:]
        return M
    for row, chk in zip(rows, checks):
        if chk[0] == "zero":
            assert np.array_equal(row, np.zeros(3))
        elif chk[0] == "r":
            assert np.abs(row.reshape(3, 3) - chk[1]).max() < 1e-12
        elif chk[0] == "v":
            assert np.abs(cv2.Rodrigues(row)[0] - chk[1]).max() < 1e-9
        elif chk[0] == "c":
            assert np.abs(T(row) - T(chk[1]) @ T(chk[2])).max() < 1e-9
        else:
            assert np.abs(T(row) @ T(chk[1]) - np.eye(4)).max() < 1e-9
