"""Binary problem file (mcba_problem_save / mcba_problem_load, csrc/mcba_io.cpp): round trips, absent sections,
rejection of damaged files. Host code only - runs on the CPU box; the GPU replay is in tests/test_gpu_facade.py."""
import os

import numpy as np
import pytest

import mcba
from mcba import synth

pytestmark = pytest.mark.skipif(not os.path.exists(mcba.LIB_PATH), reason="libmcba.so not built")


def _same_problem(a, b):
    for k in ("stage", "n_cam", "n_pose", "n_board", "n_problems"):
        assert getattr(a, k) == getattr(b, k), k
    for k in ("u", "v", "pt_idx", "bobs_ptr", "bobs_cam", "bobs_pose", "bobs_board", "pts_xyz", "cam_model",
              "cam_is_ref", "board_is_ref", "cam_problem"):
        x, y = getattr(a, k), getattr(b, k)
        if x is None:
            assert y is None, k
        else:
            assert y is not None and x.dtype == y.dtype and np.array_equal(x, y), k


@pytest.mark.parametrize("cfg,stage", [("c1", 5), ("c2", 4), ("c2", 1), ("c2", 2), ("c2", 3)])
def test_round_trip_is_bit_exact(cfg, stage, tmp_path):
    prob, init, gt, s, o = synth.config_problem(cfg, stage, n_frames=25)
    path = str(tmp_path / "p.mcba")
    mcba.save_problem(path, prob, init)
    p2, x2 = mcba.load_problem(path)
    _same_problem(prob, p2)
    for k in ("cam_pose", "pose", "board_pose", "intrinsics"):
        assert np.array_equal(getattr(init, k), getattr(x2, k)), k       # doubles, bit for bit
    # a second save of the loaded problem gives the same bytes
    path2 = str(tmp_path / "q.mcba")
    mcba.save_problem(path2, p2, x2)
    assert open(path, "rb").read() == open(path2, "rb").read()
    assert os.path.getsize(path) % 8 == 0


def test_without_parameters_and_batched_problem(tmp_path):
    prob, params, gt = synth.batched_intrinsics_problem(n_cameras=6, n_frames=5)[:3]
    path = str(tmp_path / "b.mcba")
    mcba.save_problem(path, prob)                      # no parameter section
    p2, x2 = mcba.load_problem(path)
    _same_problem(prob, p2)
    assert x2 is None and p2.cam_problem is not None and p2.n_problems == 6


def test_empty_problem_round_trips(tmp_path):
    prob, init, gt, s, o = synth.config_problem("c1", 5, n_frames=4)
    empty = mcba.Problem(5, prob.u[:0], prob.v[:0], prob.pt_idx[:0], np.zeros(1, np.int64), prob.bobs_cam[:0],
                         prob.bobs_pose[:0], prob.bobs_board[:0], prob.pts_xyz, prob.cam_model, prob.n_pose, prob.n_board,
                         prob.cam_is_ref, prob.board_is_ref)
    path = str(tmp_path / "e.mcba")
    mcba.save_problem(path, empty, init)
    p2, x2 = mcba.load_problem(path)
    _same_problem(empty, p2)
    assert p2.n_obs == 0 and p2.n_bobs == 0


def test_damaged_files_are_rejected(tmp_path):
    prob, init, gt, s, o = synth.config_problem("c1", 4, n_frames=6)
    path = str(tmp_path / "p.mcba")
    mcba.save_problem(path, prob, init)
    raw = open(path, "rb").read()
    cases = {"magic": b"XXXXXXXX" + raw[8:], "truncated": raw[:-16], "trailing": raw + b"\0" * 8,
             "count": raw[:40] + (10 ** 12).to_bytes(8, "little") + raw[48:], "empty": b""}
    for name, data in cases.items():
        bad = str(tmp_path / f"{name}.mcba")
        open(bad, "wb").write(data)
        with pytest.raises(mcba.McbaError):
            mcba.load_problem(bad)
    with pytest.raises(mcba.McbaError):
        mcba.load_problem(str(tmp_path / "does_not_exist.mcba"))
    with pytest.raises(mcba.McbaError):
        mcba.save_problem(str(tmp_path / "no_such_dir" / "x.mcba"), prob, init)
