"""Synthetic generator and the frame sharding used for N > 1: shards reproduce exactly the
frames of the full problem, and partial sums over shards equal the full-problem sums (the property the NCCL
allreduce of the reduced system relies on). The 2-rank test runs torch.distributed with the gloo backend on CPU."""
import os
import sys

import numpy as np
import pytest

import mcba
from mcba import synth
from util import dense_jacobian, layout


def test_generator_is_deterministic_and_well_formed():
    a = synth.config_problem("c2", 5, n_frames=60)[0]
    b = synth.config_problem("c2", 5, n_frames=60)[0]
    assert np.array_equal(a.u, b.u) and np.array_equal(a.pt_idx, b.pt_idx) and np.array_equal(a.bobs_ptr, b.bobs_ptr)
    assert a.bobs_ptr[0] == 0 and a.bobs_ptr[-1] == a.n_obs and np.all(np.diff(a.bobs_ptr) >= 8)
    assert np.all(np.diff(a.bobs_pose) >= 0)
    assert set(np.unique(a.cam_model)) == {0, 1} and a.n_board == 3
    assert a.u.dtype == np.float32 and a.pts_xyz.dtype == np.float32


def test_shards_reproduce_the_full_problem():
    s = synth.make_scene("c4", n_frames=500)
    full = synth.generate_observations(s, 0, 500)
    lo = synth.generate_observations(s, 0, 250)
    hi = synth.generate_observations(s, 250, 500)
    assert np.array_equal(np.concatenate([lo.u, hi.u]), full.u)
    assert np.array_equal(np.concatenate([lo.pt_idx, hi.pt_idx]), full.pt_idx)
    assert np.array_equal(np.concatenate([lo.bobs_frame, hi.bobs_frame]), full.bobs_frame)
    assert np.allclose(np.concatenate([lo.obj_init, hi.obj_init]), full.obj_init)
    # a shard that does not start on a generator chunk boundary
    mid = synth.generate_observations(s, 100, 400)
    sel = (full.bobs_frame >= 100) & (full.bobs_frame < 400)
    assert np.array_equal(mid.bobs_cam, full.bobs_cam[sel])


def test_config_shapes():
    for name, ncam, nboard in (("c1", 2, 1), ("c2", 4, 3), ("c3", 16, 5), ("c4", 64, 10)):
        s = synth.make_scene(name, n_frames=4)
        assert (s.n_cam, s.n_board) == (ncam, nboard)
        assert np.all(s.cam_gt[0] == 0) and np.all(s.board_gt[0] == 0)     # reference camera / board
    p, init, gt = synth.batched_intrinsics_problem(n_cameras=3, n_frames=5)
    assert p.n_problems == 3 and p.stage == 1


def _rank_partial(rank, world, q):
    """Each rank: its frame shard -> partial F^T F, F^T r and cost from the oracle's Jacobian."""
    import torch.distributed as dist
    import torch
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import oracle_py
    frames = 24
    per = frames // world
    s = synth.make_scene("c2", n_frames=frames)
    o = synth.generate_observations(s, rank * per, (rank + 1) * per)
    prob, init, gt = synth.make_problem(s, o, 5)
    _, cost, res, jac, _ = oracle_py.eval(prob, init)
    L = layout(prob)
    J = dense_jacobian(prob, jac)[:, L["n_e"]:]
    FF, gf = J.T @ J, J.T @ res.reshape(-1)
    t = torch.from_numpy(np.concatenate([FF.ravel(), gf, [cost, prob.n_obs]]))
    dist.all_reduce(t)
    if rank == 0:
        q.put(t.numpy())
    dist.destroy_process_group()


def test_two_rank_partial_sums_equal_full_problem(oracle):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_rank_partial, args=(r, 2, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    s = synth.make_scene("c2", n_frames=24)
    o = synth.generate_observations(s)
    prob, init, gt = synth.make_problem(s, o, 5)
    _, cost, res, jac, _ = oracle.eval(prob, init)
    L = layout(prob)
    J = dense_jacobian(prob, jac)[:, L["n_e"]:]
    want = np.concatenate([(J.T @ J).ravel(), J.T @ res.reshape(-1), [cost, prob.n_obs]])
    assert np.abs(got - want).max() <= 1e-9 * np.abs(want).max()
