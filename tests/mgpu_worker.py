"""Worker for test_gpu_multi.py, launched with torch.distributed.run: frame-sharded solve of config 2 (hybrid
Brown + Kannala rig) on WORLD_SIZE GPUs; rank 0 checks the result against the CPU oracle on the full problem."""
import ctypes
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "mc-calib_b200", "python"))
sys.path.insert(0, os.path.join(REPO, "oracle"))
import mcba  # noqa: E402
from mcba import synth  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    frames = 240
    per = frames // world
    s = synth.make_scene("c2", n_frames=frames)
    for stage in (4, 5):
        o = synth.generate_observations(s, rank * per, (rank + 1) * per)
        prob, init, gt = synth.make_problem(s, o, stage)
        h = mcba.Handle(prob, device=local)
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = ctypes.create_string_buffer(128)
            assert h.lib.mcba_nccl_unique_id(buf) == 0
            uid = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
        uid = uid.cuda()
        dist.broadcast(uid, 0)
        h.comm_init(rank, world, uid.cpu().numpy().tobytes())
        p = init.copy()
        summ, trace = h.solve(p)
        h.close()
        # gather the sharded e-blocks
        poses = [torch.zeros(per * 6, dtype=torch.float64, device="cuda") for _ in range(world)]
        dist.all_gather(poses, torch.from_numpy(p.pose.reshape(-1)).cuda())
        if rank == 0:
            import oracle_py
            of = synth.generate_observations(s)
            pf, initf, _ = synth.make_problem(s, of, stage)
            po = initf.copy()
            rc, so, to = oracle_py.solve(pf, po)
            assert rc == 0
            got = [(r.iteration, r.step_is_valid, r.step_is_successful) for r in trace]
            want = [(r.iteration, r.step_is_valid, r.step_is_successful) for r in to]
            assert got == want, (got, want)
            assert summ.termination == so.termination
            assert abs(summ.rms_px - so.rms_px) <= 1e-6, (summ.rms_px, so.rms_px)
            for a, b in zip(trace, to):
                assert abs(a.cost - b.cost) <= 1e-9 * abs(b.cost)
            allpose = torch.cat(poses).cpu().numpy().reshape(-1, 6)
            assert np.abs(allpose - po.pose).max() <= 1e-6 * max(1.0, np.abs(po.pose).max())
            assert np.abs(p.cam_pose - po.cam_pose).max() <= 1e-6
            assert np.abs(p.board_pose - po.board_pose).max() <= 1e-6
            assert np.abs(p.intrinsics - po.intrinsics).max() <= 1e-6 * np.abs(po.intrinsics).max()
            print(f"MGPU_OK stage {stage} world {world} iterations {summ.num_iterations} rms {summ.rms_px:.6f}", flush=True)
    # stage 4 then stage 5 on ONE handle (mcba_set_stage re-exchanges the peer-memory handles)
    o = synth.generate_observations(s, rank * per, (rank + 1) * per)
    prob, init, gt = synth.make_problem(s, o, 4)
    h = mcba.Handle(prob, device=local)
    uid = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        buf = ctypes.create_string_buffer(128)
        assert h.lib.mcba_nccl_unique_id(buf) == 0
        uid = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
    uid = uid.cuda()
    dist.broadcast(uid, 0)
    h.comm_init(rank, world, uid.cpu().numpy().tobytes())
    p = init.copy()
    p.intrinsics[:] = s.intr_init
    s4, _ = h.solve(p)
    h.set_stage(5)
    s5, _ = h.solve(p)
    h.close()
    if rank == 0:
        import oracle_py
        of = synth.generate_observations(s)
        pf, initf, _ = synth.make_problem(s, of, 4)
        po = initf.copy()
        po.intrinsics[:] = s.intr_init
        rc, o4, _ = oracle_py.solve(pf, po)
        rc, o5, _ = oracle_py.solve(pf.with_stage(5), po)
        assert (s4.num_iterations, s5.num_iterations) == (o4.num_iterations, o5.num_iterations)
        assert abs(s5.rms_px - o5.rms_px) <= 1e-6
        assert np.abs(p.intrinsics - po.intrinsics).max() <= 1e-6 * np.abs(po.intrinsics).max()
        print(f"MGPU_OK set_stage world {world} iterations {s4.num_iterations}+{s5.num_iterations}", flush=True)
    # the non-accept branches of the trust-region loop on 2 ranks (tests/lm_cases.py): every rank must take the same
    # decision from the gathered scalars - rejected steps, invalid steps (recovering / FAILURE), NO_CONVERGENCE, radius exit
    sys.path.insert(0, os.path.join(REPO, "tests"))
    import lm_cases
    for name in ("reject_c2_s5", "reject_c2_s4", "invalid_recover_c2_s5", "invalid_failure_c2_s5", "no_convergence_c2_s5",
                 "min_radius_c2_s5", "parameter_tolerance_c2_s5"):
        prob, p0, opt, _ = lm_cases.make_case(name)
        sub, p, (lo, hi) = synth.shard_problem(prob, p0, rank, world)
        h = mcba.Handle(sub, device=local)
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = ctypes.create_string_buffer(128)
            assert h.lib.mcba_nccl_unique_id(buf) == 0
            uid = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
        uid = uid.cuda()
        dist.broadcast(uid, 0)
        h.comm_init(rank, world, uid.cpu().numpy().tobytes())
        summ, trace = h.solve(p, opt)
        h.close()
        n_max = (prob.n_pose + world - 1) // world
        mine = torch.zeros(n_max * 6, dtype=torch.float64, device="cuda")
        mine[:(hi - lo) * 6] = torch.from_numpy(p.pose.reshape(-1)).cuda()
        poses = [torch.zeros(n_max * 6, dtype=torch.float64, device="cuda") for _ in range(world)]
        dist.all_gather(poses, mine)
        if rank == 0:
            import oracle_py
            po = p0.copy()
            rc, so, to = oracle_py.solve(prob, po, opt)
            assert rc == 0
            full = p0.copy()
            full.cam_pose[:], full.board_pose[:], full.intrinsics[:] = p.cam_pose, p.board_pose, p.intrinsics
            for r in range(world):
                a, b = (prob.n_pose * r) // world, (prob.n_pose * (r + 1)) // world
                full.pose[a:b] = poses[r][:(b - a) * 6].cpu().numpy().reshape(-1, 6)
            lm_cases.check_expectation(name, summ, trace)
            lm_cases.assert_same_run(name, (full, summ, trace), (po, so, to))
            print(f"MGPU_OK branch {name} world {world} sequence {lm_cases.sequence(trace)} termination {summ.termination}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
