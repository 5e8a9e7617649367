#!/usr/bin/env python
"""bench.py — LM iterations/s and Mobs/s of the bundle-adjustment hot path on N B200s, beside the CPU oracle.

Workload (BASELINE.json configs[3], weak scaling): the 64-camera dome with 10 boards of 48 corners, joint
intrinsics + extrinsics + object refinement (stage 5, reduced system n = 1020), FRAMES_PER_GPU = 2500 frames
(about 6.2 M observations) per GPU, frame-sharded; N = 8 is the full 20,000-frame / ~50 M-observation case.
A "step" is one complete Levenberg-Marquardt iteration: Schur solve of the normal equations, back-substitution,
candidate-cost pass, and (the step being accepted) residual + Jacobian + J^T J evaluation at the new point.
The trust-region radius is pinned small (initial = max = 1e-2) and the tolerances are off, so that every one of
the W + K iterations is a genuine accepted step doing the full work instead of hovering at the converged minimum.

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA, libmcba through the C ABI)
  python bench.py --impl reference ...                     CPU arm: the oracle port of the reference's Ceres path
                                                           (the reference needs Ceres/Eigen/OpenCV and cannot be
                                                           built in this image), all host threads, bounded sample
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(REPO, "mc-calib_b200", "python"))

FRAMES_PER_GPU = 2500
STAGE = 5
BYTES_PER_OBS = {1: 277, 2: 229, 3: 229, 4: 325, 5: 469}       # materialising resid+Jac kernel
FLOPS_PER_OBS = {4: 500 + 756, 5: 500 + 1620}                  # resid+Jac ~0.5 kFLOP + J^T J / J^T r
FP64_PEAK_TFLOPS = 36.9    # measured on this pool's B200 with tools/peak_fp64.cu (DMMA m8n8k4), profiles/r01_peak_fp64.json
# DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) from the ncu --set full captures of this workload,
# profiles/r01c_ncu_full_step_kernels.txt and profiles/r01c_ncu_full_materialize.txt (6.18 M observations per launch)
NCU_TRAFFIC_FUSED = 77.2e6 + 361.6e6
NCU_TRAFFIC_MATERIALIZE = 77.6e6 + 2712.9e6


def bench_options(mcba, **kw):
    o = mcba.default_options(max_num_iterations=1 << 30, function_tolerance=-1.0, gradient_tolerance=-1.0,
                             parameter_tolerance=-1.0, initial_trust_region_radius=1e-2,
                             max_trust_region_radius=1e-2, min_trust_region_radius=0.0)
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def measured_peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d.get("hbm_gbs", 6650.0)), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "25",
                                       "-i", str(self.idx)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.strip().lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def pinned_problem(torch, prob):
    """Observation arrays in pinned host memory (what the e2e leg hands to mcba_create)."""
    keep = []
    for name in ("u", "v", "pt_idx", "bobs_ptr", "bobs_cam", "bobs_pose", "bobs_board"):
        if not hasattr(prob, name):
            continue
        t = torch.from_numpy(getattr(prob, name)).pin_memory()
        keep.append(t)
        setattr(prob, name, t.numpy())
    prob._pinned = keep
    return prob


def run_reference(args):
    """CPU arm: oracle port of the reference's Ceres path, all host threads, bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import mcba
    from mcba import synth
    sys.path.insert(0, os.path.join(REPO, "oracle"))
    import oracle_py
    oracle_py.load()
    threads = oracle_py.max_threads()
    frames = args.ref_frames
    s = synth.make_scene("c4", n_frames=FRAMES_PER_GPU * args.gpus)
    o = synth.generate_observations(s, 0, frames)
    prob, init, gt = synth.make_problem(s, o, STAGE)
    opt = bench_options(mcba)
    oracle_py.time_iteration(prob, init, opt, max(1, args.warmup), threads)   # warm-up (page faults, thread pool)
    sec, sec_eval, sec_lin = oracle_py.time_iteration(prob, init, opt, args.steps, threads)
    value = prob.n_obs / sec / 1e6
    sample = f"first {frames} frames of the dome ({prob.n_obs} observations), {args.steps} LM iterations at the initial point"
    line = {
        "impl": "reference", "metric": "LM bundle-adjustment throughput (full LM iteration: resid+Jac, normal equations, Schur, Cholesky, candidate cost)",
        "value": value, "unit": "Mobs/s", "lm_iter_per_s": 1.0 / sec, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus, prob.n_obs, frames, sample_note=sample),
        "cpu_baseline": {"value": value, "unit": "Mobs/s", "cores": threads, "kind": "port", "sample": sample,
                         "ms_eval": sec_eval * 1e3, "ms_linear": sec_lin * 1e3,
                         # the sample is small next to the 6.18 M-observation shard of the GPU arm: the evaluation scales with
                         # the observations, the dense n = 1020 reduced solve inside ms_linear does not. Upper bound of what the
                         # CPU would reach on the full shard if ALL of ms_linear were a fixed cost (it also holds the
                         # per-frame Schur elimination, which does scale):
                         "full_shard_upper_bound": full_shard_upper_bound(prob.n_obs, sec_eval, sec_lin, args.gpus)},
        "e2e": {"value": value, "unit": "Mobs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference = CPU oracle port (oracle/): MC-Calib + ceres-solver 2.2.0 cannot be compiled in this image",
    }
    print(json.dumps(line), flush=True)


OBS_PER_FRAME = 6182724 / 2500.0      # config 4: observations per frame of the dome (measured on the 2500-frame shard)


def full_shard_upper_bound(n_obs_sample, sec_eval, sec_lin, n_gpus):
    n_full = OBS_PER_FRAME * FRAMES_PER_GPU * n_gpus
    sec = sec_eval * n_full / n_obs_sample + sec_lin
    return {"value": n_full / sec / 1e6, "unit": "Mobs/s", "observations": int(n_full),
            "model": "ms_eval x (full observations / sample observations) + ms_linear unchanged"}


def workload_config(n_gpus, n_obs_global, frames_global, sample_note=None):
    c = {"workload": f"BASELINE configs[3]: 64-camera dome, 10 boards x 48 corners, stage 5 (intrinsics+extrinsics+object), "
                     f"{FRAMES_PER_GPU} frames/GPU weak-scaled (N=8 = the full 20,000-frame ~50M-observation case)",
         "n_cameras": 64, "n_boards": 10, "frames": int(frames_global), "observations": int(n_obs_global), "reduced_system_n": 1020,
         "parallelism": f"frame-sharded x{n_gpus}, NCCL allreduce of the reduced system",
         "lm_options": "radius pinned (initial=max=1e-2), tolerances off: every timed iteration is an accepted step",
         "l2": "per-step working set (observations 75 MB + per-board-observation J^T J tiles 0.67 GB + E^T F rows 0.25 GB per GPU) exceeds the 126 MB L2; no explicit flush"}
    if sample_note:
        c["sample"] = sample_note
    return c


def main():
    # libraries (NCCL, torch) may print to stdout: keep fd 1 for the single JSON line only
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real_stdout
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="mcba", choices=["mcba", "reference"])
    ap.add_argument("--frames-per-gpu", type=int, default=FRAMES_PER_GPU)
    ap.add_argument("--ref-frames", type=int, default=200, help="frames in the CPU arm's bounded sample")
    ap.add_argument("--cpu-frames", type=int, default=100, help="frames in the cpu_baseline sample of our arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)   # timing hygiene: never fewer than 3 untimed iterations (the JSON reports the count actually run)
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import mcba
    from mcba import synth
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libmcba has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    fpg = args.frames_per_gpu
    s = synth.make_scene("c4", n_frames=fpg * world)
    o = synth.generate_observations(s, rank * fpg, (rank + 1) * fpg)
    prob, init, gt = synth.make_problem(s, o, STAGE)
    prob = pinned_problem(torch, prob)
    n_obs_local = prob.n_obs

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def comm_init(h):
        if world == 1:
            return
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = ctypes.create_string_buffer(128)
            assert h.lib.mcba_nccl_unique_id(buf) == 0
            uid = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
        uid = uid.cuda()
        dist.broadcast(uid, 0)
        h.comm_init(rank, world, bytes(uid.cpu().numpy().tobytes()))

    def allmax(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    n_obs_global = int(allsum(float(n_obs_local)))
    opt = bench_options(mcba)

    # ---- device-resident arm: K timed LM iterations, inputs already in HBM --------------------------------
    h = mcba.Handle(prob, device=local_rank)
    comm_init(h)
    p = init.copy()
    h.solve_begin(p, opt)                       # upload parameters + iteration 0 (untimed)
    for _ in range(args.warmup):
        assert h.solve_step() == 1
    barrier()
    h.reset_kernel_stats()
    l0 = h.kernel_launches()
    sampler = ClockSampler(local_rank)
    sampler.start()
    h.timer_start()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        assert h.solve_step() == 1, "LM terminated inside the timed region"
    ms_dev = h.timer_stop()
    ms_wall = (time.perf_counter() - t0) * 1e3
    barrier()
    clocks = sampler.stop()
    launches = h.kernel_launches() - l0
    stats = h.kernel_stats()
    summ = h.solve_end(p)
    ms = allmax(ms_dev)
    value = n_obs_global * args.steps / (ms * 1e-3) / 1e6
    fused_n, fused_ms = stats["resid_jac_fused"]
    fused_avg_ms = fused_ms / max(1, fused_n)
    achieved_tf = FLOPS_PER_OBS[STAGE] * n_obs_local / (fused_avg_ms * 1e-3) / 1e12 if fused_n else None

    # ---- K1 materialising variant (the HBM-roofline kernel of the north star): a few iterations, same data -----
    optm = bench_options(mcba, materialize_jacobian=1)
    pm = init.copy()
    h.solve_begin(pm, optm)
    for _ in range(2):
        h.solve_step()
    h.reset_kernel_stats()
    for _ in range(3):
        h.solve_step()
    stm = h.kernel_stats()
    h.solve_end(pm)
    mat_n, mat_ms = stm["resid_jac_materialize"]
    mat_avg_ms = mat_ms / max(1, mat_n)
    hbm_peak, hbm_src = measured_peaks()
    mat_gbs = BYTES_PER_OBS[STAGE] * n_obs_local / (mat_avg_ms * 1e-3) / 1e9 if mat_n else None
    fromj_n, fromj_ms = stm["accumulate_from_jacobian"]
    h.close()

    # ---- end-to-end arm: host buffers -> mcba_create (pack + H2D) -> K LM iterations -> parameters back ----------
    e2e = None
    if not args.no_e2e:
        e2e_s = float("inf")
        for trial in range(3):          # first trial warms allocator / page tables; best of the next two
            pe = init.copy()
            barrier()
            t0 = time.perf_counter()
            he = mcba.Handle(prob, device=local_rank)
            torch.cuda.synchronize()
            t_create = time.perf_counter() - t0
            if world > 1:
                comm_init(he)           # NCCL communicator creation: once-per-process cost, untimed
                barrier()
            t1 = time.perf_counter()
            he.solve_begin(pe, opt)
            for _ in range(args.steps):
                assert he.solve_step() == 1
            he.solve_end(pe)
            torch.cuda.synchronize()
            t_solve = time.perf_counter() - t1
            he.close()
            t_trial = allmax(t_create + t_solve)
            if trial > 0 and t_trial < e2e_s:
                e2e_s, e2e_create, e2e_solve = t_trial, t_create, t_solve
        h2d = prob.n_obs * 12 + prob.n_bobs * (16 + 8) + (init.pose.nbytes + init.cam_pose.nbytes + init.board_pose.nbytes + init.intrinsics.nbytes) * 2
        d2h = init.pose.nbytes + init.cam_pose.nbytes + init.board_pose.nbytes + init.intrinsics.nbytes + (args.steps + 1) * 2 * 12 * 8
        e2e = {"value": n_obs_global * args.steps / e2e_s / 1e6, "unit": "Mobs/s", "ms_total": e2e_s * 1e3,
               "ms_create": e2e_create * 1e3, "ms_solve": e2e_solve * 1e3,
               "h2d_bytes_per_step": int(h2d / args.steps), "d2h_bytes_per_step": int(d2h / args.steps),
               "includes": "mcba_create from pinned host arrays (validation, pair lists, H2D), iteration 0, K LM iterations "
                           "(per-iteration scalar read-backs), parameter download; NCCL communicator creation excluded"}

    # ---- pose initialisation that feeds the BA: every board observation of this rank's shard in
    # one mcba_pnp_ransac call (P3P RANSAC + pose refinement), host buffers in, poses out; shards are independent ----
    pose_init, pi_local = None, None
    try:      # local work only inside the try: a rank that fails here must still take part in the collectives below
        from mcba import pnp
        p1, _, gt1 = synth.make_problem(s, o, 1)
        P1 = pinned_problem(torch, pnp.PnpProblem(p1, s.intr_gt))
        popt = pnp.default_options(threshold_px=10.0, max_iterations=1000, seed=1)
        pnp.ransac(P1, popt, device=local_rank)
        runs = [pnp.ransac(P1, popt, device=local_rank) for _ in range(3)]
        pi_local = (min(r.kernel_ms for r in runs), min(r.total_ms for r in runs), float(P1.n_bobs),
                    int(runs[-1].kernel_launches), int(runs[-1].n_hypotheses.sum()))
    except Exception as e:   # the BA line must not depend on the extra leg
        pose_init = {"error": repr(e)}
    k_ms = allmax(pi_local[0] if pi_local else -1.0)
    t_ms = allmax(pi_local[1] if pi_local else -1.0)
    nb_global = int(allsum(pi_local[2] if pi_local else 0.0))
    n_ok = int(allsum(1.0 if pi_local else 0.0))
    if pi_local and n_ok == world:
        pose_init = {"board_observations": nb_global, "kernel_ms": k_ms, "total_ms": t_ms,
                     "bobs_per_s_kernels": nb_global / (k_ms * 1e-3), "bobs_per_s_e2e": nb_global / (t_ms * 1e-3),
                     "kernel_launches": pi_local[3], "draws_rank0": pi_local[4],
                     "note": "threshold 10 px, <= 1000 draws (the reference's YAML defaults); e2e = pinned host arrays -> device, "
                             "4 kernels, poses / inlier masks back; every rank poses its own shard, no collective; "
                             "OpenCV single-thread loop of the reference: tools/bench_pnp.py"}
    elif pose_init is None:
        pose_init = {"error": "the pose_init leg failed on another rank"}

    # ---- CPU baseline (rank 0, N = 1 only): oracle port, bounded sample ---------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(REPO, "oracle"))
        import oracle_py
        oracle_py.load()
        threads = oracle_py.max_threads()
        oc = synth.generate_observations(s, 0, args.cpu_frames)
        pc, ic, _ = synth.make_problem(s, oc, STAGE)
        oracle_py.time_iteration(pc, ic, opt, 1, threads)
        sec, sec_eval, sec_lin = oracle_py.time_iteration(pc, ic, opt, 3, threads)
        sec1, _, _ = oracle_py.time_iteration(pc, ic, opt, 1, 1)
        cpu = {"value": pc.n_obs / sec / 1e6, "unit": "Mobs/s", "cores": threads, "kind": "port",
               "sample": f"first {args.cpu_frames} frames of the same dome ({pc.n_obs} observations), 3 LM iterations at the initial point, OpenMP over all host threads",
               "resid_jac_mobs_per_s": pc.n_obs / sec_eval / 1e6, "single_thread_value": pc.n_obs / sec1 / 1e6, "single_thread_note": "the reference as shipped runs Ceres with num_threads = 1",
               "ms_eval": sec_eval * 1e3, "ms_linear": sec_lin * 1e3,
               "full_shard_upper_bound": full_shard_upper_bound(pc.n_obs, sec_eval, sec_lin, 1)}

    if rank == 0:
        line = {
            "metric": "LM bundle-adjustment throughput (full LM iteration: resid+Jac, normal equations, Schur, Cholesky, candidate cost)",
            "value": value, "unit": "Mobs/s", "lm_iter_per_s": args.steps / (ms * 1e-3), "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "ms_per_step_wall": ms_wall / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(world, n_obs_global, fpg * world),
            "clocks": clocks, "gpu_launches": int(launches), "e2e": e2e,
            "roofline": {"kernel": "k_observations<5,MODE_FUSED> (resid + Jacobian + J^T J/J^T r, DMMA)", "bound": "tensor",
                         "achieved": achieved_tf, "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s",
                         "frac": (achieved_tf / FP64_PEAK_TFLOPS) if achieved_tf else None,
                         "traffic": NCU_TRAFFIC_FUSED if fpg == FRAMES_PER_GPU else None,
                         "ms_per_launch": fused_avg_ms, "launches": fused_n,
                         "peak_source": "fp64 DMMA m8n8k4 peak measured with tools/peak_fp64.cu on this pool (profiles/r01_peak_fp64.json); MEASURED_PEAKS.json has no fp64 figure",
                         "algorithmic": f"{FLOPS_PER_OBS[STAGE]} FLOP/obs x {n_obs_local} obs per launch (SURVEY.md 8d)"},
            "roofline_hbm": {"kernel": "k_observations<5,MODE_MATERIALIZE> (resid + Jacobian written to HBM)", "bound": "hbm",
                             "achieved": mat_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": (mat_gbs / hbm_peak) if mat_gbs else None,
                             "traffic": NCU_TRAFFIC_MATERIALIZE if fpg == FRAMES_PER_GPU else None, "ms_per_launch": mat_avg_ms, "launches": mat_n, "peak_source": hbm_src,
                             "algorithmic": f"{BYTES_PER_OBS[STAGE]} B/obs x {n_obs_local} obs per launch (SURVEY.md 8d)",
                             "readback_ms_per_launch": fromj_ms / max(1, fromj_n)},
            "resid_jac": {   # the second half of BASELINE's metric: throughput of the residual+Jacobian pass alone
                "fused_mobs_per_s": (n_obs_global / (fused_avg_ms * 1e-3) / 1e6) if fused_n else None,
                "materialize_mobs_per_s": (n_obs_global / (mat_avg_ms * 1e-3) / 1e6) if mat_n else None,
                "note": "all ranks run the pass concurrently on their shards: global observations / rank-0 kernel time"},
            "kernel_ms_per_step": {k: v[1] / args.steps for k, v in stats.items() if v[0]},
            "cpu_baseline": cpu,
            "pose_init": pose_init,
            "final": {"rms_px": summ.rms_px, "cost": summ.final_cost, "iterations": summ.num_iterations},
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
