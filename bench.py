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
                                                           built in this image) on ONE GPU's shard of the same
                                                           workload (2500 frames, 6.18 M observations; at N = 1 that
                                                           is the whole workload), all host cores, K iterations

Besides the headline line's `value` / `e2e` / `roofline` / `cpu_baseline`, the line carries
  parity   a real default-options stage 4 + stage 5 solve of the N-GPU problem (reference order,
           apps/calibrate/src/calibrate.cpp:60-64) against the committed oracle fixture tests/golden/full_c4*.json
  extra    config 3 in full on one GPU, config 5 batched, config 4 strong-scaled (20 000 frames fixed)
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(REPO, "mc-calib_b200", "python"))
GOLD = os.path.join(REPO, "tests", "golden")

FRAMES_PER_GPU = 2500
FULL_FRAMES = 20000
STAGE = 5
BYTES_PER_OBS = {1: 277, 2: 229, 3: 229, 4: 325, 5: 469}       # materialising resid+Jac kernel
FLOPS_PER_OBS = {4: 500 + 756, 5: 500 + 1620}                  # resid+Jac ~0.5 kFLOP + J^T J / J^T r
FP64_PEAK_TFLOPS = 36.9    # measured on this pool's B200 with tools/peak_fp64.cu (DMMA m8n8k4), profiles/r01_peak_fp64.json
# DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the two K1 variants on the 6.18 M-observation
# shard, from ncu --set full captures of this same command; NOT measured inside the run (ncu cannot run under the timer)
NCU_TRAFFIC = {"fused": (78.0e6 + 368.4e6, "profiles/r02c_ncu_full_step_kernels.txt"),
               "materialize": (77.6e6 + 2712.9e6, "profiles/r01c_ncu_full_materialize.txt")}
METRIC = "LM bundle-adjustment throughput (full LM iteration: resid+Jac, normal equations, Schur, Cholesky, candidate cost)"


def bench_options(mcba, **kw):
    o = mcba.default_options(max_num_iterations=1 << 30, function_tolerance=-1.0, gradient_tolerance=-1.0,
                             parameter_tolerance=-1.0, initial_trust_region_radius=1e-2,
                             max_trust_region_radius=1e-2, min_trust_region_radius=0.0)
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def measured_peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d.get("hbm_gbs", 6650.0)), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def golden_c4(frames):
    """Committed oracle fixture of the stage 4 + 5 solve of the first `frames` frames of config 4, or None."""
    p = os.path.join(GOLD, "full_c4.json" if frames == FULL_FRAMES else f"full_c4_f{frames}.json")
    return json.load(open(p)) if os.path.exists(p) else None


class ClockSampler:
    """SM clock and throttle reasons while the timed region runs: NVML polled from a thread (initialised before the
    region starts, one sample every 25 ms on rank 0 only - an NVML query takes ~0.4 ms and contends with CUDA launches in the
    driver: taken from the timed loop itself six samples cost 4.6 % of the step, and a 4 ms poll on all eight ranks 15 % - the
    timed region is ~100 ms, and `nvidia-smi -lms` needs longer than that
    to produce its first line on an 8-GPU box); nvidia-smi as the fallback when the NVML binding is missing."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index, uuid=None):
        self.idx = gpu_index
        self.p = self.f = self.thread = None
        self.sm, self.reason_bits, self.mx = [], 0, None
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            h = None
            if uuid:
                try:
                    h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(uuid)) if not str(uuid).startswith("GPU-") else str(uuid))
                except Exception:
                    h = None
            self.h = h if h is not None else pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _poll(self):
        n = self.nvml
        get_reasons = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(n, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self.stop_flag:
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)))
                self.reason_bits |= int(get_reasons(self.h))
            except Exception:
                pass
            time.sleep(0.025)

    def start(self):
        if self.nvml is not None:
            import threading
            self.stop_flag = False
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        try:
            self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "25",
                                       "-i", str(self.idx)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.thread is not None:
            self.stop_flag = True
            self.thread.join(timeout=2)
            reasons = sorted(name for bit, name in self.REASONS.items() if self.reason_bits & bit)
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.mx, "reasons": reasons,
                    "samples": len(self.sm), "source": "nvml"}
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
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


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


def workload_config(n_gpus, n_obs_global, frames_global, schur_reduce=None):
    """The same dict on both arms (the driver compares them): it names the workload, not what a run sampled of it."""
    return {"workload": f"BASELINE configs[3]: 64-camera dome, 10 boards x 48 corners, stage 5 (intrinsics+extrinsics+object), "
                        f"{FRAMES_PER_GPU} frames/GPU weak-scaled (N=8 = the full 20,000-frame ~50M-observation case)",
            "n_cameras": 64, "n_boards": 10, "frames": int(frames_global), "observations": int(n_obs_global), "reduced_system_n": 1020,
            "parallelism": f"frame-sharded x{n_gpus}; reduced system summed across ranks in the Schur kernel's epilogue, pair records and "
                           "per-phase scalars exchanged by their own kernels, all over NVLink peer memory (NCCL as rendezvous and fallback)" if n_gpus > 1 else "single GPU",
            "lm_options": "radius pinned (initial=max=1e-2), tolerances off: every timed iteration is an accepted step",
            "l2": "per-step working set (observations 75 MB + per-board-observation J^T J tiles 0.42 GB x 2 + E^T F rows 0.25 GB per GPU) exceeds the 126 MB L2; no explicit flush"}


def global_observations(n_gpus, frames_per_gpu):
    """Observation count of the N-GPU problem without generating it (reference arm): from the committed fixtures."""
    g = golden_c4(frames_per_gpu * n_gpus)
    return int(g["observations"]) if g else None


def run_reference(args):
    """CPU arm: oracle port of the reference's Ceres path on ONE GPU's shard of the workload (rank 0's 2500 frames; the
    whole workload at N = 1), K timed LM iterations with all host cores (thread count set explicitly: torchrun exports
    OMP_NUM_THREADS=1), plus one single-thread iteration - the reference as shipped runs Ceres with num_threads = 1."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import mcba
    from mcba import synth
    sys.path.insert(0, os.path.join(REPO, "oracle"))
    import oracle_py
    oracle_py.load()
    threads = host_cores()
    fpg = args.frames_per_gpu
    s = synth.make_scene("c4", n_frames=fpg * args.gpus)
    o = synth.generate_observations(s, 0, fpg, workers=min(8, threads))
    prob, init, gt = synth.make_problem(s, o, STAGE)
    opt = bench_options(mcba)
    oracle_py.time_iteration(prob, init, opt, max(1, min(args.warmup, 2)), threads)   # warm-up (page faults, thread pool)
    sec, sec_eval, sec_lin = oracle_py.time_iteration(prob, init, opt, args.steps, threads)
    sec1, _, _ = oracle_py.time_iteration(prob, init, opt, 1, 1)
    value = prob.n_obs / sec / 1e6
    n_glob = global_observations(args.gpus, fpg) if args.gpus > 1 else prob.n_obs
    sample = (f"rank 0's shard of the workload: frames 0..{fpg} of the dome, {prob.n_obs} observations "
              f"({'the whole workload' if args.gpus == 1 else f'1/{args.gpus} of it: the CPU has one socket, not N'}), "
              f"{args.steps} full LM iterations at the initial point, {threads} OpenMP threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Mobs/s", "lm_iter_per_s": 1.0 / sec, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus, n_glob if n_glob else prob.n_obs * args.gpus, fpg * args.gpus),
        "cpu_baseline": {"value": value, "unit": "Mobs/s", "cores": threads, "kind": "port", "sample": sample,
                         "ms_eval": sec_eval * 1e3, "ms_linear": sec_lin * 1e3, "observations_timed": int(prob.n_obs),
                         "single_thread_value": prob.n_obs / sec1 / 1e6,
                         "single_thread_note": "one iteration with 1 thread: the reference as shipped runs Ceres with num_threads = 1"},
        "e2e": {"value": value, "unit": "Mobs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference = CPU oracle port (oracle/): MC-Calib + ceres-solver 2.2.0 cannot be compiled in this image",
    }
    print(json.dumps(line), flush=True)


def compare_with_golden(gold, summaries, traces, params):
    """parity object: our stage 4 + 5 solve against the oracle fixture (iterations, accept sequence, RMS, f-blocks)."""
    out = {"fixture": gold.get("generator", "tests/golden") + f" ({gold['config']}, {gold['frames']} frames, {gold['observations']} observations)",
           "stages": []}
    ok = True
    for st, summ, tr in zip(gold["stages"], summaries, traces):
        seq = [(r.iteration, r.step_is_valid, r.step_is_successful) for r in tr]
        want = [(r["iteration"], r["valid"], r["successful"]) for r in st["trace"]]
        same = seq == want and summ.termination == st["termination"]
        cost_rel = max((abs(a.cost - b["cost"]) / abs(b["cost"]) for a, b in zip(tr, st["trace"])), default=0.0)
        d = {"stage": st["stage"], "iterations": summ.num_iterations, "oracle_iterations": st["num_iterations"],
             "same_accept_reject_sequence_and_termination": bool(same), "rms_px": summ.rms_px,
             "rms_abs_diff_px": abs(summ.rms_px - st["rms_px"]), "max_rel_cost_diff_over_iterations": cost_rel}
        ok = ok and same and d["rms_abs_diff_px"] <= 1e-6
        out["stages"].append(d)
    last = gold["stages"][-1]
    rel = {}
    for name in ("cam_pose", "board_pose", "intrinsics"):
        a, b = getattr(params, name), np.array(last[name])
        rel[name] = float(np.abs(a - b).max() / max(1e-300, np.abs(b).max()))
    out["max_rel_diff_final"] = rel
    out["pass"] = bool(ok and max(rel.values()) <= 1e-6)
    out["bar"] = "identical iteration / accept-reject sequence and termination, RMS within 1e-6 px, f-blocks within 1e-6 relative (BASELINE.json north_star)"
    return out


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
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the config 3 / config 5 / strong-scaling lines")
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
    try:
        gpu_uuid = str(torch.cuda.get_device_properties(local_rank).uuid)     # NVML handle of THIS device whatever CUDA_VISIBLE_DEVICES says
    except Exception:
        gpu_uuid = None
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    workers = max(1, min(8, host_cores() // world))
    fpg = args.frames_per_gpu
    s = synth.make_scene("c4", n_frames=fpg * world)
    o = synth.generate_observations(s, rank * fpg, (rank + 1) * fpg, workers=workers)
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

    def timed_iterations(h, p0, options, steps, warmup, sample_clocks=False):
        """W untimed + K timed LM iterations on a handle; device time (CUDA events on the library stream), max over ranks."""
        p = p0.copy()
        h.solve_begin(p, options)                       # upload parameters + iteration 0 (untimed)
        for _ in range(warmup):
            assert h.solve_step() == 1
        # rank 0 only: eight processes polling NVML every few ms slowed the N = 8 step by 15 %. Created (NVML initialisation:
        # tens of ms) BEFORE the barrier: everything between the barrier and timer_start delays this rank alone, and the
        # other ranks - whose timers are already running - wait for it in the first exchange (the max over ranks then
        # charges that delay to the timed region: 2.77 instead of 2.13 ms per step in one N = 8 run)
        sampler = ClockSampler(local_rank, gpu_uuid) if (sample_clocks and rank == 0) else None
        h.reset_kernel_stats()
        l0 = h.kernel_launches()
        barrier()
        if sampler:
            sampler.start()
        h.timer_start()
        t0 = time.perf_counter()
        for _ in range(steps):
            assert h.solve_step() == 1, "LM terminated inside the timed region"
        ms_dev = h.timer_stop()
        ms_wall = (time.perf_counter() - t0) * 1e3
        barrier()
        clocks = sampler.stop() if sampler else None
        launches = h.kernel_launches() - l0
        stats = h.kernel_stats()
        summ = h.solve_end(p)
        return allmax(ms_dev), ms_wall, clocks, launches, stats, summ

    n_obs_global = int(allsum(float(n_obs_local)))
    opt = bench_options(mcba)

    # ---- device-resident arm: K timed LM iterations, inputs already in HBM --------------------------------
    h = mcba.Handle(prob, device=local_rank)
    comm_init(h)
    schur_reduce = h.schur_reduce_mode()
    ms, ms_wall, clocks, launches, stats, summ = timed_iterations(h, init, opt, args.steps, args.warmup, sample_clocks=True)
    value = n_obs_global * args.steps / (ms * 1e-3) / 1e6
    ms_direct = None
    graphed = world == 1 or schur_reduce == "p2p"
    if graphed:
        # the iteration is replayed as CUDA graphs: no per-kernel events there. The per-kernel breakdown (and the
        # roofline's per-launch time of the fused kernel) comes from a second handle that launches kernel by kernel.
        os.environ["MCBA_GRAPH"] = "0"
        hd = mcba.Handle(prob, device=local_rank)
        comm_init(hd)
        ms_direct, _, _, _, stats, _ = timed_iterations(hd, init, opt, args.steps, args.warmup)
        hd.close()
        del os.environ["MCBA_GRAPH"]
    fused_n, fused_ms = stats["resid_jac_fused"]
    fused_avg_ms = fused_ms / max(1, fused_n)
    achieved_tf = FLOPS_PER_OBS[STAGE] * n_obs_local / (fused_avg_ms * 1e-3) / 1e12 if fused_n else None

    # ---- K1 materialising variant (the HBM-roofline kernel of the north star): a few iterations, same data -----
    optm = bench_options(mcba, materialize_jacobian=1)
    pm = init.copy()
    os.environ["MCBA_GRAPH"] = "0"                 # per-kernel events needed
    hm = mcba.Handle(prob, device=local_rank)
    comm_init(hm)
    del os.environ["MCBA_GRAPH"]
    hm.solve_begin(pm, optm)
    for _ in range(2):
        hm.solve_step()
    hm.reset_kernel_stats()
    for _ in range(3):
        hm.solve_step()
    stm = hm.kernel_stats()
    hm.solve_end(pm)
    if hm is not h:
        hm.close()
    mat_n, mat_ms = stm["resid_jac_materialize"]
    mat_avg_ms = mat_ms / max(1, mat_n)
    hbm_peak, hbm_src = measured_peaks()
    mat_gbs = BYTES_PER_OBS[STAGE] * n_obs_local / (mat_avg_ms * 1e-3) / 1e9 if mat_n else None
    fromj_n, fromj_ms = stm["accumulate_from_jacobian"]

    # ---- parity: a REAL solve (Ceres defaults) of the N-GPU problem, stage 4 then stage 5 on the same handle, against the
    # committed oracle fixture of exactly this problem --------------------------------------------------------------------
    parity = None
    gold = golden_c4(fpg * world) if fpg == FRAMES_PER_GPU else None
    try:
        pp = init.copy()
        pp.intrinsics[:] = s.intr_init
        h.set_stage(4)
        barrier()
        t0 = time.perf_counter()
        s4, t4 = h.solve(pp)
        t_s4 = time.perf_counter() - t0
        h.set_stage(5)
        barrier()
        t0 = time.perf_counter()
        s5, t5 = h.solve(pp)
        t_s5 = time.perf_counter() - t0
        if gold is not None:
            parity = compare_with_golden(gold, [s4, s5], [t4, t5], pp)
        else:
            parity = {"fixture": None, "note": "no committed oracle fixture for this frame count"}
        parity["solve_wall_ms"] = {"stage4": allmax(t_s4) * 1e3, "stage5": allmax(t_s5) * 1e3}
        parity["schur_reduce"] = h.schur_reduce_mode()
    except Exception as e:     # never lose the headline line to the extra leg (collectives below still run on every rank)
        parity = {"error": repr(e)}
    h.close()

    # ---- end-to-end arm: host buffers -> mcba_create (pack + H2D) -> K LM iterations -> parameters back ----------
    e2e = None
    if not args.no_e2e:
        e2e_s = float("inf")
        for trial in range(4):          # first trial warms allocator / page tables; best of the next three
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
        mcba.obs_register(P1, device=local_rank)       # corners resident: what the BA that follows would share
        res_runs = [pnp.ransac(P1, popt, device=local_rank) for _ in range(3)]
        mcba.obs_release(P1, device=local_rank)
        pi_local = (min(r.kernel_ms for r in runs), min(r.total_ms for r in runs), float(P1.n_bobs),
                    int(runs[-1].kernel_launches), int(runs[-1].n_hypotheses.sum()), min(r.total_ms for r in res_runs))
    except Exception as e:   # the BA line must not depend on the extra leg
        pose_init = {"error": repr(e)}
    k_ms = allmax(pi_local[0] if pi_local else -1.0)
    t_ms = allmax(pi_local[1] if pi_local else -1.0)
    tr_ms = allmax(pi_local[5] if pi_local else -1.0)
    nb_global = int(allsum(pi_local[2] if pi_local else 0.0))
    n_ok = int(allsum(1.0 if pi_local else 0.0))
    if pi_local and n_ok == world:
        pose_init = {"board_observations": nb_global, "kernel_ms": k_ms, "total_ms": t_ms, "total_ms_corners_resident": tr_ms,
                     "bobs_per_s_kernels": nb_global / (k_ms * 1e-3), "bobs_per_s_e2e": nb_global / (t_ms * 1e-3),
                     "kernel_launches": pi_local[3], "draws_rank0": pi_local[4],
                     "note": "threshold 10 px, <= 1000 draws (the reference's YAML defaults); e2e = pinned host arrays -> device, "
                             "4 kernels, poses / inlier masks back; corners_resident = after mcba_obs_register (the copy the bundle "
                             "adjustment then shares); every rank poses its own shard, no collective; "
                             "OpenCV single-thread loop of the reference: tools/bench_pnp.py"}
    elif pose_init is None:
        pose_init = {"error": "the pose_init leg failed on another rank"}

    # ---- extra lines: the other BASELINE configs, driver-run ---------------------------------------------------------
    extra = {}
    if not args.no_extra:
        # config 4 strong-scaled: the full 20 000 frames split over the N GPUs (N = 8: that is the weak line above)
        if world * fpg == FULL_FRAMES:
            extra["config4_strong"] = {"frames": FULL_FRAMES, "n_gpus": world, "same_as": "the headline line (2 500 frames per GPU x 8)",
                                       "ms_per_step": ms / args.steps, "value": value, "unit": "Mobs/s", "scaling": "strong"}
        else:
            sp = None
            try:        # local work only: a rank that fails here must still take part in the collectives below
                per = FULL_FRAMES // world
                ss = synth.make_scene("c4", n_frames=FULL_FRAMES)
                t0 = time.perf_counter()
                os_ = synth.generate_observations(ss, rank * per, (rank + 1) * per, workers=workers)
                t_gen = time.perf_counter() - t0
                sp, si, _ = synth.make_problem(ss, os_, STAGE)
            except Exception as e:
                extra["config4_strong"] = {"error": repr(e)}
            if int(allsum(1.0 if sp is not None else 0.0)) == world:
                hs = mcba.Handle(sp, device=local_rank)
                comm_init(hs)
                k_strong = min(args.steps, 10)
                ms_s, _, _, _, st_s, _ = timed_iterations(hs, si, opt, k_strong, 3)
                n_glob_s = int(allsum(float(sp.n_obs)))
                sgold = golden_c4(FULL_FRAMES)
                pp = si.copy(); pp.intrinsics[:] = ss.intr_init
                hs.set_stage(4); s4, t4 = hs.solve(pp)
                hs.set_stage(5); s5, t5 = hs.solve(pp)
                sp_par = compare_with_golden(sgold, [s4, s5], [t4, t5], pp) if sgold is not None else None
                hs.close()
                extra["config4_strong"] = {"frames": FULL_FRAMES, "observations": n_glob_s, "n_gpus": world, "scaling": "strong",
                                           "steps": k_strong, "ms_per_step": ms_s / k_strong,
                                           "value": n_glob_s * k_strong / (ms_s * 1e-3) / 1e6, "unit": "Mobs/s",
                                           "kernel_ms_per_step": {k: v[1] / k_strong for k, v in st_s.items() if v[0]},
                                           "parity": sp_par, "generation_s": t_gen}
                del sp, si, os_
            elif "config4_strong" not in extra:
                extra["config4_strong"] = {"error": "generation failed on another rank"}
        if world == 1:
            # config 3 in full on one GPU: 16-camera ring, 5 boards, 5 000 frames, 10.19 M observations; real stage 4 + 5 solve
            try:
                t0 = time.perf_counter()
                p3, i3, _, s3, o3 = synth.config_problem("c3", 4)
                # (config_problem generates serially; c3 takes ~10 s)
                t_gen = time.perf_counter() - t0
                g3 = json.load(open(os.path.join(GOLD, "full_c3.json")))
                h3 = mcba.Handle(p3, device=local_rank)
                pp = i3.copy(); pp.intrinsics[:] = s3.intr_init
                h3.solve(pp.copy())                          # warm-up solve (module load), discarded
                t0 = time.perf_counter(); s4, t4 = h3.solve(pp); w4 = time.perf_counter() - t0
                h3.set_stage(5)
                h3.solve(pp.copy())
                t0 = time.perf_counter(); s5, t5 = h3.solve(pp); w5 = time.perf_counter() - t0
                par3 = compare_with_golden(g3, [s4, s5], [t4, t5], pp)
                ms3, _, _, _, st3, _ = timed_iterations(h3, i3, opt, 10, 3)
                h3.close()
                extra["config3_full"] = {"workload": "BASELINE configs[2]: 16-camera ring, 5 boards x 63 corners, 5 000 frames, single B200",
                                         "observations": int(p3.n_obs), "reduced_system_n": 270,
                                         "solve_wall_ms": {"stage4": w4 * 1e3, "stage5": w5 * 1e3},
                                         "solve_device_ms": {"stage4": s4.t_total_ms, "stage5": s5.t_total_ms},
                                         "iterations": {"stage4": s4.num_iterations, "stage5": s5.num_iterations},
                                         "oracle_wall_s_8_threads": {"stage4": g3["stages"][0]["wall_s"], "stage5": g3["stages"][1]["wall_s"]},
                                         "ms_per_lm_iteration_stage5": ms3 / 10, "value": p3.n_obs * 10 / (ms3 * 1e-3) / 1e6, "unit": "Mobs/s",
                                         "parity": par3, "generation_s": t_gen}
                del p3, i3, o3
            except Exception as e:
                extra["config3_full"] = {"error": repr(e)}
            # configs 1 and 2 (launch / synchronisation bound): ms per LM iteration, stage 5
            try:
                small = {}
                for cfg in ("c1", "c2"):
                    ps, is_, _, ss_, _ = synth.config_problem(cfg, 5)
                    hsm = mcba.Handle(ps, device=local_rank)
                    qs = is_.copy(); qs.intrinsics[:] = ss_.intr_init
                    best = None
                    for _ in range(3):
                        m_, w_, _, _, _, _ = timed_iterations(hsm, qs, opt, 50, 5)
                        best = w_ / 50 if best is None else min(best, w_ / 50)
                    hsm.close()
                    small[cfg] = {"observations": int(ps.n_obs), "ms_per_lm_iteration": best}
                extra["small_configs"] = small
            except Exception as e:
                extra["small_configs"] = {"error": repr(e)}
            # config 5: 4 096 independent per-camera intrinsic refinements x 300 frames, one batched launch sequence
            try:
                t0 = time.perf_counter()
                p5, i5, _ = synth.batched_intrinsics_problem(n_cameras=4096, n_frames=300, kannala_every=4)
                t_gen = time.perf_counter() - t0
                h5 = mcba.Handle(p5, device=local_rank)
                h5.solve_batched(i5.copy())                 # warm-up
                q5 = i5.copy()
                torch.cuda.synchronize()
                t0 = time.perf_counter(); sm5 = h5.solve_batched(q5); w5 = time.perf_counter() - t0
                h5.close()
                g5 = json.load(open(os.path.join(GOLD, "c5_sample.json")))
                its = np.array([x.num_iterations for x in sm5])
                cams = {}
                ok5 = True
                for c, w in g5["cameras"].items():
                    c = int(c)
                    same = (sm5[c].termination, sm5[c].num_successful_steps, sm5[c].num_unsuccessful_steps, sm5[c].num_iterations) == \
                           (w["termination"], w["num_successful_steps"], w["num_unsuccessful_steps"], w["num_iterations"])
                    rel = float(np.abs(q5.intrinsics[c] - np.array(w["intrinsics"])).max() / np.abs(np.array(w["intrinsics"])).max())
                    cams[str(c)] = {"same_iterations_and_termination": bool(same), "rms_abs_diff_px": abs(sm5[c].rms_px - w["rms_px"]),
                                    "intrinsics_max_rel_diff": rel}
                    ok5 = ok5 and same and rel <= 1e-6 and cams[str(c)]["rms_abs_diff_px"] <= 1e-6
                extra["config5_batched"] = {"workload": "BASELINE configs[4]: 4 096 cameras x 300 frames x 16 corners, stage 1, batched (mcba_solve_batched)",
                                            "observations": int(p5.n_obs), "solve_wall_ms": w5 * 1e3, "problems_per_s": 4096 / w5,
                                            "iterations_min_mean_max": [int(its.min()), float(its.mean()), int(its.max())],
                                            "converged": int(sum(x.termination == 0 for x in sm5)),
                                            "value": float(p5.n_obs / 4096 * its.sum() / w5 / 1e6), "unit": "Mobs/s (observation-iterations)",
                                            "parity": {"fixture": "tests/golden/c5_sample.json (per-camera single-problem oracle solves)",
                                                       "cameras": cams, "pass": bool(ok5)}, "generation_s": t_gen}
                del p5, i5
            except Exception as e:
                extra["config5_batched"] = {"error": repr(e)}

    # ---- CPU baseline (rank 0, N = 1 only): oracle port on the SAME shard, bounded iteration count --------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(REPO, "oracle"))
        import oracle_py
        oracle_py.load()
        threads = host_cores()
        oracle_py.time_iteration(prob, init, opt, 1, threads)
        sec, sec_eval, sec_lin = oracle_py.time_iteration(prob, init, opt, 3, threads)
        oc = synth.generate_observations(s, 0, 250)
        pc, ic, _ = synth.make_problem(s, oc, STAGE)
        sec1, _, _ = oracle_py.time_iteration(pc, ic, opt, 1, 1)
        cpu = {"value": prob.n_obs / sec / 1e6, "unit": "Mobs/s", "cores": threads, "kind": "port",
               "sample": f"the same 2500-frame shard ({prob.n_obs} observations), 3 full LM iterations at the initial point, {threads} OpenMP threads",
               "resid_jac_mobs_per_s": prob.n_obs / sec_eval / 1e6, "ms_per_step": sec * 1e3, "ms_eval": sec_eval * 1e3, "ms_linear": sec_lin * 1e3,
               "single_thread_value": pc.n_obs / sec1 / 1e6,
               "single_thread_note": f"1 thread, first 250 frames ({pc.n_obs} observations): the reference as shipped runs Ceres with num_threads = 1"}

    if rank == 0:
        nccl_ms = stats.get("nccl", (0, 0.0))[1] / args.steps
        upper = 1021 * 1022 // 2 * 8
        line = {
            "metric": METRIC, "value": value, "unit": "Mobs/s", "lm_iter_per_s": args.steps / (ms * 1e-3), "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "ms_per_step_wall": ms_wall / args.steps, "higher_is_better": True,
            "launch_mode": "CUDA graphs (two per iteration: step, assembly)" if graphed else "direct launches (NCCL fallback)",
            "ms_per_step_direct_launches": (ms_direct / args.steps) if ms_direct else None,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(world, n_obs_global, fpg * world),
            "schur_reduce": schur_reduce,
            "comm": {"schur_reduce": schur_reduce, "exchange_ms_per_step": nccl_ms,
                     "nvlink_bytes_per_step_per_gpu": (int(2 * upper * (world - 1) / world) if schur_reduce == "p2p" else None),
                     "note": "p2p: each rank pulls its 1/N slice of the upper triangle of the reduced system from the N-1 peers and pushes "
                             "the sum back (inside k_syrk_sum_p2p, timed under schur_syrk); exchange_ms = pair-record sum + the two scalar gathers "
                             "(k_allreduce_p2p, k_gather_p2p over peer memory when schur_reduce is p2p, NCCL otherwise), from the direct-launch handle"} if world > 1 else None,
            "clocks": clocks, "gpu_launches": int(launches), "e2e": e2e, "parity": parity,
            "roofline": {"kernel": "k_observations_pair<5> (resid + Jacobian + J^T J/J^T r, DMMA; csrc/mcba_obs_fused.cuh)", "bound": "tensor",
                         "achieved": achieved_tf, "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s",
                         "frac": (achieved_tf / FP64_PEAK_TFLOPS) if achieved_tf else None,
                         "traffic": NCU_TRAFFIC["fused"][0] if fpg == FRAMES_PER_GPU else None,
                         "traffic_source": f"ncu capture {NCU_TRAFFIC['fused'][1]} (not measured in this run)",
                         "ms_per_launch": fused_avg_ms, "launches": fused_n,
                         "peak_source": "fp64 DMMA m8n8k4 peak measured by the builder with tools/peak_fp64.cu on this pool (profiles/r01_peak_fp64.json); MEASURED_PEAKS.json has no fp64 figure",
                         "algorithmic": f"{FLOPS_PER_OBS[STAGE]} FLOP/obs x {n_obs_local} obs per launch (SURVEY.md 8d)"},
            "roofline_hbm": {"kernel": "k_observations<5,MODE_MATERIALIZE> (resid + Jacobian written to HBM)", "bound": "hbm",
                             "achieved": mat_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": (mat_gbs / hbm_peak) if mat_gbs else None,
                             "traffic": NCU_TRAFFIC["materialize"][0] if fpg == FRAMES_PER_GPU else None,
                             "traffic_source": f"ncu capture {NCU_TRAFFIC['materialize'][1]} (not measured in this run)",
                             "ms_per_launch": mat_avg_ms, "launches": mat_n, "peak_source": hbm_src,
                             "algorithmic": f"{BYTES_PER_OBS[STAGE]} B/obs x {n_obs_local} obs per launch (SURVEY.md 8d)",
                             "readback_ms_per_launch": fromj_ms / max(1, fromj_n)},
            "resid_jac": {   # the second half of BASELINE's metric: throughput of the residual+Jacobian pass alone
                "fused_mobs_per_s": (n_obs_global / (fused_avg_ms * 1e-3) / 1e6) if fused_n else None,
                "materialize_mobs_per_s": (n_obs_global / (mat_avg_ms * 1e-3) / 1e6) if mat_n else None,
                "note": "all ranks run the pass concurrently on their shards: global observations / rank-0 kernel time"},
            "kernel_ms_per_step": {k: v[1] / args.steps for k, v in stats.items() if v[0]},
            "kernel_ms_note": "CUDA events around each kernel family, from a second, direct-launch handle (graphs carry no per-kernel events)",
            "timed_region_note": f"{args.steps} iterations = {ms:.1f} ms of device time: a short burst-clock region, as long as a real solve of this shard (8-10 iterations)",
            "cpu_baseline": cpu,
            "pose_init": pose_init,
            "extra": extra,
            "final": {"rms_px": summ.rms_px, "cost": summ.final_cost, "iterations": summ.num_iterations},
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
