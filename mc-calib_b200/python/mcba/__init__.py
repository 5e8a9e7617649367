"""ctypes view of the libmcba C ABI (include/mcba.h) plus numpy-backed problem/parameter holders.

This is host-side plumbing for tests and bench.py: it never computes anything itself and it does
not fall back to the CPU oracle. `load_lib()` raises if the CUDA library is missing.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.abspath(os.path.join(_HERE, "..", "..", ".."))
LIB_PATH = os.path.join(REPO, "mc-calib_b200", "csrc", "libmcba.so")

c_f32p = C.POINTER(C.c_float)
c_f64p = C.POINTER(C.c_double)
c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)
c_u8p = C.POINTER(C.c_uint8)


class mcba_problem(C.Structure):
    _fields_ = [
        ("stage", C.c_int32), ("n_cam", C.c_int32), ("n_pose", C.c_int32), ("n_board", C.c_int32),
        ("n_pts", C.c_int32), ("n_obs", C.c_int64), ("n_bobs", C.c_int64),
        ("u", c_f32p), ("v", c_f32p), ("pt_idx", c_i32p), ("bobs_ptr", c_i64p),
        ("bobs_cam", c_i32p), ("bobs_pose", c_i32p), ("bobs_board", c_i32p),
        ("pts_xyz", c_f32p), ("cam_model", c_i32p), ("cam_is_ref", c_u8p), ("board_is_ref", c_u8p),
        ("cam_problem", c_i32p), ("n_problems", C.c_int32),
    ]


class mcba_params(C.Structure):
    _fields_ = [("cam_pose", c_f64p), ("pose", c_f64p), ("board_pose", c_f64p), ("intrinsics", c_f64p)]


class mcba_options(C.Structure):
    _fields_ = [
        ("max_num_iterations", C.c_int32), ("function_tolerance", C.c_double),
        ("gradient_tolerance", C.c_double), ("parameter_tolerance", C.c_double),
        ("initial_trust_region_radius", C.c_double), ("max_trust_region_radius", C.c_double),
        ("min_trust_region_radius", C.c_double), ("min_relative_decrease", C.c_double),
        ("min_lm_diagonal", C.c_double), ("max_lm_diagonal", C.c_double), ("huber_a", C.c_double),
        ("jacobi_scaling", C.c_int32), ("max_num_consecutive_invalid_steps", C.c_int32),
        ("verbose", C.c_int32), ("materialize_jacobian", C.c_int32),
    ]


class mcba_summary(C.Structure):
    _fields_ = [
        ("termination", C.c_int32), ("num_successful_steps", C.c_int32),
        ("num_unsuccessful_steps", C.c_int32), ("num_iterations", C.c_int32),
        ("initial_cost", C.c_double), ("final_cost", C.c_double), ("rms_px", C.c_double),
        ("t_total_ms", C.c_double), ("t_jacobian_ms", C.c_double), ("t_schur_ms", C.c_double),
        ("t_chol_ms", C.c_double), ("t_cost_ms", C.c_double), ("t_comm_ms", C.c_double),
    ]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


class mcba_trace_row(C.Structure):
    _fields_ = [
        ("iteration", C.c_int32), ("step_is_valid", C.c_int32), ("step_is_successful", C.c_int32),
        ("_pad", C.c_int32), ("cost", C.c_double), ("cost_change", C.c_double),
        ("gradient_max_norm", C.c_double), ("step_norm", C.c_double),
        ("relative_decrease", C.c_double), ("trust_region_radius", C.c_double),
    ]


def default_options(max_num_iterations=1000, **kw):
    """Ceres 2.2.0 Solver::Options defaults + the reference's max_num_iterations (YAML number_iterations)."""
    o = mcba_options(max_num_iterations, 1e-6, 1e-10, 1e-8, 1e4, 1e16, 1e-32, 1e-3, 1e-6, 1e32, 1.0, 1, 5, 0, 0)
    for k, v in kw.items():
        setattr(o, k, v)
    return o


STAGE_K = {1: 15, 2: 12, 3: 12, 4: 18, 5: 27}


def _ptr(a, t):
    return a.ctypes.data_as(t) if a is not None else t()


class Problem:
    """Owns the numpy arrays behind a mcba_problem (keeps them alive and contiguous)."""

    def __init__(self, stage, u, v, pt_idx, bobs_ptr, bobs_cam, bobs_pose, bobs_board, pts_xyz,
                 cam_model, n_pose, n_board, cam_is_ref=None, board_is_ref=None, cam_problem=None,
                 n_problems=1):
        self.stage = int(stage)
        self.u = np.ascontiguousarray(u, np.float32)
        self.v = np.ascontiguousarray(v, np.float32)
        self.pt_idx = np.ascontiguousarray(pt_idx, np.int32)
        self.bobs_ptr = np.ascontiguousarray(bobs_ptr, np.int64)
        self.bobs_cam = np.ascontiguousarray(bobs_cam, np.int32)
        self.bobs_pose = np.ascontiguousarray(bobs_pose, np.int32)
        self.bobs_board = np.ascontiguousarray(bobs_board, np.int32)
        self.pts_xyz = np.ascontiguousarray(pts_xyz, np.float32).reshape(-1, 3)
        self.cam_model = np.ascontiguousarray(cam_model, np.int32)
        self.n_cam = len(self.cam_model)
        self.n_pose = int(n_pose)
        self.n_board = int(n_board)
        self.cam_is_ref = None if cam_is_ref is None else np.ascontiguousarray(cam_is_ref, np.uint8)
        self.board_is_ref = None if board_is_ref is None else np.ascontiguousarray(board_is_ref, np.uint8)
        self.cam_problem = None if cam_problem is None else np.ascontiguousarray(cam_problem, np.int32)
        self.n_problems = int(n_problems)

    @property
    def n_obs(self):
        return len(self.u)

    @property
    def n_bobs(self):
        return len(self.bobs_cam)

    def with_stage(self, stage):
        p = Problem.__new__(Problem)
        p.__dict__.update(self.__dict__)
        p.stage = int(stage)
        return p

    def c(self):
        s = mcba_problem()
        s.stage, s.n_cam, s.n_pose, s.n_board = self.stage, self.n_cam, self.n_pose, self.n_board
        s.n_pts, s.n_obs, s.n_bobs = len(self.pts_xyz), self.n_obs, self.n_bobs
        s.u, s.v, s.pt_idx = _ptr(self.u, c_f32p), _ptr(self.v, c_f32p), _ptr(self.pt_idx, c_i32p)
        s.bobs_ptr = _ptr(self.bobs_ptr, c_i64p)
        s.bobs_cam, s.bobs_pose, s.bobs_board = (_ptr(self.bobs_cam, c_i32p), _ptr(self.bobs_pose, c_i32p),
                                                 _ptr(self.bobs_board, c_i32p))
        s.pts_xyz, s.cam_model = _ptr(self.pts_xyz, c_f32p), _ptr(self.cam_model, c_i32p)
        s.cam_is_ref, s.board_is_ref = _ptr(self.cam_is_ref, c_u8p), _ptr(self.board_is_ref, c_u8p)
        s.cam_problem, s.n_problems = _ptr(self.cam_problem, c_i32p), self.n_problems
        return s


class Params:
    def __init__(self, cam_pose, pose, board_pose, intrinsics):
        f = lambda a, w: np.ascontiguousarray(np.asarray(a, np.float64).reshape(-1, w)).copy()
        self.cam_pose, self.pose = f(cam_pose, 6), f(pose, 6)
        self.board_pose, self.intrinsics = f(board_pose, 6), f(intrinsics, 9)

    def copy(self):
        return Params(self.cam_pose, self.pose, self.board_pose, self.intrinsics)

    def c(self):
        return mcba_params(_ptr(self.cam_pose, c_f64p), _ptr(self.pose, c_f64p),
                           _ptr(self.board_pose, c_f64p), _ptr(self.intrinsics, c_f64p))


_lib = None


def load_lib():
    """Load libmcba.so (the CUDA product). Raises if it has not been built: there is no CPU fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python __graft_entry__.py` (build()) first; "
                           "libmcba has no CPU fallback")
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    vp, vpp = C.c_void_p, C.POINTER(C.c_void_p)
    L.mcba_default_options.argtypes = [C.POINTER(mcba_options)]
    L.mcba_default_options.restype = None
    L.mcba_create.argtypes = [vpp, C.POINTER(mcba_problem), C.c_int]
    L.mcba_set_stage.argtypes = [vp, C.c_int]
    L.mcba_obs_register.argtypes = [c_f32p, c_f32p, c_i32p, C.c_int64, C.c_int]
    L.mcba_obs_release.argtypes = [c_f32p, C.c_int]
    L.mcba_schur_reduce_mode.argtypes = [vp]
    L.mcba_nccl_unique_id.argtypes = [vp]
    L.mcba_comm_init.argtypes = [vp, C.c_int, C.c_int, vp]
    L.mcba_solve.argtypes = [vp, C.POINTER(mcba_params), C.POINTER(mcba_options), C.POINTER(mcba_summary),
                             C.POINTER(mcba_trace_row), C.c_int]
    L.mcba_solve_batched.argtypes = [vp, C.POINTER(mcba_params), C.POINTER(mcba_options), C.POINTER(mcba_summary)]
    L.mcba_solve_begin.argtypes = [vp, C.POINTER(mcba_params), C.POINTER(mcba_options)]
    L.mcba_solve_step.argtypes = [vp]
    L.mcba_solve_end.argtypes = [vp, C.POINTER(mcba_params), C.POINTER(mcba_summary)]
    L.mcba_eval.argtypes = [vp, C.POINTER(mcba_params), C.POINTER(mcba_options), c_f64p, c_f64p, c_f64p]
    L.mcba_jacobian_width.argtypes = [C.c_int]
    L.mcba_reprojection_errors.argtypes = [vp, C.POINTER(mcba_params), c_f64p]
    L.mcba_reprojection_errors_f32.argtypes = [vp, C.POINTER(mcba_params), c_f32p]
    L.mcba_destroy.argtypes = [vp]
    L.mcba_destroy.restype = None
    L.mcba_last_error.argtypes = [vp]
    L.mcba_last_error.restype = C.c_char_p
    L.mcba_num_kernels.argtypes = []
    L.mcba_kernel_name.argtypes = [C.c_int]
    L.mcba_kernel_name.restype = C.c_char_p
    L.mcba_kernel_stats.argtypes = [vp, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_double)]
    L.mcba_reset_kernel_stats.argtypes = [vp]
    L.mcba_kernel_launches.argtypes = [vp]
    L.mcba_kernel_launches.restype = C.c_int64
    L.mcba_timer_start.argtypes = [vp]
    L.mcba_timer_stop.argtypes = [vp, c_f64p]
    L.mcba_debug_fetch.argtypes = [vp, C.c_char_p, c_f64p, C.c_int64, C.POINTER(C.c_int64)]
    L.mcba_debug_set_k1.argtypes = [C.c_int]
    L.mcba_problem_save.argtypes = [C.c_char_p, C.POINTER(mcba_problem), C.POINTER(mcba_params)]
    L.mcba_problem_load.argtypes = [C.c_char_p, C.POINTER(mcba_problem), C.POINTER(mcba_params), vpp]
    L.mcba_problem_free.argtypes = [vp]
    L.mcba_problem_free.restype = None
    L.mcba_io_last_error.restype = C.c_char_p
    L.mcba_debug_cholesky.argtypes = [vp, C.c_int, c_f64p, c_f64p, c_f64p, C.c_int, C.c_int, c_f64p, C.POINTER(C.c_int)]
    _lib = L
    return L


class McbaError(RuntimeError):
    pass


def save_problem(path, problem, params=None):
    """Binary dump of a packed problem (+ parameter blocks): mcba_problem_save, the "MCBAPRB1" format."""
    L = load_lib()
    cp = problem.c()
    cq = params.c() if params is not None else None
    rc = L.mcba_problem_save(os.fsencode(path), C.byref(cp), C.byref(cq) if cq is not None else None)
    if rc != 0:
        raise McbaError(f"mcba_problem_save rc={rc}: {L.mcba_io_last_error().decode()}")


def load_problem(path):
    """-> (Problem, Params or None) from a file written by mcba_problem_save (copies out of the library's storage)."""
    L = load_lib()
    cp, cq, st = mcba_problem(), mcba_params(), C.c_void_p()
    rc = L.mcba_problem_load(os.fsencode(path), C.byref(cp), C.byref(cq), C.byref(st))
    if rc != 0:
        raise McbaError(f"mcba_problem_load rc={rc}: {L.mcba_io_last_error().decode()}")
    try:
        def arr(ptr, n, dt):
            if not ptr:
                return None
            return np.ctypeslib.as_array(ptr, shape=(int(n),)).astype(dt, copy=True) if n else np.zeros(0, dt)
        nb, no, nc = cp.n_bobs, cp.n_obs, cp.n_cam
        prob = Problem(cp.stage, arr(cp.u, no, np.float32), arr(cp.v, no, np.float32), arr(cp.pt_idx, no, np.int32),
                       arr(cp.bobs_ptr, nb + 1, np.int64), arr(cp.bobs_cam, nb, np.int32), arr(cp.bobs_pose, nb, np.int32),
                       arr(cp.bobs_board, nb, np.int32), arr(cp.pts_xyz, 3 * cp.n_pts, np.float32),
                       arr(cp.cam_model, nc, np.int32), cp.n_pose, cp.n_board, arr(cp.cam_is_ref, nc, np.uint8),
                       arr(cp.board_is_ref, cp.n_board, np.uint8), arr(cp.cam_problem, nc, np.int32), cp.n_problems)
        params = None
        if cq.pose or cq.intrinsics or cq.cam_pose or cq.board_pose:
            z = lambda n, w: np.zeros((n, w))
            params = Params(arr(cq.cam_pose, 6 * nc, np.float64) if cq.cam_pose else z(nc, 6),
                            arr(cq.pose, 6 * cp.n_pose, np.float64) if cq.pose else z(cp.n_pose, 6),
                            arr(cq.board_pose, 6 * cp.n_board, np.float64) if cq.board_pose else z(cp.n_board, 6),
                            arr(cq.intrinsics, 9 * nc, np.float64) if cq.intrinsics else z(nc, 9))
    finally:
        L.mcba_problem_free(st)
    return prob, params


def obs_register(problem, device=0):
    """Upload u / v / pt_idx once; later Handle(problem) / pnp.ransac calls on the same arrays reuse the device copy."""
    rc = load_lib().mcba_obs_register(_ptr(problem.u, c_f32p), _ptr(problem.v, c_f32p), _ptr(problem.pt_idx, c_i32p),
                                      problem.n_obs, device)
    if rc != 0:
        raise McbaError(f"mcba_obs_register failed rc={rc}")


def obs_release(problem, device=0):
    rc = load_lib().mcba_obs_release(_ptr(problem.u, c_f32p), device)
    if rc != 0:
        raise McbaError(f"mcba_obs_release failed rc={rc} (unknown arrays, or a live handle still uses them)")


class Handle:
    """RAII wrapper of a libmcba handle; mirrors the call sequence of a reference refine* stage:
    build the problem (mcba_create) then solve (mcba_solve)."""

    def __init__(self, problem, device=0):
        self.lib = load_lib()
        self.problem = problem
        self.h = C.c_void_p()
        cp = problem.c()
        rc = self.lib.mcba_create(C.byref(self.h), C.byref(cp), device)
        if rc != 0:
            msg = self.lib.mcba_last_error(self.h)
            raise McbaError(f"mcba_create failed rc={rc}: {msg.decode() if msg else ''}")

    def _check(self, rc, what):
        if rc < 0:
            msg = self.lib.mcba_last_error(self.h)
            raise McbaError(f"{what} failed rc={rc}: {msg.decode() if msg else ''}")
        return rc

    def set_stage(self, stage):
        self._check(self.lib.mcba_set_stage(self.h, stage), "mcba_set_stage")
        self.problem = self.problem.with_stage(stage)

    def comm_init(self, rank, n_ranks, uid_bytes):
        buf = C.create_string_buffer(bytes(uid_bytes), 128)
        self._check(self.lib.mcba_comm_init(self.h, rank, n_ranks, buf), "mcba_comm_init")

    def schur_reduce_mode(self):
        return {0: "single", 1: "nccl", 2: "p2p"}[self.lib.mcba_schur_reduce_mode(self.h)]

    def solve(self, params, opt=None, trace_cap=2048):
        opt = opt or default_options()
        summ = mcba_summary()
        trace = (mcba_trace_row * trace_cap)()
        cp = params.c()
        self._check(self.lib.mcba_solve(self.h, C.byref(cp), C.byref(opt), C.byref(summ), trace, trace_cap),
                    "mcba_solve")
        return summ, [trace[i] for i in range(min(trace_cap, summ.num_iterations))]

    def solve_batched(self, params, opt=None):
        opt = opt or default_options()
        summ = (mcba_summary * self.problem.n_problems)()
        cp = params.c()
        self._check(self.lib.mcba_solve_batched(self.h, C.byref(cp), C.byref(opt), summ), "mcba_solve_batched")
        return list(summ)

    def solve_begin(self, params, opt):
        cp = params.c()
        self._check(self.lib.mcba_solve_begin(self.h, C.byref(cp), C.byref(opt)), "mcba_solve_begin")

    def solve_step(self):
        return self._check(self.lib.mcba_solve_step(self.h), "mcba_solve_step")

    def solve_end(self, params):
        summ = mcba_summary()
        cp = params.c()
        self._check(self.lib.mcba_solve_end(self.h, C.byref(cp), C.byref(summ)), "mcba_solve_end")
        return summ

    def eval(self, params, opt=None, want_jac=True):
        opt = opt or default_options()
        K = STAGE_K[self.problem.stage]
        n = self.problem.n_obs
        cost = C.c_double()
        res = np.zeros((n, 2))
        jac = np.zeros((n, 2, K)) if want_jac else None
        cp = params.c()
        self._check(self.lib.mcba_eval(self.h, C.byref(cp), C.byref(opt), C.byref(cost), _ptr(res, c_f64p),
                                       _ptr(jac, c_f64p)), "mcba_eval")
        return cost.value, res, jac

    def reprojection_errors(self, params):
        err = np.zeros(self.problem.n_obs)
        cp = params.c()
        self._check(self.lib.mcba_reprojection_errors(self.h, C.byref(cp), _ptr(err, c_f64p)),
                    "mcba_reprojection_errors")
        return err

    def reprojection_errors_f32(self, params):
        err = np.zeros(self.problem.n_obs, np.float32)
        cp = params.c()
        self._check(self.lib.mcba_reprojection_errors_f32(self.h, C.byref(cp), _ptr(err, c_f32p)),
                    "mcba_reprojection_errors_f32")
        return err

    def kernel_stats(self):
        out = {}
        for k in range(self.lib.mcba_num_kernels()):
            n, ms = C.c_int64(), C.c_double()
            self.lib.mcba_kernel_stats(self.h, k, C.byref(n), C.byref(ms))
            out[self.lib.mcba_kernel_name(k).decode()] = (n.value, ms.value)
        return out

    def kernel_launches(self):
        return self.lib.mcba_kernel_launches(self.h)

    def timer_start(self):
        self._check(self.lib.mcba_timer_start(self.h), "mcba_timer_start")

    def timer_stop(self):
        ms = C.c_double()
        self._check(self.lib.mcba_timer_stop(self.h, C.byref(ms)), "mcba_timer_stop")
        return ms.value

    def debug_fetch(self, name):
        n = C.c_int64()
        self._check(self.lib.mcba_debug_fetch(self.h, name.encode(), c_f64p(), 0, C.byref(n)), "mcba_debug_fetch")
        out = np.zeros(n.value)
        self._check(self.lib.mcba_debug_fetch(self.h, name.encode(), _ptr(out, c_f64p), n.value, C.byref(n)),
                    "mcba_debug_fetch")
        return out

    def debug_cholesky(self, A, rhs, variant=-1, iters=1):
        """Solve A z = rhs with the reduced-system kernels alone -> (z, L, mean kernel ms, fail flag)."""
        n = len(rhs)
        S = np.ascontiguousarray(np.vstack([np.asarray(A, np.float64), np.asarray(rhs, np.float64)[None, :]]))
        z, L = np.zeros(n), np.zeros((n + 1, n))
        ms, fail = C.c_double(), C.c_int()
        self._check(self.lib.mcba_debug_cholesky(self.h, n, _ptr(S, c_f64p), _ptr(z, c_f64p), _ptr(L, c_f64p), variant, iters,
                                                 C.byref(ms), C.byref(fail)), "mcba_debug_cholesky")
        return z, L, ms.value, fail.value

    def reset_kernel_stats(self):
        self.lib.mcba_reset_kernel_stats(self.h)

    def close(self):
        if self.h:
            self.lib.mcba_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
