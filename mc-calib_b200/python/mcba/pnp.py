"""ctypes view of include/mcba_pnp.h (batched P3P RANSAC + pose refinement on the GPU). Host-side plumbing only:
nothing here computes, and there is no CPU fallback."""
import ctypes as C

import numpy as np

from . import load_lib, McbaError, c_f32p, c_f64p, c_i32p, c_i64p, c_u8p, _ptr


class mcba_pnp_problem(C.Structure):
    _fields_ = [("n_obs", C.c_int64), ("u", c_f32p), ("v", c_f32p), ("pt_idx", c_i32p),
                ("n_bobs", C.c_int64), ("bobs_ptr", c_i64p), ("bobs_cam", c_i32p),
                ("n_pts", C.c_int32), ("pts_xyz", c_f32p),
                ("n_cam", C.c_int32), ("cam_model", c_i32p), ("intrinsics", c_f64p)]


class mcba_pnp_options(C.Structure):
    _fields_ = [("threshold_px", C.c_double), ("max_iterations", C.c_int32), ("confidence", C.c_double),
                ("refine", C.c_int32), ("refine_max_iterations", C.c_int32), ("seed", C.c_uint64)]


class mcba_pnp_result(C.Structure):
    _fields_ = [("pose", c_f64p), ("n_inliers", c_i32p), ("inlier_mask", c_u8p), ("best_hypothesis", c_i32p),
                ("n_hypotheses", c_i32p), ("unrefined_pose", c_f64p), ("kernel_ms", C.c_double),
                ("total_ms", C.c_double), ("kernel_launches", C.c_int64)]


def _lib():
    L = load_lib()
    if not getattr(L, "_pnp_ready", False):
        L.mcba_pnp_default_options.argtypes = [C.POINTER(mcba_pnp_options)]
        L.mcba_pnp_default_options.restype = None
        L.mcba_pnp_ransac.argtypes = [C.POINTER(mcba_pnp_problem), C.POINTER(mcba_pnp_options), C.c_int,
                                      C.POINTER(mcba_pnp_result)]
        L.mcba_pnp_last_error.argtypes = []
        L.mcba_pnp_last_error.restype = C.c_char_p
        L._pnp_ready = True
    return L


def default_options(**kw):
    o = mcba_pnp_options()
    _lib().mcba_pnp_default_options(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


class PnpProblem:
    """Owns the arrays behind a mcba_pnp_problem. `prob` is any object with the observation fields of mcba.Problem."""

    def __init__(self, prob, intrinsics):
        self.u, self.v, self.pt_idx = prob.u, prob.v, prob.pt_idx
        self.bobs_ptr, self.bobs_cam = prob.bobs_ptr, prob.bobs_cam
        self.pts_xyz, self.cam_model = prob.pts_xyz, prob.cam_model
        self.intrinsics = np.ascontiguousarray(np.asarray(intrinsics, np.float64).reshape(-1, 9)).copy()
        assert len(self.intrinsics) == len(self.cam_model)

    @property
    def n_obs(self):
        return len(self.u)

    @property
    def n_bobs(self):
        return len(self.bobs_cam)

    def c(self):
        s = mcba_pnp_problem()
        s.n_obs, s.u, s.v, s.pt_idx = self.n_obs, _ptr(self.u, c_f32p), _ptr(self.v, c_f32p), _ptr(self.pt_idx, c_i32p)
        s.n_bobs, s.bobs_ptr, s.bobs_cam = self.n_bobs, _ptr(self.bobs_ptr, c_i64p), _ptr(self.bobs_cam, c_i32p)
        s.n_pts, s.pts_xyz = len(self.pts_xyz), _ptr(self.pts_xyz, c_f32p)
        s.n_cam, s.cam_model, s.intrinsics = len(self.cam_model), _ptr(self.cam_model, c_i32p), _ptr(self.intrinsics, c_f64p)
        return s


class PnpResult:
    def __init__(self, n_bobs, n_obs):
        self.pose = np.zeros((n_bobs, 6))
        self.unrefined_pose = np.zeros((n_bobs, 6))
        self.n_inliers = np.zeros(n_bobs, np.int32)
        self.inlier_mask = np.zeros(n_obs, np.uint8)
        self.best_hypothesis = np.zeros(n_bobs, np.int32)
        self.n_hypotheses = np.zeros(n_bobs, np.int32)
        self.kernel_ms = self.total_ms = 0.0
        self.kernel_launches = 0


def ransac(problem, opt=None, device=0):
    """mcba_pnp_ransac on `device`; raises McbaError on failure (e.g. no sm_100 GPU)."""
    L = _lib()
    opt = opt or default_options()
    out = PnpResult(problem.n_bobs, problem.n_obs)
    r = mcba_pnp_result(_ptr(out.pose, c_f64p), _ptr(out.n_inliers, c_i32p), _ptr(out.inlier_mask, c_u8p),
                        _ptr(out.best_hypothesis, c_i32p), _ptr(out.n_hypotheses, c_i32p),
                        _ptr(out.unrefined_pose, c_f64p), 0.0, 0.0, 0)
    cp = problem.c()
    rc = L.mcba_pnp_ransac(C.byref(cp), C.byref(opt), int(device), C.byref(r))
    if rc != 0:
        raise McbaError(f"mcba_pnp_ransac failed ({rc}): {L.mcba_pnp_last_error().decode()}")
    out.kernel_ms, out.total_ms, out.kernel_launches = r.kernel_ms, r.total_ms, r.kernel_launches
    return out
