"""ctypes wrapper of oracle/liboracle.so — TEST INFRASTRUCTURE (see oracle/oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C
import os
import subprocess
import sys
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(_HERE, "..", "mc-calib_b200", "python"))
import mcba  # noqa: E402  (struct definitions only)

_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def load():
    global _lib
    if _lib is not None:
        return _lib
    path = os.path.join(_HERE, "liboracle.so")
    if not os.path.exists(path):
        build()
    L = C.CDLL(path)
    P, PR, O, S, T = (C.POINTER(mcba.mcba_problem), C.POINTER(mcba.mcba_params), C.POINTER(mcba.mcba_options),
                      C.POINTER(mcba.mcba_summary), C.POINTER(mcba.mcba_trace_row))
    d = mcba.c_f64p
    L.oracle_eval.argtypes = [P, PR, O, d, d, d, d]
    L.oracle_solve.argtypes = [P, PR, O, S, T, C.c_int, C.c_int]
    L.oracle_time_iteration.argtypes = [P, PR, O, C.c_int, C.c_int, d, d, d]
    L.oracle_functor.argtypes = [C.c_int, C.c_int, C.c_int, d, d, d, d, d, d, d, d]
    L.oracle_angle_axis_rotate.argtypes = [d, d, d, d, d]
    L.oracle_angle_axis_rotate.restype = None
    L.oracle_huber.argtypes = [C.c_double, C.c_double, d]
    L.oracle_huber.restype = None
    _lib = L
    return L


def _p(a):
    return a.ctypes.data_as(mcba.c_f64p) if a is not None else mcba.c_f64p()


def eval(problem, params, opt=None):
    L = load()
    opt = opt or mcba.default_options()
    K = mcba.STAGE_K[problem.stage]
    cost, g = C.c_double(), C.c_double()
    res = np.zeros((problem.n_obs, 2))
    jac = np.zeros((problem.n_obs, 2, K))
    cp, cq = problem.c(), params.c()
    rc = L.oracle_eval(C.byref(cp), C.byref(cq), C.byref(opt), C.byref(cost), _p(res), _p(jac), C.byref(g))
    return rc, cost.value, res, jac, g.value


def solve(problem, params, opt=None, n_threads=1, trace_cap=2048):
    """Runs the CPU restatement of ceres::Solve in place on `params`. Returns (rc, summary, trace rows)."""
    L = load()
    opt = opt or mcba.default_options()
    summ = mcba.mcba_summary()
    trace = (mcba.mcba_trace_row * trace_cap)()
    cp, cq = problem.c(), params.c()
    rc = L.oracle_solve(C.byref(cp), C.byref(cq), C.byref(opt), C.byref(summ), trace, trace_cap, n_threads)
    return rc, summ, [trace[i] for i in range(min(trace_cap, summ.num_iterations))]


def time_iteration(problem, params, opt=None, n_iters=1, n_threads=1):
    L = load()
    opt = opt or mcba.default_options()
    a, b, c = C.c_double(), C.c_double(), C.c_double()
    cp, cq = problem.c(), params.c()
    rc = L.oracle_time_iteration(C.byref(cp), C.byref(cq), C.byref(opt), n_iters, n_threads, C.byref(a),
                                 C.byref(b), C.byref(c))
    if rc != 0:
        raise RuntimeError(f"oracle_time_iteration rc={rc}")
    return a.value, b.value, c.value


def functor(stage, model, flags, cam_pose, pose, board_pose, intr, xyz, uv):
    L = load()
    K = mcba.STAGE_K[stage]
    r, J = np.zeros(2), np.zeros((2, K))
    f = lambda a: None if a is None else np.ascontiguousarray(a, np.float64)
    a = [f(cam_pose), f(pose), f(board_pose), f(intr), f(xyz), f(uv)]
    L.oracle_functor(stage, model, flags, *[_p(x) for x in a], _p(r), _p(J))
    return r, J


def angle_axis_rotate(aa, pt):
    L = load()
    aa, pt = np.ascontiguousarray(aa, np.float64), np.ascontiguousarray(pt, np.float64)
    out, d_aa, d_pt = np.zeros(3), np.zeros((3, 3)), np.zeros((3, 3))
    L.oracle_angle_axis_rotate(_p(aa), _p(pt), _p(out), _p(d_aa), _p(d_pt))
    return out, d_aa, d_pt


def huber(a, s):
    L = load()
    rho = np.zeros(3)
    L.oracle_huber(a, s, _p(rho))
    return rho


def max_threads():
    return load().oracle_max_threads()
