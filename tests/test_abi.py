"""The C-ABI library loads on a CPU-only box and exports every function include/mcba.h declares
(no compute call is made here: there is no CPU fallback to call)."""
import ctypes
import os
import re

import pytest

import mcba

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    fns = set()
    for h in ("mcba.h", "mcba_pnp.h"):
        src = open(os.path.join(REPO, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        fns |= set(re.findall(r"\b(mcba_[a-z0-9_]+)\s*\(", src))
    return sorted(fns)


def test_header_declares_the_expected_entry_points():
    fns = declared_functions()
    for name in ("mcba_create", "mcba_solve", "mcba_eval", "mcba_solve_begin", "mcba_solve_step", "mcba_solve_end",
                 "mcba_comm_init", "mcba_destroy", "mcba_last_error", "mcba_reprojection_errors",
                 "mcba_pnp_ransac", "mcba_pnp_default_options", "mcba_pnp_last_error"):
        assert name in fns


def test_library_exports_every_declared_symbol():
    if not os.path.exists(mcba.LIB_PATH):
        pytest.skip("libmcba.so not built yet (python __graft_entry__.py)")
    lib = ctypes.CDLL(mcba.LIB_PATH)
    missing = [f for f in declared_functions() if not hasattr(lib, f)]
    assert not missing, missing


def test_struct_sizes_match_the_header():
    """ctypes mirrors in mcba/__init__.py against sizes computed from the C header by the compiler."""
    import subprocess
    import tempfile
    code = '#include "mcba.h"\n#include <stdio.h>\nint main(){printf("%zu %zu %zu %zu %zu\\n", sizeof(mcba_problem), sizeof(mcba_params), sizeof(mcba_options), sizeof(mcba_summary), sizeof(mcba_trace_row));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(code)
        subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(REPO, "include"), "-o", os.path.join(d, "t"), os.path.join(d, "t.c")])
        out = subprocess.check_output([os.path.join(d, "t")]).decode().split()
    sizes = [ctypes.sizeof(x) for x in (mcba.mcba_problem, mcba.mcba_params, mcba.mcba_options, mcba.mcba_summary, mcba.mcba_trace_row)]
    assert [int(x) for x in out] == sizes
    from mcba import pnp
    code = '#include "mcba_pnp.h"\n#include <stdio.h>\nint main(){printf("%zu %zu %zu\\n", sizeof(mcba_pnp_problem), sizeof(mcba_pnp_options), sizeof(mcba_pnp_result));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(code)
        subprocess.check_call(["/usr/bin/gcc", "-I", os.path.join(REPO, "include"), "-o", os.path.join(d, "t"), os.path.join(d, "t.c")])
        out = subprocess.check_output([os.path.join(d, "t")]).decode().split()
    assert [int(x) for x in out] == [ctypes.sizeof(x) for x in (pnp.mcba_pnp_problem, pnp.mcba_pnp_options, pnp.mcba_pnp_result)]


def test_no_cpu_fallback_without_gpu():
    """On a box without a CUDA device mcba_create must fail loudly (MCBA_ERR_CUDA), never compute on the CPU."""
    import torch
    if torch.cuda.is_available() or not os.path.exists(mcba.LIB_PATH):
        pytest.skip("needs a CPU-only box with the library built")
    from mcba import synth
    s = synth.make_scene("c1", n_frames=3)
    o = synth.generate_observations(s)
    prob, init, gt = synth.make_problem(s, o, 4)
    with pytest.raises(mcba.McbaError) as e:
        mcba.Handle(prob)
    assert "rc=-2" in str(e.value)
