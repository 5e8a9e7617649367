"""The two reduced-system kernels on their own (mcba_debug_cholesky): k_chol_coop (grid-synchronous panels) and
k_chol_tiles (dataflow, one CTA per 64x64 tile) against numpy on random SPD systems of every shape class: a single
partial tile, exact multiples of 64 (the rhs row is a tile row of its own), several tiles with a ragged last one, the
bench's n = 1020, and a non-positive-definite matrix (failure flag, z = 0)."""
import numpy as np
import pytest

import mcba
from mcba import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def handle():
    s = synth.make_scene("c1", n_frames=4)
    o = synth.generate_observations(s)
    prob, init, gt = synth.make_problem(s, o, 4)
    h = mcba.Handle(prob)
    yield h
    h.close()


def spd(n, seed, cond=1e6):
    rng = np.random.default_rng(seed)
    Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    ev = np.logspace(0, -np.log10(cond), n)
    A = (Q * ev) @ Q.T
    return 0.5 * (A + A.T), rng.standard_normal(n)


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("n", [1, 5, 36, 63, 64, 65, 78, 128, 200, 270, 449, 512, 1020, 1024, 1100, 1408])
def test_factor_and_solve_match_numpy(handle, variant, n):
    A, b = spd(n, 100 + n)
    z, L, ms, fail = handle.debug_cholesky(A, b, variant=variant)
    assert fail == 0
    zr = np.linalg.solve(A, b)
    assert np.abs(z - zr).max() <= 1e-8 * np.abs(zr).max()
    Lg = np.tril(L[:n])
    assert np.abs(Lg @ Lg.T - A).max() <= 1e-12 * np.abs(A).max()
    y = np.linalg.solve(Lg, b)                      # row n of the factor = L^-1 rhs (forward substitution)
    assert np.abs(L[n] - y).max() <= 1e-9 * np.abs(y).max()


@pytest.mark.parametrize("variant", [0, 1])
def test_not_positive_definite_sets_the_failure_flag(handle, variant):
    A, b = spd(200, 7)
    A[150, 150] = -1.0
    z, L, ms, fail = handle.debug_cholesky(A, b, variant=variant)
    assert fail == 1 and np.array_equal(z, np.zeros(200))


def test_the_two_kernels_agree_and_are_reproducible(handle):
    A, b = spd(1020, 3)
    z0, _, _, _ = handle.debug_cholesky(A, b, variant=0)
    z1, _, _, _ = handle.debug_cholesky(A, b, variant=1)
    z2, _, _, _ = handle.debug_cholesky(A, b, variant=1)
    assert np.array_equal(z1, z2)
    assert np.abs(z0 - z1).max() <= 1e-9 * np.abs(z0).max()
