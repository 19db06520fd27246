"""GPU parity of the shifted family (-m gpu): shifted_lopbicg_switching through the C ABI (same prototype as
shifted_switching_solver.h:12) against the oracle restatement, which is pinned bitwise to the reference's own compiled sources
(tests/test_oracle_golden.py).  Tolerances as for the un-shifted solvers: seed residual history, iterations 1..10, <= 1e-10
relative; iteration count and per-shift stopping iterations within 2; same seed-switching sequence outcome; every shifted
system solved to the reference's EPS."""
import numpy as np
import pytest

from helpers import SHIFTED_CASES, global_csr, shifted_problem

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,kind,g,p0,L,scale,seed", SHIFTED_CASES)
def test_shifted_matches_oracle(B, O, name, kind, g, p0, L, scale, seed):
    B.set_options(quiet=1, cache=1, shift_tol=1e-12, shift_max_iter=1000)
    blk, n, ptr, col, val = global_csr(B, kind, g, p0)
    sigma, b = shifted_problem(O, n, ptr, col, val, L, scale, seed)
    ref = O.shifted_solve(n, ptr, col, val, b, sigma, seed, tol=1e-12, max_iter=1000)
    x = np.zeros((L, n))
    r = b.copy()
    ret = B.shifted_lopbicg_switching(blk, x, r, sigma, seed)
    hist = B.last_history()
    end_seed, stop = B.last_shift_info(L)
    assert abs(ret - ref["ret"]) <= 2, (ret, ref["ret"])
    m = min(10, ret - 1, ref["ret"] - 1)
    got, want = np.sqrt(hist[1:m + 1]), np.sqrt(ref["hist"][1:m + 1])
    assert np.all(np.abs(got - want) <= 1e-10 * want + 1e-15), np.abs(got - want) / want
    assert end_seed == ref["seed"]
    assert np.all(np.abs(stop - ref["stop_iter"]) <= 2), (stop, ref["stop_iter"])
    for j in range(L):
        res = O.spmv(n, ptr, col, val, x[j]) + sigma[j] * x[j] - b
        assert np.linalg.norm(res) <= 1e-10 * np.linalg.norm(b), (j, np.linalg.norm(res) / np.linalg.norm(b))
        assert np.abs(x[j] - ref["x"][j]).max() <= 1e-8 * np.abs(ref["x"][j]).max()
    # the returned r is the seed system's recursive residual
    assert abs(np.dot(r, r) / np.dot(b, b) - hist[ret - 1]) <= 1e-8 * max(hist[ret - 1], 1e-300)


def test_shifted_many_shifts_medium_size(B, O):
    """64 shifts on a 250 k-row matrix (T' family, 63^3): the multi-vector update kernel with a full coefficient table; every
    sampled x_j against the oracle's and against its own shifted system."""
    B.set_options(quiet=1, shift_tol=1e-10, shift_max_iter=1000)
    blk, n, ptr, col, val = global_csr(B, "stencil15", 63, 14.0)
    L, seed = 64, 0
    sigma, b = shifted_problem(O, n, ptr, col, val, L, 0.5 / L, seed)
    ref = O.shifted_solve(n, ptr, col, val, b, sigma, seed, tol=1e-10, max_iter=1000)
    assert ref["ret"] < 1000                                          # the case does converge
    x = np.zeros((L, n))
    r = b.copy()
    ret = B.shifted_lopbicg_switching(blk, x, r, sigma, seed)
    B.set_options(shift_tol=1e-12)
    assert abs(ret - ref["ret"]) <= max(2, int(0.02 * ref["ret"])), (ret, ref["ret"])
    for j in (0, 1, 31, 63):
        res = O.spmv(n, ptr, col, val, x[j]) + sigma[j] * x[j] - b
        res_ref = O.spmv(n, ptr, col, val, ref["x"][j]) + sigma[j] * ref["x"][j] - b
        assert np.linalg.norm(res) <= max(10 * np.linalg.norm(res_ref), 1e-8 * np.linalg.norm(b)), (j, np.linalg.norm(res), np.linalg.norm(res_ref))
        assert np.abs(x[j] - ref["x"][j]).max() <= 1e-6 * np.abs(ref["x"][j]).max()


def test_shifted_stdout_contract(B, O, capfd):
    name, kind, g, p0, L, scale, seed = SHIFTED_CASES[1]
    B.set_options(quiet=0)
    blk, n, ptr, col, val = global_csr(B, kind, g, p0)
    sigma, b = shifted_problem(O, n, ptr, col, val, L, scale, seed)
    x = np.zeros((L, n)); r = b.copy()
    ret = B.shifted_lopbicg_switching(blk, x, r, sigma, seed)
    B.lib.bicg_synchronize()
    import ctypes
    ctypes.CDLL(None).fflush(None)
    out = capfd.readouterr().out
    B.set_options(quiet=1)
    assert f"Total iter   : {ret - 1}" in out and "Total time   : " in out and "Avg time/iter: " in out
    assert "seed: " in out and "remain: " in out and "sigma[" in out            # the seed-switch lines (:518-526)
