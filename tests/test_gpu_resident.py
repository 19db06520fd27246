"""Resident mode of the persistent kernel (-m gpu; csrc/mega.cu): a CTA whose matrix slice fits into its shared memory --
8-byte values, 16-bit offsets into the CTA's own / ghost column windows, row pointers -- loads it once per solve instead of
streaming it through the TMA ring in every SpMV (the strong-scaling regime: T' over 8 GPUs x 148 CTAs).  The two modes must
agree with each other and with the oracle; `resident_ctas` says how many CTAs really took the resident path."""
import numpy as np
import pytest
import scipy.sparse as sp

from helpers import METHODS, RR, SMALL_CASES, global_csr, rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-10


@pytest.fixture(autouse=True)
def _opts(B):
    B.set_options(quiet=1, tol=TOL, max_iter=1000, cache=1, mega=1, resident=1)
    yield
    B.set_options(resident=1, tol=1e-15, mega=1)


def _solve(B, dm, method, n, resident):
    B.set_options(resident=resident)
    b = dm.spmv(np.ones(n))
    x = np.zeros(n)
    kw = RR if method.endswith("rr") else {}
    it, st = dm.solve(method, x, b, **kw)
    assert st["kernel_launches"] <= 8                         # the loop ran as the persistent kernel
    return it, x, np.sqrt(B.last_history()), dm.resident_ctas()


def _agree(run1, run0, ref, n_expected_resident=None):
    it1, x1, h1, nres1 = run1
    it0, x0, h0, nres0 = run0
    assert nres0 == 0
    if n_expected_resident is not None:
        assert nres1 == n_expected_resident
    assert nres1 > 0
    want = np.sqrt(ref["hist"])
    for it, h in ((it1, h1), (it0, h0)):
        m = min(10, it, ref["iters"])
        assert np.all(np.abs(h[1:m + 1] - want[1:m + 1]) <= 1e-10 * want[1:m + 1] + 1e-15), (h[1:m + 1], want[1:m + 1])
        assert abs(it - ref["iters"]) <= max(2, int(0.02 * ref["iters"]))
    assert abs(it1 - it0) <= 2
    assert np.abs(x1 - 1).max() < 1e-6 and np.abs(x0 - 1).max() < 1e-6


@pytest.mark.parametrize("method", METHODS)
@pytest.mark.parametrize("name,kind,g,p0", SMALL_CASES)
def test_resident_matches_streaming_and_oracle(B, O, name, kind, g, p0, method):
    blk, n, ptr, col, val = global_csr(B, kind, g, p0)
    kw = RR if method.endswith("rr") else {}
    ref = O.solve(method, n, ptr, col, val, O.spmv(n, ptr, col, val, np.ones(n)), tol=TOL, max_iter=1000, **kw)
    dm = B.DeviceMatrix(blk)
    try:
        run1 = _solve(B, dm, method, n, 1)
        run0 = _solve(B, dm, method, n, 0)
    finally:
        dm.destroy()
    _agree(run1, run0, ref)


def test_resident_at_the_per_gpu_size_of_the_8_gpu_benchmark(B, O):
    """A 58^3 member of the T' family = 195 112 rows / 2.9 M entries: what one GPU holds of T' at 8 ranks.  Every CTA's slice
    (~ 20 k entries = 200 KB in the resident layout) must fit, and the solve must agree with the streaming mode and the oracle."""
    blk, n, ptr, col, val = global_csr(B, "stencil15", 58, 14.0)
    B.set_options(tol=1e-8)
    ref = O.solve("bicgstab", n, ptr, col, val, O.spmv(n, ptr, col, val, np.ones(n)), tol=1e-8, max_iter=1000)
    dm = B.DeviceMatrix(blk)
    try:
        run1 = _solve(B, dm, "bicgstab", n, 1)
        run0 = _solve(B, dm, "bicgstab", n, 0)
    finally:
        dm.destroy()
    it1, x1, h1, nres1 = run1
    it0, x0, h0, nres0 = run0
    assert nres0 == 0 and nres1 >= 140                        # (almost) every CTA of the 148 owns rows and keeps them resident
    want = np.sqrt(ref["hist"])
    for it, h in ((it1, h1), (it0, h0)):
        assert np.all(np.abs(h[1:11] - want[1:11]) <= 1e-10 * want[1:11] + 1e-15)
        # iterations-to-tol of this family are chaotic in the summation order (the oracle itself: 146 vs 134 when the entries of
        # every row are merely stored in reverse order), so only a loose bound here; the history above is the sharp check
        assert abs(it - ref["iters"]) <= int(0.2 * ref["iters"]), (it, ref["iters"])
    assert np.abs(x1 - 1).max() < 1e-5 and np.abs(x0 - 1).max() < 1e-5


def test_ctas_with_wide_column_windows_keep_streaming(B, O):
    """Band matrix whose first rows also reference a column 100 000 places away: the CTAs that own those rows cannot express
    their columns as 16-bit offsets and stream their slice, the others keep theirs resident -- in the same launch."""
    n = 150000
    rng = np.random.default_rng(11)
    off = [-300, -1, 1, 300]
    diags = [-(0.2 + 0.6 * rng.random(n - abs(o))) for o in off]
    A = sp.diags(diags, off, shape=(n, n), format="lil")
    far = np.arange(0, 3000)
    for i in far:
        A[i, i + 100000] = -0.5
    A = sp.csr_matrix(A)
    A = sp.csr_matrix(A + sp.diags(np.asarray(abs(A).sum(axis=1)).ravel() + 0.5 + rng.random(n)))   # row sums vary: b = A 1 is no eigenvector
    A.sort_indices()
    blk = B.blocks_from_csr(n, A.indptr, A.indices, A.data)
    ptr, col, val = A.indptr, A.indices, A.data
    x = rng.standard_normal(n)
    ref = O.solve("bicgstab", n, ptr, col, val, O.spmv(n, ptr, col, val, np.ones(n)), tol=TOL, max_iter=1000)
    dm = B.DeviceMatrix(blk)
    try:
        assert rel_err(dm.spmv(x), O.spmv(n, ptr, col, val, x, long_double=True)) <= 1e-13
        run1 = _solve(B, dm, "bicgstab", n, 1)
        run0 = _solve(B, dm, "bicgstab", n, 0)
        run1c = _solve(B, dm, "ca_bicgstab", n, 1)
    finally:
        dm.destroy()
    _agree(run1, run0, ref)
    assert 0 < run1[3] < 148                                  # mixed: some CTAs resident, some streaming
    assert abs(run1c[0] - ref["iters"]) <= max(2, int(0.05 * ref["iters"])) and np.abs(run1c[1] - 1).max() < 1e-6
