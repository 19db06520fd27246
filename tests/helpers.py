"""Shared test inputs: small members of the BASELINE.json matrix families (csrc/gen.cpp)."""
import numpy as np

# (name, kind, g, p0): sizes the oracle finishes in well under a second
SMALL_CASES = [
    ("stencil15_g12", "stencil15", 12, 14.0),      # Transport-like T', not diagonally dominant
    ("convdiff_g40", "convdiff", 40, 1.5),         # nonsymmetric convection-diffusion
    ("laplace5_g37", "laplace5", 37, 0.0),         # cfg 3 family, odd n
    ("random_n3001_k8", "random", 3001, 8),        # cfg 5 family, odd n
]
METHODS = ["bicgstab", "ca_bicgstab", "pipe_bicgstab", "pipe_bicgstab_rr"]
RR = dict(krr=10, nrr=3)


def global_csr(B, kind, g, p0, seed=12345):
    blk = B.gen_block(kind, g, p0, seed)
    ptr, col, val = B.block_to_global_csr(blk)
    return blk, blk.n, ptr, col, val


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)
