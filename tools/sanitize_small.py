#!/usr/bin/env python
"""Small end-to-end exercise of every kernel family for `compute-sanitizer --tool memcheck` (one tool per
gpurun call, smallest case that shows what is needed -- B200_PROFILING.md)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vbc_b200 as vb
from vbc_b200 import synth

rng = np.random.default_rng(0)
A, pi, phi = synth.config_c2(n=4000, S=7)
mats = [vb.SparseMatrixVBC[4, 4](A, pi, phi), vb.SparseMatrix1DVBC[4](A, phi),
        vb.SparseMatrixVBC[8, 8](A, synth.variable_partition(A.m, 8, 1), synth.variable_partition(A.n, 8, 2)),
        vb.SparseMatrix1DVBC[5](A, vb.RandomChunker(5, 3))]
S = A.to_scipy()
for B in mats:
    x, xt = rng.random(A.n), rng.random(A.m)
    assert np.allclose(vb.mul_(np.empty(A.m), B, x), S @ x)
    assert np.allclose(vb.mul_(np.empty(A.n), B.T, xt), S.T @ xt)
    X = rng.random((A.m, 5))
    assert np.allclose(vb.mul_(np.empty((A.n, 5)), B.T, X), S.T @ X)
    Xn = rng.random((A.n, 5))
    assert np.allclose(vb.mul_(np.empty((A.m, 5)), B, Xn), S @ Xn)
assert np.allclose(vb.TrSpMV_(np.empty(A.n), A, xt), S.T @ xt)
T, tpi, tphi = synth.config_c4_triangular(n=4000, S=7)
Bt = vb.SparseMatrixVBC[4, 4](T, tpi, tphi)
b = rng.random(T.n)
xs = vb.ldiv_lower_(np.empty(T.n), Bt.T, b)
import scipy.sparse as sp
assert np.allclose(sp.tril(T.to_scipy().T) @ xs, b)
print("sanitize_small ok")
