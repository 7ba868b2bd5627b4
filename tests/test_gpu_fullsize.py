"""Full-size property tests on BASELINE.json's headline config (configs[1]: 2D-VBC F64, n = 1M,
nnz ≈ 52M, U = W = 4), where the oracle is too slow for an element-wise compare in a unit test:
size-independent properties instead -- linearity, one-hot probes against the generator's closed
form, adjoint identity <Ax, z> == <x, A'z>, and a checksum against the CSC comparator kernel."""
import numpy as np
import pytest

import vbc_b200 as vb
from vbc_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c2():
    A, pi, phi = synth.config_c2()
    B = vb.SparseMatrixVBC[4, 4](A, pi, phi)
    return A, B


def test_c2_sizes(c2):
    A, B = c2
    assert A.shape == (1_000_000, 1_000_000)
    assert B.nval == A.nnz  # uniform dense blocks: no zero fill
    assert B.nidx == A.nnz // 16
    assert 3_100_000 < B.nidx < 3_300_000


def test_c2_onehot_probes_match_generator(c2):
    A, B = c2
    n = A.n
    rng = np.random.default_rng(0)
    for i in rng.integers(0, A.m, size=8):
        e = np.zeros(A.m); e[i] = 1.0
        y = vb.mul_(np.empty(n), B.T, e)  # row i of A
        cols = np.flatnonzero(y)
        expect = synth.entry_values(np.full(len(cols), i, dtype=np.uint64), cols.astype(np.uint64), n)
        assert np.array_equal(y[cols], expect)
        assert len(cols) in range(4 * 7, 4 * 13 + 1)
    for j in rng.integers(0, n, size=8):
        e = np.zeros(n); e[j] = 1.0
        y = vb.mul_(np.empty(A.m), B, e)  # column j of A
        lo, hi = A.colptr[j] - 1, A.colptr[j + 1] - 1
        dense = np.zeros(A.m); dense[A.rowval[lo:hi] - 1] = A.nzval[lo:hi]
        assert np.array_equal(y, dense)


def test_c2_linearity_and_adjoint_identity(c2):
    A, B = c2
    x1, x2, z = synth.vector(A.m, 1), synth.vector(A.m, 2), synth.vector(A.n, 3)
    y1 = vb.mul_(np.empty(A.n), B.T, x1)
    y2 = vb.mul_(np.empty(A.n), B.T, x2)
    y12 = vb.mul_(np.empty(A.n), B.T, 2.0 * x1 - 0.5 * x2)
    assert np.allclose(y12, 2.0 * y1 - 0.5 * y2, rtol=1e-12, atol=1e-12)
    Az = vb.mul_(np.empty(A.m), B, z)
    lhs, rhs = float(np.dot(Az, x1)), float(np.dot(z, y1))
    assert abs(lhs - rhs) <= 1e-11 * abs(rhs)


def test_c2_matches_csc_comparator(c2):
    A, B = c2
    x = synth.vector(A.m, 5)
    y = vb.mul_(np.empty(A.n), B.T, x)
    yc = vb.TrSpMV_(np.empty(A.n), A, x)
    assert np.allclose(y, yc, rtol=1e-12, atol=0)
    S = A.to_scipy()
    assert np.allclose(y, S.T @ x, rtol=1e-12, atol=0)
