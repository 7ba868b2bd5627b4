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


def test_c3_spmm_k32_fullsize():
    """configs[2]: 1D-VBC SpMM, Float64, k = 32, n = 1M, W = 8, 50 rows per stripe (nnz = 50M): every column of
    Y = B'X must equal the single-vector multiply of that column, and a one-hot panel must reproduce A's rows."""
    import torch
    K, L = 1_000_000, 125_000
    A, _, phi = synth.banded_blocks(K, L, 1, 8, np.arange(-25, 25) * 37)
    assert A.nnz > 49_000_000
    B = vb.SparseMatrix1DVBC[8](A, phi)
    k = 32
    X = torch.rand(A.m, k, dtype=torch.float64, device="cuda")
    Y = torch.empty(A.n, k, dtype=torch.float64, device="cuda")
    vb.mul_(Y, B.T, X)
    for c in (0, 17, 31):
        y = torch.empty(A.n, dtype=torch.float64, device="cuda")
        vb.mul_(y, B.T, X[:, c].contiguous())
        assert torch.allclose(Y[:, c], y, rtol=1e-12, atol=1e-12)
    # Julia's column-major layout gives the same panel
    Xc = X.t().contiguous().t()
    Yc = torch.empty(k, A.n, dtype=torch.float64, device="cuda").t()
    vb.mul_(Yc, B.T, Xc)
    assert torch.equal(Yc, Y)
    # forward orientation, linearity in the panel
    Z = torch.rand(A.n, 4, dtype=torch.float64, device="cuda")
    F1 = vb.mul_(torch.empty(A.m, 4, dtype=torch.float64, device="cuda"), B, Z)
    F2 = vb.mul_(torch.empty(A.m, 4, dtype=torch.float64, device="cuda"), B, 2.0 * Z)
    assert torch.allclose(F2, 2.0 * F1, rtol=1e-12, atol=1e-12)


def test_c4_triangular_solve_fullsize():
    """configs[3], BASELINE wording: lower-triangular VBC solve, Float64, n = 2M.  Residual of tril(A') x = b
    against scipy's CSC product (the reference has no solve: parity unpinned), plus the reference meaning of
    "TrSpMV" on the same matrix: the adjoint multiply vs the CSC comparator."""
    import scipy.sparse as sp
    A, pi, phi = synth.config_c4_triangular()
    assert A.shape == (2_000_000, 2_000_000)
    B = vb.SparseMatrixVBC[4, 4](A, pi, phi)
    levels = vb.trsv_analyse(B.T)
    assert 1000 < levels < 20000
    b = synth.vector(A.n, 9)
    x = vb.ldiv_lower_(np.empty(A.n), B.T, b)
    T = sp.tril(A.to_scipy().T.tocsr()).tocsr()
    r = T @ x - b
    assert np.max(np.abs(r)) <= 1e-12 * np.max(np.abs(b)) * 200
    xt = synth.vector(A.m, 10)
    y = vb.mul_(np.empty(A.n), B.T, xt)
    assert np.allclose(y, vb.TrSpMV_(np.empty(A.n), A, xt), rtol=1e-12, atol=0)


def test_c5_slab_float32():
    """configs[4] per-rank shape: 2D-VBC Float32 / Int32, U = W = 4, 10 blocks per stripe (here n = 2M, one slab):
    one-hot probes against the generator, linearity, agreement with the CSC comparator within 1e-5."""
    n = 2_000_000
    offs = [0, 1, -1, 2, -2, 57, -57, 58, -58, 3249]
    A, pi, phi = synth.banded_blocks(n // 4, n // 4, 4, 4, offs, dtype=np.float32, ti=np.int32)
    B = vb.SparseMatrixVBC[4, 4](A, pi, phi)
    assert B.Tv == np.float32 and B.Ti == np.int32 and B.nval == A.nnz
    rng = np.random.default_rng(1)
    for i in rng.integers(0, A.m, size=4):
        e = np.zeros(A.m, dtype=np.float32); e[i] = 1.0
        y = vb.mul_(np.empty(A.n, dtype=np.float32), B.T, e)
        cols = np.flatnonzero(y)
        assert np.array_equal(y[cols], synth.entry_values(np.full(len(cols), i, dtype=np.uint64), cols.astype(np.uint64), n, dtype=np.float32))
    x = synth.vector(A.m, 3, dtype=np.float32)
    y = vb.mul_(np.empty(A.n, dtype=np.float32), B.T, x)
    yc = vb.TrSpMV_(np.empty(A.n, dtype=np.float32), A, x)
    assert np.allclose(y, yc, rtol=1e-5, atol=1e-6)
    y2 = vb.mul_(np.empty(A.n, dtype=np.float32), B.T, (0.5 * x).astype(np.float32))
    assert np.allclose(y2, 0.5 * y, rtol=1e-6, atol=1e-7)
