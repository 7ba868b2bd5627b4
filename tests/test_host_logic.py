"""CPU-only tests of the host-side logic: partition stand-ins and synthetic generators."""
import numpy as np
import pytest
import scipy.sparse as sp

import vbc_b200 as vb
from vbc_b200 import synth


def test_equi_chunker():
    A = vb.SparseMatrixCSC.from_scipy(sp.random(5, 10, 0.3, random_state=0, format="csc"))
    assert vb.pack_stripe(A, vb.EquiChunker(4)).spl.tolist() == [1, 5, 9, 11]
    assert vb.pack_stripe(A, vb.EquiChunker(5)).spl.tolist() == [1, 6, 11]
    E = vb.SparseMatrixCSC(3, 0, np.array([1], dtype=np.int64), np.array([], dtype=np.int64), np.array([]))
    assert vb.pack_stripe(E, vb.EquiChunker(4)).spl.tolist() == [1]


def test_strict_chunker_groups_identical_patterns():
    # columns: [a a a a a b b c]; w_max = 2 -> (a a)(a a)(a)(b b)(c)
    pat = {"a": [0, 2], "b": [1], "c": []}
    cols = "aaaaabbc"
    rows, cc = [], []
    for j, ch in enumerate(cols):
        for r in pat[ch]:
            rows.append(r); cc.append(j)
    M = sp.csc_matrix((np.ones(len(rows)), (rows, cc)), shape=(3, len(cols)))
    A = vb.SparseMatrixCSC.from_scipy(M)
    assert vb.pack_stripe(A, vb.StrictChunker(2)).spl.tolist() == [1, 3, 5, 6, 8, 9]
    assert vb.pack_stripe(A, vb.StrictChunker(8)).spl.tolist() == [1, 6, 8, 9]
    # same count but different rows must not merge
    M2 = sp.csc_matrix(np.array([[1, 0], [0, 1.0]]))
    assert vb.pack_stripe(vb.SparseMatrixCSC.from_scipy(M2), vb.StrictChunker(4)).spl.tolist() == [1, 2, 3]


@pytest.mark.parametrize("n,w", [(1, 4), (17, 4), (100, 8), (5, 1)])
def test_random_chunker_is_a_valid_split(n, w):
    A = vb.SparseMatrixCSC.from_scipy(sp.random(3, n, 0.5, random_state=1, format="csc"))
    spl = vb.pack_stripe(A, vb.RandomChunker(w, seed=n)).spl
    assert spl[0] == 1 and spl[-1] == n + 1
    d = np.diff(spl)
    assert d.min() >= 1 and d.max() <= w


def test_alternating_packer_partitions_rows_and_columns():
    A = vb.SparseMatrixCSC.from_scipy(sp.random(9, 13, 0.3, random_state=2, format="csc"))
    # reference order: first method -> columns (Φ), second -> rows (Π)  (bin/test_table.jl:89)
    pi, phi = vb.pack_plaid(A, vb.AlternatingPacker(vb.EquiChunker(3), vb.EquiChunker(4)))
    assert pi.spl.tolist() == [1, 5, 9, 10] and phi.spl.tolist() == [1, 4, 7, 10, 13, 14]


def test_synth_banded_blocks_structure():
    A, Pi, Phi = synth.banded_blocks(K=50, L=50, u=2, w=3, offsets=[0, 1, -1, 7, -7])
    assert A.shape == (100, 150) and len(Pi) == 50 and len(Phi) == 50
    S = A.to_scipy().tocoo()
    kb, lb = S.row // 2, S.col // 3
    assert set(np.unique(kb - lb)) == {-7, -1, 0, 1, 7}
    # every block is dense: nnz = (#blocks) * u * w
    nblocks = len(set(zip(kb.tolist(), lb.tolist())))
    assert A.nnz == nblocks * 6
    assert A.nzval.min() >= 0.0 and A.nzval.max() < 1.0
    # regenerating any entry from its coordinates gives the stored value (slab independence)
    v = synth.entry_values(S.row.astype(np.uint64), S.col.astype(np.uint64), 150)
    assert np.array_equal(v, S.data)
    # rows ascending inside every column (CSC invariant the pack kernel relies on)
    for j in range(A.n):
        seg = A.rowval[A.colptr[j] - 1:A.colptr[j + 1] - 1]
        assert np.all(np.diff(seg) > 0)


def test_synth_random_blocks_counts():
    A, Pi, Phi = synth.random_blocks(K=200, L=25, u=1, w=8, per_stripe=10, seed=3)
    assert A.shape == (200, 200) and A.nnz == 25 * 10 * 8
    A2, _, _ = synth.random_blocks(K=6, L=4, u=2, w=2, per_stripe=5, seed=3)
    assert A2.nnz == 4 * 5 * 4


def test_config_c1_shape():
    A, Phi = synth.config_c1()
    assert A.shape == (10_000, 10_000) and A.nnz == 1_000_000 and len(Phi) == 1_250


def test_fem_stencil_has_13_offsets():
    assert len(synth.fem_stencil_offsets(63)) == 13


def test_time_model_fit_recovers_coefficients_1d():
    """costs.jl:112-131 restated in costs.fit_1d: relative LSQ on the one-hot design recovers a planted model."""
    from vbc_b200 import costs
    W = 4
    rng = np.random.default_rng(0)
    a_row, a_col, b_col = 2e-12, np.array([1e-9, 1.5e-9, 2e-9, 2.2e-9]), np.array([3e-10, 4e-10, 6e-10, 7e-10])
    ms, Ls, ws, qs, T = [], [], [], [], []
    for w in range(W, 0, -1):
        L0 = 10_000 * w
        for (m, L, q) in ((L0 * w, L0, 8 * L0), (L0 * w, L0 // 2, 8 * L0), (L0 * w // 2, L0, 8 * L0), (L0 * w, L0, 4 * L0)):
            ms.append(m); Ls.append(L); ws.append(w); qs.append(q)
            T.append(a_row * m + a_col[w - 1] * L + b_col[w - 1] * q)
    ar, ac, bc = costs.fit_1d(W, ms, Ls, ws, qs, T)
    assert np.allclose(ac, a_col, rtol=1e-6) and np.allclose(bc, b_col, rtol=1e-6) and np.isclose(ar, a_row, rtol=1e-6)
    # monotonisation: a dip is lifted to its predecessor
    T2 = list(T)
    ar, ac, bc = costs.fit_1d(W, ms, Ls, ws, qs, [t * (0.5 if w == 3 else 1.0) for t, w in zip(T2, ws)])
    assert np.all(np.diff(ac) >= 0) and np.all(np.diff(bc) >= 0)


def test_time_model_fit_2d_rank_and_monotone():
    from vbc_b200 import costs
    U = W = 3
    br, bcv = np.array([1.0, 1.6, 2.0]), np.array([2e-10, 3e-10, 3.5e-10])  # rank-1 planted beta
    a_row, a_col = np.array([1e-10, 1e-10, 1e-10]), np.array([1e-9, 1.2e-9, 1.4e-9])
    Ks, Ls, us, ws, qs, T = [], [], [], [], [], []
    for u in range(U, 0, -1):
        for w in range(W, 0, -1):
            L0 = 5000
            K0 = L0 * w // u
            for (K, L, q) in ((K0, L0, 8 * L0), (K0, L0 // 2, 8 * L0), (K0 // 2, L0, 8 * L0), (K0, L0, 4 * L0)):
                Ks.append(K); Ls.append(L); us.append(u); ws.append(w); qs.append(q)
                T.append(a_row[u - 1] * K + a_col[w - 1] * L + br[u - 1] * bcv[w - 1] * q)
    ar, ac, brow, bcol, beta = costs.fit_2d(1, U, W, Ks, Ls, us, ws, qs, T)
    assert np.allclose(beta, np.outer(br, bcv), rtol=1e-6)
    recon = np.outer(brow[0], bcol[0])
    assert np.allclose(recon, beta, rtol=1e-6)
    assert np.allclose(ar, a_row, rtol=1e-5) and np.allclose(ac, a_col, rtol=1e-5)


def test_memory_models_match_reference_formulas():
    from vbc_b200 import costs
    m1 = costs.model_SparseMatrix1DVBC_memory(np.float64, np.int64)
    assert m1.stripe_value(4, 10) == 3 * 8 + 10 * (8 + 4 * 8)  # costs.jl:10
    m2 = costs.model_SparseMatrixVBC_memory(np.float32, np.int32)
    assert m2.block_value(4, 4) == 4 + 4 * 4 * 4  # |Ti| + u*w*|Tv|  (costs.jl:140)
    assert costs.model_SparseMatrix1DVBC_blocks().stripe_value(7, 13) == 13  # costs.jl:8


def _brute_force_dp(A, stripe_cost, W):
    """min over all contiguous partitions with widths <= W of sum stripe_cost(j0, j1); exponential-free O(n W) reference."""
    n = A.n
    best = [0.0] + [None] * n
    for b in range(1, n + 1):
        best[b] = min(best[b - w] + stripe_cost(b - w, b) for w in range(1, min(W, b) + 1))
    return best[n]


def test_window_distinct_counts_and_dynamic_total_chunker():
    from vbc_b200 import costs
    from vbc_b200.partition import DynamicTotalChunker, window_weight_sums
    rng = np.random.default_rng(3)
    M = sp.random(40, 31, 0.15, random_state=5, format="csc")
    A = vb.SparseMatrixCSC.from_scipy(M)
    S = M.tocsc()
    W = 5
    D = window_weight_sums(A, W)[:, :, 0]
    for b in range(0, A.n + 1):
        for w in range(1, min(W, b) + 1):
            assert D[b, w - 1] == len(np.unique(S[:, b - w:b].tocoo().row)), (b, w)
    # 1D: block-count and memory models against an exhaustive recurrence with set-based stripe costs
    for mdl in (costs.model_SparseMatrix1DVBC_blocks(), costs.model_SparseMatrix1DVBC_memory(np.float64, np.int64)):
        phi = vb.pack_stripe(A, DynamicTotalChunker(mdl, W))
        spl = phi.spl.astype(np.int64) - 1
        assert spl[0] == 0 and spl[-1] == A.n and np.all(np.diff(spl) >= 1) and np.diff(spl).max() <= W
        cost = lambda a, b: mdl.stripe_value(b - a, len(np.unique(S[:, a:b].tocoo().row)))
        got = sum(cost(a, b) for a, b in zip(spl[:-1], spl[1:]))
        assert np.isclose(got, _brute_force_dp(A, cost, W))
    # the memory model never does worse than Strict or Equi partitions of the same width limit
    mem = costs.model_SparseMatrix1DVBC_memory(np.float64, np.int64)
    def total(phi):
        s_ = phi.spl.astype(np.int64) - 1
        return sum(mem.stripe_value(b - a, len(np.unique(S[:, a:b].tocoo().row))) for a, b in zip(s_[:-1], s_[1:]))
    best = total(vb.pack_stripe(A, DynamicTotalChunker(mem, W)))
    assert best <= total(vb.pack_stripe(A, vb.StrictChunker(W))) and best <= total(vb.pack_stripe(A, vb.EquiChunker(W)))
    # 2D: columns given a row partition, weighted distinct-part sums
    pi = vb.pack_stripe(A.transpose(), vb.RandomChunker(4, 2))
    m2 = costs.model_SparseMatrixVBC_memory(np.float64, np.int64)
    phi2 = vb.pack_stripe(A, DynamicTotalChunker(m2, W), other=pi)
    ps = pi.spl.astype(np.int64) - 1
    heights = np.diff(ps)
    def cost2(a, b):
        parts = np.unique(np.searchsorted(ps, S[:, a:b].tocoo().row, side="right") - 1)
        return 3 * 8 + sum(m2.block_value(heights[k], b - a) for k in parts)
    s2 = phi2.spl.astype(np.int64) - 1
    assert np.isclose(sum(cost2(a, b) for a, b in zip(s2[:-1], s2[1:])), _brute_force_dp(A, cost2, W))
    # AlternatingPacker in the reference's order: columns, rows | columns, columns | rows ...
    pi3, phi3 = vb.pack_plaid(A, vb.AlternatingPacker(DynamicTotalChunker(costs.model_SparseMatrix1DVBC_blocks(), W), vb.EquiChunker(1)))
    assert len(pi3) == A.m and np.diff(phi3.spl).max() <= W
    pi4, phi4 = vb.pack_plaid(A, vb.AlternatingPacker(vb.EquiChunker(1), vb.EquiChunker(1), DynamicTotalChunker(m2, W),
                                                       DynamicTotalChunker(vb.partition.permutedims(m2), 4), DynamicTotalChunker(m2, W)))
    assert np.diff(pi4.spl).max() <= 4 and np.diff(phi4.spl).max() <= W and pi4.spl[-1] == A.m + 1 and phi4.spl[-1] == A.n + 1


def test_overlap_chunker_assumed_definition():
    # columns: a a' b b c   with a' = a plus one extra row (Jaccard 2/3), b identical twice, c empty
    pats = [[0, 2], [0, 2, 5], [1, 3], [1, 3], []]
    rows, cc = [], []
    for j, p in enumerate(pats):
        for r in p:
            rows.append(r); cc.append(j)
    A = vb.SparseMatrixCSC.from_scipy(sp.csc_matrix((np.ones(len(rows)), (rows, cc)), shape=(6, len(pats))))
    assert vb.pack_stripe(A, vb.OverlapChunker(0.9, 4)).spl.tolist() == [1, 2, 3, 5, 6]   # only the identical pair merges
    assert vb.pack_stripe(A, vb.OverlapChunker(0.6, 4)).spl.tolist() == [1, 3, 5, 6]      # a, a' merge at 2/3 >= 0.6
    assert vb.pack_stripe(A, vb.OverlapChunker(0.0, 2)).spl.tolist() == [1, 3, 5, 6]      # width limit
    assert vb.pack_stripe(A, vb.OverlapChunker(1.0, 8)).spl.tolist() == vb.pack_stripe(A, vb.StrictChunker(8)).spl.tolist()
    M = sp.random(60, 200, 0.1, random_state=3, format="csc")
    B = vb.SparseMatrixCSC.from_scipy(M)
    spl = vb.pack_stripe(B, vb.OverlapChunker(0.3, 4)).spl
    assert spl[0] == 1 and spl[-1] == 201 and np.diff(spl).min() >= 1 and np.diff(spl).max() <= 4
