"""Pins the oracle (oracle/vbc_oracle.c) against everything the reference's own tests hold for
this path -- CPU only.

  * test/matrices.jl:4-9       six literal matrices (tests/golden/fixtures.npz)
  * test/runtests.jl:29-53     1D one-hot invariants, exact ==, both orientations
  * test/runtests.jl:63-87     2D one-hot invariants
  * test/runtests.jl:14-16     the 11x11 size grid of random matrices (f64 / Bool / Int32 values)
  * bin/test_table.jl:42,:84   random-x isapprox against the CSC product
  * SURVEY.md Appendix A / B   structural known answers (independent restatement)
"""
import json
import os

import numpy as np
import pytest

import oracle
import vbc_b200 as vb
from conftest import ROOT, SIZES, sprand, sprand_typed

KA = json.load(open(os.path.join(ROOT, "tests", "golden", "known_answers.json")))
FIXTURE_NAMES = [k for k in KA if not k.startswith("_") and k != "appendix_a"]


def partitions_1d(A):
    yield "strict4", vb.pack_stripe(A, vb.StrictChunker(4)), True
    yield "equi4", vb.pack_stripe(A, vb.EquiChunker(4)), False
    yield "rand4", vb.pack_stripe(A, vb.RandomChunker(4, seed=A.nnz)), False


def partitions_2d(A):
    for name, ch in (("strict4", vb.StrictChunker(4)), ("equi4", vb.EquiChunker(4)),
                     ("rand4", vb.RandomChunker(4, seed=A.nnz + 1))):
        pi, phi = vb.pack_plaid(A, vb.AlternatingPacker(ch, ch))
        yield name, pi, phi


def onehot_check(A, B):
    """runtests.jl:29-53: for every unit vector, mul!(y, B, e) == mul!(y, A, e), exactly, and the
    same for the adjoint."""
    m, n = A.shape
    tv = A.nzval.dtype
    x = np.zeros(n, dtype=tv)
    for j in range(n):
        x[j] = 1
        y_ref = oracle.csc_spmv(m, n, A.colptr, A.rowval, A.nzval, x)
        y_test = oracle.mul(B, x, trans=False, alpha=True, beta=False, y=np.full(m, 7, dtype=tv))
        assert np.array_equal(y_ref, y_test), f"forward, column {j + 1}"
        x[j] = 0
    x = np.zeros(m, dtype=tv)
    for i in range(m):
        x[i] = 1
        y_ref = oracle.csc_trspmv(m, n, A.colptr, A.rowval, A.nzval, x)
        y_test = oracle.mul(B, x, trans=True, alpha=True, beta=False, y=np.full(n, 7, dtype=tv), nthreads=2)
        assert np.array_equal(y_ref, y_test), f"adjoint, row {i + 1}"
        x[i] = 0


@pytest.mark.parametrize("name", FIXTURE_NAMES)
def test_fixture_known_answers(fixtures, name):
    A = fixtures[name]
    ka = KA[name]
    for key, ch in (("equi", vb.EquiChunker(4)), ("strict", vb.StrictChunker(4))):
        phi = vb.pack_stripe(A, ch)
        B = oracle.pack_1d(A.m, A.n, A.colptr, A.rowval, A.nzval, phi.spl, 4)
        assert [B.L, len(B.idx), len(B.val)] == ka[f"{key}1d"]
        pi, phi2 = vb.pack_plaid(A, vb.AlternatingPacker(ch, ch))
        C = oracle.pack_2d(A.m, A.n, A.colptr, A.rowval, A.nzval, pi.spl, phi2.spl, 4, 4)
        assert [C.K, C.L, len(C.idx), len(C.val)] == ka[f"{key}2d"]


def test_appendix_a_worked_example():
    a = KA["appendix_a"]
    i64, f64 = np.int64, np.float64
    colptr, rowval = np.array(a["colptr"], i64), np.array(a["rowval"], i64)
    nzval = np.array(a["nzval"], f64)
    B = oracle.pack_1d(a["m"], a["n"], colptr, rowval, nzval, np.array(a["phi_spl"], i64), 4)
    for f in ("pos", "idx", "ofs", "val"):
        assert np.array_equal(getattr(B, f), np.array(a["d1"][f])), f
    C = oracle.pack_2d(a["m"], a["n"], colptr, rowval, nzval, np.array(a["pi_spl"], i64),
                       np.array(a["phi_spl"], i64), 4, 4)
    for f in ("pos", "idx", "ofs", "val"):
        assert np.array_equal(getattr(C, f), np.array(a["d2"][f])), f
    # empty stripes still store zeros into their y slice in the adjoint (Appendix A note)
    y = oracle.mul(B, np.ones(6), trans=True, y=np.full(7, 5.0))
    assert np.array_equal(y, np.array([4, 6, 0, 0, 19, 7, 9], dtype=f64))


@pytest.mark.parametrize("name", FIXTURE_NAMES)
def test_fixture_onehot_1d(fixtures, name):
    A = fixtures[name]
    for pname, phi, strict in partitions_1d(A):
        B = oracle.pack_1d(A.m, A.n, A.colptr, A.rowval, A.nzval, phi.spl, 4)
        if strict:  # the StrictChunker constructor (constructors_1DVBC.jl:94-143) must agree
            B2 = oracle.pack_1d(A.m, A.n, A.colptr, A.rowval, A.nzval, phi.spl, 4, strict=True)
            for f in ("pos", "idx", "ofs", "val"):
                assert np.array_equal(getattr(B, f), getattr(B2, f)), (pname, f)
        onehot_check(A, B)


@pytest.mark.parametrize("name", FIXTURE_NAMES)
def test_fixture_onehot_2d(fixtures, name):
    A = fixtures[name]
    for pname, pi, phi in partitions_2d(A):
        B = oracle.pack_2d(A.m, A.n, A.colptr, A.rowval, A.nzval, pi.spl, phi.spl, 4, 4)
        onehot_check(A, B)


@pytest.mark.parametrize("kind", ["f64", "bool", "int32"])
def test_size_grid_onehot(kind):
    """runtests.jl:14-16: m, n in SIZES, density 0.2 (one trial per cell here; four in the reference)."""
    rng = np.random.default_rng(0xDEADBEEF)
    for m in SIZES:
        for n in SIZES:
            A = sprand(m, n, 0.2, rng, kind)
            for _, phi, _s in partitions_1d(A):
                onehot_check(A, oracle.pack_1d(m, n, A.colptr, A.rowval, A.nzval, phi.spl, 4))
            for _, pi, phi in partitions_2d(A):
                onehot_check(A, oracle.pack_2d(m, n, A.colptr, A.rowval, A.nzval, pi.spl, phi.spl, 4, 4))


@pytest.mark.parametrize("tv", [np.int32, np.int64])
@pytest.mark.parametrize("ti", [np.int64, np.int32])
def test_integer_element_types_wrap_like_julia(tv, ti):
    """runtests.jl:16 (`sprand(Int32, ...)`) with the element type kept: the oracle's integer instantiations pack the same arrays as
    the floating-point ones and multiply with wrapping arithmetic (checked against numpy's modular integer arithmetic)."""
    rng = np.random.default_rng(16)
    info = np.iinfo(tv)
    for m, n in [(1, 1), (5, 9), (17, 16), (40, 33)]:
        A = sprand_typed(m, n, 0.3, rng, tv, ti)
        D = np.zeros((m, n), dtype=tv)
        for j in range(n):
            for q in range(A.colptr[j] - 1, A.colptr[j + 1] - 1):
                D[A.rowval[q] - 1, j] = A.nzval[q]
        Af = vb.SparseMatrixCSC(m, n, A.colptr, A.rowval, np.arange(1, A.nnz + 1, dtype=np.float64))
        phi = vb.pack_stripe(A, vb.RandomChunker(4, seed=m + n))
        pi, phi2 = vb.pack_plaid(A, vb.AlternatingPacker(vb.RandomChunker(4, seed=m), vb.RandomChunker(4, seed=n)))
        H1 = oracle.pack_1d(m, n, A.colptr, A.rowval, A.nzval, phi.spl, 4)
        H2 = oracle.pack_2d(m, n, A.colptr, A.rowval, A.nzval, pi.spl, phi2.spl, 4, 4)
        F1 = oracle.pack_1d(m, n, Af.colptr, Af.rowval, Af.nzval, phi.spl, 4)
        F2 = oracle.pack_2d(m, n, Af.colptr, Af.rowval, Af.nzval, pi.spl, phi2.spl, 4, 4)
        for H, F in ((H1, F1), (H2, F2)):
            assert H.val.dtype == np.dtype(tv)
            for f in ("pos", "idx", "ofs"):
                assert np.array_equal(getattr(H, f), getattr(F, f)), f
            # the value slots hold the nonzeros in the same places (F carries the nonzero's ordinal)
            slot = F.val.astype(np.int64)
            assert np.array_equal(H.val[slot > 0], A.nzval[slot[slot > 0] - 1]) and not H.val[slot == 0].any()
            with np.errstate(over="ignore"):
                x = rng.integers(info.min, info.max, size=n, dtype=tv, endpoint=True)
                assert np.array_equal(oracle.mul(H, x), D @ x)
                xt = rng.integers(info.min, info.max, size=m, dtype=tv, endpoint=True)
                assert np.array_equal(oracle.mul(H, xt, trans=True), D.T @ xt)
                assert np.array_equal(oracle.csc_trspmv(m, n, A.colptr, A.rowval, A.nzval, xt), D.T @ xt)
                assert np.array_equal(oracle.csc_spmv(m, n, A.colptr, A.rowval, A.nzval, x), D @ x)


@pytest.mark.parametrize("tv,ti,tol", [(np.float64, np.int64, 1e-12), (np.float64, np.int32, 1e-12),
                                        (np.float32, np.int64, 1e-5), (np.float32, np.int32, 1e-5)])
def test_random_x_isapprox_all_types(fixtures, tv, ti, tol):
    """bin/test_table.jl:42/:84/:126 `@assert y ≈ z`, for every (Tv, Ti) instantiation."""
    rng = np.random.default_rng(7)
    for name, A64 in fixtures.items():
        A = A64.astype(tv, ti)
        S = A64.to_scipy()
        absS = abs(S)
        x, xt = rng.random(A.n), rng.random(A.m)
        phi = vb.pack_stripe(A, vb.RandomChunker(4, 3))
        pi, phi2 = vb.pack_plaid(A, vb.AlternatingPacker(vb.RandomChunker(4, 5), vb.RandomChunker(4, 6)))
        for B in (oracle.pack_1d(A.m, A.n, A.colptr, A.rowval, A.nzval, phi.spl, 4),
                  oracle.pack_2d(A.m, A.n, A.colptr, A.rowval, A.nzval, pi.spl, phi2.spl, 4, 4)):
            assert B.pos.dtype == ti and B.val.dtype == tv
            y = oracle.mul(B, x.astype(tv))
            assert np.all(np.abs(y - S @ x) <= 4 * tol * (absS @ x) + 1e-300)
            yt = oracle.mul(B, xt.astype(tv), trans=True, nthreads=3)
            assert np.all(np.abs(yt - S.T @ xt) <= 4 * tol * (absS.T @ xt) + 1e-300)


def test_reference_quirks_alpha_beta():
    """forward computes y <- A x + beta y (alpha ignored); adjoint overwrites y."""
    A = sprand(9, 7, 0.4, np.random.default_rng(1))
    phi = vb.pack_stripe(A, vb.EquiChunker(3))
    B = oracle.pack_1d(A.m, A.n, A.colptr, A.rowval, A.nzval, phi.spl, 3)
    S = A.to_scipy()
    x, y0 = np.arange(1.0, 8.0), np.arange(1.0, 10.0)
    y = oracle.mul(B, x, alpha=5.0, beta=2.0, y=y0.copy())
    assert np.allclose(y, S @ x + 2.0 * y0)
    xt, yt0 = np.arange(1.0, 10.0), np.arange(1.0, 8.0)
    yt = oracle.mul(B, xt, trans=True, alpha=5.0, beta=2.0, y=yt0.copy())
    assert np.allclose(yt, S.T @ xt)


def test_errors():
    A = sprand(8, 8, 0.3, np.random.default_rng(2))
    phi = vb.pack_stripe(A, vb.EquiChunker(4))
    with pytest.raises(oracle.OracleError) as e:  # @assert w <= W  constructors_1DVBC.jl:46
        oracle.pack_1d(8, 8, A.colptr, A.rowval, A.nzval, phi.spl, 3)
    assert e.value.code == 1
    with pytest.raises(oracle.OracleError) as e:  # @assert u <= U  constructors_VBC.jl:58-60
        oracle.pack_2d(8, 8, A.colptr, A.rowval, A.nzval, phi.spl, phi.spl, 2, 4)
    assert e.value.code == 2
    B = oracle.pack_1d(8, 8, A.colptr, A.rowval, A.nzval, phi.spl, 4)
    with pytest.raises(oracle.OracleError) as e:  # DimensionMismatch multiply_1DVBC.jl:44-45
        oracle.mul(B, np.zeros(7))
    assert e.value.code == 3
    with pytest.raises(oracle.OracleError):  # TrSpMV.jl:3-4
        oracle.csc_trspmv(8, 8, A.colptr, A.rowval, A.nzval, np.zeros(5))


def test_memory_cost_models(fixtures):
    """costs.jl:10 / :140 on packed fixtures == the reference's format accounting minus the
    per-array +1 entries (test_table.jl:78: sizeof(Φ)+sizeof(pos)+sizeof(idx)+sizeof(ofs)+sizeof(val))."""
    A = fixtures["HB__west0132"]
    phi = vb.pack_stripe(A, vb.EquiChunker(4))
    B = oracle.pack_1d(A.m, A.n, A.colptr, A.rowval, A.nzval, phi.spl, 4)
    cost, extra = oracle.memory_cost(B)
    assert extra == 0
    assert cost.sum() == 3 * 8 * B.L + 8 * len(B.idx) + 8 * len(B.val)
    pi, phi2 = vb.pack_plaid(A, vb.AlternatingPacker(vb.EquiChunker(4), vb.EquiChunker(4)))
    C = oracle.pack_2d(A.m, A.n, A.colptr, A.rowval, A.nzval, pi.spl, phi2.spl, 4, 4)
    cost, extra = oracle.memory_cost(C)
    assert extra == 8 * C.K
    assert cost.sum() == 3 * 8 * C.L + 8 * len(C.idx) + 8 * len(C.val)


def reference_methods_1d():
    """The method list of test/runtests.jl:19-25 (W = 4)."""
    from vbc_b200 import costs
    return [vb.StrictChunker(4), vb.OverlapChunker(0.9, 4),
            vb.DynamicTotalChunker(costs.model_SparseMatrix1DVBC_blocks(), 4),
            vb.DynamicTotalChunker(costs.model_SparseMatrix1DVBC_memory(np.float64, np.int64), 4)]


def reference_methods_2d():
    """test/runtests.jl:56-59 (U = W = 4)."""
    return [vb.AlternatingPacker(vb.StrictChunker(4), vb.StrictChunker(4)),
            vb.AlternatingPacker(vb.OverlapChunker(0.9, 4), vb.OverlapChunker(0.9, 4))]


@pytest.mark.parametrize("name", FIXTURE_NAMES)
def test_runtests_jl_method_list_on_the_oracle(fixtures, name):
    """test/runtests.jl:19-88 with its own partitioner list (stand-ins for the un-vendored ChainPartitioners; the
    invariants are partition-agnostic, so they pin the pack + multiply, not the partitions)."""
    A = fixtures[name]
    for method in reference_methods_1d():
        phi = vb.pack_stripe(A, method)
        assert np.diff(phi.spl).max() <= 4
        onehot_check(A, oracle.pack_1d(A.m, A.n, A.colptr, A.rowval, A.nzval, phi.spl, 4))
    for method in reference_methods_2d():
        pi, phi = vb.pack_plaid(A, method)
        onehot_check(A, oracle.pack_2d(A.m, A.n, A.colptr, A.rowval, A.nzval, pi.spl, phi.spl, 4, 4))


# ---- the format DEFINITION, stated with sets instead of merges, against the oracle's packers (property test) ----------
def _split(rng_ints, dim, cap):
    """split points of a random contiguous partition of 1..dim with parts <= cap (from a list of random ints)"""
    spl, pos, i = [1], 0, 0
    while pos < dim:
        step = 1 + rng_ints[i % len(rng_ints)] % cap
        i += 1
        pos = min(dim, pos + step)
        spl.append(pos + 1)
    return np.array(spl)


def _definition_pack(m, n, colptr, rowval, nzval, pi_spl, phi_spl):
    """SparseMatrixVBCs.jl:36-43 / :62-70 read as a definition: per stripe, the ascending distinct rows (1D) or row
    parts (2D) that hold a STORED entry; every unit is a dense row-major block, zero where nothing is stored."""
    dense = {}
    for j in range(n):
        for t in range(colptr[j] - 1, colptr[j + 1] - 1):
            dense[(int(rowval[t]), j + 1)] = nzval[t]
    part_of = None
    if pi_spl is not None:
        part_of = {}
        for k in range(len(pi_spl) - 1):
            for i in range(pi_spl[k], pi_spl[k + 1]):
                part_of[int(i)] = k + 1
    pos, ofs, idx, val = [1], [1], [], []
    for l in range(len(phi_spl) - 1):
        cols = range(int(phi_spl[l]), int(phi_spl[l + 1]))
        rows = sorted({i for (i, j) in dense if j in cols})
        units = rows if part_of is None else sorted({part_of[i] for i in rows})
        for uid in units:
            idx.append(uid)
            urows = [uid] if part_of is None else range(int(pi_spl[uid - 1]), int(pi_spl[uid]))
            for i in urows:
                for j in cols:
                    val.append(dense.get((int(i), j), 0))
        pos.append(len(idx) + 1)
        ofs.append(len(val) + 1)
    return np.array(pos), np.array(idx, dtype=np.int64), np.array(ofs), np.array(val, dtype=np.float64)


def test_packers_match_the_set_definition_of_the_format():
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=120, deadline=None, derandomize=True)
    @given(m=st.integers(1, 13), n=st.integers(1, 13), seed=st.integers(0, 2**31 - 1), density=st.sampled_from([0.0, 0.1, 0.3, 0.7, 1.0]),
           ints=st.lists(st.integers(0, 1000), min_size=4, max_size=12), cap=st.integers(1, 5), ti=st.sampled_from([np.int64, np.int32]))
    def check(m, n, seed, density, ints, cap, ti):
        rng = np.random.default_rng(seed)
        A = sprand(m, n, density, rng)
        colptr, rowval = A.colptr.astype(ti), A.rowval.astype(ti)
        phi = _split(ints, n, cap).astype(ti)
        pi = _split(ints[::-1], m, cap).astype(ti)
        H1 = oracle.pack_1d(m, n, colptr, rowval, A.nzval, phi, cap)
        p, i, o, v = _definition_pack(m, n, colptr, rowval, A.nzval, None, phi)
        assert np.array_equal(H1.pos, p) and np.array_equal(H1.idx, i) and np.array_equal(H1.ofs, o) and np.array_equal(H1.val, v)
        H2 = oracle.pack_2d(m, n, colptr, rowval, A.nzval, pi, phi, cap, cap)
        p, i, o, v = _definition_pack(m, n, colptr, rowval, A.nzval, pi, phi)
        assert np.array_equal(H2.pos, p) and np.array_equal(H2.idx, i) and np.array_equal(H2.ofs, o) and np.array_equal(H2.val, v)
        assert H1.pos.dtype == np.dtype(ti) and H2.idx.dtype == np.dtype(ti)

    check()
