"""GPU parity tests: the CUDA path (through the C ABI of libvbc.so) against the CPU oracle.

Bar (BASELINE.json north_star): packed index arrays bit-exact; y within 1e-12 relative (Float64) /
1e-5 (Float32), measured componentwise against |A||x| because summation order differs.
"""
import os

import numpy as np
import pytest

import oracle
import vbc_b200 as vb
from conftest import SIZES, sprand, sprand_typed
from vbc_b200 import _lib, synth

pytestmark = pytest.mark.gpu

TOL = {np.dtype(np.float64): 1e-12, np.dtype(np.float32): 1e-5}


def parts_1d(A):
    return [("strict4", vb.pack_stripe(A, vb.StrictChunker(4))), ("equi4", vb.pack_stripe(A, vb.EquiChunker(4))),
            ("rand4", vb.pack_stripe(A, vb.RandomChunker(4, seed=A.nnz)))]


def parts_2d(A):
    out = []
    for name, ch in (("strict4", vb.StrictChunker(4)), ("equi4", vb.EquiChunker(4)), ("rand4", vb.RandomChunker(4, seed=A.nnz + 1))):
        pi, phi = vb.pack_plaid(A, vb.AlternatingPacker(ch, ch))
        out.append((name, pi, phi))
    return out


def assert_packed_equal(B, H):
    d = B.download()
    for f in ("pos", "idx", "ofs"):
        assert d[f].dtype == getattr(H, f).dtype
        assert np.array_equal(d[f], getattr(H, f)), f
    assert d["val"].dtype == H.val.dtype
    assert d["val"].tobytes() == H.val.tobytes(), "val (bitwise)"


def onehot_check(A, B, H):
    """test/runtests.jl:29-53 / :63-87 with the device matrix in place of B; also == the oracle."""
    m, n = A.shape
    tv = A.nzval.dtype
    x = np.zeros(n, dtype=tv)
    y = np.empty(m, dtype=tv)
    for j in range(n):
        x[j] = 1
        y.fill(7)
        vb.mul_(y, B, x, True, False)
        assert np.array_equal(y, oracle.csc_spmv(m, n, A.colptr, A.rowval, A.nzval, x)), f"forward col {j + 1}"
        x[j] = 0
    x = np.zeros(m, dtype=tv)
    y = np.empty(n, dtype=tv)
    for i in range(m):
        x[i] = 1
        y.fill(7)
        vb.mul_(y, B.T, x, True, False)
        assert np.array_equal(y, oracle.csc_trspmv(m, n, A.colptr, A.rowval, A.nzval, x)), f"adjoint row {i + 1}"
        assert np.array_equal(y, oracle.mul(H, x, trans=True)), f"adjoint row {i + 1} vs oracle"
        x[i] = 0


def randx_check(A, B, H, rng):
    tv = A.nzval.dtype
    tol = TOL[np.dtype(tv)]
    S = A.to_scipy().astype(np.float64)
    absS = abs(S)
    for trans in (False, True):
        xlen, ylen = (A.m, A.n) if trans else (A.n, A.m)
        x = rng.random(xlen).astype(tv)
        y = vb.mul_(np.empty(ylen, dtype=tv), B.T if trans else B, x)
        yo = oracle.mul(H, x, trans=trans)
        Sx = (S.T if trans else S) @ x.astype(np.float64)
        bound = (absS.T if trans else absS) @ np.abs(x.astype(np.float64))
        assert np.all(np.abs(y - Sx) <= 4 * tol * bound + 1e-300), "vs exact product"
        assert np.all(np.abs(y - yo) <= 8 * tol * bound + 1e-300), "vs oracle"


@pytest.mark.parametrize("ti", [np.int64, np.int32])
@pytest.mark.parametrize("tv", [np.float64, np.float32])
def test_fixtures_pack_bitexact_and_multiply(fixtures, tv, ti):
    rng = np.random.default_rng(11)
    for name, A64 in fixtures.items():
        A = A64.astype(tv, ti)
        for pname, phi in parts_1d(A):
            H = oracle.pack_1d(A.m, A.n, A.colptr, A.rowval, A.nzval, phi.spl, 4)
            B = vb.SparseMatrix1DVBC[4](A, phi)
            assert B.shape == A.shape and B.L == H.L
            assert_packed_equal(B, H)
            randx_check(A, B, H, rng)
        for pname, pi, phi in parts_2d(A):
            H = oracle.pack_2d(A.m, A.n, A.colptr, A.rowval, A.nzval, pi.spl, phi.spl, 4, 4)
            B = vb.SparseMatrixVBC[4, 4](A, pi, phi)
            assert_packed_equal(B, H)
            randx_check(A, B, H, rng)


@pytest.mark.parametrize("name", ["LPnetlib__lpi_itest6", "HB__west0132", "LPnetlib__lp_blend", "Pajek__GD99_c"])
def test_fixtures_onehot_exact(fixtures, name):
    A = fixtures[name]
    for pname, phi in parts_1d(A):
        H = oracle.pack_1d(A.m, A.n, A.colptr, A.rowval, A.nzval, phi.spl, 4)
        onehot_check(A, vb.SparseMatrix1DVBC[4](A, phi), H)
    for pname, pi, phi in parts_2d(A):
        H = oracle.pack_2d(A.m, A.n, A.colptr, A.rowval, A.nzval, pi.spl, phi.spl, 4, 4)
        onehot_check(A, vb.SparseMatrixVBC[4, 4](A, pi, phi), H)


@pytest.mark.parametrize("kind", ["f64", "bool", "int32"])
def test_size_grid(kind):
    """runtests.jl:14-16 size grid (incl. 1x1, single rows/columns, empty columns)."""
    rng = np.random.default_rng(0xDEADBEEF)
    for m in SIZES:
        for n in SIZES:
            A = sprand(m, n, 0.2, rng, kind)
            phi = vb.pack_stripe(A, vb.RandomChunker(4, seed=m * 31 + n))
            H = oracle.pack_1d(m, n, A.colptr, A.rowval, A.nzval, phi.spl, 4)
            B = vb.SparseMatrix1DVBC[4](A, phi)
            assert_packed_equal(B, H)
            pi, phi2 = vb.pack_plaid(A, vb.AlternatingPacker(vb.RandomChunker(4, seed=m), vb.RandomChunker(4, seed=n)))
            H2 = oracle.pack_2d(m, n, A.colptr, A.rowval, A.nzval, pi.spl, phi2.spl, 4, 4)
            B2 = vb.SparseMatrixVBC[4, 4](A, pi, phi2)
            assert_packed_equal(B2, H2)
            if m <= 9 and n <= 9:
                onehot_check(A, B, H)
                onehot_check(A, B2, H2)
            else:
                randx_check(A, B, H, rng)
                randx_check(A, B2, H2, rng)


def test_empty_and_degenerate():
    i64 = np.int64
    # all-zero matrix: every stripe empty; the adjoint must still store zeros
    A = vb.SparseMatrixCSC(5, 6, np.ones(7, dtype=i64), np.array([], dtype=i64), np.array([], dtype=np.float64))
    B = vb.SparseMatrix1DVBC[4](A, vb.EquiChunker(4))
    assert B.nidx == 0 and B.nval == 0
    y = vb.mul_(np.full(6, 3.0), B.T, np.ones(5))
    assert np.array_equal(y, np.zeros(6))
    y = vb.mul_(np.full(5, 3.0), B, np.ones(6))
    assert np.array_equal(y, np.zeros(5))
    B2 = vb.SparseMatrixVBC[4, 4](A, vb.AlternatingPacker(vb.EquiChunker(4), vb.EquiChunker(4)))
    assert np.array_equal(vb.mul_(np.full(6, 3.0), B2.T, np.ones(5)), np.zeros(6))
    # zero-sized dimensions
    E = vb.SparseMatrixCSC(4, 0, np.ones(1, dtype=i64), np.array([], dtype=i64), np.array([], dtype=np.float64))
    Be = vb.SparseMatrix1DVBC[4](E, vb.EquiChunker(4))
    assert Be.shape == (4, 0)
    assert np.array_equal(vb.mul_(np.full(4, 2.0), Be, np.zeros(0)), np.zeros(4))
    assert vb.mul_(np.zeros(0), Be.T, np.ones(4)).shape == (0,)


def test_wide_and_odd_widths_all_paths():
    """Widths 1..12 in f64 and f32 hit every EPV/CPR class incl. the generic and 'wide' bodies."""
    rng = np.random.default_rng(5)
    for tv in (np.float64, np.float32):
        for w in (1, 2, 3, 4, 5, 6, 7, 8, 12, 16, 24):
            A = sprand(40, 3 * w + 1, 0.3, rng).astype(tv)
            phi = vb.pack_stripe(A, vb.EquiChunker(w))
            H = oracle.pack_1d(A.m, A.n, A.colptr, A.rowval, A.nzval, phi.spl, w)
            B = vb.SparseMatrix1DVBC[w](A, phi)
            assert_packed_equal(B, H)
            for g in (8, 32):
                B.set_option(_lib.OPT_ADJ_GROUP, g)
                B.set_option(_lib.OPT_FWD_GROUP, g)
                B.set_option(_lib.OPT_FWD_MODE, 1 if g == 8 else 2)  # atomic scatter kernel / transposed-index kernel
                randx_check(A, B, H, rng)
            for u in (1, 2, 3, 4, 8):
                pi = vb.pack_stripe(A.transpose(), vb.EquiChunker(u))
                H2 = oracle.pack_2d(A.m, A.n, A.colptr, A.rowval, A.nzval, pi.spl, phi.spl, u, w)
                B2 = vb.SparseMatrixVBC[u, w](A, pi, phi)
                assert_packed_equal(B2, H2)
                for g in (4, 8, 16, 32):
                    B2.set_option(_lib.OPT_ADJ_GROUP, g)
                    B2.set_option(_lib.OPT_FWD_GROUP, g if g in (8, 32) else 0)
                    B2.set_option(_lib.OPT_FWD_MODE, {4: 1, 8: 0, 16: 2, 32: 3}[g])  # atomic scatter / auto / transposed index / transposed copy
                    randx_check(A, B2, H2, rng)


def test_alpha_beta_blas_semantics():
    rng = np.random.default_rng(9)
    A = sprand(33, 29, 0.3, rng)
    S = A.to_scipy()
    B = vb.SparseMatrixVBC[4, 4](A, vb.AlternatingPacker(vb.EquiChunker(4), vb.EquiChunker(4)))
    B1 = vb.SparseMatrix1DVBC[4](A, vb.EquiChunker(4))
    Ba = vb.SparseMatrixVBC[4, 4](A, vb.AlternatingPacker(vb.EquiChunker(4), vb.EquiChunker(4)))
    Ba.set_option(_lib.OPT_FWD_MODE, 1)  # forward through the atomic scatter kernel
    for M in (B, B1, Ba):
        x, y0 = rng.random(29), rng.random(33)
        y = vb.mul_(y0.copy(), M, x, 2.5, -0.5)
        assert np.allclose(y, 2.5 * (S @ x) - 0.5 * y0, rtol=1e-13, atol=1e-13)
        xt, yt0 = rng.random(33), rng.random(29)
        yt = vb.mul_(yt0.copy(), M.T, xt, 2.5, -0.5)
        assert np.allclose(yt, 2.5 * (S.T @ xt) - 0.5 * yt0, rtol=1e-13, atol=1e-13)
        # beta == 0 must not propagate NaNs from y (BLAS convention; `fill!` in multiply_1DVBC.jl:51)
        assert np.all(np.isfinite(vb.mul_(np.full(33, np.nan), M, x, True, False)))
        assert np.all(np.isfinite(vb.mul_(np.full(29, np.nan), M.T, xt, True, False)))
        assert np.allclose(M @ x, S @ x) and np.allclose(M.T @ xt, S.T @ xt)


def test_upload_host_packed_and_parity_mode(fixtures):
    rng = np.random.default_rng(3)
    A = fixtures["LPnetlib__lp_blend"]
    phi = vb.pack_stripe(A, vb.RandomChunker(4, 1))
    H = oracle.pack_1d(A.m, A.n, A.colptr, A.rowval, A.nzval, phi.spl, 4)
    B = vb.SparseMatrix1DVBC.from_packed(4, A.m, A.n, H.spl, H.pos, H.idx, H.ofs, H.val)
    assert_packed_equal(B, H)
    randx_check(A, B, H, rng)
    B.set_option(_lib.OPT_PARITY_MODE, 1)
    randx_check(A, B, H, rng)
    pi, phi2 = vb.pack_plaid(A, vb.AlternatingPacker(vb.RandomChunker(4, 2), vb.RandomChunker(4, 3)))
    H2 = oracle.pack_2d(A.m, A.n, A.colptr, A.rowval, A.nzval, pi.spl, phi2.spl, 4, 4)
    B2 = vb.SparseMatrixVBC.from_packed(4, 4, A.m, A.n, H2.pi_spl, H2.spl, H2.pos, H2.idx, H2.ofs, H2.val)
    assert_packed_equal(B2, H2)
    randx_check(A, B2, H2, rng)
    B2.set_option(_lib.OPT_PARITY_MODE, 1)
    randx_check(A, B2, H2, rng)


def test_errors_map_to_reference_exceptions():
    A = sprand(8, 8, 0.3, np.random.default_rng(2))
    phi = vb.pack_stripe(A, vb.EquiChunker(4))
    with pytest.raises(AssertionError, match="w <= W"):  # constructors_1DVBC.jl:46
        vb.SparseMatrix1DVBC[3](A, phi)
    with pytest.raises(AssertionError, match="u <= U"):  # constructors_VBC.jl:58-60
        vb.SparseMatrixVBC[2, 4](A, phi, phi)
    with pytest.raises(vb.ArgumentError):
        vb.SparseMatrix1DVBC[4](A, vb.SplitPartition(np.array([1, 5, 8], dtype=np.int64)))  # does not cover 1:n
    B = vb.SparseMatrix1DVBC[4](A, phi)
    with pytest.raises(vb.DimensionMismatch):  # multiply_1DVBC.jl:44-45
        vb.mul_(np.zeros(8), B, np.zeros(7))
    with pytest.raises(vb.DimensionMismatch):  # multiply_1DVBC.jl:139-140
        vb.mul_(np.zeros(7), B.T, np.zeros(8))
    with pytest.raises(vb.DimensionMismatch):  # TrSpMV.jl:3-4
        vb.TrSpMV_(np.zeros(8), A, np.zeros(5))


def test_csc_trspmv(fixtures):
    rng = np.random.default_rng(4)
    for tv, ti in ((np.float64, np.int64), (np.float32, np.int32)):
        for name, A64 in fixtures.items():
            A = A64.astype(tv, ti)
            x = rng.random(A.m).astype(tv)
            y = vb.TrSpMV_(np.empty(A.n, dtype=tv), A, x)
            yo = oracle.csc_trspmv(A.m, A.n, A.colptr, A.rowval, A.nzval, x)
            bound = abs(A64.to_scipy()).T @ np.abs(x.astype(np.float64))
            assert np.all(np.abs(y - yo) <= 8 * TOL[np.dtype(tv)] * bound + 1e-300)


def test_device_tensor_path_and_memory_cost(fixtures):
    import torch
    A = fixtures["HB__can_292"]
    pi, phi = vb.pack_plaid(A, vb.AlternatingPacker(vb.EquiChunker(4), vb.EquiChunker(4)))
    B = vb.SparseMatrixVBC[4, 4](A, pi, phi)
    H = oracle.pack_2d(A.m, A.n, A.colptr, A.rowval, A.nzval, pi.spl, phi.spl, 4, 4)
    x = torch.rand(A.m, dtype=torch.float64, device="cuda")
    y = torch.empty(A.n, dtype=torch.float64, device="cuda")
    before = B.launch_count()
    vb.mul_(y, B.T, x)
    torch.cuda.synchronize()
    assert B.launch_count() == before + 1
    yo = oracle.mul(H, x.cpu().numpy(), trans=True)
    assert np.allclose(y.cpu().numpy(), yo, rtol=1e-12, atol=1e-12)
    cost, row_term = B.memory_cost()
    co, ro = oracle.memory_cost(H)
    assert np.array_equal(cost, co) and row_term == ro
    ref_bytes, adj_bytes, _ = B.format_bytes()
    assert ref_bytes == 8 * (3 * (H.L + 1) + (H.K + 1) + len(H.idx)) + 8 * len(H.val)
    assert adj_bytes < ref_bytes


def test_config_c1_1d_vbc():
    """BASELINE configs[0]: 1D-VBC F64, n = 10k, nnz = 1M, W = 8 -- bit-exact pack, y within 1e-12."""
    A, phi = synth.config_c1()
    H = oracle.pack_1d(A.m, A.n, A.colptr, A.rowval, A.nzval, phi.spl, 8)
    B = vb.SparseMatrix1DVBC[8](A, phi)
    assert B.nval == 1_000_000 and B.nidx == 125_000
    assert_packed_equal(B, H)
    randx_check(A, B, H, np.random.default_rng(1))


def test_reduced_c2_and_variable_blocks():
    """configs[1] shape at n = 40k (oracle finishes in seconds) + the C2v variable-block variant."""
    A, pi, phi = synth.config_c2(n=40_000, S=21)
    H = oracle.pack_2d(A.m, A.n, A.colptr, A.rowval, A.nzval, pi.spl, phi.spl, 4, 4)
    B = vb.SparseMatrixVBC[4, 4](A, pi, phi)
    assert_packed_equal(B, H)
    rng = np.random.default_rng(2)
    randx_check(A, B, H, rng)
    pv = synth.variable_partition(A.m, 8, seed=1)
    fv = synth.variable_partition(A.n, 8, seed=2)
    Hv = oracle.pack_2d(A.m, A.n, A.colptr, A.rowval, A.nzval, pv.spl, fv.spl, 8, 8)
    Bv = vb.SparseMatrixVBC[8, 8](A, pv, fv)
    assert_packed_equal(Bv, Hv)
    randx_check(A, Bv, Hv, rng)
    A32 = A.astype(np.float32, np.int32)
    H32 = oracle.pack_2d(A.m, A.n, A32.colptr, A32.rowval, A32.nzval, pi.spl.astype(np.int32), phi.spl.astype(np.int32), 4, 4)
    B32 = vb.SparseMatrixVBC[4, 4](A32, pi, phi)
    assert_packed_equal(B32, H32)
    randx_check(A32, B32, H32, rng)


def test_peer_exchange_two_ranks_on_one_gpu():
    """The fused multiply + all-gather (vbc_peer_*) with two "ranks" driven by one process on one GPU:
    signal and wait are issued separately so no launch waits for a later launch of the same stream."""
    import ctypes
    import torch
    from vbc_b200 import dist as vdist
    n, u, w, P = 8_000, 4, 4, 2
    A, pi, phi = synth.config_c2(n=n, S=9)
    S = A.to_scipy()
    L = n // w
    cost = np.asarray([3 * 8 + (A.colptr[(l + 1) * w] - A.colptr[l * w]) // w * (8 // u) + (A.colptr[(l + 1) * w] - A.colptr[l * w]) * 8 for l in range(L)])
    b = vdist.split_by_cost(cost, P)
    b[1] += 37  # force unequal slices so the padded layout is exercised
    layout = vdist.PaddedLayout(b * w)
    mats, peers = [], []
    Lh = _lib.lib()
    for r in range(P):
        Ar, pir, phir = synth.config_c2(n=n, S=9, stripes=(int(b[r]), int(b[r + 1])))
        Ar = vdist.remap_rows_to_padded(Ar, layout, u)
        mats.append(vb.SparseMatrixVBC[u, w](Ar, vdist.padded_row_partition(layout, u, np.int64), phir))
        h = ctypes.c_void_p()
        _lib.check(Lh.vbc_peer_create(ctypes.byref(h), _lib.VBC_F64, layout.padded_len, r, P, 0, None))
        peers.append(h)
    ptrs = (ctypes.c_void_p * (P * 3))()
    for r in range(P):
        for k in range(3):
            p = ctypes.c_void_p()
            _lib.check(Lh.vbc_peer_buffer(peers[r], k, ctypes.byref(p)))
            ptrs[r * 3 + k] = p.value
    for r in range(P):
        _lib.check(Lh.vbc_peer_connect_local(peers[r], ptrs))
    x0 = synth.vector(n, 3)
    xp = layout.scatter(x0)
    for r in range(P):  # load x into buffer 0 of every rank
        _lib.check(Lh.vbc_peer_buffer(peers[r], 0, ctypes.byref(p)))
        assert torch.cuda.current_device() == 0
        ctypes.CDLL("libcudart.so").cudaMemcpy(ctypes.c_void_p(p.value), ctypes.c_void_p(xp.ctypes.data), ctypes.c_size_t(xp.nbytes), 1)
    x_ref = x0.copy()
    for it in range(3):
        for r in range(P):
            _lib.check(Lh.vbc_peer_spmv_step(peers[r], mats[r]._h, 0.05, r * layout.S, 0))
        for r in range(P):
            _lib.check(Lh.vbc_peer_barrier(peers[r], None, 1))
        for r in range(P):
            _lib.check(Lh.vbc_peer_barrier(peers[r], None, 2))
        x_ref = 0.05 * (S.T @ x_ref)
    torch.cuda.synchronize()
    out = np.empty(layout.padded_len)
    for r in range(P):
        cur, to = ctypes.c_int(), ctypes.c_int()
        _lib.check(Lh.vbc_peer_current(peers[r], ctypes.byref(cur)))
        assert cur.value == 1  # three flips from 0
        _lib.check(Lh.vbc_peer_status(peers[r], ctypes.byref(to)))
        assert to.value == 0
        _lib.check(Lh.vbc_peer_buffer(peers[r], cur.value, ctypes.byref(p)))
        ctypes.CDLL("libcudart.so").cudaMemcpy(ctypes.c_void_p(out.ctypes.data), ctypes.c_void_p(p.value), ctypes.c_size_t(out.nbytes), 2)
        assert np.allclose(layout.gather(out), x_ref, rtol=1e-12, atol=0), f"rank {r}"
    for h in peers:
        Lh.vbc_peer_destroy(h)


def test_spmm_matches_column_by_column_oracle(fixtures):
    """north_star (c): k right-hand sides.  Oracle = k independent `mul!`s (the reference's matrix `*` is
    non-functional)."""
    rng = np.random.default_rng(21)
    A = fixtures["LPnetlib__lp_blend"]
    S = A.to_scipy()
    absS = abs(S)
    phi = vb.pack_stripe(A, vb.RandomChunker(8, 4))
    pi, phi2 = vb.pack_plaid(A, vb.AlternatingPacker(vb.EquiChunker(4), vb.EquiChunker(4)))
    piv, phiv = vb.pack_plaid(A, vb.AlternatingPacker(vb.RandomChunker(12, 8), vb.RandomChunker(4, 7)))  # columns <= 12, rows <= 4
    mats = [(vb.SparseMatrix1DVBC[8](A, phi), oracle.pack_1d(A.m, A.n, A.colptr, A.rowval, A.nzval, phi.spl, 8)),
            (vb.SparseMatrixVBC[4, 4](A, pi, phi2), oracle.pack_2d(A.m, A.n, A.colptr, A.rowval, A.nzval, pi.spl, phi2.spl, 4, 4)),
            (vb.SparseMatrixVBC[4, 12](A, piv, phiv), oracle.pack_2d(A.m, A.n, A.colptr, A.rowval, A.nzval, piv.spl, phiv.spl, 4, 12))]
    for B, H in mats:
        for k in (1, 3, 32, 40, 70):
            B.set_option(_lib.OPT_SPMM_SIMT, 1 if k in (3, 40) else 0)  # Float64 adjoint: SIMT kernel / DMMA tiles
            for order in ("C", "F"):
                for trans in (False, True):
                    xr, yr = (A.m, A.n) if trans else (A.n, A.m)
                    X = np.asarray(rng.random((xr, k)), order=order)
                    Y0 = np.asarray(rng.random((yr, k)), order=order)
                    Y = vb.mul_(Y0.copy(order=order), B.T if trans else B, X, 1.5, -0.25)
                    Sx = (S.T if trans else S) @ X
                    bound = (absS.T if trans else absS) @ np.abs(X) + np.abs(Y0)
                    assert np.all(np.abs(Y - (1.5 * Sx - 0.25 * Y0)) <= 1e-11 * bound), (k, order, trans)
                    for c in (0, k - 1):
                        yo = oracle.mul(H, np.ascontiguousarray(X[:, c]), trans=trans)
                        Yb = vb.mul_(np.full((yr, k), np.nan, order=order), B.T if trans else B, X)
                        assert np.all(np.abs(Yb[:, c] - yo) <= 8e-12 * ((absS.T if trans else absS) @ np.abs(X[:, c])) + 1e-300)
    # device panels + `@` sugar
    import torch
    B, H = mats[1]
    Xd = torch.rand(A.m, 32, dtype=torch.float64, device="cuda")
    Yd = B.T @ Xd
    assert np.allclose(Yd.cpu().numpy(), S.T @ Xd.cpu().numpy(), rtol=1e-11, atol=1e-12)
    with pytest.raises(vb.DimensionMismatch):
        vb.mul_(np.zeros((A.m, 3)), B, np.zeros((A.n + 1, 3)))


def test_triangular_solve_extension():
    """north_star (d): tril(A') x = b.  No reference counterpart (parity unpinned): oracle = scipy's
    forward substitution on the CSC matrix."""
    import scipy.sparse as sp
    from scipy.sparse.linalg import spsolve_triangular
    rng = np.random.default_rng(31)
    cases = []
    A, pi, phi = synth.config_c4_triangular(n=12_000, S=13)
    cases.append(("2D 4x4 f64", A, lambda: vb.SparseMatrixVBC[4, 4](A, pi, phi), 1e-11))
    cases.append(("1D w=4 f64", A, lambda: vb.SparseMatrix1DVBC[4](A, phi), 1e-11))
    pv, fv = synth.variable_partition(A.m, 8, 3), synth.variable_partition(A.n, 8, 4)
    cases.append(("2D variable f64", A, lambda: vb.SparseMatrixVBC[8, 8](A, pv, fv), 1e-11))
    A32 = A.astype(np.float32, np.int32)
    cases.append(("2D 4x4 f32", A32, lambda: vb.SparseMatrixVBC[4, 4](A32, pi, phi), 2e-5))
    Ar = sprand(300, 300, 0.05, rng)
    Sr = sp.triu(Ar.to_scipy()) + sp.identity(300) * 10.0  # A upper triangular <=> A' lower triangular
    Ar = vb.SparseMatrixCSC.from_scipy(sp.csc_matrix(Sr))
    cases.append(("1D random w<=3", Ar, lambda: vb.SparseMatrix1DVBC[3](Ar, vb.RandomChunker(3, 9)), 1e-11))
    for name, M, ctor, tol in cases:
        B = ctor()
        tv = M.nzval.dtype
        T = sp.tril(M.to_scipy().T.astype(np.float64)).tocsr()
        b = rng.random(M.n).astype(tv)
        x = vb.ldiv_lower_(np.empty(M.n, dtype=tv), B.T, b)
        xo = spsolve_triangular(T, b.astype(np.float64), lower=True)
        assert np.all(np.isfinite(x)), name
        assert np.max(np.abs(x - xo)) <= tol * max(1.0, np.max(np.abs(xo))), name
        assert np.max(np.abs(T @ x.astype(np.float64) - b)) <= 50 * tol * np.max(np.abs(b)) * 40, name
        assert vb.trsv_analyse(B.T) >= 1
        # second solve on the same plan (epoch flags) and the device-vector path
        import torch
        bd = torch.from_numpy(b).cuda()
        xd = torch.empty_like(bd)
        vb.ldiv_lower_(xd, B.T, bd)
        B.sync()
        assert np.array_equal(xd.cpu().numpy(), x), name
    with pytest.raises(vb.DimensionMismatch):
        vb.ldiv_lower_(np.zeros(5), B.T, np.zeros(5))


def test_peer_single_rank_interior_and_boundary():
    """The fused step with a single rank: interior range + claimed boundary runs + the mask path all execute,
    there is nobody to wait for; two interior choices and the auto-derived one give the same iterate."""
    import ctypes
    import torch
    n, u, w = 12_000, 4, 4
    A, pi, phi = synth.config_c2(n=n, S=11)
    S = A.to_scipy()
    B = vb.SparseMatrixVBC[u, w](A, pi, phi)
    Lh = _lib.lib()
    rt = ctypes.CDLL("libcudart.so")
    x0 = synth.vector(n, 4)
    outs = []
    for interior in ((1000, 2001), (0, 0), None):
        h = ctypes.c_void_p()
        _lib.check(Lh.vbc_peer_create(ctypes.byref(h), _lib.VBC_F64, n, 0, 1, 0, None))
        mask = np.ones((n + 127) // 128, dtype=np.uint8)
        _lib.check(Lh.vbc_peer_set_mask(h, mask.ctypes.data_as(ctypes.c_void_p), len(mask), 7))
        if interior is None:
            i0, i1 = ctypes.c_int64(), ctypes.c_int64()
            _lib.check(Lh.vbc_peer_auto_interior(h, B._h, 0, ctypes.byref(i0), ctypes.byref(i1)))
            assert (i0.value, i1.value) == (0, B.L)
        else:
            _lib.check(Lh.vbc_peer_set_interior(h, *interior))
        p = ctypes.c_void_p()
        _lib.check(Lh.vbc_peer_buffer(h, 0, ctypes.byref(p)))
        rt.cudaMemcpy(ctypes.c_void_p(p.value), ctypes.c_void_p(x0.ctypes.data), ctypes.c_size_t(x0.nbytes), 1)
        ref = x0.copy()
        for _ in range(3):
            _lib.check(Lh.vbc_peer_spmv_step(h, B._h, 0.04, 0, 3))
            ref = 0.04 * (S.T @ ref)
        torch.cuda.synchronize()
        cur, to = ctypes.c_int(), ctypes.c_int()
        _lib.check(Lh.vbc_peer_current(h, ctypes.byref(cur)))
        _lib.check(Lh.vbc_peer_status(h, ctypes.byref(to)))
        assert cur.value == 1 and to.value == 0
        _lib.check(Lh.vbc_peer_buffer(h, 1, ctypes.byref(p)))
        out = np.empty(n)
        rt.cudaMemcpy(ctypes.c_void_p(out.ctypes.data), ctypes.c_void_p(p.value), ctypes.c_size_t(out.nbytes), 2)
        assert np.allclose(out, ref, rtol=1e-12, atol=0)
        st = (ctypes.c_uint64 * 4)()
        _lib.check(Lh.vbc_peer_wait_stats(h, st, 0))
        assert st[0] == (0 if interior is None else 3)  # steps published (no boundary claims <=> nothing to publish)
        outs.append(out)
        Lh.vbc_peer_destroy(h)
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])  # same stripe bodies, same bits


def test_pack_from_device_resident_csc(fixtures):
    """vbc_pack_csc_dev: CSC arrays and partitions already in HBM (matrices generated on the device)."""
    import torch
    for tv, ti in ((np.float64, np.int64), (np.float32, np.int32)):
        A = fixtures["LPnetlib__lp_etamacro"].astype(tv, ti)
        d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
        phi = vb.pack_stripe(A, vb.RandomChunker(4, 11))
        H = oracle.pack_1d(A.m, A.n, A.colptr, A.rowval, A.nzval, phi.spl, 4)
        B = vb.SparseMatrix1DVBC.from_device_csc(4, A.m, A.n, d(A.colptr), d(A.rowval), d(A.nzval), d(phi.spl))
        assert_packed_equal(B, H)
        pi, phi2 = vb.pack_plaid(A, vb.AlternatingPacker(vb.RandomChunker(4, 12), vb.RandomChunker(4, 13)))
        H2 = oracle.pack_2d(A.m, A.n, A.colptr, A.rowval, A.nzval, pi.spl, phi2.spl, 4, 4)
        B2 = vb.SparseMatrixVBC.from_device_csc(4, 4, A.m, A.n, d(A.colptr), d(A.rowval), d(A.nzval), d(pi.spl), d(phi2.spl))
        assert_packed_equal(B2, H2)
        randx_check(A, B2, H2, np.random.default_rng(5))


def test_gpu_time_model_autotuner_small():
    """the costs.jl time-model sweep + fit, on the device (small cache budget so it runs in seconds)."""
    from vbc_b200 import costs
    d = costs.model_SparseMatrix1DVBC_TrSpMV_time_data(4, np.float64, np.int64, np.float64, cache_bytes=4 << 20, use_cache=False)
    assert len(d["T"]) == 16 and all(t > 0 for t in d["T"])
    mdl = costs.model_SparseMatrix1DVBC_TrSpMV_time(4, np.float64, np.int64, np.float64, cache_bytes=4 << 20)
    assert np.all(np.diff(np.asarray(mdl.alpha_col)) >= 0) and np.all(np.diff(np.asarray(mdl.beta_col)) >= 0)
    m2 = costs.model_SparseMatrixVBC_TrSpMV_time(2, 2, 2, np.float64, np.int64, np.float64, cache_bytes=4 << 20, use_cache=False)
    assert len(m2.beta_row) == 2 and len(m2.beta_col) == 2
    # evaluate models on a packed matrix (`total_value`, bin/test_table.jl:82/:124)
    A, pi, phi = synth.config_c2(n=4000, S=7)
    B = vb.SparseMatrixVBC[4, 4](A, pi, phi)
    ref_bytes = B.format_bytes()[0]
    mem = costs.total_value(B, costs.model_SparseMatrixVBC_memory(np.float64, np.int64))
    assert mem == ref_bytes - 8 * 4  # the model has no "+1" entries: Π, Φ, pos, ofs each one element longer
    B1 = vb.SparseMatrix1DVBC[4](A, phi)
    assert costs.total_value(B1, costs.model_SparseMatrix1DVBC_blocks()) == B1.nidx
    assert costs.total_value(B1, costs.model_SparseMatrix1DVBC_memory(np.float64, np.int64)) == B1.format_bytes()[0] - 8 * 3


@pytest.mark.parametrize("name", ["LPnetlib__lpi_itest6", "HB__west0132", "LPnetlib__lp_blend", "Pajek__GD99_c"])
def test_runtests_jl_mirror(fixtures, name):
    """test/runtests.jl:19-88 as written there: `B = SparseMatrix1DVBC{4}(A, method)` for its four 1D methods and
    `SparseMatrixVBC{4, 4}(A, method)` for its two 2D packers, then `mul!(y_test, B, e_j, true, false) == mul!(y_ref, A, e_j, ...)`
    for every unit vector, both orientations, exact equality."""
    from test_oracle_golden import reference_methods_1d, reference_methods_2d
    A = fixtures[name]
    for method in reference_methods_1d():
        B = vb.SparseMatrix1DVBC[4](A, method)
        H = oracle.pack_1d(A.m, A.n, A.colptr, A.rowval, A.nzval, B.Phi.spl, 4)
        assert_packed_equal(B, H)
        onehot_check(A, B, H)
    for method in reference_methods_2d():
        B = vb.SparseMatrixVBC[4, 4](A, method)
        H = oracle.pack_2d(A.m, A.n, A.colptr, A.rowval, A.nzval, B.Pi.spl, B.Phi.spl, 4, 4)
        assert_packed_equal(B, H)
        onehot_check(A, B, H)


def test_float32_matrix_with_float64_vectors_accumulates_in_float64():
    """eltype(y) decides the arithmetic: values and x are converted to eltype(y) before multiplying
    (multiply_1DVBC.jl:23/27/34, :102; multiply_VBC.jl:40-45, :131).  A Float32 matrix applied to Float64
    vectors must therefore agree with the Float64 product of the (exactly representable) Float32 values to
    Float64 accuracy -- a Float32 accumulation would miss this bound by seven orders of magnitude."""
    import torch
    rng = np.random.default_rng(21)
    for (m, n, u, w) in ((57, 41, 4, 4), (40, 37, 3, 5), (64, 64, 1, 8), (30, 50, 8, 2)):
        A = sprand(m, n, 0.3, rng).astype(np.float32)
        S = A.to_scipy().astype(np.float64)
        absS = abs(S)
        phi = vb.pack_stripe(A, vb.EquiChunker(w))
        pi = vb.pack_stripe(A.transpose(), vb.EquiChunker(u))
        pir, phir = vb.pack_plaid(A, vb.AlternatingPacker(vb.RandomChunker(w, seed=m), vb.RandomChunker(u, seed=n)))
        mats = [vb.SparseMatrix1DVBC[w](A, phi), vb.SparseMatrixVBC[u, w](A, pi, phi), vb.SparseMatrixVBC[max(u, w), max(u, w)](A, pir, phir)]
        for B in mats:
            assert B.Tv == np.dtype(np.float32)
            for trans in (False, True):
                xlen, ylen = (m, n) if trans else (n, m)
                op, Sop, aSop = (B.T, S.T, absS.T) if trans else (B, S, absS)
                x = rng.random(xlen)
                y0 = rng.random(ylen)
                bound = aSop @ np.abs(x)
                y = vb.mul_(np.full(ylen, np.nan), op, x, True, False)
                assert y.dtype == np.float64
                assert np.all(np.abs(y - Sop @ x) <= 1e-12 * bound + 1e-300)
                # BLAS alpha / beta on the same path
                y = vb.mul_(y0.copy(), op, x, 2.0, -0.25)
                assert np.all(np.abs(y - (2.0 * (Sop @ x) - 0.25 * y0)) <= 1e-12 * (2 * bound + np.abs(y0)) + 1e-300)
                # a Float32 x is widened like the values are
                x32 = x.astype(np.float32)
                y = vb.mul_(np.empty(ylen), op, x32)
                assert np.all(np.abs(y - Sop @ x32.astype(np.float64)) <= 1e-12 * bound + 1e-300)
                # device tensors
                xd = torch.from_numpy(x).cuda()
                yd = torch.full((ylen,), float("nan"), dtype=torch.float64, device="cuda")
                vb.mul_(yd, op, xd)
                torch.cuda.synchronize()
                assert np.all(np.abs(yd.cpu().numpy() - Sop @ x) <= 1e-12 * bound + 1e-300)
            # same-type call on the same handle still runs the Float32 kernels
            x = rng.random(n).astype(np.float32)
            y = vb.mul_(np.empty(m, dtype=np.float32), B, x)
            assert np.all(np.abs(y - S @ x.astype(np.float64)) <= 4e-5 * (absS @ np.abs(x.astype(np.float64))) + 1e-30)
            with pytest.raises(vb.DimensionMismatch):
                vb.mul_(np.empty(m + 1), B, rng.random(n))
    # narrowing (Float64 values, Float32 vectors) is refused at the ABI
    A64 = sprand(9, 9, 0.5, rng)
    B64 = vb.SparseMatrix1DVBC[4](A64, vb.EquiChunker(4))
    x32, y32 = np.ones(9, dtype=np.float32), np.empty(9, dtype=np.float32)
    rc = _lib.lib().vbc_spmv_mixed(B64._h, 0, 1.0, x32.ctypes.data, 9, 0.0, y32.ctypes.data, 9, _lib.VBC_F32, 0)
    assert rc == _lib.VBC_EARG


def _spd_blocked(n, rng):
    import scipy.sparse as sp
    R = sp.random(n, n, density=6.0 / n, random_state=np.random.RandomState(int(rng.integers(1 << 30))), format="csr")
    band = sp.diags([rng.random(n - 1), rng.random(n - 4)], [1, 4], format="csr")
    S = R + band
    S = S + S.T
    d = np.asarray(abs(S).sum(axis=1)).ravel() + 1.0
    d[7] = 5.0 * d.max()  # one dominant entry: a clear spectral gap for the power iteration
    return (S + sp.diags(d)).tocsc()


@pytest.mark.parametrize("graph", [False, True])
def test_device_resident_cg_and_power_iteration(graph):
    """the adjoint multiply inside device-resident iterations (eager and CUDA-graph replay)."""
    import scipy.sparse.linalg as spla
    import torch
    from vbc_b200 import solvers
    rng = np.random.default_rng(33)
    n = 3000
    S = _spd_blocked(n, rng)
    A = vb.SparseMatrixCSC.from_scipy(S)
    for B in (vb.SparseMatrixVBC[4, 4](A, vb.AlternatingPacker(vb.EquiChunker(4), vb.EquiChunker(4))), vb.SparseMatrix1DVBC[8](A, vb.EquiChunker(8))):
        b = rng.random(n)
        x, its, res = solvers.cg(B.T, b, iters=500, rtol=1e-12, check_every=5, graph=graph)
        torch.cuda.synchronize()
        assert res <= 1e-12 and its < 500
        xs = spla.spsolve(S.T.tocsc(), b)
        assert np.allclose(x.cpu().numpy(), xs, rtol=1e-9, atol=1e-12)
        lam, v, its = solvers.power_iteration(B.T, iters=400, tol=1e-13, check_every=20, graph=graph)
        top = spla.eigsh(S, k=1, which="LA", return_eigenvectors=False)[0]
        assert abs(lam - top) <= 1e-6 * top
        assert abs(float(torch.linalg.vector_norm(v)) - 1.0) < 1e-12


def test_host_vector_pipeline_matches_plain_upload():
    """VBC_OPT_E2E_PIPELINE (default on): x uploaded in pieces, chunks of stripes started as their rows arrive, y ranges
    copied back while later chunks run -- same bits as upload-everything-first, for a banded matrix, for one whose first
    stripes reach the LAST rows (the first piece must then be all of x) and for a slab that reads only a row window
    (only that window is uploaded)."""
    import scipy.sparse as sp
    rng = np.random.default_rng(3)

    def both(B, x, n, alpha=True, beta=False, y0=None):
        outs = []
        for mode in (0, 1):
            B.set_option(_lib.OPT_E2E_PIPELINE, mode)
            y = np.full(n, np.nan) if y0 is None else y0.copy()
            outs.append(vb.mul_(y, B.T, x, alpha, beta))
        return outs

    # banded 2D blocks, large enough for the chunked path (y >= 1 MiB)
    A, pi, phi = synth.config_c2(n=200_000, S=23)
    B = vb.SparseMatrixVBC[4, 4](A, pi, phi)
    assert B.get_option(_lib.OPT_E2E_PIPELINE) == 1
    x = rng.random(A.m)
    ya, yb = both(B, x, A.n)
    assert np.array_equal(ya, yb)
    assert B.get_option(_lib.OPT_E2E_UPLOAD_ELEMS) == A.m
    assert np.allclose(ya, A.to_scipy().T @ x, rtol=1e-12)
    y0 = rng.random(A.n)
    ya, yb = both(B, x, A.n, 2.0, -0.5, y0)
    assert np.array_equal(ya, yb)
    # far-reaching first stripes
    n = 400_000
    S = sp.diags([rng.random(n), rng.random(n - 3)], [0, -3], format="lil")
    S[n - 1, 0] = 2.0
    S[n - 2, 5] = 3.0
    A = vb.SparseMatrixCSC.from_scipy(S.tocsc())
    B = vb.SparseMatrix1DVBC[4](A, vb.EquiChunker(4))
    x = rng.random(n)
    ya, yb = both(B, x, n)
    assert np.array_equal(ya, yb) and np.allclose(ya, S.T @ x, rtol=1e-12)
    # a slab like one rank of the row-partitioned multiply: gathers only from rows [100k, 300k) of 400k
    m, n = 400_000, 200_000
    S = sp.diags([rng.random(n), rng.random(n), rng.random(n)], [-100_000, -100_007, -99_990], shape=(m, n), format="csc")
    A = vb.SparseMatrixCSC.from_scipy(S)
    pi, phi = vb.pack_plaid(A, vb.AlternatingPacker(vb.EquiChunker(4), vb.EquiChunker(4)))
    B = vb.SparseMatrixVBC[4, 4](A, pi, phi)
    x = rng.random(m)
    ya, yb = both(B, x, n)
    assert np.array_equal(ya, yb) and np.allclose(ya, S.T @ x, rtol=1e-12)
    assert B.get_option(_lib.OPT_E2E_UPLOAD_ELEMS) < 0.6 * m


def test_flat_slab_bodies_short_long_and_oversized_stripes():
    """Unaligned stripes (odd widths, odd / non-multiple-of-four slab starts) with 0 ... 700 stored rows: short ones, ones that
    need several batches, and ones whose x row does not fit the group's shared-memory row (16 * G values) and fall back to the
    per-element bodies inside the same kernel -- every group size, Float64 and Float32, 1D and variable 2D, against the oracle."""
    rng = np.random.default_rng(2024)
    for tv in (np.float64, np.float32):
        for m, dens in ((700, 0.95), (700, 0.3), (333, 0.04)):
            for w in (3, 5, 6, 7):
                n = 7 * w + 2  # the last stripe is narrower: its slab starts wherever the others end
                A = sprand(m, n, dens, rng).astype(tv)
                phi = vb.pack_stripe(A, vb.EquiChunker(w))
                H = oracle.pack_1d(A.m, A.n, A.colptr, A.rowval, A.nzval, phi.spl, w)
                B = vb.SparseMatrix1DVBC[w](A, phi)
                for g in (8, 16, 32):
                    B.set_option(_lib.OPT_ADJ_GROUP, g)
                    randx_check(A, B, H, rng)
                B.close()
        A = sprand(300, 211, 0.4, rng).astype(tv)
        pv, fv = synth.variable_partition(A.m, 8, 5), synth.variable_partition(A.n, 8, 6)
        H2 = oracle.pack_2d(A.m, A.n, A.colptr, A.rowval, A.nzval, pv.spl, fv.spl, 8, 8)
        B2 = vb.SparseMatrixVBC[8, 8](A, pv, fv)
        for g in (8, 16, 32):
            B2.set_option(_lib.OPT_ADJ_GROUP, g)
            B2.set_option(_lib.OPT_FWD_GROUP, g if g != 16 else 0)
            randx_check(A, B2, H2, rng)
        B2.close()


def test_long_stripes_take_the_cta_per_stripe_kernel():
    """A dense column group among sparse ones (adjoint) and a dense row (forward through the transposed copy: a dense row of A
    is a long stripe of the copy): stripes with more than 2 K values and more than 16 times the average are multiplied by one
    CTA each (k_spmv_adj_long).  1D and 2D, uniform and variable blocks, both value types, alpha / beta, against scipy."""
    import scipy.sparse as sp
    rng = np.random.default_rng(99)
    m = n = 24_000
    M = sp.random(m, n, density=2.0 / n, random_state=np.random.RandomState(3), format="lil")
    M[:, 4000:4004] = rng.random((m, 4))     # a dense column group: 96 000 values in one stripe
    M[9001, :] = rng.random(n)               # a dense row
    S = sp.csc_matrix(M)
    for tv, tol in ((np.float64, 1e-12), (np.float32, 2e-5)):
        A = vb.SparseMatrixCSC.from_scipy(S).astype(tv)
        Sd = A.to_scipy().astype(np.float64)
        absS = abs(Sd)
        pv, fv = synth.variable_partition(A.m, 8, 3), synth.variable_partition(A.n, 8, 4)
        mats = [vb.SparseMatrix1DVBC[4](A, vb.pack_stripe(A, vb.EquiChunker(4))),
                vb.SparseMatrixVBC[4, 4](A, vb.pack_stripe(A.transpose(), vb.EquiChunker(4)), vb.pack_stripe(A, vb.EquiChunker(4))),
                vb.SparseMatrixVBC[8, 8](A, pv, fv)]
        for B in mats:
            x, y0 = synth.vector(A.m, 3, dtype=tv), synth.vector(A.n, 4, dtype=tv)
            y = vb.mul_(y0.copy(), B.T, x, 1.5, -0.5)
            want = 1.5 * (Sd.T @ x.astype(np.float64)) - 0.5 * y0
            assert np.all(np.abs(y - want) <= 8 * tol * (absS.T @ np.abs(x.astype(np.float64)) + np.abs(y0))), "adjoint"
            xf, yf0 = synth.vector(A.n, 5, dtype=tv), synth.vector(A.m, 6, dtype=tv)
            yf = vb.mul_(yf0.copy(), B, xf, 1.5, -0.5)
            wantf = 1.5 * (Sd @ xf.astype(np.float64)) - 0.5 * yf0
            assert np.all(np.abs(yf - wantf) <= 8 * tol * (absS @ np.abs(xf.astype(np.float64)) + np.abs(yf0))), "forward"
            B.close()


def test_forward_variable_blocks_through_transposed_copy():
    """2D blocks with iid heights and widths (2..8): the forward multiply in auto mode (0) and with the copy forced (3) runs the
    adjoint kernel on a rows-mode transposed copy built from the canonical block arrays -- against the oracle, scipy and the
    atomic scatter kernel (1); empty row parts and stripes, both value and index types, alpha / beta."""
    import scipy.sparse as sp
    rng = np.random.default_rng(123)
    for tv, ti in ((np.float64, np.int64), (np.float32, np.int32), (np.float64, np.int32)):
        M = sp.random(611, 467, density=0.03, random_state=np.random.RandomState(7), format="lil")
        M[100:140, :] = 0.0      # row parts without any block
        M[:, 200:230] = 0.0      # empty stripes
        M[300:308, 50:58] = 1.25 # a dense spot
        A = vb.SparseMatrixCSC.from_scipy(sp.csc_matrix(M)).astype(tv, ti)
        pv, fv = synth.variable_partition(A.m, 8, 3, ti=ti), synth.variable_partition(A.n, 8, 4, ti=ti)
        H = oracle.pack_2d(A.m, A.n, A.colptr, A.rowval, A.nzval, pv.spl, fv.spl, 8, 8)
        S = A.to_scipy().astype(np.float64)
        tol = TOL[np.dtype(tv)]
        ys = {}
        for mode in (0, 3, 1):
            B = vb.SparseMatrixVBC[8, 8](A, pv, fv)
            B.set_option(_lib.OPT_FWD_MODE, mode)
            randx_check(A, B, H, rng)
            x = synth.vector(A.n, 5, dtype=tv)
            y0 = synth.vector(A.m, 6, dtype=tv)
            y = vb.mul_(y0.copy(), B, x, 1.5, -0.25)
            want = 1.5 * (S @ x.astype(np.float64)) - 0.25 * y0
            assert np.all(np.abs(y - want) <= 8 * tol * (abs(S) @ np.abs(x.astype(np.float64)) + np.abs(y0)))
            ys[mode] = vb.mul_(np.empty(A.m, dtype=tv), B, x)
            fb = B.format_bytes()
            if mode != 1:
                assert fb[2] != fb[1], "the forward multiply should read the transposed copy"  # its own meta / descriptors
            B.close()
        assert np.array_equal(ys[0], ys[3])  # same copy, same kernel: bit-identical and run-to-run deterministic
        assert np.allclose(ys[0], ys[1], rtol=50 * tol, atol=50 * tol)


def test_spmm_uniform_widths_ragged_stripes():
    """Float64 adjoint SpMM of matrices whose stripes all have width 4 or 8: stripes with 0, 1, odd and many stored rows, empty
    leading / trailing stripes, k below, at and above one 32-column panel, alpha / beta, against scipy, tensor-tile kernel
    against the SIMT kernel.  Parity of SpMM is unpinned by the reference: the oracle is k independent products."""
    import scipy.sparse as sp
    import torch
    rng = np.random.default_rng(77)
    for w, L, m, dens in ((8, 37, 300, 0.02), (8, 700, 900, 0.04), (4, 1201, 500, 0.01), (8, 3000, 4000, 0.012), (4, 64, 40, 0.5)):
        n = w * L
        M = sp.random(m, n, density=dens, random_state=np.random.RandomState(int(rng.integers(1 << 30))), format="lil")
        M[:, : 3 * w] = 0.0                      # empty stripes at the start,
        M[:, n - 2 * w:] = 0.0                   # at the end,
        M[:, 10 * w: 11 * w] = 0.0               # and inside
        M[5, 12 * w] = 1.5                       # a stripe with exactly one stored row
        A = vb.SparseMatrixCSC.from_scipy(sp.csc_matrix(M))
        S = A.to_scipy()
        absS = abs(S)
        B = vb.SparseMatrix1DVBC[w](A, vb.pack_stripe(A, vb.EquiChunker(w)))
        for k in (2, 32, 34, 64):
            X = torch.rand(m, k, dtype=torch.float64, device="cuda")
            Y0 = torch.rand(n, k, dtype=torch.float64, device="cuda")
            want = S.T @ X.cpu().numpy()
            bound = absS.T @ np.abs(X.cpu().numpy()) + 1.0
            outs = []
            for mode in (0, 1, 2):  # auto (TMA-fed tensor tiles for width 8), SIMT, tensor tiles fed by per-lane loads
                B.set_option(_lib.OPT_SPMM_SIMT, mode)
                Y = torch.full((n, k), float("nan"), dtype=torch.float64, device="cuda")
                vb.mul_(Y, B.T, X)
                assert np.all(np.abs(Y.cpu().numpy() - want) <= 1e-12 * bound), (w, L, k, mode)
                Yb = vb.mul_(Y0.clone(), B.T, X, 1.5, -0.25)
                assert np.all(np.abs(Yb.cpu().numpy() - (1.5 * want - 0.25 * Y0.cpu().numpy())) <= 2e-12 * bound), (w, L, k, mode)
                outs.append(Y)
            for o in outs[1:]:
                assert torch.allclose(outs[0], o, rtol=1e-12, atol=1e-13)
        B.set_option(_lib.OPT_SPMM_SIMT, 0)
        # a strided view (ldx > k) and an odd k
        Xw = torch.rand(m, 40, dtype=torch.float64, device="cuda")
        Y = vb.mul_(torch.empty(n, 6, dtype=torch.float64, device="cuda"), B.T, Xw[:, 2:8])
        assert np.allclose(Y.cpu().numpy(), S.T @ Xw[:, 2:8].cpu().numpy(), rtol=1e-11, atol=1e-12)
        Y = vb.mul_(torch.empty(n, 5, dtype=torch.float64, device="cuda"), B.T, Xw[:, 1:6].contiguous())
        assert np.allclose(Y.cpu().numpy(), S.T @ Xw[:, 1:6].cpu().numpy(), rtol=1e-11, atol=1e-12)


def test_spmm_host_panels_with_leading_dimension_padding():
    """a host panel with ld > inner holds (outer-1)*ld + inner elements; the copies must not touch more."""
    rng = np.random.default_rng(8)
    A = sprand(60, 45, 0.2, rng)
    B = vb.SparseMatrix1DVBC[4](A, vb.EquiChunker(4))
    k, ldx, ldy = 5, 9, 11
    Xbuf = rng.random((A.m - 1) * ldx + k)            # exactly the BLAS-sized buffers: one element more would fault under a sanitizer
    Ybuf = np.full((A.n - 1) * ldy + k, -7.0)
    X = np.lib.stride_tricks.as_strided(Xbuf, shape=(A.m, k), strides=(8 * ldx, 8))
    L = _lib.lib()
    import ctypes
    _lib.check(L.vbc_spmm(B._h, 1, k, 1.0, ctypes.c_void_p(Xbuf.ctypes.data), ldx, 0.0, ctypes.c_void_p(Ybuf.ctypes.data), ldy, 0, 0))
    Y = np.lib.stride_tricks.as_strided(Ybuf, shape=(A.n, k), strides=(8 * ldy, 8))
    assert np.allclose(Y, A.to_scipy().T @ X, rtol=1e-12)
    pad = np.ones(len(Ybuf), dtype=bool)
    for r in range(A.n):
        pad[r * ldy: r * ldy + k] = False
    assert np.all(Ybuf[pad] == -7.0)  # the gaps between the rows of Y were not written


def test_device_generator_matches_host_generator():
    """vbc_gen_banded_csc: the slab generated on the device equals synth.banded_blocks bit for bit (structure and values),
    for both element / index types, a slab in the middle and the clipped edges; and packs to the same VBC arrays."""
    import torch
    K = L = 3000
    for tv, ti in ((np.float64, np.int64), (np.float32, np.int32)):
        for stripes in ((0, 700), (1100, 1900), (2500, 3000), None):
            H, pi, phi = synth.banded_blocks(K, L, 4, 4, synth.C5_OFFSETS, dtype=tv, ti=ti, stripes=stripes, diag_boost=2.5)
            D = synth.banded_blocks_device(K, L, 4, 4, synth.C5_OFFSETS, dtype=tv, ti=ti, stripes=stripes, diag_boost=2.5)
            assert (D.m, D.n, D.nnz) == (H.m, H.n, H.nnz)
            assert np.array_equal(D.colptr.cpu().numpy(), H.colptr)
            assert np.array_equal(D.rowval.cpu().numpy(), H.rowval)
            assert np.array_equal(D.nzval.cpu().numpy(), H.nzval)
            d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
            B = vb.SparseMatrixVBC.from_device_csc(4, 4, D.m, D.n, D.colptr, D.rowval, D.nzval, d(pi.spl), d(phi.spl))
            Hp = oracle.pack_2d(H.m, H.n, H.colptr, H.rowval, H.nzval, pi.spl, phi.spl, 4, 4)
            assert_packed_equal(B, Hp)
            D.free()


def test_malformed_csc_is_refused_on_the_device():
    """rowval outside 1:m, descending rows, a broken colptr -> ArgumentError before any kernel indexes with them."""
    import torch
    A = sprand(40, 30, 0.3, np.random.default_rng(2))
    phi = vb.pack_stripe(A, vb.EquiChunker(4))
    pi = vb.pack_stripe(A.transpose(), vb.EquiChunker(4))
    bad = []
    rv = A.rowval.copy(); rv[5] = A.m + 1; bad.append((A.colptr, rv))
    rv = A.rowval.copy(); rv[7] = 0; bad.append((A.colptr, rv))
    j = int(np.argmax(np.diff(A.colptr) >= 2)); b = A.colptr[j] - 1
    rv = A.rowval.copy(); rv[b], rv[b + 1] = rv[b + 1], rv[b]; bad.append((A.colptr, rv))
    cp = A.colptr.copy(); cp[3] = cp[-1] + 5; bad.append((cp, A.rowval))
    for cpb, rvb in bad:
        M = vb.SparseMatrixCSC(A.m, A.n, cpb, rvb, A.nzval)
        for ctor in (lambda: vb.SparseMatrix1DVBC[4](M, phi), lambda: vb.SparseMatrixVBC[4, 4](M, pi, phi)):
            with pytest.raises(vb.ArgumentError):
                ctor()
        d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
        with pytest.raises(vb.ArgumentError):
            vb.SparseMatrix1DVBC.from_device_csc(4, A.m, A.n, d(cpb), d(rvb), d(A.nzval), d(phi.spl))
    B = vb.SparseMatrix1DVBC[4](A, phi)  # the device is still healthy
    assert np.allclose(vb.mul_(np.empty(A.n), B.T, np.ones(A.m)), A.to_scipy().T @ np.ones(A.m))


def test_host_vector_multiply_replays_a_graph_for_repeated_pinned_buffers():
    """VBC_OPT_E2E_GRAPH: the iterative caller's `mul!(y, A', x)` with the same pinned x / y is captured at the second call and
    replayed afterwards -- the results follow the CONTENTS of the buffers; pageable buffers and changing pointers stay eager."""
    import torch
    rng = np.random.default_rng(17)
    A, pi, phi = synth.config_c2(n=200_000, S=23)
    S = A.to_scipy()
    B = vb.SparseMatrixVBC[4, 4](A, pi, phi)
    x = torch.empty(A.m, dtype=torch.float64).pin_memory().numpy()
    y = torch.empty(A.n, dtype=torch.float64).pin_memory().numpy()
    for it in range(5):
        x[:] = rng.random(A.m)
        y[:] = np.nan
        vb.mul_(y, B.T, x)
        assert np.allclose(y, S.T @ x, rtol=1e-12), it
        assert B.get_option(_lib.OPT_E2E_GRAPH) == (2 if it >= 1 else 1)
    eager = vb.mul_(np.empty(A.n), B.T, x.copy())          # pageable buffers: eager path, same bits
    assert np.array_equal(eager, y) and B.get_option(_lib.OPT_E2E_GRAPH) == 1
    y0 = rng.random(A.n)
    for it in range(3):                                        # alpha / beta are part of the captured work
        y[:] = y0
        vb.mul_(y, B.T, x, 2.0, -0.5)
        assert np.allclose(y, 2.0 * (S.T @ x) - 0.5 * y0, rtol=1e-12)
    B.set_option(_lib.OPT_E2E_GRAPH, 0)
    y[:] = np.nan
    vb.mul_(y, B.T, x)
    assert np.array_equal(y, eager) and B.get_option(_lib.OPT_E2E_GRAPH) == 0


def test_triangular_solve_graph_replay_wide_stripes_and_singular():
    """a solve captured in a CUDA graph must stay correct when replayed (x carries the dependencies: no host-side
    epoch); stripes up to 32 wide; a zero diagonal is an ArgumentError at analysis."""
    import scipy.sparse as sp
    import torch
    from scipy.sparse.linalg import spsolve_triangular
    rng = np.random.default_rng(5)
    A, pi, phi = synth.config_c4_triangular(n=16_000, S=11)
    B = vb.SparseMatrixVBC[4, 4](A, pi, phi)
    T = sp.tril(A.to_scipy().T).tocsr()
    bd = torch.from_numpy(rng.random(A.n)).cuda()
    xd = torch.empty_like(bd)
    vb.ldiv_lower_(xd, B.T, bd)  # analysis + warm-up outside capture
    torch.cuda.synchronize()
    side, g = torch.cuda.Stream(), torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        vb.ldiv_lower_(xd, B.T, bd)
    for rep in range(3):
        bd.copy_(torch.from_numpy(rng.random(A.n)))
        xd.zero_()  # stale, finite x from an earlier solve must not satisfy any dependency
        g.replay()
        torch.cuda.synchronize()
        xo = spsolve_triangular(T, bd.cpu().numpy(), lower=True)
        assert np.max(np.abs(xd.cpu().numpy() - xo)) <= 1e-11 * max(1.0, np.max(np.abs(xo))), rep
    B.sync()
    # wide stripes (one lane per unknown of a row block)
    for w in (12, 16, 32):
        n = 640
        S = sp.random(n, n, density=0.05, random_state=np.random.RandomState(w), format="csc")
        S = sp.triu(S) + sp.identity(n) * 8.0   # A upper triangular <=> A' lower triangular
        M = vb.SparseMatrixCSC.from_scipy(sp.csc_matrix(S))
        Bw = vb.SparseMatrix1DVBC[w](M, vb.EquiChunker(w))
        b = rng.random(n)
        x = vb.ldiv_lower_(np.empty(n), Bw.T, b)
        xo = spsolve_triangular(sp.tril(S.T).tocsr(), b, lower=True)
        assert np.max(np.abs(x - xo)) <= 1e-11 * max(1.0, np.max(np.abs(xo))), w
    # singular: a missing diagonal entry
    S = sp.identity(16, format="lil") * 2.0
    S[5, 5] = 0.0
    S[2, 9] = 1.0
    M = vb.SparseMatrixCSC.from_scipy(sp.csc_matrix(S))
    Bs = vb.SparseMatrix1DVBC[4](M, vb.EquiChunker(4))
    with pytest.raises(vb.ArgumentError):
        vb.trsv_analyse(Bs.T)
    with pytest.raises(vb.ArgumentError):
        xb = np.ones(16)
        vb.ldiv_lower_(xb, Bs.T, xb)  # in place is refused


def test_benchmark_table_on_matrix_market_inputs(fixtures, tmp_path):
    """the reference's benchmark table (bin/test_table.jl:27-129) over Matrix Market files -- two of the
    SuiteSparse matrices the reference's own tests hold (test/matrices.jl:5-6), written as .mtx and read back through the
    script's reader (MatrixDepot itself needs the network): every method row packs, multiplies `y = B'x` and passes the
    script's `y ≈ z` check (test_table.jl:42/:84/:126); memory is the reference's format accounting."""
    import json
    import subprocess
    import sys
    import scipy.io
    paths = []
    for name in ("HB__west0132", "LPnetlib__lp_etamacro"):
        pth = str(tmp_path / (name + ".mtx"))
        scipy.io.mmwrite(pth, fixtures[name].to_scipy())
        paths.append(pth)
    out = str(tmp_path / "table.json")
    env = dict(os.environ, TABLE_OUT=out, TABLE_CACHE_BYTES=str(2 << 20))
    from conftest import ROOT
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "test_table.py")] + paths, capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    table = json.load(open(out))
    assert set(table) == {"HB__west0132.mtx", "LPnetlib__lp_etamacro.mtx"}
    for name, rows in table.items():
        methods = [row["method"] for row in rows]
        assert methods[0].startswith("reference") and len(methods) == 1 + 5 + 7
        assert all(row["runtime"] > 0 and row["memory"] > 0 for row in rows)
        strict = next(row for row in rows if row["method"] == "strict")
        minmem = next(row for row in rows if row["method"] == "min memory")
        assert minmem["memory"] <= strict["memory"]  # the DP under the memory model cannot do worse than the strict chunker


def test_device_layout_limits_answer_elimit():
    """Limits the reference does not have must be refused with VBC_ELIMIT, not overflow: a stripe wider than the pack
    kernel holds (32 columns) and dimensions beyond the compact layout's 32-bit fields (2^31)."""
    import ctypes
    A = sprand(20, 70, 0.3, np.random.default_rng(3))
    phi = vb.SplitPartition(np.array([1, 34, 60, 71], dtype=np.int64))  # widths 33, 26, 11
    with pytest.raises(_lib.VBCError) as ei:
        vb.SparseMatrix1DVBC[40](A, phi)
    assert ei.value.code == _lib.VBC_ELIMIT and "32" in str(ei.value)
    with pytest.raises(AssertionError):  # the reference's own `@assert w <= W` (constructors_1DVBC.jl:46) still comes first
        vb.SparseMatrix1DVBC[8](A, phi)
    one = np.array([1], dtype=np.int64)
    L = _lib.lib()
    h = ctypes.c_void_p()
    vp = lambda a: ctypes.c_void_p(a.ctypes.data)
    rc = L.vbc_pack_csc(ctypes.byref(h), _lib.VBC_F64, _lib.VBC_I64, 1 << 31, 0, 0, 4, vp(one), vp(one), vp(one), None, 0, vp(one), 0, 0)
    assert rc == _lib.VBC_ELIMIT and b"32-bit" in L.vbc_last_error() and not h.value
    rc = L.vbc_upload(ctypes.byref(h), _lib.VBC_F64, _lib.VBC_I64, 1 << 31, 0, 0, 4, None, 0, vp(one), 0, vp(one), vp(one), vp(one), vp(one), 0)
    assert rc == _lib.VBC_ELIMIT and not h.value


def _dense_of(A):
    D = np.zeros((A.m, A.n), dtype=A.nzval.dtype)
    for j in range(A.n):
        for q in range(A.colptr[j] - 1, A.colptr[j + 1] - 1):
            D[A.rowval[q] - 1, j] = A.nzval[q]
    return D


@pytest.mark.parametrize("ti", [np.int64, np.int32])
@pytest.mark.parametrize("tv", [np.int32, np.int64])
def test_integer_element_types_exact_with_wrapping(tv, ti):
    """runtests.jl:16 with the element type kept (Int32; Int64 is Julia's default Int): pack bit-exact against the oracle's integer
    instantiation, every multiply exactly equal to Julia's wrapping arithmetic (numpy's modular integer matmul) in both directions,
    1D and 2D, uniform and variable blocks, narrow and wide stripes, host arrays and device tensors, alpha / beta included."""
    import torch
    rng = np.random.default_rng(1616)
    info = np.iinfo(tv)
    tdt = torch.int32 if tv == np.int32 else torch.int64
    cases = [(m, n, 4, 4, "rand") for m in (1, 3, 8, 17) for n in (1, 4, 9, 16)]
    cases += [(64, 48, 4, 4, "strict"), (66, 50, 4, 4, "strict"), (90, 75, 8, 12, "rand"), (300, 211, 5, 3, "rand")]
    for m, n, U, W, how in cases:
        A = sprand_typed(m, n, 0.3, rng, tv, ti)
        D = _dense_of(A)
        ch = (lambda w, s: vb.StrictChunker(w)) if how == "strict" else (lambda w, s: vb.RandomChunker(w, seed=s))
        phi = vb.pack_stripe(A, ch(W, m + n))
        pi, phi2 = vb.pack_plaid(A, vb.AlternatingPacker(ch(W, n), ch(U, m)))  # columns first
        H1 = oracle.pack_1d(m, n, A.colptr, A.rowval, A.nzval, phi.spl, W)
        H2 = oracle.pack_2d(m, n, A.colptr, A.rowval, A.nzval, pi.spl, phi2.spl, U, W)
        B1 = vb.SparseMatrix1DVBC[W](A, phi)
        B2 = vb.SparseMatrixVBC[U, W](A, pi, phi2)
        assert B1.Tv == np.dtype(tv) and B2.Tv == np.dtype(tv)
        for B, H in ((B1, H1), (B2, H2)):
            assert_packed_equal(B, H)
            with np.errstate(over="ignore"):
                for trans in (False, True):
                    xlen, ylen = (m, n) if trans else (n, m)
                    Dop = D.T if trans else D
                    x = rng.integers(info.min, info.max, size=xlen, dtype=tv, endpoint=True)
                    y0 = rng.integers(info.min, info.max, size=ylen, dtype=tv, endpoint=True)
                    op = B.T if trans else B
                    y = vb.mul_(y0.copy(), op, x)
                    assert np.array_equal(y, Dop @ x), (m, n, trans, "product")
                    assert np.array_equal(y, oracle.mul(H, x, trans=trans)), (m, n, trans, "oracle")
                    y = vb.mul_(y0.copy(), op, x, 3, -2)
                    assert np.array_equal(y, np.asarray(3, dtype=tv) * (Dop @ x) + np.asarray(-2, dtype=tv) * y0), (m, n, trans, "alpha, beta")
                    xd, yd = torch.from_numpy(x).cuda(), torch.from_numpy(y0.copy()).cuda()
                    assert xd.dtype == tdt
                    vb.mul_(yd, op, xd, True, True)
                    torch.cuda.synchronize()
                    assert np.array_equal(yd.cpu().numpy(), Dop @ x + y0), (m, n, trans, "device tensors, beta = 1")
                    assert np.array_equal((op @ x), Dop @ x)
                xt = rng.integers(info.min, info.max, size=m, dtype=tv, endpoint=True)
                assert np.array_equal(vb.TrSpMV_(np.empty(n, dtype=tv), A, xt), D.T @ xt)
    # scalars that are not integers of the element type: the InexactError of convert(eltype(y), alpha)
    with pytest.raises(vb.ArgumentError, match="InexactError"):
        vb.mul_(np.zeros(m, dtype=tv), B1, np.zeros(n, dtype=tv), 0.5, 0)
    # floating-point-only entry points refuse integer matrices instead of reinterpreting their bits
    with pytest.raises(vb.ArgumentError):
        vb.mul_(np.zeros((m, 2), dtype=tv), B1, np.zeros((n, 2), dtype=tv))
    Asq = sprand_typed(12, 12, 0.3, rng, tv, ti)
    with pytest.raises(vb.ArgumentError):
        vb.trsv_analyse(vb.SparseMatrix1DVBC[4](Asq, vb.pack_stripe(Asq, vb.StrictChunker(4))).T)


def test_bool_matrices_onehot_like_runtests_jl():
    """runtests.jl:15 (`sprand(Bool, m, n, 0.2)`) over the whole size grid with Bool vectors: one-hot products in both directions equal
    the matrix' columns / rows (:29-53, :63-87); val downloads as Bool; a sum above 1 is an InexactError as in Julia."""
    rng = np.random.default_rng(15)
    for m in SIZES:
        for n in SIZES:
            A = sprand_typed(m, n, 0.2, rng, np.bool_)
            D = _dense_of(A)
            A32 = vb.SparseMatrixCSC(m, n, A.colptr, A.rowval, A.nzval.astype(np.int32))
            phi = vb.pack_stripe(A, vb.StrictChunker(4))
            pi, phi2 = vb.pack_plaid(A, vb.AlternatingPacker(vb.StrictChunker(4), vb.StrictChunker(4)))
            H1 = oracle.pack_1d(m, n, A32.colptr, A32.rowval, A32.nzval, phi.spl, 4)
            H2 = oracle.pack_2d(m, n, A32.colptr, A32.rowval, A32.nzval, pi.spl, phi2.spl, 4, 4)
            for B, H in ((vb.SparseMatrix1DVBC[4](A, phi), H1), (vb.SparseMatrixVBC[4, 4](A, pi, phi2), H2)):
                assert B.eltype == np.dtype(np.bool_)
                d = B.download()
                assert d["val"].dtype == np.bool_ and np.array_equal(d["val"], H.val.astype(np.bool_))
                for f in ("pos", "idx", "ofs"):
                    assert np.array_equal(d[f], getattr(H, f)), f
                x = np.zeros(n, dtype=np.bool_)
                y = np.empty(m, dtype=np.bool_)
                for j in range(n):
                    x[j] = True
                    y.fill(True)
                    vb.mul_(y, B, x, True, False)
                    assert y.dtype == np.bool_ and np.array_equal(y, D[:, j]), f"forward col {j + 1}"
                    x[j] = False
                x = np.zeros(m, dtype=np.bool_)
                y = np.empty(n, dtype=np.bool_)
                for i in range(m):
                    x[i] = True
                    y.fill(True)
                    vb.mul_(y, B.T, x, True, False)
                    assert np.array_equal(y, D[i, :]), f"adjoint row {i + 1}"
                    x[i] = False
    A = sprand_typed(9, 9, 0.9, rng, np.bool_)
    B = vb.SparseMatrix1DVBC[4](A, vb.pack_stripe(A, vb.StrictChunker(4)))
    with pytest.raises(vb.InexactError):
        vb.mul_(np.zeros(9, dtype=np.bool_), B, np.ones(9, dtype=np.bool_))
    xt = np.zeros(9, dtype=np.bool_)
    xt[2] = True
    assert np.array_equal(vb.TrSpMV_(np.empty(9, dtype=np.bool_), A, xt), _dense_of(A)[2, :])
