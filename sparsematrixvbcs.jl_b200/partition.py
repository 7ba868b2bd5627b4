"""Host-side inputs of the pack path: CSC container, SplitPartition, and stand-ins for the
ChainPartitioners chunkers the reference's tests and cost-model experiments use.

ChainPartitioners.jl (v1.1.6, un-vendored dependency; Manifest.toml:23-29) is NOT under
/root/reference, so only chunkers that are fully determined by their definition are provided:

  EquiChunker(w)        fixed-width chunks (costs.jl:83, :220; bin/test_table.jl:89)
  StrictChunker(w_max)  adjacent columns with identical patterns, width <= w_max
                        (test/runtests.jl:20; relied on by constructors_1DVBC.jl:94-143)
  DynamicTotalChunker(model, w_max)  minimum-total-cost contiguous partition, stripes <= w_max wide
                        (test/runtests.jl:22-23, bin/test_table.jl:66-69 `DynamicTotalChunker(ConstrainedCost(
                        mdl, VertexCount(), w_max))`).  Restated from its definition: exact per-window
                        distinct-row (or weighted distinct-row-part) counts + the O(n W) recurrence
                        (libvbc's host function vbc_dp_chunk); tie-breaking parity unpinned.
  OverlapChunker(rho, w_max)  greedy Jaccard-overlap grouping (test/runtests.jl:21) -- ASSUMED definition, unpinned
  AlternatingPacker(m1, m2, m3, ...)  the reference's order: m1 partitions the COLUMNS, m2 the ROWS
                        given the columns, m3 the columns again given the rows, ... (bin/test_table.jl:
                        89-111: `AlternatingPacker(DynamicTotalChunker(..1D..), EquiChunker(1))` is
                        "1D columns, unit rows"; test/runtests.jl:57)
  RandomChunker(w_max, seed)  any valid SplitPartition with widths in 1..w_max (fuzzing: the
                        pack/multiply contract is partition-agnostic)

The partition is host INPUT to the device pack kernel (north_star (a)); nothing here runs on
the hot path.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


class SparseMatrixCSC:
    """Julia's `SparseMatrixCSC{Tv,Ti}`: m, n, colptr (n+1), rowval (nnz), nzval (nnz); indices
    1-based, rows ascending inside a column."""

    def __init__(self, m, n, colptr, rowval, nzval):
        self.m, self.n = int(m), int(n)
        self.colptr = np.ascontiguousarray(colptr)
        self.rowval = np.ascontiguousarray(rowval, dtype=self.colptr.dtype)
        self.nzval = np.ascontiguousarray(nzval)
        if self.colptr.dtype not in (np.dtype(np.int32), np.dtype(np.int64)):
            raise TypeError("Ti must be int32 or int64")
        if len(self.colptr) != self.n + 1:
            raise ValueError("colptr must have n+1 entries")

    @property
    def shape(self):
        return (self.m, self.n)

    @property
    def nnz(self):
        return int(self.colptr[-1]) - 1

    @classmethod
    def from_scipy(cls, A, ti=np.int64, tv=None):
        import scipy.sparse as sp
        A = sp.csc_matrix(A)
        A.sort_indices()
        A.sum_duplicates()
        tv = A.dtype if tv is None else tv
        return cls(A.shape[0], A.shape[1], (A.indptr + 1).astype(ti), (A.indices + 1).astype(ti),
                   A.data.astype(tv))

    def to_scipy(self):
        import scipy.sparse as sp
        return sp.csc_matrix((self.nzval, self.rowval.astype(np.int64) - 1, self.colptr.astype(np.int64) - 1),
                             shape=(self.m, self.n))

    def transpose(self):
        """`permutedims(A)` (bin/test_table.jl:27) as a new CSC matrix."""
        return SparseMatrixCSC.from_scipy(self.to_scipy().T.tocsc(), ti=self.colptr.dtype.type,
                                          tv=self.nzval.dtype.type)

    def astype(self, tv=None, ti=None):
        tv = self.nzval.dtype if tv is None else tv
        ti = self.colptr.dtype if ti is None else ti
        return SparseMatrixCSC(self.m, self.n, self.colptr.astype(ti), self.rowval.astype(ti),
                               self.nzval.astype(tv))


class SplitPartition:
    """ChainPartitioners' `SplitPartition{Ti}`: part k = spl[k]:spl[k+1]-1 (1-based),
    spl[1] = 1, spl[K+1] = n+1 (SparseMatrixVBCs.jl:39, multiply_1DVBC.jl:54,:59)."""

    def __init__(self, spl):
        self.spl = np.ascontiguousarray(spl)

    def __len__(self):
        return len(self.spl) - 1

    def astype(self, ti):
        return SplitPartition(self.spl.astype(ti))

    def widths(self):
        return np.diff(self.spl)

    def __repr__(self):
        return f"SplitPartition(K={len(self)})"


@dataclass
class EquiChunker:
    w: int = 1


@dataclass
class StrictChunker:
    w_max: int


@dataclass
class RandomChunker:
    w_max: int
    seed: int = 0


@dataclass
class OverlapChunker:
    """`OverlapChunker(ρ, w_max)` (test/runtests.jl:21).  ASSUMED definition -- ChainPartitioners is not vendored:
    greedy, the next column joins while Jaccard(pattern, union of the stripe's patterns) >= ρ and width < w_max."""
    rho: float
    w_max: int


@dataclass
class DynamicTotalChunker:
    model: object  # costs.AffineConnectivityModel | ColumnBlockComponentCostModel | BlockComponentCostModel
    w_max: int


class AlternatingPacker:
    """`AlternatingPacker(m1, m2, ...)`: m1 -> Φ (columns), m2 -> Π (rows) given Φ, m3 -> Φ given Π, ..."""

    def __init__(self, *methods):
        if len(methods) < 2:
            raise TypeError("AlternatingPacker needs at least a column method and a row method")
        self.methods = methods


def permutedims(model):
    """`permutedims(mdl)` of a BlockComponentCostModel: the same model seen from the rows."""
    from .costs import BlockComponentCostModel
    if isinstance(model, BlockComponentCostModel):
        return BlockComponentCostModel(model.alpha_col, model.alpha_row, model.beta_col, model.beta_row)
    return model


def _equi(n, w, ti):
    spl = np.arange(1, n + 1, w, dtype=ti)
    return SplitPartition(np.append(spl, ti(n + 1)).astype(ti))


def _strict(A: SparseMatrixCSC, w_max):
    n, ti = A.n, A.colptr.dtype.type
    if n == 0:
        return SplitPartition(np.array([1], dtype=ti))
    cp = A.colptr.astype(np.int64) - 1
    cnt = np.diff(cp)
    same = np.zeros(n, dtype=bool)  # same[j]: column j has the pattern of column j-1
    cand = np.flatnonzero(cnt[1:] == cnt[:-1]) + 1
    if len(cand):
        lens = cnt[cand]
        tot = int(lens.sum())
        if tot:
            seg = np.repeat(np.arange(len(cand)), lens)
            off = np.arange(tot) - np.repeat(np.cumsum(lens) - lens, lens)
            a = A.rowval[cp[cand][seg] + off]
            b = A.rowval[cp[cand - 1][seg] + off]
            diff = np.zeros(len(cand), dtype=np.int64)
            np.add.at(diff, seg, (a != b).astype(np.int64))
            same[cand] = diff == 0
        else:
            same[cand] = True
        same[cand[lens == 0]] = True
    # position inside a run of identical columns; a new stripe every w_max columns
    run_start = ~same
    run_id = np.cumsum(run_start) - 1
    first = np.flatnonzero(run_start)
    pos_in_run = np.arange(n) - first[run_id]
    starts = np.flatnonzero(pos_in_run % w_max == 0)
    return SplitPartition(np.append(starts + 1, n + 1).astype(ti))


def _random(n, w_max, seed, ti):
    rng = np.random.default_rng(seed)
    spl = [1]
    while spl[-1] <= n:
        spl.append(min(n + 1, spl[-1] + int(rng.integers(1, w_max + 1))))
    return SplitPartition(np.array(spl, dtype=ti))


def window_weight_sums(A: SparseMatrixCSC, W, pi: SplitPartition = None, weights=None):
    """D[b, w-1, r] = sum over the DISTINCT units touched by columns [b-w, b) of weight_r(unit), for b in 0..n and
    w in 1..W (zero where b < w).  Units are rows (pi is None) or row parts of `pi`; weights is an (R, #units)
    array (default: one all-ones row => plain distinct counts, `hst`-style as constructors_1DVBC.jl:26-30).

    Exact and O(nnz + n W R): a unit's entry in column j counts for every window [j, b) with b <= next column in
    which the same unit appears, so per column the entries are binned by that gap."""
    n, m = A.n, A.m
    cp = A.colptr.astype(np.int64) - 1
    rows = A.rowval.astype(np.int64) - 1
    cols = np.repeat(np.arange(n, dtype=np.int64), np.diff(cp))
    if pi is not None:
        unit = np.searchsorted(pi.spl.astype(np.int64) - 1, rows, side="right") - 1
        nunits = len(pi)
        key = unit * n + cols                      # one entry per (part, column)
        key = np.unique(key)
        unit, cols = key // n, key % n
    else:
        unit, nunits = rows, m
    weights = np.ones((1, nunits)) if weights is None else np.asarray(weights, dtype=np.float64)
    R = weights.shape[0]
    order = np.lexsort((cols, unit))               # by unit, then column
    u_s, c_s = unit[order], cols[order]
    gap = np.full(len(u_s), W, dtype=np.int64)     # distance to the unit's next column, capped at W
    same = u_s[1:] == u_s[:-1]
    gap[:-1][same] = np.minimum(c_s[1:][same] - c_s[:-1][same], W)
    # c[j, t-1, r] = weight of column j's entries whose gap >= t
    cgt = np.zeros((n, W, R))
    for r in range(R):
        hist = np.zeros((n, W))
        np.add.at(hist, (c_s, gap - 1), weights[r][u_s])
        cgt[:, :, r] = np.cumsum(hist[:, ::-1], axis=1)[:, ::-1]
    D = np.zeros((n + 1, W, R))
    for w in range(1, W + 1):                      # D[b, w] = D[b, w-1] + c[b-w][w]
        prev = D[w:, w - 2, :] if w > 1 else 0.0
        D[w:, w - 1, :] = prev + cgt[: n + 1 - w, w - 1, :]
    return D


def _dp_cost_table(A, method: DynamicTotalChunker, other: SplitPartition = None):
    from .costs import AffineConnectivityModel, BlockComponentCostModel, ColumnBlockComponentCostModel
    mdl, W = method.model, int(method.w_max)
    ev = ColumnBlockComponentCostModel._ev
    ws = np.arange(1, W + 1)
    if isinstance(mdl, BlockComponentCostModel):
        if other is None:  # no row partition yet: every row is its own part
            other = SplitPartition(np.arange(1, A.m + 2, dtype=np.int64))
        heights = np.diff(other.spl).astype(np.int64)
        wts = np.stack([np.array([ev(br, u) for u in heights], dtype=np.float64) for br in mdl.beta_row])
        D = window_weight_sums(A, W, other, wts)            # (n+1, W, R)
        bc = np.stack([np.array([ev(b, w) for w in ws], dtype=np.float64) for b in mdl.beta_col], axis=1)  # (W, R)
        ac = np.array([ev(mdl.alpha_col, w) for w in ws], dtype=np.float64)
        return ac[None, :] + (D * bc[None, :, :]).sum(axis=2)
    D = window_weight_sums(A, W)[:, :, 0]                   # distinct rows per window
    if isinstance(mdl, AffineConnectivityModel):
        return mdl.alpha + mdl.beta_width * ws[None, :] + mdl.beta_net * D
    if isinstance(mdl, ColumnBlockComponentCostModel):
        ac = np.array([ev(mdl.alpha_col, w) for w in ws], dtype=np.float64)
        bc = np.array([ev(mdl.beta_col, w) for w in ws], dtype=np.float64)
        return ac[None, :] + D * bc[None, :]
    raise TypeError(f"unsupported cost model {mdl!r}")


def _dynamic_total(A: SparseMatrixCSC, method: DynamicTotalChunker, other: SplitPartition = None):
    import ctypes
    from . import _lib
    ti = A.colptr.dtype.type
    cost = np.ascontiguousarray(_dp_cost_table(A, method, other), dtype=np.float64)
    spl = np.empty(A.n + 1, dtype=np.int64)
    L = ctypes.c_int64()
    _lib.check(_lib.lib().vbc_dp_chunk(A.n, int(method.w_max), cost.ctypes.data_as(ctypes.c_void_p),
                                       spl.ctypes.data_as(ctypes.c_void_p), ctypes.byref(L)))
    return SplitPartition(spl[: L.value + 1].astype(ti))


def pack_stripe(A: SparseMatrixCSC, method, other: SplitPartition = None) -> SplitPartition:
    """`pack_stripe(A, method)`: contiguous partition of A's COLUMNS (constructors_1DVBC.jl:5, :100;
    costs.jl:83).  `other` is the partition of the other dimension when called by an AlternatingPacker."""
    ti = A.colptr.dtype.type
    if isinstance(method, SplitPartition):
        return method
    if isinstance(method, EquiChunker):
        return _equi(A.n, method.w, ti)
    if isinstance(method, StrictChunker):
        return _strict(A, method.w_max)
    if isinstance(method, RandomChunker):
        return _random(A.n, method.w_max, method.seed, ti)
    if isinstance(method, DynamicTotalChunker):
        return _dynamic_total(A, method, other)
    if isinstance(method, OverlapChunker):
        import ctypes
        from . import _lib
        cp = np.ascontiguousarray(A.colptr.astype(np.int64) - 1)
        rv = np.ascontiguousarray(A.rowval.astype(np.int64) - 1)
        spl = np.empty(A.n + 1, dtype=np.int64)
        L = ctypes.c_int64()
        _lib.check(_lib.lib().vbc_overlap_chunk(A.n, cp.ctypes.data_as(ctypes.c_void_p), rv.ctypes.data_as(ctypes.c_void_p),
                                                float(method.rho), int(method.w_max), spl.ctypes.data_as(ctypes.c_void_p), ctypes.byref(L)))
        return SplitPartition(spl[: L.value + 1].astype(ti))
    raise TypeError(f"unsupported partitioner {method!r} (ChainPartitioners is not vendored; "
                    "pass a SplitPartition computed on the Julia side)")


def pack_plaid(A: SparseMatrixCSC, method):
    """`pack_plaid(A, method)` -> (Π, Φ) (constructors_VBC.jl:11): columns first, then rows given the columns,
    then columns given the rows, ... one method per step."""
    if not isinstance(method, AlternatingPacker):
        raise TypeError(f"unsupported packer {method!r}")
    At = A.transpose()
    pi = phi = None
    for step, mth in enumerate(method.methods):
        if step % 2 == 0:
            phi = pack_stripe(A, mth, other=pi)
        else:
            pi = pack_stripe(At, mth, other=phi)  # the caller passes `permutedims(model)` for row steps, as the reference does
    return pi, phi
