"""Host-side inputs of the pack path: CSC container, SplitPartition, and stand-ins for the
ChainPartitioners chunkers the reference's tests and cost-model experiments use.

ChainPartitioners.jl (v1.1.6, un-vendored dependency; Manifest.toml:23-29) is NOT under
/root/reference, so only chunkers that are fully determined by their definition are provided:

  EquiChunker(w)        fixed-width chunks (costs.jl:83, :220; bin/test_table.jl:89)
  StrictChunker(w_max)  adjacent columns with identical patterns, width <= w_max
                        (test/runtests.jl:20; relied on by constructors_1DVBC.jl:94-143)
  AlternatingPacker(row_chunker, col_chunker)  -- for the two chunkers above one alternation
                        is a fixed point, so it reduces to (rows of A', columns of A)
                        (test/runtests.jl:57)
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
class AlternatingPacker:
    row: object
    col: object


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


def pack_stripe(A: SparseMatrixCSC, method) -> SplitPartition:
    """`pack_stripe(A, method)`: contiguous partition of A's COLUMNS
    (constructors_1DVBC.jl:5, :100; costs.jl:83)."""
    ti = A.colptr.dtype.type
    if isinstance(method, EquiChunker):
        return _equi(A.n, method.w, ti)
    if isinstance(method, StrictChunker):
        return _strict(A, method.w_max)
    if isinstance(method, RandomChunker):
        return _random(A.n, method.w_max, method.seed, ti)
    raise TypeError(f"unsupported partitioner {method!r} (ChainPartitioners is not vendored; "
                    "pass a SplitPartition computed on the Julia side)")


def pack_plaid(A: SparseMatrixCSC, method):
    """`pack_plaid(A, method)` -> (Π, Φ) (constructors_VBC.jl:11)."""
    if not isinstance(method, AlternatingPacker):
        raise TypeError(f"unsupported packer {method!r}")
    At = A.transpose()
    return pack_stripe(At, method.row), pack_stripe(A, method.col)
