"""Synthetic block-structured matrices for the benchmark configs (host side, numpy).

Modelled on the reference's own generators for its time-model experiments
(/root/reference/src/costs.jl:63-85 1D, :200-222 2D): dense w-wide row segments / u x w blocks at
distinct (row part, stripe) positions, values `rand(Tv)` in [0, 1), partition `EquiChunker`.
Values come from a counter-based hash (splitmix64 of seed and the entry's (row, col)), so any
slab of the matrix can be regenerated independently (by another rank, or on the device) and
always gives the same numbers.  Seed default 0xDEADBEEF = the reference's test seed
(test/runtests.jl:11).
"""
from __future__ import annotations

import numpy as np

from .partition import SparseMatrixCSC, SplitPartition

SEED = 0xDEADBEEF
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(z):
    """Vectorised splitmix64 finaliser on uint64 arrays."""
    with np.errstate(over="ignore"):
        z = (z + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def entry_values(rows0, cols0, n, seed=SEED, dtype=np.float64):
    """Value of entry (row, col) (0-based): uniform in [0, 1) with 24 (f32) / 53 (f64) random bits."""
    with np.errstate(over="ignore"):
        key = rows0.astype(np.uint64) * np.uint64(n) + cols0.astype(np.uint64)
        h = splitmix64(key ^ np.uint64(seed))
    if np.dtype(dtype) == np.float32:
        return ((h >> np.uint64(40)).astype(np.float32) * np.float32(2.0 ** -24)).astype(np.float32)
    return (h >> np.uint64(11)).astype(np.float64) * (2.0 ** -53)


def vector(n, seed=1, dtype=np.float64):
    """Deterministic dense vector in [0, 1)."""
    return entry_values(np.arange(n, dtype=np.uint64), np.zeros(n, dtype=np.uint64), 1, seed=seed ^ 0x5851F42D, dtype=dtype)


def _csc_from_stripe_blocks(K, L, u, w, blk_stripe_counts, blk_parts, seed, dtype, ti, col0=0, n_global=None, diag_boost=0.0):
    """Uniform u x w blocks; stripe l owns blk_parts[bpos[l]:bpos[l+1]] (ascending part ids).
    col0 / n_global: the L stripes are a slab starting at global column col0 of an n_global-column
    matrix (values are hashed with GLOBAL coordinates, so slabs agree with the full matrix)."""
    m, n = K * u, L * w
    n_global = n if n_global is None else n_global
    nb = blk_stripe_counts.astype(np.int64)
    bpos = np.concatenate([[0], np.cumsum(nb)])
    # rows of a stripe, stripe-major: every block contributes u consecutive rows
    rows_stripe = (blk_parts.astype(np.int64)[:, None] * u + np.arange(u, dtype=np.int64)[None, :]).reshape(-1)
    col_stripe = np.repeat(np.arange(L, dtype=np.int64), w)          # stripe of each column
    col_len = (nb * u)[col_stripe]                                     # nnz per column
    colptr0 = np.concatenate([[0], np.cumsum(col_len)])
    nnz = int(colptr0[-1])
    seg_start = (bpos[:-1] * u)[col_stripe]                            # start of the stripe's row list
    gather = np.repeat(seg_start - colptr0[:-1], col_len) + np.arange(nnz, dtype=np.int64)
    rowval0 = rows_stripe[gather]
    del gather
    cols0 = np.repeat(np.arange(n, dtype=np.int64), col_len)
    nzval = entry_values(rowval0, cols0 + col0, n_global, seed=seed, dtype=dtype)
    if diag_boost:
        nzval[rowval0 == cols0 + col0] += np.dtype(dtype).type(diag_boost)
    del cols0
    A = SparseMatrixCSC(m, n, (colptr0 + 1).astype(ti), (rowval0 + 1).astype(ti), nzval)
    Pi = SplitPartition(np.arange(1, m + 2, u, dtype=ti))
    Phi = SplitPartition(np.arange(1, n + 2, w, dtype=ti))
    return A, Pi, Phi


def fem_stencil_offsets(S=63):
    """13-offset block stencil {0, ±1, ±2, ±S, ±(S±1), ±S²} of SURVEY.md 8(d) config C2 (a 3-D
    grid numbering with S³ ≈ K)."""
    offs = [0, 1, -1, 2, -2, S, -S, S + 1, -(S + 1), S - 1, -(S - 1), S * S, -(S * S)]
    return np.array(sorted(offs), dtype=np.int64)


def banded_blocks(K, L, u, w, offsets, seed=SEED, dtype=np.float64, ti=np.int64, stripes=None, diag_boost=0.0):
    """Block-banded matrix: stripe l has a dense u x w block at every row part l*K//L + δ,
    δ in `offsets`, clipped to [0, K).  Returns (A::SparseMatrixCSC, Π, Φ) with Π = Equi(u),
    Φ = Equi(w).  stripes=(l0, l1): only the column slab of stripes l0 <= l < l1 (A then has
    (l1-l0)*w columns and all K*u rows) -- what one rank of the row-partitioned multiply owns."""
    offsets = np.array(sorted(set(int(o) for o in offsets)), dtype=np.int64)
    l0, l1 = (0, L) if stripes is None else stripes
    center = (np.arange(l0, l1, dtype=np.int64) * K) // L
    kk = center[:, None] + offsets[None, :]
    ok = (kk >= 0) & (kk < K)
    counts = ok.sum(axis=1)
    parts = kk[ok]  # row-major flatten keeps each stripe's part ids ascending
    return _csc_from_stripe_blocks(K, l1 - l0, u, w, counts, parts, seed, dtype, ti, col0=l0 * w, n_global=L * w, diag_boost=diag_boost)


def random_blocks(K, L, u, w, per_stripe, seed=SEED, dtype=np.float64, ti=np.int64):
    """The reference's autotuner generator (costs.jl:63-85 / :200-222): `per_stripe` distinct
    random row parts in every stripe, dense u x w blocks.  u = 1 gives the 1D generator."""
    rng = np.random.default_rng(seed)
    per_stripe = min(per_stripe, K)
    if per_stripe * 4 <= K:
        # rejection-free for sparse stripes: sample, sort, fix duplicates by re-drawing whole rows
        parts = np.sort(rng.integers(0, K, size=(L, per_stripe), dtype=np.int64), axis=1)
        bad = np.flatnonzero((np.diff(parts, axis=1) == 0).any(axis=1))
        while len(bad):
            parts[bad] = np.sort(rng.integers(0, K, size=(len(bad), per_stripe), dtype=np.int64), axis=1)
            bad = bad[(np.diff(parts[bad], axis=1) == 0).any(axis=1)]
    else:
        parts = np.stack([np.sort(rng.choice(K, size=per_stripe, replace=False)) for _ in range(L)]).astype(np.int64)
    counts = np.full(L, per_stripe, dtype=np.int64)
    return _csc_from_stripe_blocks(K, L, u, w, counts, parts.reshape(-1), seed, dtype, ti)


def config_c1(dtype=np.float64, ti=np.int64):
    """C1: 1D-VBC, m = n = 10 000, W = 8, 100 distinct random rows per stripe => nnz = 1 000 000."""
    A, _, Phi = random_blocks(10_000, 1_250, 1, 8, 100, dtype=dtype, ti=ti)
    return A, Phi


def config_c2(n=1_000_000, u=4, w=4, S=63, dtype=np.float64, ti=np.int64, seed=SEED, stripes=None):
    """C2: 2D-VBC, m = n (default 1 000 000), U = W = 4, 13-block FEM-like stencil => nnz ≈ 52 M."""
    K, L = n // u, n // w
    return banded_blocks(K, L, u, w, fem_stencil_offsets(S), seed=seed, dtype=dtype, ti=ti, stripes=stripes)


def variable_partition(n, w_max, seed, ti=np.int64, w_min=2):
    """iid widths in [w_min, w_max] (config C2v: variable blocks)."""
    rng = np.random.default_rng(seed)
    widths = rng.integers(w_min, w_max + 1, size=n // w_min + 2)
    spl = np.concatenate([[1], 1 + np.cumsum(widths)])
    spl = spl[spl <= n]
    return SplitPartition(np.append(spl, n + 1).astype(ti) if spl[-1] != n + 1 else spl.astype(ti))


def config_c4_triangular(n=2_000_000, u=4, w=4, S=79, dtype=np.float64, ti=np.int64, seed=SEED):
    """C4 (BASELINE wording): square block matrix A whose adjoint A' is block lower triangular -- blocks at
    part offsets {0, -(S-1), -S, -(S+1), -S^2} (the lower half of the FEM stencil without the +-1, +-2
    neighbours, so the dependency depth is ~K/(S-1) row blocks instead of K), diagonal boosted so the
    solve tril(A') x = b is well conditioned."""
    K, L = n // u, n // w
    offs = [0, -(S - 1), -S, -(S + 1), -(S * S)]
    return banded_blocks(K, L, u, w, offs, seed=seed, dtype=dtype, ti=ti, diag_boost=float(8 * len(offs) * u))


def variable_block_matrix(n, per_stripe=8, g_max=6, band=2000, seed=SEED, dtype=np.float64, ti=np.int64):
    """Square matrix with NATURAL variable blocks (FEM-like supernodes): columns fall into groups of 1..g_max
    adjacent columns with identical patterns, rows likewise, and every column group holds `per_stripe` dense
    blocks at random row groups within +-band rows of the diagonal.  Returns (A, natural Π, natural Φ)."""
    rng = np.random.default_rng(seed)

    def groups(total):
        w = rng.integers(1, g_max + 1, size=total)
        s = np.concatenate([[0], np.cumsum(w)])
        s = s[s < total]
        return np.append(s, total).astype(np.int64)
    cs, rs = groups(n), groups(n)
    L, K = len(cs) - 1, len(rs) - 1
    rows_l, cols_l = [], []
    centre = np.searchsorted(rs, cs[:-1], side="right") - 1         # row group under the stripe's first column
    half = max(1, int(band / ((g_max + 1) / 2)))
    cand = centre[:, None] + rng.integers(-half, half + 1, size=(L, per_stripe))
    cand[:, 0] = centre                                             # keep the diagonal block
    cand = np.clip(cand, 0, K - 1)
    import scipy.sparse as sp
    lk = np.unique(np.stack([np.repeat(np.arange(L), per_stripe), cand.reshape(-1)], axis=1), axis=0)
    ls, ks = lk[:, 0], lk[:, 1]
    u, w = (rs[ks + 1] - rs[ks]), (cs[ls + 1] - cs[ls])
    cnt = u * w
    tot = int(cnt.sum())
    blk = np.repeat(np.arange(len(ls)), cnt)
    off = np.arange(tot) - np.repeat(np.cumsum(cnt) - cnt, cnt)
    rr = rs[ks][blk] + off // w[blk]
    cc = cs[ls][blk] + off % w[blk]
    vals = entry_values(rr, cc, n, seed=seed, dtype=dtype)
    A = SparseMatrixCSC.from_scipy(sp.csc_matrix((vals, (rr, cc)), shape=(n, n)), ti=ti, tv=dtype)
    return A, SplitPartition((rs + 1).astype(ti)), SplitPartition((cs + 1).astype(ti))


class DeviceCSC:
    """A SparseMatrixCSC slab living in device memory, produced by libvbc's generator (`vbc_gen_banded_csc`):
    colptr / rowval / nzval are torch CUDA tensors viewing memory owned by this object."""

    def __init__(self, m, n, nnz, ptrs, tv, ti, device):
        import torch
        self.m, self.n, self.nnz, self._ptrs, self.device = int(m), int(n), int(nnz), ptrs, int(device)
        self.tv, self.ti = np.dtype(tv), np.dtype(ti)

        def view(ptr, count, dt):
            class _W:
                pass
            o = _W()
            o.__cuda_array_interface__ = {"shape": (max(int(count), 1),), "typestr": np.dtype(dt).str, "data": (int(ptr), False), "version": 2}
            return torch.as_tensor(o, device=f"cuda:{device}")[: int(count)]
        self.colptr = view(ptrs[0], self.n + 1, ti)
        self.rowval = view(ptrs[1], self.nnz, ti)
        self.nzval = view(ptrs[2], self.nnz, tv)

    def free(self):
        if self._ptrs is not None:
            import ctypes

            from . import _lib
            self.colptr = self.rowval = self.nzval = None
            _lib.lib().vbc_gen_free(ctypes.c_void_p(self._ptrs[0]), ctypes.c_void_p(self._ptrs[1]), ctypes.c_void_p(self._ptrs[2]), self.device)
            self._ptrs = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def banded_blocks_device(K, L, u, w, offsets, seed=SEED, dtype=np.float64, ti=np.int64, stripes=None, diag_boost=0.0, device=0):
    """`banded_blocks` generated on the device (same matrix, same values): -> DeviceCSC slab of stripes [l0, l1)."""
    import ctypes

    from . import _lib
    offs = np.array(sorted(set(int(o) for o in offsets)), dtype=np.int64)
    l0, l1 = (0, L) if stripes is None else stripes
    vt = _lib.VBC_F64 if np.dtype(dtype) == np.dtype(np.float64) else _lib.VBC_F32
    it = _lib.VBC_I64 if np.dtype(ti) == np.dtype(np.int64) else _lib.VBC_I32
    cp, rv, nz = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
    nnz = ctypes.c_int64()
    _lib.check(_lib.lib().vbc_gen_banded_csc(vt, it, int(K), int(L), int(u), int(w), ctypes.c_void_p(offs.ctypes.data), len(offs), int(l0), int(l1),
                                            ctypes.c_uint64(seed), float(diag_boost), ctypes.byref(cp), ctypes.byref(rv), ctypes.byref(nz),
                                            ctypes.byref(nnz), int(device)))
    return DeviceCSC(K * u, (l1 - l0) * w, nnz.value, (cp.value, rv.value, nz.value), dtype, ti, device)


C5_OFFSETS = (0, 1, -1, 2, -2, 57, -57, 58, -58, 3249)  # BASELINE.json configs[4]: 10 blocks per stripe, banded
