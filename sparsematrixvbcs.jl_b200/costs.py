"""Cost models of the VBC formats and the GPU time-model autotuner -- the device-side analogue of
/root/reference/src/costs.jl ("next" row N1 of SURVEY.md 8f; a CALLER of the hot path, host logic only).

    reference (costs.jl)                                  here
    ---------------------------------------------------   ------------------------------------------
    Line(a, b)                                   :1-6      Line
    model_SparseMatrix1DVBC_blocks()             :8        model_SparseMatrix1DVBC_blocks()
    model_SparseMatrix1DVBC_memory(Tv, Ti)       :10       model_SparseMatrix1DVBC_memory(Tv, Ti)
    model_SparseMatrix1DVBC_TrSpMV_time_data     :14-107   model_SparseMatrix1DVBC_TrSpMV_time_data   (times mul_(y, B.T, x) on the GPU)
    model_SparseMatrix1DVBC_TrSpMV_time_params   :109-136  model_SparseMatrix1DVBC_TrSpMV_time_params (relative LSQ, monotonise)
    model_SparseMatrixVBC_blocks / _memory       :138-140  model_SparseMatrixVBC_blocks / _memory
    model_SparseMatrixVBC_TrSpMV_time_data       :144-247  model_SparseMatrixVBC_TrSpMV_time_data
    model_SparseMatrixVBC_TrSpMV_time_params     :249-290  model_SparseMatrixVBC_TrSpMV_time_params   (+ rank-R SVD)

The reference sizes its synthetic matrices against the CPU's L2 cache (`cache=:L2Cache`, half of it, or twice it
with `exceed=true`) and times `mul!(y, B', x)` with BenchmarkTools (minimum sample).  Here the cache is the
GPU's L2, the timing is a CUDA graph of repeated `mul_(y, B.T, x)` calls (best replay), and results are
memoised on disk per GPU model (the reference memoises per CPU `arch_id()`, util.jl:52).

The fitted model predicts  time(stripe partition) = sum_l alpha_col[w_l] + rows_l * beta_col[w_l]  (1D) and
sum_k alpha_row[u_k] + sum_l alpha_col[w_l] + sum_blocks sum_r beta_row[r][u] * beta_col[r][w]  (2D) -- the
ColumnBlockComponentCostModel / BlockComponentCostModel shapes ChainPartitioners' DynamicTotalChunker consumes.
"""
from __future__ import annotations

import hashlib
import json
import os
from dataclasses import dataclass, field

import numpy as np


@dataclass
class Line:
    """`Line(a, b)(x) = a + b*x` (costs.jl:1-6)."""
    a: float
    b: float

    def __call__(self, x):
        return self.a + self.b * x


@dataclass
class AffineConnectivityModel:
    """cost(stripe) = a + b*w + c*rows*w?  The reference only instantiates (0, 0, 0, 1): the number of
    distinct rows of the stripe (block count), costs.jl:8."""
    alpha: float = 0
    beta_width: float = 0
    beta_work: float = 0
    beta_net: float = 1

    def stripe_value(self, w, rows):
        return self.alpha + self.beta_width * w + self.beta_net * rows


@dataclass
class ColumnBlockComponentCostModel:
    """cost(stripe) = alpha_col(w) + rows * beta_col(w); alpha_col / beta_col are numbers, Lines in w, or
    per-width tables indexed 1..W (costs.jl:10, :12)."""
    alpha_col: object
    beta_col: object

    @staticmethod
    def _ev(f, w):
        if isinstance(f, Line):
            return f(w)
        if np.ndim(f) == 0:
            return float(f)
        return float(np.asarray(f)[int(w) - 1])

    def stripe_value(self, w, rows):
        return self._ev(self.alpha_col, w) + rows * self._ev(self.beta_col, w)


@dataclass
class BlockComponentCostModel:
    """cost = sum_k alpha_row(u_k) + sum_l alpha_col(w_l) + sum_blocks sum_r beta_row[r](u) * beta_col[r](w)
    (costs.jl:138-142)."""
    alpha_row: object
    alpha_col: object
    beta_row: tuple = field(default_factory=tuple)
    beta_col: tuple = field(default_factory=tuple)

    def block_value(self, u, w):
        ev = ColumnBlockComponentCostModel._ev
        return sum(ev(br, u) * ev(bc, w) for br, bc in zip(self.beta_row, self.beta_col))


def model_SparseMatrix1DVBC_blocks():
    return AffineConnectivityModel(0, 0, 0, 1)


def model_SparseMatrix1DVBC_memory(Tv, Ti):
    tv, ti = np.dtype(Tv).itemsize, np.dtype(Ti).itemsize
    return ColumnBlockComponentCostModel(3 * ti, Line(ti, tv))


def model_SparseMatrixVBC_blocks():
    return BlockComponentCostModel(0, 0, (1,), (1,))


def model_SparseMatrixVBC_memory(Tv, Ti):
    tv, ti = np.dtype(Tv).itemsize, np.dtype(Ti).itemsize
    return BlockComponentCostModel(ti, 3 * ti, (Line(1, 0), Line(0, 1)), (Line(ti, 0), Line(0, tv)))


def total_value(B, model):
    """Model cost of a packed matrix (`total_value(A, Φ, mdl)` / `total_value(A, Π, Φ, mdl)` +
    `row_component_value`, bin/test_table.jl:60, :82, :124), from the packed arrays."""
    d = B.download()
    widths = np.diff(B.Phi.spl).astype(np.int64)
    units = np.diff(d["pos"]).astype(np.int64)
    if isinstance(model, (ColumnBlockComponentCostModel, AffineConnectivityModel)):
        return float(sum(model.stripe_value(w, r) for w, r in zip(widths, units)))
    ev = ColumnBlockComponentCostModel._ev
    heights = np.diff(B.Pi.spl).astype(np.int64)
    total = sum(ev(model.alpha_row, u) for u in heights) + sum(ev(model.alpha_col, w) for w in widths)
    stripe_of_block = np.repeat(np.arange(len(widths)), units)
    bu = heights[d["idx"].astype(np.int64) - 1]
    bw = widths[stripe_of_block]
    for u in np.unique(bu):
        for w in np.unique(bw):
            total += int(((bu == u) & (bw == w)).sum()) * model.block_value(u, w)
    return float(total)


# ---- measurement ------------------------------------------------------------------------------------
def _cache_dir():
    d = os.path.join(os.environ.get("XDG_CACHE_HOME", os.path.join(os.path.expanduser("~"), ".cache")), "vbc_b200", "autotune")
    os.makedirs(d, exist_ok=True)
    return d


def arch_id():
    """Cache key of the device the timings belong to (the reference hashes cpuinfo(), util.jl:52)."""
    import torch
    p = torch.cuda.get_device_properties(0)
    return hashlib.sha256(f"{p.name}|{p.multi_processor_count}|{p.total_memory}|{p.major}.{p.minor}".encode()).hexdigest()[:16]


def _memo(name, key, compute, use_cache=True):
    path = os.path.join(_cache_dir(), f"{name}_{hashlib.sha256(json.dumps(key, sort_keys=True).encode()).hexdigest()[:20]}.json")
    if use_cache and os.path.exists(path):
        with open(path) as f:
            return json.load(f)
    out = compute()
    with open(path, "w") as f:
        json.dump(out, f)
    return out


def _l2_bytes():
    import torch
    return int(torch.cuda.get_device_properties(0).L2_cache_size)


def _time_adjoint(B, m, n, Tu, reps=20):
    """seconds per `mul_(y, B.T, x)`: minimum over 3 replays of a CUDA graph of `reps` calls (≙ BenchmarkTools' minimum)."""
    import torch
    from .matrix import mul_
    tdt = torch.float64 if np.dtype(Tu) == np.float64 else torch.float32
    x = torch.ones(m, dtype=tdt, device="cuda")
    y = torch.ones(n, dtype=tdt, device="cuda")
    for _ in range(3):
        mul_(y, B.T, x)
    torch.cuda.synchronize()
    side, g = torch.cuda.Stream(), torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        for _ in range(reps):
            mul_(y, B.T, x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = float("inf")
    for _ in range(3):
        with torch.cuda.stream(side):
            e0.record(); g.replay(); e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3 / reps)
    return best


def model_SparseMatrix1DVBC_TrSpMV_time_data(W, Tv, Ti, Tu, cache_bytes=None, exceed=False, use_cache=True, seed=0):
    """Synthetic sweep of costs.jl:14-107: for w = W..1, four matrices around the cache-sized shape
    (m0, L0, q0), q random distinct (row, stripe) positions each a dense w-wide row segment; returns
    dict(ms, ns, Ls, ws, qs, T)."""
    from .matrix import SparseMatrix1DVBC
    from .partition import EquiChunker
    from . import synth
    C0 = _l2_bytes() if cache_bytes is None else int(cache_bytes)
    tv, ti, tu = np.dtype(Tv).itemsize, np.dtype(Ti).itemsize, np.dtype(Tu).itemsize
    key = dict(kind="1d", W=W, Tv=str(np.dtype(Tv)), Ti=str(np.dtype(Ti)), Tu=str(np.dtype(Tu)), C=C0, exceed=exceed, arch=arch_id(), seed=seed)

    def compute():
        out = dict(ms=[], ns=[], Ls=[], ws=[], qs=[], T=[])
        d = 8
        for w in range(W, 0, -1):
            per_L = (3 + d) * ti + (2 * w) * tu + (d * w) * tv  # costs.jl:29-43
            if exceed:
                L0 = int(np.ceil(2 * C0 / per_L))
                space = ((L0 * w, L0, L0 * d), (L0 * w, 2 * L0, L0 * d), (2 * L0 * w, L0, L0 * d), (L0 * w, L0, 2 * L0 * d))
            else:
                L0 = int(np.floor(C0 / 2 / per_L))
                space = ((L0 * w, L0, L0 * d), (L0 * w, L0 // 2, L0 * d), (L0 * w // 2, L0, L0 * d), (L0 * w, L0, L0 * d // 2))
            assert L0 >= 4
            for (m, L, q) in space:
                per_stripe = max(1, q // L)
                A, _, Phi = synth.random_blocks(m, L, 1, w, per_stripe, seed=seed + w, dtype=Tv, ti=Ti)
                B = SparseMatrix1DVBC[W](A, Phi)
                t = _time_adjoint(B, A.m, A.n, Tu)
                B.close()
                out["ms"].append(int(A.m)); out["ns"].append(int(A.n)); out["Ls"].append(int(L))
                out["ws"].append(int(w)); out["qs"].append(int(per_stripe * L)); out["T"].append(float(t))
        return out

    return _memo("1DVBC_TrSpMV_timings", key, compute, use_cache)


def fit_1d(W, ms, Ls, ws, qs, T):
    """costs.jl:112-131: one-hot design [m | L at width w | q at width w], relative least squares
    (rows scaled by 1/T, target 1), then monotonise alpha_col, beta_col in w."""
    T = np.asarray(T, dtype=np.float64)
    D = np.zeros((len(T), 1 + 2 * W))
    for i in range(len(T)):
        D[i, 0] = ms[i]
        D[i, ws[i]] = Ls[i]
        D[i, W + ws[i]] = qs[i]
    P, *_ = np.linalg.lstsq(D / T[:, None], np.ones(len(T)), rcond=None)
    alpha_row, alpha_col, beta_col = P[0], P[1:1 + W].copy(), P[1 + W:].copy()
    for w in range(1, W):
        alpha_col[w] = max(alpha_col[w], alpha_col[w - 1])
        beta_col[w] = max(beta_col[w], beta_col[w - 1])
    return alpha_row, alpha_col, beta_col


def model_SparseMatrix1DVBC_TrSpMV_time_params(W, Tv, Ti, Tu, **kw):
    d = model_SparseMatrix1DVBC_TrSpMV_time_data(W, Tv, Ti, Tu, **kw)
    _, alpha_col, beta_col = fit_1d(W, d["ms"], d["Ls"], d["ws"], d["qs"], d["T"])
    return alpha_col, beta_col


def model_SparseMatrix1DVBC_TrSpMV_time(W, Tv, Ti, Tu, **kw):
    """`ColumnBlockComponentCostModel{Float64}(params...)` (costs.jl:12) with GPU-fitted coefficients (seconds)."""
    return ColumnBlockComponentCostModel(*model_SparseMatrix1DVBC_TrSpMV_time_params(W, Tv, Ti, Tu, **kw))


def model_SparseMatrixVBC_TrSpMV_time_data(U, W, Tv, Ti, Tu, cache_bytes=None, exceed=False, use_cache=True, seed=0):
    """Synthetic sweep of costs.jl:144-247 over (u, w): q random distinct (part, stripe) dense u x w blocks."""
    from .matrix import SparseMatrixVBC
    from . import synth
    C0 = _l2_bytes() if cache_bytes is None else int(cache_bytes)
    tv, ti, tu = np.dtype(Tv).itemsize, np.dtype(Ti).itemsize, np.dtype(Tu).itemsize
    key = dict(kind="2d", U=U, W=W, Tv=str(np.dtype(Tv)), Ti=str(np.dtype(Ti)), Tu=str(np.dtype(Tu)), C=C0, exceed=exceed, arch=arch_id(), seed=seed)

    def compute():
        out = dict(ms=[], ns=[], Ks=[], Ls=[], us=[], ws=[], qs=[], T=[])
        d = 8
        for u in range(U, 0, -1):
            for w in range(W, 0, -1):
                per_L = (3 + d + w / u) * ti + (2 * w) * tu + (d * u * w) * tv  # costs.jl:160-176
                if exceed:
                    L0 = int(np.ceil(2 * C0 / per_L)); K0 = -(-(L0 * w) // u)
                    space = ((K0, L0, L0 * d), (K0, 2 * L0, L0 * d), (2 * K0, L0, L0 * d), (K0, L0, 2 * L0 * d))
                else:
                    L0 = int(np.floor(C0 / 2 / per_L)); K0 = (L0 * w) // u
                    space = ((K0, L0, L0 * d), (K0, L0 // 2, L0 * d), (K0 // 2, L0, L0 * d), (K0, L0, L0 * d // 2))
                assert L0 >= 4 and K0 >= 4
                for (K, L, q) in space:
                    per_stripe = max(1, min(K, q // L))
                    A, Pi, Phi = synth.random_blocks(K, L, u, w, per_stripe, seed=seed + 100 * u + w, dtype=Tv, ti=Ti)
                    B = SparseMatrixVBC[U, W](A, Pi, Phi)
                    t = _time_adjoint(B, A.m, A.n, Tu)
                    B.close()
                    for k_, v_ in (("ms", A.m), ("ns", A.n), ("Ks", K), ("Ls", L), ("us", u), ("ws", w), ("qs", per_stripe * L)):
                        out[k_].append(int(v_))
                    out["T"].append(float(t))
        return out

    return _memo("VBC_TrSpMV_timings", key, compute, use_cache)


def fit_2d(R, U, W, Ks, Ls, us, ws, qs, T):
    """costs.jl:252-281: one-hot design [K at height u | L at width w | q at (u, w)], relative least squares,
    monotonise beta along u and w, rank-R SVD  beta ~ sum_r beta_row[r] beta_col[r]'."""
    T = np.asarray(T, dtype=np.float64)
    D = np.zeros((len(T), U + W + U * W))
    for i in range(len(T)):
        D[i, us[i] - 1] = Ks[i]
        D[i, U + ws[i] - 1] = Ls[i]
        D[i, U + W + (ws[i] - 1) * U + (us[i] - 1)] = qs[i]  # column-major (u, w), like reshape(d_block, :)
    P, *_ = np.linalg.lstsq(D / T[:, None], np.ones(len(T)), rcond=None)
    alpha_row, alpha_col = P[:U].copy(), P[U:U + W].copy()
    beta = P[U + W:].reshape(W, U).T.copy()
    for w in range(1, W):
        beta[0, w] = max(beta[0, w], beta[0, w - 1])
    for u in range(1, U):
        beta[u, 0] = max(beta[u, 0], beta[u - 1, 0])
        for w in range(1, W):
            beta[u, w] = max(beta[u, w], beta[u - 1, w - 1], beta[u, w - 1])
    Um, S, Vt = np.linalg.svd(beta)
    beta_row = tuple(Um[:, r].copy() for r in range(R))
    beta_col = tuple((S[r] * Vt[r, :]).copy() for r in range(R))
    return alpha_row, alpha_col, beta_row, beta_col, beta


def model_SparseMatrixVBC_TrSpMV_time_params(R, U, W, Tv, Ti, Tu, **kw):
    d = model_SparseMatrixVBC_TrSpMV_time_data(U, W, Tv, Ti, Tu, **kw)
    return fit_2d(R, U, W, d["Ks"], d["Ls"], d["us"], d["ws"], d["qs"], d["T"])[:4]


def model_SparseMatrixVBC_TrSpMV_time(R, U, W, Tv, Ti, Tu, **kw):
    return BlockComponentCostModel(*model_SparseMatrixVBC_TrSpMV_time_params(R, U, W, Tv, Ti, Tu, **kw))
