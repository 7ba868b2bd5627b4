#!/usr/bin/env python
"""Per-kernel throughput table over the workloads of BASELINE.json (run on a B200 via gpurun).
Every row: kernel time from one CUDA-event pair per launch (median of 30), algorithmic bytes of the
device layout + vectors, achieved GB/s and useful GFLOP/s.  Written to gpurun_out/perf_table.json."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vbc_b200 as vb  # noqa: E402
from vbc_b200 import _lib, synth  # noqa: E402


def tk(fn, reps=50):
    """seconds per call: `reps` calls captured in one CUDA graph, replayed 3x, best replay (no host launch gaps)."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        for _ in range(reps):
            fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(3):
        with torch.cuda.stream(side):
            e0.record(); g.replay(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / reps * 1e-3)
    return sorted(ts)[1], min(ts)


rows = []


def tk_eager(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(reps):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e-3)
    return sorted(ts)[len(ts) // 2], min(ts)


def report(name, B, A, groups=(0,)):
    tdt = torch.float64 if B.Tv == np.float64 else torch.float32
    es = 8 if B.Tv == np.float64 else 4
    ref_b, adj_b, fwd_b = B.format_bytes()
    xm = torch.rand(A.m, dtype=tdt, device="cuda"); yn = torch.empty(A.n, dtype=tdt, device="cuda")
    xn = torch.rand(A.n, dtype=tdt, device="cuda"); ym = torch.empty(A.m, dtype=tdt, device="cuda")
    for g in groups:
        B.set_option(_lib.OPT_ADJ_GROUP, g); B.set_option(_lib.OPT_FWD_GROUP, g if g not in (4, 16) else 0)
        for kind, fn, nb in (("adjoint", lambda: vb.mul_(yn, B.T, xm), adj_b + es * (A.m + A.n)),
                             ("forward", lambda: vb.mul_(ym, B, xn), fwd_b + es * (A.n + 2 * A.m)),
                             ("fwd_tindex", lambda: vb.mul_(ym, B, xn), None)):
            if kind == "forward" and g in (4, 16):
                continue
            if kind == "fwd_tindex":
                if g != groups[0]:
                    continue
                B.set_option(_lib.OPT_FWD_MODE, 2)
                vb.mul_(ym, B, xn)  # builds the index outside graph capture
                nb = B.format_bytes()[2] + es * (A.n + A.m)
            elif kind == "forward":
                B.set_option(_lib.OPT_FWD_MODE, 1)
            med, mn = tk(fn)
            r = dict(workload=name, kernel=kind, group=g, us_med=med * 1e6, us_min=mn * 1e6, bytes=nb, gbs=nb / med / 1e9,
                     gflops=2.0 * A.nnz / med / 1e9, nnz=A.nnz, nval=B.nval, ref_format_bytes=ref_b)
            rows.append(r)
            print(f"{name:34s} {kind:8s} G={g:2d} {med * 1e6:8.1f} us  {r['gbs']:7.0f} GB/s  {r['gflops']:8.1f} GFLOP/s", flush=True)


def csc_row(name, A):
    tdt = torch.float64 if A.nzval.dtype == np.float64 else torch.float32
    es = A.nzval.dtype.itemsize
    C = vb.CuSparseMatrixCSC(A)
    xm = torch.rand(A.m, dtype=tdt, device="cuda"); yn = torch.empty(A.n, dtype=tdt, device="cuda")
    med, mn = tk(lambda: vb.TrSpMV_(yn, C, xm))
    nb = A.colptr.nbytes + A.rowval.nbytes + A.nzval.nbytes + es * (A.m + A.n)
    rows.append(dict(workload=name, kernel="csc_trspmv", group=0, us_med=med * 1e6, us_min=mn * 1e6, bytes=nb, gbs=nb / med / 1e9,
                     gflops=2.0 * A.nnz / med / 1e9, nnz=A.nnz))
    print(f"{name:34s} {'csc':8s}      {med * 1e6:8.1f} us  {nb / med / 1e9:7.0f} GB/s  {2.0 * A.nnz / med / 1e9:8.1f} GFLOP/s", flush=True)


def pack_time(name, ctor):
    torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); B = ctor(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
        if _ < 2:
            B.close()
    print(f"{name:34s} pack (H2D + kernels, wall)  {min(ts) * 1e3:8.1f} ms", flush=True)
    rows.append(dict(workload=name, kernel="pack_wall", us_med=min(ts) * 1e6))
    return B


def main():
    which = os.environ.get("PERF_ONLY", "").split(",") if os.environ.get("PERF_ONLY") else None

    def want(k):
        return which is None or k in which
    if want("c2"):
        A, pi, phi = synth.config_c2()
        B = pack_time("C2 2D f64 4x4 n=1M", lambda: vb.SparseMatrixVBC[4, 4](A, pi, phi))
        report("C2 2D f64 4x4 n=1M", B, A, groups=(4, 8, 16, 32))
        csc_row("C2 matrix as CSC f64/i64", A)
        B.close()
        A32 = A.astype(np.float32, np.int32)
        B = pack_time("C2 2D f32/i32 4x4 n=1M", lambda: vb.SparseMatrixVBC[4, 4](A32, pi, phi))
        report("C2 2D f32/i32 4x4 n=1M", B, A32, groups=(4, 8, 16, 32))
        csc_row("C2 matrix as CSC f32/i32", A32)
        B.close()
        del A, A32
    if want("c3"):
        # C3's matrix: 1D-VBC, n = 1M, W = 8, 50 rows per stripe, banded
        K, L = 1_000_000, 125_000
        offs = np.arange(-25, 25) * 37
        A, _, phi = synth.banded_blocks(K, L, 1, 8, offs)
        B = pack_time("C3 1D f64 w=8 n=1M 50 rows/stripe", lambda: vb.SparseMatrix1DVBC[8](A, phi))
        report("C3 1D f64 w=8 n=1M 50 rows/stripe", B, A, groups=(8, 16, 32))
        # SpMM, k = 32 (BASELINE configs[2]); row-major panels and Julia's column-major
        for order, k in (("rowmajor", 32), ("colmajor", 32), ("rowmajor", 8)):
            if order == "rowmajor":
                X = torch.rand(A.m, k, dtype=torch.float64, device="cuda"); Y = torch.empty(A.n, k, dtype=torch.float64, device="cuda")
            else:
                X = torch.rand(k, A.m, dtype=torch.float64, device="cuda").t(); Y = torch.empty(k, A.n, dtype=torch.float64, device="cuda").t()
            med, mn = tk(lambda: vb.mul_(Y, B.T, X), reps=10) if order == "rowmajor" else tk_eager(lambda: vb.mul_(Y, B.T, X))
            nb = B.format_bytes()[1] + 8 * k * (A.m + A.n)
            rows.append(dict(workload="C3 1D f64 w=8 n=1M", kernel=f"spmm_adj k={k} {order}", group=32, us_med=med * 1e6, us_min=mn * 1e6, bytes=nb,
                             gbs=nb / med / 1e9, gflops=2.0 * A.nnz * k / med / 1e9, nnz=A.nnz))
            print(f"{'C3 SpMM adjoint k=%d %s' % (k, order):34s}          {med * 1e6:8.1f} us  {nb / med / 1e9:7.0f} GB/s  {2.0 * A.nnz * k / med / 1e9:8.1f} GFLOP/s", flush=True)
            del X, Y
        B.close()
        del A
    if want("c3b"):
        # same sizes as C3 but a CONTIGUOUS 50-row band per stripe (adjacent stripes share 42 of their 50 rows)
        K, L = 1_000_000, 125_000
        A, _, phi = synth.banded_blocks(K, L, 1, 8, np.arange(-25, 25))
        B = pack_time("C3b 1D f64 w=8 n=1M contiguous band", lambda: vb.SparseMatrix1DVBC[8](A, phi))
        report("C3b 1D f64 w=8 n=1M contiguous band", B, A, groups=(8,))
        for k, simt in ((32, 0), (64, 0), (32, 1)):
            B.set_option(_lib.OPT_SPMM_SIMT, simt)
            X = torch.rand(A.m, k, dtype=torch.float64, device="cuda"); Y = torch.empty(A.n, k, dtype=torch.float64, device="cuda")
            med, mn = tk(lambda: vb.mul_(Y, B.T, X), reps=10)
            nb = B.format_bytes()[1] + 8 * k * (A.m + A.n)
            rows.append(dict(workload="C3b 1D f64 w=8 n=1M contiguous band", kernel=f"spmm_adj k={k} rowmajor {'simt' if simt else 'dmma'}", group=32, us_med=med * 1e6, us_min=mn * 1e6,
                             bytes=nb, gbs=nb / med / 1e9, gflops=2.0 * A.nnz * k / med / 1e9, nnz=A.nnz))
            print(f"{'C3b SpMM adjoint k=%d %s' % (k, 'simt' if simt else 'dmma'):34s}          {med * 1e6:8.1f} us  {nb / med / 1e9:7.0f} GB/s  {2.0 * A.nnz * k / med / 1e9:8.1f} GFLOP/s", flush=True)
            del X, Y
        B.close()
        del A
    if want("c5"):
        n5 = 4_000_000
        A, pi, phi = synth.banded_blocks(n5 // 4, n5 // 4, 4, 4, [0, 1, -1, 2, -2, 57, -57, 58, -58, 3249], dtype=np.float32, ti=np.int32)
        B = pack_time("C5 slab 2D f32/i32 4x4 10 blk/stripe n=4M", lambda: vb.SparseMatrixVBC[4, 4](A, pi, phi))
        report("C5 slab 2D f32/i32 4x4 10 blk/stripe n=4M", B, A, groups=(4, 8, 16, 32))
        B.close()
        del A
    if want("c1"):
        A, phi = synth.config_c1()
        B = pack_time("C1 1D f64 w=8 n=10k (L2-resident)", lambda: vb.SparseMatrix1DVBC[8](A, phi))
        report("C1 1D f64 w=8 n=10k (L2-resident)", B, A, groups=(8, 16, 32))
        csc_row("C1 matrix as CSC", A)
        B.close()
    if want("c2v"):
        A, pi, phi = synth.config_c2(n=400_000, S=41)
        pv, fv = synth.variable_partition(A.m, 8, 1), synth.variable_partition(A.n, 8, 2)
        B = pack_time("C2v 2D f64 variable 2..8 n=400k", lambda: vb.SparseMatrixVBC[8, 8](A, pv, fv))
        report("C2v 2D f64 variable 2..8 n=400k", B, A, groups=(4, 8, 16, 32))
        B.close()
    if want("c4"):
        A, pi, phi = synth.config_c4_triangular()
        B = pack_time("C4 triangular 2D f64 4x4 n=2M", lambda: vb.SparseMatrixVBC[4, 4](A, pi, phi))
        levels = vb.trsv_analyse(B.T)
        b = torch.rand(A.n, dtype=torch.float64, device="cuda"); xs = torch.empty_like(b)
        for _ in range(2):
            vb.ldiv_lower_(xs, B.T, b)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ts = []
        for _ in range(5):
            e0.record(); vb.ldiv_lower_(xs, B.T, b); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e-3)
        B.sync()
        med = sorted(ts)[2]
        # residual check
        r = torch.empty_like(b)
        nb = B.format_bytes()[1] + 8 * 3 * A.n
        rows.append(dict(workload="C4 triangular 2D f64 4x4 n=2M", kernel="trsv_lower", levels=levels, us_med=med * 1e6, bytes=nb, gbs=nb / med / 1e9,
                         gflops=2.0 * A.nnz / med / 1e9, nnz=A.nnz))
        print(f"{'C4 trsv lower n=2M levels=%d' % levels:34s}          {med * 1e6:8.1f} us  {nb / med / 1e9:7.0f} GB/s  {2.0 * A.nnz / med / 1e9:8.1f} GFLOP/s", flush=True)
        report("C4 matrix, adjoint multiply", B, A, groups=(8,))
        B.close()
    out = os.path.join(ROOT, "gpurun_out", "perf_table.json")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    json.dump(rows, open(out, "w"), indent=0)


if __name__ == "__main__":
    main()
