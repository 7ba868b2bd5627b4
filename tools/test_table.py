#!/usr/bin/env python
"""Benchmark table ≙ /root/reference/bin/test_table.jl, with GPU rows.

For each input matrix (Matrix Market files given on the command line, else built-in synthetic ones), after the
reference's `A = permutedims(sparse(mdopen(mtx).A))` (test_table.jl:27):
  "reference"  CSC `TrSpMV!(y, A, x)`                                        (test_table.jl:33-43)
  1D methods   strict, min blocks, min memory, min time                      (test_table.jl:64-86)
  2D methods   1D 2D, strict 2D, dynamic blocks 2D, dynamic memory 2D        (test_table.jl:88-127)
columns: setup time (partition + pack, s), memory (bytes, the reference's format accounting), run time of
`mul_(y, B.T, x, True, False)` (s, CUDA-graph minimum) and the time model's prediction (s); `y ≈ z` is asserted
against scipy like test_table.jl:42/:84/:126.  The partitioners are stand-ins restated from their definitions
(ChainPartitioners is not vendored; OverlapChunker's definition is assumed), "dynamic time 2D" is left out."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vbc_b200 as vb  # noqa: E402
from vbc_b200 import costs, synth  # noqa: E402
from vbc_b200.costs import _time_adjoint  # noqa: E402

W_MAX = 8


def run_matrix(name, A, out):
    import torch
    rows = []
    S = A.to_scipy()
    x = np.random.default_rng(0).random(A.m)
    z = S.T @ x
    # reference row: CSC TrSpMV
    C = vb.CuSparseMatrixCSC(A)
    y = vb.TrSpMV_(np.empty(A.n), C, x)
    assert np.allclose(y, z)
    xd, yd = torch.from_numpy(x).cuda(), torch.empty(A.n, dtype=torch.float64, device="cuda")
    for _ in range(3):
        vb.TrSpMV_(yd, C, xd)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(10):
        e0.record(); vb.TrSpMV_(yd, C, xd); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e-3)
    rows.append(["reference (CSC TrSpMV!)", 0.0, A.colptr.nbytes + A.rowval.nbytes + A.nzval.nbytes, min(ts), 0.0])

    mdl_blocks_1d = costs.model_SparseMatrix1DVBC_blocks()
    mdl_memory_1d = costs.model_SparseMatrix1DVBC_memory(np.float64, np.int64)
    tkw = {"cache_bytes": int(os.environ["TABLE_CACHE_BYTES"]), "use_cache": False} if os.environ.get("TABLE_CACHE_BYTES") else {}
    mdl_time_1d = costs.model_SparseMatrix1DVBC_TrSpMV_time(W_MAX, np.float64, np.int64, np.float64, exceed=True, **tkw)
    mdl_blocks_2d = costs.model_SparseMatrixVBC_blocks()
    mdl_memory_2d = costs.model_SparseMatrixVBC_memory(np.float64, np.int64)
    DP = vb.DynamicTotalChunker

    for key, method in (("strict", vb.StrictChunker(W_MAX)), ("overlap", vb.OverlapChunker(0.9, W_MAX)), ("min blocks", DP(mdl_blocks_1d, W_MAX)),
                        ("min memory", DP(mdl_memory_1d, W_MAX)), ("min time", DP(mdl_time_1d, W_MAX))):
        t0 = time.perf_counter()
        B = vb.SparseMatrix1DVBC[W_MAX](A, method)
        B.sync()
        setup = time.perf_counter() - t0
        y = vb.mul_(np.empty(A.n), B.T, x, True, False)
        assert np.allclose(y, z), key
        rows.append([key, setup, B.format_bytes()[0], _time_adjoint(B, A.m, A.n, np.float64), costs.total_value(B, mdl_time_1d)])
        B.close()

    pd = vb.permutedims
    for key, method in (("1D 2D", vb.AlternatingPacker(DP(mdl_blocks_1d, W_MAX), vb.EquiChunker(1))),
                        ("strict 2D", vb.AlternatingPacker(vb.StrictChunker(W_MAX), vb.StrictChunker(W_MAX))),
                        ("overlap 2D 0.9", vb.AlternatingPacker(vb.OverlapChunker(0.9, W_MAX), vb.OverlapChunker(0.9, W_MAX))),
                        ("overlap 2D 0.8", vb.AlternatingPacker(vb.OverlapChunker(0.8, W_MAX), vb.OverlapChunker(0.8, W_MAX))),
                        ("overlap 2D 0.7", vb.AlternatingPacker(vb.OverlapChunker(0.7, W_MAX), vb.OverlapChunker(0.7, W_MAX))),
                        ("dynamic blocks 2D", vb.AlternatingPacker(DP(mdl_blocks_1d, W_MAX), DP(pd(mdl_blocks_2d), W_MAX), DP(mdl_blocks_2d, W_MAX))),
                        ("dynamic memory 2D", vb.AlternatingPacker(vb.EquiChunker(1), vb.EquiChunker(1), DP(mdl_memory_2d, W_MAX),
                                                                   DP(pd(mdl_memory_2d), W_MAX), DP(mdl_memory_2d, W_MAX)))):
        t0 = time.perf_counter()
        B = vb.SparseMatrixVBC[W_MAX, W_MAX](A, method)
        B.sync()
        setup = time.perf_counter() - t0
        y = vb.mul_(np.empty(A.n), B.T, x, True, False)
        assert np.allclose(y, z), key
        rows.append([key, setup, B.format_bytes()[0], _time_adjoint(B, A.m, A.n, np.float64), 0.0])
        B.close()

    print(f"\n{name}: {A.m} x {A.n}, nnz = {A.nnz}")
    print(f"{'method':26s} {'setuptime s':>12s} {'memory B':>12s} {'runtime us':>11s} {'model us':>10s}")
    for r in rows:
        print(f"{r[0]:26s} {r[1]:12.4f} {r[2]:12d} {r[3] * 1e6:11.1f} {r[4] * 1e6:10.1f}")
    out[name] = [dict(method=r[0], setuptime=r[1], memory=int(r[2]), runtime=r[3], model=r[4]) for r in rows]


def main():
    out = {}
    if len(sys.argv) > 1:
        import scipy.io
        for path in sys.argv[1:]:
            M = scipy.io.mmread(path).tocsc()
            run_matrix(os.path.basename(path), vb.SparseMatrixCSC.from_scipy(M.T.tocsc()), out)  # permutedims, test_table.jl:27
    else:
        A, _, _ = synth.variable_block_matrix(int(os.environ.get("TABLE_N", "300000")))
        run_matrix("synthetic supernodal (natural blocks 1..6)", A, out)
        A, _, _ = synth.config_c2(n=int(os.environ.get("TABLE_N2", "400000")), S=41)
        run_matrix("synthetic FEM band (4x4 blocks)", A, out)
    p = os.environ.get("TABLE_OUT") or os.path.join(ROOT, "gpurun_out", "test_table.json")
    os.makedirs(os.path.dirname(p), exist_ok=True)
    json.dump(out, open(p, "w"), indent=1)


if __name__ == "__main__":
    main()
