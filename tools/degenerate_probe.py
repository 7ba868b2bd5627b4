#!/usr/bin/env python
"""Degenerate partitions of the configs[1] matrix (n = 400k): 1D stripes of width 1 (the CSC-like case of constructors_1DVBC.jl:47-55),
2, 3, 16 and 32 -- adjoint and forward time, GB/s of the format actually read, parity vs scipy."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vbc_b200 as vb
from vbc_b200 import _lib, synth
from bench import timed_graph
A, _, _ = synth.config_c2(n=400_000, S=41)
S = A.to_scipy()
x = synth.vector(A.m, 7); xd, yd = torch.from_numpy(x).cuda(), torch.empty(A.n, dtype=torch.float64, device="cuda")
xn = synth.vector(A.n, 9); xnd, ymd = torch.from_numpy(xn).cuda(), torch.empty(A.m, dtype=torch.float64, device="cuda")
yref, fref = S.T @ x, S @ xn
out = {}
for w in (1, 2, 3, 16, 32):
    B = vb.SparseMatrix1DVBC[w](A, vb.pack_stripe(A, vb.EquiChunker(w)))
    med, _ = timed_graph(lambda: vb.mul_(yd, B.T, xd), 30)
    ea = float(np.max(np.abs(yd.cpu().numpy() - yref) / (abs(S).T @ np.abs(x))))
    vb.mul_(ymd, B, xnd)
    medf, _ = timed_graph(lambda: vb.mul_(ymd, B, xnd), 30)
    ef = float(np.max(np.abs(ymd.cpu().numpy() - fref) / (abs(S) @ np.abs(xn))))
    fb = B.format_bytes()
    out[f"w{w}"] = dict(adj_us=round(med * 1e6, 1), adj_GBps=round((fb[1] + 16 * A.n) / med / 1e9), adj_err=ea, fwd_us=round(medf * 1e6, 1), fwd_err=ef, nval=B.nval)
    B.close()
print(json.dumps(out))
