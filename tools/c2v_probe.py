#!/usr/bin/env python
"""C2v (variable 2..8 blocks over the configs[1] matrix) adjoint multiply: time per group size; parity against scipy."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, ROOT)
import vbc_b200 as vb
from vbc_b200 import _lib, synth
from bench import timed_graph
A, _, _ = synth.config_c2()
DT = np.float32 if os.environ.get("PROBE_F32") else np.float64
if DT == np.float32:
    A = A.astype(np.float32, np.int32)
TDT = torch.float32 if DT == np.float32 else torch.float64
pv, fv = synth.variable_partition(A.m, 8, 3), synth.variable_partition(A.n, 8, 4)
B = vb.SparseMatrixVBC[8, 8](A, pv, fv)
x = synth.vector(A.m, 7, dtype=DT)
xd, yd = torch.from_numpy(x).cuda(), torch.empty(A.n, dtype=TDT, device="cuda")
S = A.to_scipy().astype(np.float64)
yref = S.T @ x.astype(np.float64)
out = {"lib": os.environ.get("VBC_LIBRARY", "product"), "dtype": str(np.dtype(DT)), "no_flat": bool(os.environ.get("VBC_NO_FLAT")), "adj_bytes": B.format_bytes()[1] + A.nzval.dtype.itemsize * (A.m + A.n)}
for g in (0, 8, 16, 32):
    B.set_option(_lib.OPT_ADJ_GROUP, g)
    med, mn = timed_graph(lambda: vb.mul_(yd, B.T, xd), 30)
    err = float(np.max(np.abs(yd.cpu().numpy() - yref) / (abs(S).T @ np.abs(x))))
    out[f"G{g}_us"] = round(med * 1e6, 1)
    out[f"G{g}_err"] = err
print(json.dumps(out))
# forward multiply y = B x: auto mode (variable blocks: the rows-mode transposed copy), the atomic scatter kernel, the unit index
xn = synth.vector(A.n, 9, dtype=DT)
xnd, ymd = torch.from_numpy(xn).cuda(), torch.empty(A.m, dtype=TDT, device="cuda")
fref = S @ xn.astype(np.float64)
fb = abs(S) @ np.abs(xn)
fwd = {}
for mode, name in ((0, "auto"), (1, "atomic"), (2, "unit_index")):
    B.set_option(_lib.OPT_FWD_MODE, mode)
    B.set_option(_lib.OPT_FWD_GROUP, 0)
    vb.mul_(ymd, B, xnd)
    med, mn = timed_graph(lambda: vb.mul_(ymd, B, xnd), 30)
    fwd[f"fwd_{name}_us"] = round(med * 1e6, 1)
    fwd[f"fwd_{name}_err"] = float(np.max(np.abs(ymd.cpu().numpy() - fref) / fb))
fwd["fwd_bytes"] = B.format_bytes()[2] + A.nzval.dtype.itemsize * (A.n + A.m)
print(json.dumps(fwd))
