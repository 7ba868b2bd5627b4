import json, os, sys
import numpy as np, torch
sys.path.insert(0, "/root/repo")
import vbc_b200 as vb
from vbc_b200 import _lib, synth
from bench import timed_graph
A, pi, phi = synth.variable_block_matrix(300000)
out = {}
for name, ctor in (("1D natural", lambda: vb.SparseMatrix1DVBC[8](A, phi)), ("2D natural", lambda: vb.SparseMatrixVBC[8, 8](A, pi, phi))):
    B = ctor()
    x = synth.vector(A.m, 7)
    xd, yd = torch.from_numpy(x).cuda(), torch.empty(A.n, dtype=torch.float64, device="cuda")
    S = A.to_scipy(); yref = S.T @ x
    for g in (0, 4, 8, 16, 32):
        B.set_option(_lib.OPT_ADJ_GROUP, g)
        med, mn = timed_graph(lambda: vb.mul_(yd, B.T, xd), 50)
        err = float(np.max(np.abs(yd.cpu().numpy() - yref) / (abs(S).T @ np.abs(x) + 1e-300)))
        out[f"{name} G{g}"] = (round(med * 1e6, 1), err)
    xn = synth.vector(A.n, 9); xnd, ymd = torch.from_numpy(xn).cuda(), torch.empty(A.m, dtype=torch.float64, device="cuda")
    vb.mul_(ymd, B, xnd)
    med, mn = timed_graph(lambda: vb.mul_(ymd, B, xnd), 50)
    out[f"{name} forward auto"] = (round(med * 1e6, 1), float(np.max(np.abs(ymd.cpu().numpy() - S @ xn) / (abs(S) @ np.abs(xn) + 1e-300))))
    out[f"{name} bytes"] = B.format_bytes()[1] + 8 * (A.m + A.n)
    out[f"{name} L"] = B.L
    B.close()
print(json.dumps(out))
