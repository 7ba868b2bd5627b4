#!/usr/bin/env python
"""Single-GPU timing of the adjoint kernel variants used by the multi-GPU step (nranks = 1, so nothing
waits): plain, peer epilogue + flag kernel, peer epilogue with in-kernel sync (with/without an interior
run).  200 steps in one CUDA graph each."""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vbc_b200 as vb
from vbc_b200 import _lib, synth

A, pi, phi = synth.config_c2()
B = vb.SparseMatrixVBC[4, 4](A, pi, phi)
n = A.n
Lh = _lib.lib()
x = torch.rand(n, dtype=torch.float64, device="cuda"); y = torch.empty(n, dtype=torch.float64, device="cuda")


def graph_time(fn, steps=200):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    side = torch.cuda.Stream(); g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        for _ in range(steps):
            fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        with torch.cuda.stream(side):
            e0.record(); g.replay(); e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / steps)
    return best * 1e3


print(f"plain mul_                     {graph_time(lambda: vb.mul_(y, B.T, x)):7.1f} us/step")
for name, mask, fused, i0, i1 in (("peer full, flag kernel", False, 0, 0, 0), ("peer masked, flag kernel", True, 0, 0, 0),
                                   ("peer full, in-kernel sync", False, 1, 0, 0), ("peer masked, in-kernel, interior", True, 1, 4000, 246000),
                                   ("peer masked, in-kernel, no interior", True, 1, 0, 0), ("peer masked, split launches", True, 2, 4000, 246000)):
    h = ctypes.c_void_p()
    _lib.check(Lh.vbc_peer_create(ctypes.byref(h), _lib.VBC_F64, n, 0, 1, 0, None))
    if mask:
        m = np.ones((n + 127) // 128, dtype=np.uint8)
        _lib.check(Lh.vbc_peer_set_mask(h, m.ctypes.data_as(ctypes.c_void_p), len(m), 7))
    _lib.check(Lh.vbc_peer_set_fused_sync(h, fused, i0, i1))

    def step():
        B._use_torch_stream()
        _lib.check(Lh.vbc_peer_spmv_step(h, B._h, 0.04, 0, 3))
    print(f"{name:36s} {graph_time(step):7.1f} us/step")
    Lh.vbc_peer_destroy(h)
