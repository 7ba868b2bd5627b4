#!/usr/bin/env python
"""Short program for ncu: a few launches of the plain adjoint kernel (configs[1]) and of the fused multiply + exchange
kernel with a middle rank's boundary (single rank: nothing waits), no CUDA graph, so every launch is its own kernel node."""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vbc_b200 as vb
from vbc_b200 import _lib, synth
A, pi, phi = synth.config_c2()
B = vb.SparseMatrixVBC[4, 4](A, pi, phi)
n = A.n
x = torch.rand(n, dtype=torch.float64, device="cuda"); y = torch.empty(n, dtype=torch.float64, device="cuda")
for _ in range(6):
    vb.mul_(y, B.T, x)
torch.cuda.synchronize()
Lh = _lib.lib()
h = ctypes.c_void_p()
_lib.check(Lh.vbc_peer_create(ctypes.byref(h), _lib.VBC_F64, n, 0, 1, 0, None))
mask = np.ones((n + 127) // 128, dtype=np.uint8)
_lib.check(Lh.vbc_peer_set_mask(h, mask.ctypes.data_as(ctypes.c_void_p), len(mask), 7))
_lib.check(Lh.vbc_peer_set_interior(h, 3970, B.L - 3970))
p = ctypes.c_void_p()
_lib.check(Lh.vbc_peer_buffer(h, 0, ctypes.byref(p)))
ctypes.CDLL("libcudart.so").cudaMemcpy(ctypes.c_void_p(p.value), ctypes.c_void_p(x.data_ptr()), ctypes.c_size_t(8 * n), 3)
B._use_torch_stream()
for _ in range(6):
    _lib.check(Lh.vbc_peer_spmv_step(h, B._h, 0.04, 0, 3))
torch.cuda.synchronize()
Lh.vbc_peer_destroy(h)
print("ok")
