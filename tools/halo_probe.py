#!/usr/bin/env python
"""Single-GPU dissection of the fused multiply + exchange kernel (k_spmv_adj_halo) against the plain adjoint kernel on the
configs[1] matrix: (a) everything interior -- the cost of the kernel's structure alone; (b) a boundary of the size a middle
rank has (3970 stripes at each end), results stored to this rank only -- boundary handling without NVLink; for several
boundary weights.  K launches per CUDA graph, median of 5 replays."""
import ctypes, json, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vbc_b200 as vb
from vbc_b200 import _lib, synth

K = int(os.environ.get("PROBE_STEPS", "20"))


def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    side, g = torch.cuda.Stream(), torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        for _ in range(K):
            fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(side):
            e0.record(); g.replay(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / K * 1e3)
    return sorted(ts)[2]


A, pi, phi = synth.config_c2()
B = vb.SparseMatrixVBC[4, 4](A, pi, phi)
n = A.n
x = torch.rand(n, dtype=torch.float64, device="cuda")
y = torch.empty(n, dtype=torch.float64, device="cuda")
out = {"plain_us": timed(lambda: vb.mul_(y, B.T, x, 0.04, False))}
_flip = [0]
_bufs = [x.clone(), y.clone()]


def pingpong():  # the plain kernel in the access pattern of the iteration: the x of a step is the y of the previous one
    a, b = _bufs[_flip[0]], _bufs[1 - _flip[0]]
    vb.mul_(b, B.T, a, 0.04, False)
    _flip[0] ^= 1


out["plain_pingpong_us"] = timed(pingpong)
import ctypes as _ct
_rt = _ct.CDLL("libcudart.so")
_pa, _pb = _ct.c_void_p(), _ct.c_void_p()
_rt.cudaMalloc(_ct.byref(_pa), _ct.c_size_t(8 * n)); _rt.cudaMalloc(_ct.byref(_pb), _ct.c_size_t(8 * n))
_rt.cudaMemcpy(_pa, _ct.c_void_p(x.data_ptr()), _ct.c_size_t(8 * n), 3)


def raw_pingpong():  # same, on buffers from cudaMalloc (as the exchange buffers are) instead of torch's allocator
    B._use_torch_stream()
    a, b = (_pa, _pb) if _flip[0] == 0 else (_pb, _pa)
    _lib.check(_lib.lib().vbc_spmv(B._h, 1, 0.04, a, n, 0.0, b, n, 1))
    _flip[0] ^= 1


_flip[0] = 0
out["plain_pingpong_cudamalloc_us"] = timed(raw_pingpong)
Lh = _lib.lib()


def peer_case(i0, i1, masked=True):
    h = ctypes.c_void_p()
    _lib.check(Lh.vbc_peer_create(ctypes.byref(h), _lib.VBC_F64, n, 0, 1, 0, None))
    if masked:
        mask = np.ones((n + 127) // 128, dtype=np.uint8)
        _lib.check(Lh.vbc_peer_set_mask(h, mask.ctypes.data_as(ctypes.c_void_p), len(mask), 7))
    _lib.check(Lh.vbc_peer_set_interior(h, i0, i1))
    p = ctypes.c_void_p()
    _lib.check(Lh.vbc_peer_buffer(h, 0, ctypes.byref(p)))
    ctypes.CDLL("libcudart.so").cudaMemcpy(ctypes.c_void_p(p.value), ctypes.c_void_p(x.data_ptr()), ctypes.c_size_t(8 * n), 3)

    def step():
        B._use_torch_stream()
        _lib.check(Lh.vbc_peer_spmv_step(h, B._h, 0.04, 0, 3))
    t = timed(step)
    torch.cuda.synchronize()
    Lh.vbc_peer_destroy(h)
    return t


out["halo_all_interior_us"] = peer_case(0, B.L)
out["halo_boundary_3970_each_end_us"] = peer_case(3970, B.L - 3970)
out["halo_boundary_3970_one_end_us"] = peer_case(0, B.L - 3970)
out["halo_everything_boundary_self_only_us"] = peer_case(0, 0)
out["boundary_weight"] = os.environ.get("VBC_HALO_BOUNDARY_WEIGHT", "2 (default)")
print(json.dumps(out))
