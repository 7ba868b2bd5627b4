#!/usr/bin/env python
"""Kernel-variant sweep on the headline config (run on a B200 via gpurun).

Loads every build/variants/libvbc_<tag>.so side by side, packs the same configs[1] matrix with
each, and times the adjoint kernel (one CUDA-event pair per launch, device pointers) for every
(group size, CTAs/SM) setting.  Prints one table; nothing here is a bench value."""
import ctypes
import glob
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vbc_b200 as vb  # noqa: E402
from vbc_b200 import synth  # noqa: E402


def vp(a):
    return ctypes.c_void_p(a.ctypes.data)


def time_kernel(fn, reps=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    pairs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in pairs:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    t = sorted(a.elapsed_time(b) for a, b in pairs)
    return t[0], t[len(t) // 2], sum(t) / len(t)


def main():
    n = int(os.environ.get("TUNE_N", "1000000"))
    dtype = np.float32 if os.environ.get("TUNE_F32") else np.float64
    A, pi, phi = synth.config_c2(n=n, dtype=dtype)
    tdt = torch.float32 if dtype == np.float32 else torch.float64
    x = torch.rand(A.m, dtype=tdt, device="cuda")
    y = torch.empty(A.n, dtype=tdt, device="cuda")
    libs = sorted(glob.glob(os.path.join(ROOT, "build", "variants", "libvbc_*.so")))
    only = os.environ.get("TUNE_ONLY")
    rows = []
    for path in libs:
        tag = os.path.basename(path)[len("libvbc_"):-3]
        if only and tag not in only.split(","):
            continue
        L = ctypes.CDLL(path)
        h = ctypes.c_void_p()
        i64, ci = ctypes.c_int64, ctypes.c_int
        L.vbc_pack_csc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ci, ci, i64, i64, ci, ci] + [ctypes.c_void_p] * 4 + [i64, ctypes.c_void_p, i64, ci]
        L.vbc_spmv.argtypes = [ctypes.c_void_p, ci, ctypes.c_double, ctypes.c_void_p, i64, ctypes.c_double, ctypes.c_void_p, i64, ci]
        L.vbc_set_option.argtypes = [ctypes.c_void_p, ci, i64]
        L.vbc_set_stream.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.vbc_format_bytes.argtypes = [ctypes.c_void_p, ctypes.POINTER(i64)]
        L.vbc_destroy.argtypes = [ctypes.c_void_p]
        L.vbc_last_error.restype = ctypes.c_char_p
        rc = L.vbc_pack_csc(ctypes.byref(h), 0 if dtype == np.float32 else 1, 1, A.m, A.n, 4, 4, vp(A.colptr), vp(A.rowval), vp(A.nzval),
                            vp(pi.spl), len(pi), vp(phi.spl), len(phi), 0)
        assert rc == 0, L.vbc_last_error()
        L.vbc_set_stream(h, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        fb = (i64 * 3)()
        L.vbc_format_bytes(h, fb)
        nbytes = fb[1] + x.element_size() * (A.m + A.n)
        for G in (8, 16, 32):
            for mult in (0,):
                L.vbc_set_option(h, 1, G)
                L.vbc_set_option(h, 3, mult)

                def fn():
                    rc = L.vbc_spmv(h, 1, 1.0, ctypes.c_void_p(x.data_ptr()), A.m, 0.0, ctypes.c_void_p(y.data_ptr()), A.n, 1)
                    assert rc == 0, L.vbc_last_error()
                tmin, tmed, tavg = time_kernel(fn)
                rows.append(dict(tag=tag, G=G, mult=mult, us_min=1e3 * tmin, us_med=1e3 * tmed, us_avg=1e3 * tavg, gbs_med=nbytes / tmed / 1e6))
                print(f"{tag:14s} G={G:2d} mult={mult} min {1e3 * tmin:7.1f} us  med {1e3 * tmed:7.1f} us  avg {1e3 * tavg:7.1f}  {nbytes / tmed / 1e6:7.0f} GB/s", flush=True)
        L.vbc_destroy(h)
    out = os.path.join(ROOT, "gpurun_out", os.environ.get("TUNE_OUT", "tune.json"))
    os.makedirs(os.path.dirname(out), exist_ok=True)
    json.dump(rows, open(out, "w"), indent=0)


if __name__ == "__main__":
    main()
