#!/usr/bin/env python
"""What bounds the host-vector (e2e) multiply on this box: pinned H2D / D2H bandwidth for 8 MB vectors, alone and both
directions at once, with the process on the GPU's NUMA-local CPUs and elsewhere; then the e2e multiply itself."""
import json, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vbc_b200 as vb
from vbc_b200 import synth, _lib

out = {}
try:
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    ncpu = os.cpu_count()
    mask = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
    local = [i for i in range(ncpu) if (mask[i // 64] >> (i % 64)) & 1]
    out["gpu_local_cpus"] = f"{local[0]}-{local[-1]} ({len(local)})" if local else None
    out["link_gen_width"] = [pynvml.nvmlDeviceGetCurrPcieLinkGeneration(h), pynvml.nvmlDeviceGetCurrPcieLinkWidth(h),
                             pynvml.nvmlDeviceGetMaxPcieLinkGeneration(h), pynvml.nvmlDeviceGetMaxPcieLinkWidth(h)]
except Exception as e:
    local = []
    out["nvml_error"] = repr(e)
out["affinity_now"] = len(os.sched_getaffinity(0))
out["nproc"] = os.cpu_count()

def bw(nbytes=8_000_000, reps=20):
    hx = torch.empty(nbytes, dtype=torch.uint8).pin_memory(); hy = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    hx.fill_(1)
    dx = torch.empty(nbytes, dtype=torch.uint8, device="cuda"); dy = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    res = {}
    for name in ("h2d", "d2h", "both"):
        ts = []
        for _ in range(reps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            if name in ("h2d", "both"):
                with torch.cuda.stream(s1): dx.copy_(hx, non_blocking=True)
            if name in ("d2h", "both"):
                with torch.cuda.stream(s2): hy.copy_(dy, non_blocking=True)
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        ts.sort()
        res[name + "_us"] = ts[len(ts) // 2] * 1e6
        res[name + "_GBps"] = nbytes / ts[len(ts) // 2] / 1e9
    return res

all_cpus = sorted(os.sched_getaffinity(0))
out["bw_default_affinity"] = bw()
if local:
    os.sched_setaffinity(0, set(local) & set(all_cpus) or set(all_cpus))
    out["bw_gpu_local_cpus"] = bw()
    far = [c for c in all_cpus if c not in local]
    if far:
        os.sched_setaffinity(0, far)
        out["bw_far_cpus"] = bw()
    os.sched_setaffinity(0, all_cpus)

A, pi, phi = synth.config_c2()
B = vb.SparseMatrixVBC[4, 4](A, pi, phi)
def e2e(reps=30):
    x = torch.rand(A.m, dtype=torch.float64).pin_memory().numpy(); y = torch.empty(A.n, dtype=torch.float64).pin_memory().numpy()
    for _ in range(3): vb.mul_(y, B.T, x)
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); vb.mul_(y, B.T, x); ts.append(time.perf_counter() - t0)
    ts.sort()
    return [ts[len(ts) // 2] * 1e6, ts[0] * 1e6]
out["e2e_us_med_min_default_affinity"] = e2e()
B.set_option(_lib.OPT_E2E_PIPELINE, 0)
out["e2e_us_med_min_no_pipeline"] = e2e()
B.set_option(_lib.OPT_E2E_PIPELINE, 1)
if local:
    os.sched_setaffinity(0, set(local) & set(all_cpus) or set(all_cpus))
    out["e2e_us_med_min_gpu_local_cpus"] = e2e()
    os.sched_setaffinity(0, all_cpus)
print(json.dumps(out, indent=1))
