#!/usr/bin/env python
"""e2e (host vectors) adjoint multiply time on configs[1] for different numbers of overlap chunks (VBC_E2E_CHUNKS)."""
import os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1:
    sys.path.insert(0, ROOT)
    import numpy as np, torch
    import vbc_b200 as vb
    from vbc_b200 import synth
    A, pi, phi = synth.config_c2()
    B = vb.SparseMatrixVBC[4, 4](A, pi, phi)
    xh = torch.from_numpy(synth.vector(A.m, 1)).pin_memory(); yh = torch.empty(A.n, dtype=torch.float64).pin_memory()
    x, y = xh.numpy(), yh.numpy()
    for _ in range(5):
        vb.mul_(y, B.T, x)
    best = 1e9
    for rep in range(5):
        t0 = time.perf_counter()
        for _ in range(50):
            vb.mul_(y, B.T, x)
        best = min(best, (time.perf_counter() - t0) / 50)
    # copy-only floor: H2D + D2H of the same buffers
    xd = torch.empty(A.m, dtype=torch.float64, device="cuda"); yd = torch.empty(A.n, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(50):
        xd.copy_(xh, non_blocking=True); yh.copy_(yd, non_blocking=True); torch.cuda.synchronize()
    floor = (time.perf_counter() - t0) / 50
    print(f"chunks={os.environ.get('VBC_E2E_CHUNKS')}  e2e {best * 1e6:7.1f} us/step  {2.0 * A.nnz / best / 1e9:6.1f} GFLOP/s   (H2D+D2H alone {floor * 1e6:6.1f} us)", flush=True)
else:
    for nc in ("1", "2", "3", "4", "8"):
        subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=dict(os.environ, VBC_E2E_CHUNKS=nc))
