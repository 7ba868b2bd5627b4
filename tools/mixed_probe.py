"""Times the mixed-type adjoint (Float32 matrix, Float64 vectors) and the Int32 adjoint on the configs[1] matrix (CUDA graph, 50 launches)."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
import vbc_b200 as vb  # noqa: E402
from vbc_b200 import synth  # noqa: E402
from vbc_b200.partition import SplitPartition  # noqa: E402

peak = bench.measured_peak()
peak = peak[0] if isinstance(peak, tuple) else peak
A, _, _ = synth.config_c2()
pi = SplitPartition(np.arange(1, A.m + 2, 4, dtype=np.int32)); phi = SplitPartition(np.arange(1, A.n + 2, 4, dtype=np.int32))
A32 = A.astype(np.float32, np.int32)
B32 = vb.SparseMatrixVBC[4, 4](A32, pi, phi)
S32 = A32.to_scipy().astype(np.float64)
xm = synth.vector(A.m, 15)
xmd, ymd = torch.from_numpy(xm).cuda(), torch.empty(A.n, dtype=torch.float64, device="cuda")
vb.mul_(ymd, B32.T, xmd)
med, mn = bench.timed_graph(lambda: vb.mul_(ymd, B32.T, xmd), 50)
nb = B32.format_bytes()[1] + 8 * (A.m + A.n)
ref = S32.T @ xm
err = float(np.max(np.abs(ymd.cpu().numpy() - ref) / np.maximum(1e-12 * (abs(S32).T @ np.abs(xm)), 1e-300)))
print(json.dumps({"mixed_adjoint_us": med * 1e6, "us_min": mn * 1e6, "GBps": nb / med / 1e9, "frac": nb / med / 1e9 / peak, "err_over_bound_1e-12": err}))
