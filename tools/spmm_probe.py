#!/usr/bin/env python
"""configs[2] adjoint SpMM (1D-VBC, Float64, k = 32, n = 1M): the DMMA kernel against the SIMT kernel
(PROBE_MODES = comma list of VBC_OPT_SPMM_SIMT values, first one is the reference: 0 DMMA tiles, 1 SIMT; the round-1 variants
with other X feeds were validated in round 2 -- profiles/r02_round2_checks.json -- and removed) -- results compared first,
then CUDA-graph timing.  PROBE_BAND=1 swaps
the strided rows for a contiguous band; PROBE_NCU=1 only launches each kernel twice (for an `ncu -k regex:k_spmm_adj`
capture).  Writes gpurun_out/spmm_probe*.json.  Nothing here is a bench value."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import vbc_b200 as vb  # noqa: E402
from vbc_b200 import _lib, synth  # noqa: E402
from perf_table import tk  # noqa: E402


def main():
    ncu = bool(os.environ.get("PROBE_NCU"))
    K, L, k = 1_000_000, 125_000, 32
    # PROBE_BAND=1: contiguous 50-row band (adjacent stripes share 42 of their 50 x rows) instead of the strided rows of configs[2]
    band = bool(os.environ.get("PROBE_BAND"))
    A, _, phi = synth.banded_blocks(K, L, 1, 8, np.arange(-25, 25) * (1 if band else 37))
    B = vb.SparseMatrix1DVBC[8](A, phi)
    X = torch.rand(A.m, k, dtype=torch.float64, device="cuda")
    out, ys = {}, {}
    modes = {"0": (0, "auto"), "1": (1, "simt"), "2": (2, "dmma"), "3": (3, "tma")}
    picked = [modes[m] for m in os.environ.get("PROBE_MODES", "2,3,1").split(",")]
    tag = "_".join(n for _, n in picked[1:])
    for mode, name in picked:
        B.set_option(_lib.OPT_SPMM_SIMT, mode)
        Y = torch.full((A.n, k), float("nan"), dtype=torch.float64, device="cuda")
        vb.mul_(Y, B.T, X)
        vb.mul_(Y, B.T, X)
        torch.cuda.synchronize()
        ys[name] = Y
        if len(ys) > 1:  # results first: a wrong kernel is not worth timing
            ref = ys[picked[0][1]]
            d = float(((ref - Y).abs() / (ref.abs() + 1e-300)).max())
            print(name, "max rel diff vs", picked[0][1], d, flush=True)
            assert d < 1e-11, d
        if not ncu:
            med, mn = tk(lambda: vb.mul_(Y, B.T, X), reps=10)
            nb = B.format_bytes()[1] + 8 * k * (A.m + A.n)
            out[name] = dict(us_med=med * 1e6, us_min=mn * 1e6, tflops=2.0 * A.nnz * k / med / 1e12, gbs=nb / med / 1e9)
            print(name, out[name], flush=True)
    a, b = ys[picked[0][1]], ys[picked[-1][1]]
    err = float(((a - b).abs() / (a.abs() + 1e-300)).max())
    out["max_rel_diff_between_variants"] = err
    # alpha / beta on the vector-load variant
    # fewer right-hand sides than one 32-column pass, last variant against the first
    X8 = X[:, :8].contiguous()
    y8 = []
    for mode, _ in (picked[0], picked[-1]):
        B.set_option(_lib.OPT_SPMM_SIMT, mode)
        Y8 = torch.full((A.n, 8), float("nan"), dtype=torch.float64, device="cuda")
        vb.mul_(Y8, B.T, X8)
        y8.append(Y8)
    torch.cuda.synchronize()
    out["k8_max_abs_diff"] = float((y8[0] - y8[1]).abs().max())
    assert out["k8_max_abs_diff"] == 0.0 or out["k8_max_abs_diff"] < 1e-9, out
    Y0 = torch.rand(A.n, k, dtype=torch.float64, device="cuda")
    Y1 = Y0.clone()
    vb.mul_(Y1, B.T, X, 2.0, -0.5)
    err2 = float((Y1 - (2.0 * a - 0.5 * Y0)).abs().max() / a.abs().max())
    out["alpha_beta_rel_err"] = err2
    print(json.dumps(out))
    assert err < 1e-11 and err2 < 1e-11, (err, err2)
    if not ncu:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", ("spmm_probe_band" if band else "spmm_probe") + ("_" + tag if tag not in ("dmma_256bit_row_loads",) else "") + ".json"), "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
