#!/usr/bin/env python
"""First GPU call of round 2: run every opt-in path that round 1 wrote after its GPU budget was spent, against the
default path on the same inputs, and time both.  One B200, ~1 minute.  Writes gpurun_out/round2_checks.json.

  1. adjoint SpMM kernels VBC_OPT_SPMM_SIMT = 3 (256-bit X-row loads), 4 (bulk-copy fed), 5 (cp.async fed, NEVER RUN YET)
     against 2 (default) on small 1D / 2D matrices with odd widths, odd slab starts, empty stripes, k in {2, 8, 32, 34, 64},
     then on configs[2] with timing;
  2. VBC_OPT_E2E_PIPELINE = 1 (x uploaded in pieces, NEVER RUN YET) against 0 on configs[1] (banded: should pipeline),
     on a matrix whose first stripes reach the last rows (must degrade to the plain upload order, same result), with
     host-to-host timing from pinned vectors.

Nothing here is a bench value; what passes gets a pytest case and, if faster, becomes the default."""
import json
import os
import sys
import time
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import vbc_b200 as vb  # noqa: E402
from vbc_b200 import _lib, synth  # noqa: E402
from perf_table import tk  # noqa: E402

out = {}


def section(name):
    def deco(fn):
        t0 = time.time()
        try:
            out[name] = fn()
        except Exception as e:  # keep going: every section is independent
            out[name] = {"error": repr(e), "trace": traceback.format_exc()[-1500:]}
        out[name]["seconds"] = round(time.time() - t0, 1)
        print(name, json.dumps(out[name])[:600], flush=True)
        return fn
    return deco


def sprand(m, n, density, rng):
    import scipy.sparse as sp
    S = sp.random(m, n, density=density, random_state=np.random.RandomState(int(rng.integers(1 << 30))), format="csc")
    S.data = rng.random(S.nnz) + 0.5
    return vb.SparseMatrixCSC.from_scipy(S)


@section("spmm_variants_small")
def _():
    rng = np.random.default_rng(7)
    res = {"cases": 0, "max_abs_diff": {3: 0.0, 4: 0.0, 5: 0.0}}
    for (m, n, u, w) in ((300, 257, 1, 8), (300, 255, 1, 5), (200, 96, 4, 4), (150, 90, 3, 6), (64, 40, 8, 2), (500, 16, 1, 8)):
        A = sprand(m, n, 0.15, rng)
        phi = vb.pack_stripe(A, vb.EquiChunker(w))
        pi = vb.pack_stripe(A.transpose(), vb.EquiChunker(u))
        mats = [vb.SparseMatrix1DVBC[w](A, phi)]
        if u > 1:
            mats.append(vb.SparseMatrixVBC[u, w](A, pi, phi))
        for B in mats:
            for k in (2, 8, 32, 34, 64):
                X = torch.rand(m, k, dtype=torch.float64, device="cuda")
                Y0 = torch.rand(n, k, dtype=torch.float64, device="cuda")
                ref = None
                for mode in (2, 3, 4, 5):
                    B.set_option(_lib.OPT_SPMM_SIMT, mode)
                    Y = Y0.clone()
                    vb.mul_(Y, B.T, X, 1.5, -0.25)
                    torch.cuda.synchronize()
                    if mode == 2:
                        ref = Y
                    else:
                        d = float((Y - ref).abs().max())
                        res["max_abs_diff"][mode] = max(res["max_abs_diff"][mode], d)
                res["cases"] += 1
    res["ok"] = all(v < 1e-12 for v in res["max_abs_diff"].values())
    return res


@section("spmm_variants_configs2")
def _():
    K, L, k = 1_000_000, 125_000, 32
    A, _, phi = synth.banded_blocks(K, L, 1, 8, np.arange(-25, 25) * 37)
    B = vb.SparseMatrix1DVBC[8](A, phi)
    X = torch.rand(A.m, k, dtype=torch.float64, device="cuda")
    res, ref = {}, None
    for mode in (2, 3, 4, 5):
        B.set_option(_lib.OPT_SPMM_SIMT, mode)
        Y = torch.full((A.n, k), float("nan"), dtype=torch.float64, device="cuda")
        vb.mul_(Y, B.T, X)
        torch.cuda.synchronize()
        if mode == 2:
            ref = Y
        d = float((Y - ref).abs().max())
        if not d < 1e-12:
            res[str(mode)] = {"max_abs_diff": d, "skipped_timing": True}
            continue
        med, mn = tk(lambda: vb.mul_(Y, B.T, X), reps=10)
        res[str(mode)] = {"max_abs_diff": d, "us_med": med * 1e6, "tflops": 2.0 * A.nnz * k / med / 1e12}
    return res


def e2e_time(B, x, y, reps=20):
    for _ in range(3):
        vb.mul_(y, B.T, x)
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        vb.mul_(y, B.T, x)
        ts.append(time.perf_counter() - t0)
    ts.sort()
    return ts[len(ts) // 2] * 1e6, ts[0] * 1e6


@section("e2e_pipeline_configs1")
def _():
    A, pi, phi = synth.config_c2()
    B = vb.SparseMatrixVBC[4, 4](A, pi, phi)
    x = torch.rand(A.m, dtype=torch.float64).pin_memory().numpy()
    y0 = np.empty(A.n)
    y1 = np.empty(A.n)
    yp0 = torch.from_numpy(y0).pin_memory().numpy()
    yp1 = torch.from_numpy(y1).pin_memory().numpy()
    res = {}
    B.set_option(_lib.OPT_E2E_PIPELINE, 0)
    res["plain_us_med_min"] = e2e_time(B, x, yp0)
    B.set_option(_lib.OPT_E2E_PIPELINE, 1)
    res["pipelined_us_med_min"] = e2e_time(B, x, yp1)
    res["uploaded_x_elements"] = B.get_option(_lib.OPT_E2E_UPLOAD_ELEMS)
    res["max_abs_diff"] = float(np.abs(yp0 - yp1).max())
    # alpha / beta through the pipelined path
    yb = yp1.copy()
    vb.mul_(yb, B.T, x, 2.0, -0.5)
    res["alpha_beta_max_abs_diff"] = float(np.abs(yb - (2.0 * yp0 - 0.5 * yp1)).max())
    res["ok"] = res["max_abs_diff"] == 0.0 and res["alpha_beta_max_abs_diff"] < 1e-9
    return res


@section("e2e_pipeline_far_reaching_rows")
def _():
    # stripes of the first chunk gather from the LAST rows: the first piece must be the whole x
    rng = np.random.default_rng(3)
    import scipy.sparse as sp
    n = 400_000
    S = sp.diags([rng.random(n), rng.random(n - 3)], [0, -3], format="lil")
    S[n - 1, 0] = 2.0
    S[n - 2, 5] = 3.0
    A = vb.SparseMatrixCSC.from_scipy(S.tocsc())
    B = vb.SparseMatrix1DVBC[4](A, vb.EquiChunker(4))
    x = rng.random(n)
    B.set_option(_lib.OPT_E2E_PIPELINE, 0)
    ya = vb.mul_(np.empty(n), B.T, x)
    B.set_option(_lib.OPT_E2E_PIPELINE, 1)
    yb = vb.mul_(np.empty(n), B.T, x)
    return {"max_abs_diff": float(np.abs(ya - yb).max()), "vs_scipy": float(np.abs(ya - S.T @ x).max()), "ok": bool(np.array_equal(ya, yb))}


@section("e2e_pipeline_row_window")
def _():
    # a slab like a rank of the row-partitioned multiply: the stripes gather only from rows [100k, 300k) of 400k
    rng = np.random.default_rng(4)
    import scipy.sparse as sp
    m, n = 400_000, 200_000
    S = sp.diags([rng.random(n), rng.random(n), rng.random(n)], [-100_000, -100_007, -99_990], shape=(m, n), format="csc")
    A = vb.SparseMatrixCSC.from_scipy(S)
    pi, phi = vb.pack_plaid(A, vb.AlternatingPacker(vb.EquiChunker(4), vb.EquiChunker(4)))
    B = vb.SparseMatrixVBC[4, 4](A, pi, phi)
    x = rng.random(m)
    B.set_option(_lib.OPT_E2E_PIPELINE, 0)
    ya = vb.mul_(np.full(n, np.nan), B.T, x)
    B.set_option(_lib.OPT_E2E_PIPELINE, 1)
    yb = vb.mul_(np.full(n, np.nan), B.T, x)
    up = B.get_option(_lib.OPT_E2E_UPLOAD_ELEMS)
    return {"max_abs_diff": float(np.abs(ya - yb).max()), "vs_scipy": float(np.abs(ya - S.T @ x).max()), "uploaded_x_elements": up,
            "ok": bool(np.array_equal(ya, yb)) and up < 0.6 * m}


if __name__ == "__main__":
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "round2_checks.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("ALL OK" if all(v.get("ok", True) and "error" not in v for v in out.values()) else "SOME CHECKS FAILED")
