#!/usr/bin/env python
"""bench.py -- VBC SpMV throughput on B200 (BASELINE.json metric: GFLOP/s + achieved HBM GB/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one adjoint multiply y <- A' x (`mul!(y, B', x)`, the reference's benchmarked kernel,
bin/test_table.jl:122, costs.jl:224) over the synthetic matrix of BASELINE.json configs[1]:
2D-VBC, Float64, m = n = 1 000 000, U = W = 4, 13-block FEM-like band, nnz ≈ 52 M.  At N > 1
(torchrun, one rank per GPU) every rank owns a 1M-column slab of an (N·1M)² matrix (weak scaling)
and each step is ONE kernel launch per rank that also moves the new x to the ranks that read it
(NVLink peer stores + per-step flags, csrc/peer.cu).

value  = useful GFLOP/s (2·nnz per multiply), whole job, inputs resident in HBM, CUDA events.
e2e    = same metric through the public API with HOST vectors (H2D x + kernel + D2H y per step).
roofline.achieved = algorithmic bytes per launch / average launch time.
`--impl reference` times the CPU restatement of the reference's `mul!(y, B', x)` (oracle/, all
host threads) on the same matrix -- Julia is not installed, see DESIGN.md.

The N = 1 line also carries `extra.configs`: the other BASELINE.json configs (C1 cold / warm, C2v, C3 SpMM, C4 adjoint and
triangular solve), each with its time, GB/s, fraction of the HBM peak and parity error; the N > 1 line carries
`exchange_parity` (fused exchange vs NCCL all-gather path, must be 0) and `configs4` (BASELINE.json configs[4] itself:
Float32 / Int32 slab of the n = 50 M matrix, generated on the device).
"""
import argparse
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_LOCAL = 1_000_000
METRIC = "vbc_spmv_gflops"
UNIT = "GFLOP/s"


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        self.active = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self._stop.is_set():
            if not self.active:   # only while a timed region is in flight
                time.sleep(0.0002)
                continue
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.0005)

    def start(self):
        if self.nv and self._t is None:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        return self

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join()

    def summary(self):
        return {"sm_mhz": (statistics.median(self.samples) if self.samples else None),
                "sm_max_mhz": self.max_mhz, "samples": len(self.samples), "reasons": sorted(self.reasons)}


def cpu_reference_setup(A, pi, phi):
    import oracle
    H = oracle.pack_2d(A.m, A.n, A.colptr, A.rowval, A.nzval, pi.spl, phi.spl, 4, 4)
    return oracle, H


def host_threads():
    """All the host threads this process may use.  (torchrun exports OMP_NUM_THREADS=1, which would silently make the
    reference arm single-threaded; the oracle's `num_threads` clause overrides it.)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def time_cpu_reference(oracle, H, x, reps, warm, threads):
    y = np.empty(H.n, dtype=H.val.dtype)
    for _ in range(warm):
        oracle.mul(H, x, trans=True, y=y, nthreads=threads)
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        oracle.mul(H, x, trans=True, y=y, nthreads=threads)
        ts.append(time.perf_counter() - t0)
    return ts, y


def run_reference(args, rank, world):
    """Reference arm: the CPU restatement of `mul!(y, B', x)` on the box's host cores (rank 0 only)."""
    if rank != 0:
        return
    from vbc_b200 import synth
    A, pi, phi = synth.config_c2(n=N_LOCAL)
    oracle, H = cpu_reference_setup(A, pi, phi)
    threads = host_threads()
    x = synth.vector(A.m, 1)
    ts, _ = time_cpu_reference(oracle, H, x, args.steps, args.warmup, threads)
    total = sum(ts)
    gflops = 2.0 * A.nnz * len(ts) / total / 1e9
    oracle.set_static_schedule(True)   # context: the same loops without the reference's shared stripe counter
    ts_static, _ = time_cpu_reference(oracle, H, x, max(3, args.steps // 4), 1, threads)
    oracle.set_static_schedule(False)
    line = {
        "impl": "reference", "metric": METRIC, "value": gflops, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(ts), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(1, A.nnz),
        "cpu_baseline": {"value": gflops, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"full configs[1] matrix (nnz={A.nnz}), {args.steps} adjoint multiplies, OpenMP dynamic,1 over stripes "
                                   f"(the reference's shared-counter @threads loop, multiply_VBC.jl:182-189), "
                                   f"min {2.0 * A.nnz / min(ts) / 1e9:.2f} GFLOP/s; C restatement of the reference CPU path (no Julia in image)",
                         "value_with_static_schedule": 2.0 * A.nnz * len(ts_static) / sum(ts_static) / 1e9},
        "e2e": {"value": gflops, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "same_config_as_repo_arm": (args.gpus == 1),
        "note": (None if args.gpus == 1 else f"the repo arm at N = {args.gpus} multiplies N slabs of this size (weak scaling); both values are rates"),
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(world, nnz_local):
    return {"workload": "configs[1]: 2D-VBC adjoint SpMV y=A'x, Float64, U=W=4, 13-block FEM-like band (S=63), "
                        f"n={N_LOCAL} per GPU, nnz={nnz_local} per GPU",
            "n_global": N_LOCAL * world, "partition": "EquiChunker(4) rows and columns",
            "index_types": "Ti=Int64 canonical arrays; kernel reads 16-B stripe meta + Int32 block descriptors",
            "l2": "inputs (431 MB/GPU) larger than L2 (126 MB); no explicit flush",
            "timing": "K steps captured in one CUDA graph; the graph is replayed R times, each replay between two CUDA events "
                      "that open after a device-side rendezvous of all ranks; ms_per_step = median replay / K, max over ranks "
                      "(every replay is listed in 'replays_ms_per_step'; --exchange nccl: plain launch loop between the events)",
            "parallelism": f"row-block partition over {world} GPU(s)" + (", x exchange fused into the multiply kernel (see 'exchange')" if world > 1 else "")}


def timed_graph(fn, reps, replays=3):
    """seconds per call: `reps` calls captured in one CUDA graph, replayed `replays` times -> (median, min)"""
    import torch
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        for _ in range(reps):
            fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(replays):
        with torch.cuda.stream(side):
            e0.record(); g.replay(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / reps * 1e-3)
    return sorted(ts)[len(ts) // 2], min(ts)


def extra_configs(peak):
    """The other BASELINE.json configs on one GPU (device-resident, CUDA events), each with a parity error against the CSC
    product of the same matrix on the host (scipy; the oracle is pinned to it in tests/)."""
    import torch
    import vbc_b200 as vb
    from vbc_b200 import synth
    out = {}

    def bound_err(y, yref, absAx, tol):
        return float(np.max(np.abs(y - yref) / np.maximum(tol * absAx, 1e-300)))  # <= 1 passes at `tol`

    def adj_record(B, A, reps, cold_flush=None):
        x = synth.vector(A.m, 7)
        xd, yd = torch.from_numpy(x).cuda(), torch.empty(A.n, dtype=torch.float64, device="cuda")
        med, mn = timed_graph(lambda: vb.mul_(yd, B.T, xd), reps)
        S = A.to_scipy()
        yref = S.T @ x
        err = bound_err(yd.cpu().numpy(), yref, abs(S).T @ np.abs(x), 1e-12)
        nb = B.format_bytes()[1] + 8 * (A.m + A.n)
        rec = {"us": med * 1e6, "us_min": mn * 1e6, "algorithmic_bytes": nb, "GBps": nb / med / 1e9, "frac_of_hbm_peak": nb / med / 1e9 / peak,
               "gflops": 2.0 * A.nnz / med / 1e9, "nnz": A.nnz, "parity_err_over_bound_1e-12": err, "launches_per_multiply": 1}
        if cold_flush is not None:  # L2 flushed before every launch: one event pair per launch
            ts = []
            for _ in range(20):
                cold_flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); vb.mul_(yd, B.T, xd); e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e-3)
            ts.sort()
            rec["cold_us"] = ts[len(ts) // 2] * 1e6
            rec["cold_GBps"] = nb / ts[len(ts) // 2] / 1e9
            rec["cold_frac_of_hbm_peak"] = rec["cold_GBps"] / peak
            rec["cold_note"] = "512 MB buffer written before every launch (L2 flush); one CUDA-event pair per launch, median of 20"
        return rec

    # C1: 1D-VBC Float64, n = 10 000, W = 8, nnz = 1 000 000 (L2-resident when warm: launch-latency regime)
    t0 = time.perf_counter()
    A, phi = synth.config_c1()
    B = vb.SparseMatrix1DVBC[8](A, phi)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    out["C1_1D_f64_n10k_w8"] = adj_record(B, A, 200, cold_flush=flush)
    out["C1_1D_f64_n10k_w8"]["warm_note"] = "9.2 MB working set stays in the 126 MB L2 between launches (200 launches in one graph)"
    del flush
    B.close()
    # C2v: variable 2..8 blocks (U = W = 8) on the C2 matrix
    A, _, _ = synth.config_c2()
    pv, fv = synth.variable_partition(A.m, 8, 3), synth.variable_partition(A.n, 8, 4)
    B = vb.SparseMatrixVBC[8, 8](A, pv, fv)
    out["C2v_2D_f64_variable_2to8_blocks"] = adj_record(B, A, 50)
    # ... and its forward multiply (the rows-mode transposed copy of the variable blocks, built at the first call)
    xv = synth.vector(A.n, 11)
    xvd, yvd = torch.from_numpy(xv).cuda(), torch.empty(A.m, dtype=torch.float64, device="cuda")
    vb.mul_(yvd, B, xvd)
    med, mn = timed_graph(lambda: vb.mul_(yvd, B, xvd), 50)
    Sv = A.to_scipy()
    nb = B.format_bytes()[2] + 8 * (A.m + A.n)
    out["C2v_forward_2D_f64_variable_2to8_blocks"] = {"us": med * 1e6, "us_min": mn * 1e6, "algorithmic_bytes": nb, "GBps": nb / med / 1e9, "frac_of_hbm_peak": nb / med / 1e9 / peak,
                                                      "gflops": 2.0 * A.nnz / med / 1e9, "parity_err_over_bound_1e-12": bound_err(yvd.cpu().numpy(), Sv @ xv, abs(Sv) @ np.abs(xv), 1e-12)}
    del Sv
    B.close()
    # forward multiply on the C2 matrix (the reference's serial scatter orientation, multiply_VBC.jl:3-87)
    from vbc_b200.partition import SplitPartition
    pi = SplitPartition(np.arange(1, A.m + 2, 4, dtype=np.int64)); phi = SplitPartition(np.arange(1, A.n + 2, 4, dtype=np.int64))
    B = vb.SparseMatrixVBC[4, 4](A, pi, phi)
    x = synth.vector(A.n, 9)
    xd, yd = torch.from_numpy(x).cuda(), torch.empty(A.m, dtype=torch.float64, device="cuda")
    vb.mul_(yd, B, xd)  # builds the transposed index outside graph capture
    med, mn = timed_graph(lambda: vb.mul_(yd, B, xd), 50)
    S = A.to_scipy()
    nb = B.format_bytes()[2] + 8 * (A.m + A.n)
    out["C2_forward_2D_f64"] = {"us": med * 1e6, "us_min": mn * 1e6, "algorithmic_bytes": nb, "GBps": nb / med / 1e9, "frac_of_hbm_peak": nb / med / 1e9 / peak,
                                "gflops": 2.0 * A.nnz / med / 1e9, "parity_err_over_bound_1e-12": bound_err(yd.cpu().numpy(), S @ x, abs(S) @ np.abs(x), 1e-12)}
    B.close()
    # the same matrix stored in Float32 / Int32, applied to Float64 vectors (the reference converts values and x to eltype(y): vbc_spmv_mixed)
    A32 = A.astype(np.float32, np.int32)
    B32 = vb.SparseMatrixVBC[4, 4](A32, pi.astype(np.int32), phi.astype(np.int32))
    S32 = A32.to_scipy().astype(np.float64)
    xm = synth.vector(A.m, 15)
    xmd, ymd = torch.from_numpy(xm).cuda(), torch.empty(A.n, dtype=torch.float64, device="cuda")
    vb.mul_(ymd, B32.T, xmd)
    med, mn = timed_graph(lambda: vb.mul_(ymd, B32.T, xmd), 50)
    nb = B32.format_bytes()[1] + 8 * (A.m + A.n)
    out["C2_mixed_f32_matrix_f64_vectors_adjoint"] = {"us": med * 1e6, "us_min": mn * 1e6, "algorithmic_bytes": nb, "GBps": nb / med / 1e9, "frac_of_hbm_peak": nb / med / 1e9 / peak,
                                                      "gflops": 2.0 * A.nnz / med / 1e9, "parity_err_over_bound_1e-12": bound_err(ymd.cpu().numpy(), S32.T @ xm, abs(S32).T @ np.abs(xm), 1e-12)}
    B32.close()
    # the same matrix with Int32 elements (test/runtests.jl:16): wrapping arithmetic, exact -- a semantics path (csrc/inttypes.cu), timed for the record
    Ai = vb.SparseMatrixCSC(A.m, A.n, A32.colptr, A32.rowval, (A.nzval * 65536.0 - 32768.0).astype(np.int32))
    Bi = vb.SparseMatrixVBC[4, 4](Ai, pi.astype(np.int32), phi.astype(np.int32))
    xi = (synth.vector(A.m, 17) * 65536.0 - 32768.0).astype(np.int32)
    xid, yid = torch.from_numpy(xi).cuda(), torch.empty(A.n, dtype=torch.int32, device="cuda")
    vb.mul_(yid, Bi.T, xid)
    med, mn = timed_graph(lambda: vb.mul_(yid, Bi.T, xid), 20)
    Si = Ai.to_scipy().astype(np.int64)
    exact = bool(np.array_equal(yid.cpu().numpy(), (Si.T @ xi.astype(np.int64)).astype(np.int32)))  # int64 sums (no overflow at these magnitudes) wrapped to Int32
    nb = Bi.format_bytes()[1] + 4 * (A.m + A.n)
    out["C2_int32_elements_adjoint"] = {"us": med * 1e6, "us_min": mn * 1e6, "algorithmic_bytes": nb, "GBps": nb / med / 1e9, "frac_of_hbm_peak": nb / med / 1e9 / peak,
                                        "exact_vs_wrapped_int64_product": exact, "note": "Int32 sums overflow and wrap as in Julia; kernel k_int_adj (16-byte loads, 8 lanes per stripe)"}
    xf = (synth.vector(A.n, 19) * 65536.0 - 32768.0).astype(np.int32)
    xfd, yfd = torch.from_numpy(xf).cuda(), torch.empty(A.m, dtype=torch.int32, device="cuda")
    vb.mul_(yfd, Bi, xfd)  # builds the transposed copy outside graph capture
    fmed, _ = timed_graph(lambda: vb.mul_(yfd, Bi, xfd), 20)
    out["C2_int32_elements_adjoint"]["forward_us"] = fmed * 1e6
    out["C2_int32_elements_adjoint"]["forward_exact"] = bool(np.array_equal(yfd.cpu().numpy(), (Si @ xf.astype(np.int64)).astype(np.int32)))
    Bi.close()
    del A, S, A32, S32, Ai, Si
    # C3: 1D-VBC SpMM, k = 32, n = 1M, W = 8, 50 rows per stripe (parity unpinned by the reference: its matrix `*` cannot run)
    K, L, k = 1_000_000, 125_000, 32
    A, _, phi = synth.banded_blocks(K, L, 1, 8, np.arange(-25, 25) * 37)
    B = vb.SparseMatrix1DVBC[8](A, phi)
    X = torch.rand(A.m, k, dtype=torch.float64, device="cuda")
    Y = torch.empty(A.n, k, dtype=torch.float64, device="cuda")
    med, mn = timed_graph(lambda: vb.mul_(Y, B.T, X), 10)
    S = A.to_scipy()
    # the same 1D matrix: adjoint and forward SpMV (a6 / a7 of SURVEY 8a)
    out["C3m_1D_f64_w8_adjoint"] = adj_record(B, A, 30)
    xf = synth.vector(A.n, 13)
    xfd, yfd = torch.from_numpy(xf).cuda(), torch.empty(A.m, dtype=torch.float64, device="cuda")
    vb.mul_(yfd, B, xfd)  # builds the transposed copy outside graph capture
    fmed, fmn = timed_graph(lambda: vb.mul_(yfd, B, xfd), 30)
    fnb = B.format_bytes()[2] + 8 * (A.m + A.n)
    out["C3m_1D_f64_w8_forward"] = {"us": fmed * 1e6, "us_min": fmn * 1e6, "algorithmic_bytes": fnb, "GBps": fnb / fmed / 1e9, "frac_of_hbm_peak": fnb / fmed / 1e9 / peak,
                                    "gflops": 2.0 * A.nnz / fmed / 1e9, "parity_err_over_bound_1e-12": bound_err(yfd.cpu().numpy(), S @ xf, abs(S) @ np.abs(xf), 1e-12)}
    Xh = X.cpu().numpy()
    err = bound_err(Y.cpu().numpy(), S.T @ Xh, abs(S).T @ np.abs(Xh), 1e-12)
    nb = B.format_bytes()[1] + 8 * k * (A.m + A.n)
    out["C3_SpMM_1D_f64_k32"] = {"us": med * 1e6, "us_min": mn * 1e6, "algorithmic_bytes": nb, "GBps": nb / med / 1e9, "frac_of_hbm_peak": nb / med / 1e9 / peak,
                                 "tflops": 2.0 * A.nnz * k / med / 1e12, "kernel": "k_spmm_adj_tma (FP64 mma.sync m8n8k4 tiles fed by TMA tile::gather4 row gathers)", "parity_err_over_bound_1e-12": err,
                                 "parity": "unpinned by the reference (SURVEY.md R3); checked against k independent CSC products"}
    B.close()
    del A, S, X, Y, Xh
    # C4 (i) the reference's meaning of TrSpMV: the adjoint multiply at n = 2M; (ii) BASELINE's wording: lower-triangular solve
    A, pi, phi = synth.config_c2(n=2_000_000, S=79)
    B = vb.SparseMatrixVBC[4, 4](A, pi, phi)
    out["C4i_TrSpMV_adjoint_2D_f64_n2M"] = adj_record(B, A, 30)
    B.close()
    del A
    A, pi, phi = synth.config_c4_triangular()
    B = vb.SparseMatrixVBC[4, 4](A, pi, phi)
    nlev = vb.trsv_analyse(B.T)
    b = synth.vector(A.n, 11)
    bd, xd = torch.from_numpy(b).cuda(), torch.empty(A.n, dtype=torch.float64, device="cuda")
    for _ in range(2):
        vb.ldiv_lower_(xd, B.T, bd)
    B.sync()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); vb.ldiv_lower_(xd, B.T, bd); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    import scipy.sparse as sp
    T = sp.tril(A.to_scipy().T).tocsr()
    xh = xd.cpu().numpy()
    resid = float(np.max(np.abs(T @ xh - b)) / np.max(np.abs(b)))
    out["C4ii_triangular_solve_2D_f64_n2M"] = {"ms": ts[len(ts) // 2], "ms_min": ts[0], "levels": nlev, "us_per_level": 1e3 * ts[len(ts) // 2] / max(nlev, 1),
                                               "residual_max_rel": resid, "parity": "extension, unpinned by the reference (SURVEY.md R2); residual of tril(A')x=b"}
    B.close()
    out["_seconds"] = time.perf_counter() - t0
    return out


def configs4_slab(rank, world, dev, steps, peak):
    """BASELINE.json configs[4] itself: row-partitioned 2D-VBC Float32 / Int32, n = 50 M, U = W = 4, 10 blocks per stripe
    (nnz = 2.0 B), every rank's slab of 50 M / N columns GENERATED ON THE DEVICE (vbc_gen_banded_csc) and packed there."""
    import torch
    import torch.distributed as dist
    import vbc_b200 as vb
    from vbc_b200 import dist as vdist
    from vbc_b200 import synth
    u = w = 4
    n = 50_000_000
    n_local = n // world
    K = L = n // 4
    Lloc = n_local // 4
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    D = synth.banded_blocks_device(K, L, u, w, synth.C5_OFFSETS, dtype=np.float32, ti=np.int32, stripes=(rank * Lloc, (rank + 1) * Lloc), device=dev)
    t_gen = time.perf_counter() - t0
    pi_spl = torch.arange(1, n + 2, u, dtype=torch.int32, device="cuda")
    phi_spl = torch.arange(1, n_local + 2, w, dtype=torch.int32, device="cuda")
    t0 = time.perf_counter()
    B = vb.SparseMatrixVBC.from_device_csc(u, w, n, n_local, D.colptr, D.rowval, D.nzval, pi_spl, phi_spl, device=dev)
    torch.cuda.synchronize()
    t_pack = time.perf_counter() - t0
    nnz_local = D.nnz
    layout = vdist.PaddedLayout(np.arange(world + 1, dtype=np.int64) * n_local)
    alpha = 1.0 / 160.0
    peer = vdist.PeerExchangeOperator(B, layout, rank, world, dev, alpha=alpha, halo=True)
    # x0 generated on the device too (same hash as synth.vector): [0, 1)
    g = torch.Generator(device="cuda"); g.manual_seed(1234)
    x0 = torch.rand(n, dtype=torch.float32, device="cuda", generator=g)  # same seed on every rank -> same vector
    peer._as_tensor(peer.current()).copy_(x0)
    torch.cuda.synchronize()
    dist.barrier()
    # sampled parity BEFORE the CSC arrays are freed: column j of the slab, dot(nzval, x[rowval]) in Float64 on the device
    peer.step(); peer.finish()
    torch.cuda.synchronize()
    x1 = peer._as_tensor(peer.current())[rank * n_local: (rank + 1) * n_local]
    gs = torch.Generator(device="cuda"); gs.manual_seed(99 + rank)
    cols = torch.randint(0, n_local, (4096,), device="cuda", generator=gs)
    cols[0], cols[1] = 0, n_local - 1
    cp = D.colptr.long()
    b0, b1 = cp[cols] - 1, cp[cols + 1] - 1
    maxlen = int((b1 - b0).max())
    ar = torch.arange(maxlen, device="cuda")[None, :]
    idx = torch.minimum(b0[:, None] + ar, b1[:, None] - 1).clamp_(min=0)
    valid = ar < (b1 - b0)[:, None]
    vals = D.nzval[idx].double() * valid
    xs = x0[(D.rowval[idx].long() - 1)].double()
    ref = alpha * (vals * xs).sum(dim=1)
    absref = alpha * (vals.abs() * xs.abs()).sum(dim=1)
    err = float(((x1[cols].double() - ref).abs() / (1e-5 * absref).clamp_(min=1e-30)).max())
    D.free()
    del cp, idx, vals, xs
    for _ in range(3):
        peer.step()
    torch.cuda.synchronize()
    side, gr = torch.cuda.Stream(), torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr, stream=side):
        for _ in range(steps):
            peer.step()
    torch.cuda.synchronize()
    tok = torch.zeros(1, device="cuda")
    reps = []
    for _ in range(3):
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(side):
            torch.cuda._sleep(2_000_000)
            dist.all_reduce(tok)          # device-side rendezvous: every rank's window opens together
            e0.record(); gr.replay(); e1.record()
        torch.cuda.synchronize()
        reps.append(e0.elapsed_time(e1) / steps)
    t = torch.tensor(reps + [float(nnz_local), err], device="cuda", dtype=torch.float64)
    mx = t.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    sm = t.clone(); dist.all_reduce(sm)
    reps_max = sorted(float(v) for v in mx[:3])
    ms = reps_max[1]
    ref_b, adj_b, _ = B.format_bytes()
    alg = adj_b + 4 * 2 * n_local
    rec = {"workload": "BASELINE.json configs[4]: row-partitioned 2D-VBC adjoint SpMV, Float32 values / Int32 indices, n = 50 000 000, U = W = 4, "
                       "10 blocks per stripe, slabs generated and packed on the device",
           "n_gpus": world, "n_local": n_local, "nnz_total": float(sm[3]), "steps": steps, "ms_per_step": ms, "replays_ms_per_step": reps_max,
           "gflops": 2.0 * float(sm[3]) / (ms * 1e-3) / 1e9, "per_gpu_algorithmic_bytes": alg, "per_gpu_GBps": alg / (ms * 1e-3) / 1e9,
           "frac_of_hbm_peak_per_gpu": alg / (ms * 1e-3) / 1e9 / peak, "kernel": "k_spmv_adj_halo<float,%d,DESC_BLOCKS>" % 4,
           "sampled_parity_err_over_bound_1e-5": float(mx[4]), "sampled_parity": "4096 columns per rank: dot(nzval, x[rowval]) of the generated CSC slab in Float64 vs one fused step",
           "gen_seconds": t_gen, "pack_seconds": t_pack, "interior": list(peer.interior), "sent_fraction": peer.sent_fraction, "timed_out": bool(peer.timed_out())}
    peer.close()
    B.close()
    return rec


def run_ours(args, rank, world, local_rank):
    import torch
    import vbc_b200 as vb
    from vbc_b200 import dist as vdist
    from vbc_b200 import synth

    torch.cuda.set_device(local_rank)
    dev = local_rank
    if world > 1:
        import torch.distributed as dist
        # NCCL's own log (the driver reads the communicator's rank count from it) must stay on, but off stdout: this
        # process prints exactly one JSON line there.  NCCL_DEBUG_FILE would hide it from the driver, so stdout is
        # pointed at stderr while NCCL initialises and speaks, and restored for the JSON line.
        if os.environ.get("NCCL_DEBUG", "").upper() not in ("INFO", "TRACE"):   # VERSION / WARN / unset: raise it to INFO for the init lines
            os.environ["NCCL_DEBUG"] = "INFO"
            os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT")
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        tok = torch.zeros(1, device="cuda")
        dist.all_reduce(tok)  # forces communicator creation now
        torch.cuda.synchronize()
    clk = ClockSampler(dev)

    u = w = 4
    n_glob = N_LOCAL * world
    L_loc = N_LOCAL // w
    # weak scaling: the uniform band makes equal stripe counts the cost-balanced split (checked below)
    A, pi, phi = synth.config_c2(n=n_glob, stripes=(rank * L_loc, (rank + 1) * L_loc))
    layout = vdist.PaddedLayout(np.arange(world + 1, dtype=np.int64) * N_LOCAL)
    if world > 1:
        A = vdist.remap_rows_to_padded(A, layout, u)
        pi = vdist.padded_row_partition(layout, u, A.colptr.dtype.type)
    t0 = time.perf_counter()
    B = vb.SparseMatrixVBC[u, w](A, pi, phi, device=dev)
    pack_s = time.perf_counter() - t0
    nnz_local = A.nnz
    ref_bytes, adj_bytes, _ = B.format_bytes()
    if world > 1:
        cost, _ = B.memory_cost()
        tot = torch.tensor([float(cost.sum())], device="cuda")
        lst = [torch.zeros_like(tot) for _ in range(world)]
        dist.all_gather(lst, tot)
        imbalance = max(t.item() for t in lst) / (sum(t.item() for t in lst) / world)
    else:
        imbalance = 1.0

    # x_{t+1} <- alpha * A' x_t : alpha ~ 1 / (row sum) keeps the iterates in range at N > 1
    alpha_step = 1.0 if world == 1 else 1.0 / 26.0
    op = vdist.RowPartitionedOperator(lambda y, x: vb.mul_(y, B.T, x, alpha_step, False), layout, rank, world, torch.float64)
    xg = synth.vector(n_glob, 1)
    op.set_x(xg)
    peer = None
    if world > 1 and args.exchange in ("peer", "halo"):
        peer = vdist.PeerExchangeOperator(B, layout, rank, world, dev, alpha=alpha_step, halo=(args.exchange == "halo"), sync_mode=args.sync_mode)
        peer.set_x(xg)
        dist.barrier()

    def step():
        if peer is not None:
            peer.step()          # ONE launch: multiply + exchange of the new x (+ a flag kernel with --sync-mode 1)
        elif world > 1:
            op.step()            # multiply, then NCCL all-gather
        else:
            op.local_multiply()

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    # The K timed steps are captured into ONE CUDA graph (libvbc launches on torch's current stream,
    # so they are captured like any other work): a timed region holds exactly K steps with no host launch
    # latency in between.  The graph is replayed R times; each replay is its own timed region of K steps.
    use_graph = not (world > 1 and peer is None)  # NCCL collectives are left out of graph capture (plain loop)
    side = torch.cuda.Stream()
    launches_before = B.launch_count()
    K = args.steps
    replays = max(1, args.replays)
    rep_ms = []
    if peer is not None:
        peer.wait_stats(reset=True)
    if use_graph:
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            for _ in range(K):
                step()
        flag_launches = 0 if (peer is None or args.sync_mode == 0) else 1
        gpu_launches = B.launch_count() - launches_before + K * flag_launches
        torch.cuda.synchronize()
    clk.start()
    for rep in range(replays):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        clk.active = True
        with torch.cuda.stream(side):
            if world > 1:
                # device-side rendezvous: a short sleep lets this rank's host enqueue everything below, then a tiny
                # all-reduce releases all ranks together, so every rank's timed window opens at the same moment
                torch.cuda._sleep(2_000_000)
                dist.all_reduce(tok)
            ev0.record()
            if use_graph:
                graph.replay()
            else:
                for _ in range(K):
                    step()
            ev1.record()
        torch.cuda.synchronize()
        clk.active = False
        rep_ms.append(ev0.elapsed_time(ev1))
    clk.stop()
    if not use_graph:
        gpu_launches = (B.launch_count() - launches_before) // replays
    if world > 1:
        dist.barrier()
    if peer is not None and peer.timed_out():
        raise RuntimeError("peer flag wait timed out")
    wait_stats = peer.wait_stats() if peer is not None else None

    # duration of the plain single-GPU kernel on this slab (no exchange): one event pair per launch, so host launch
    # latency between launches does not count
    for _ in range(3):
        op.local_multiply()
    torch.cuda.synchronize()
    pairs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(max(K, 20))]
    for a, b in pairs:
        a.record()
        op.local_multiply()
        b.record()
    torch.cuda.synchronize()
    kt = sorted(a.elapsed_time(b) for a, b in pairs)
    ms_kernel = sum(kt) / len(kt)
    ms_kernel_min, ms_kernel_med = kt[0], kt[len(kt) // 2]
    # ... and the same plain kernel timed the way the step is (K launches in one graph): what the exchange is compared with
    ms_plain_graph = 1e3 * timed_graph(op.local_multiply, K)[0] if world > 1 else None

    # ---- parity of the distributed iteration: the fused exchange and the NCCL all-gather path, same x0, same number of
    # steps, must end with bit-identical x (every stripe runs the same body in both); and against the host CSC product
    exchange_parity = None
    if world > 1 and peer is not None:
        S_par = 6
        peer.set_x(xg)
        op.set_x(xg)
        dist.barrier()
        for _ in range(S_par):
            peer.step()
        peer.finish()
        for _ in range(S_par):
            op.step()
        torch.cuda.synchronize()
        xa, xb = peer.x_global(), op.x_global()
        # every rank also checks the halo it RECEIVED: its whole current buffer against the all-gather path's x where it reads
        need = B.read_chunks(7).astype(bool)
        mine = peer._as_tensor(peer.current()).cpu().numpy()
        full = op.x.cpu().numpy()
        rd = np.repeat(need, 128)[: len(mine)]
        halo_bad = torch.tensor([float(np.count_nonzero(mine[rd] != full[rd]))], device="cuda")
        dist.all_reduce(halo_bad)
        exchange_parity = {"max_abs_diff_vs_nccl_allgather_path": float(np.max(np.abs(xa - xb))), "steps": S_par,
                           "entries_read_by_some_rank_that_differ": float(halo_bad.item()),
                           "x_abs_sum": float(np.abs(xa).sum())}

    # ---- e2e: the public API with host vectors (pinned), H2D + kernel + D2H every step
    xh = torch.from_numpy(layout.scatter(xg)).pin_memory()
    yh = torch.empty(N_LOCAL, dtype=torch.float64).pin_memory()
    xh_np, yh_np = xh.numpy(), yh.numpy()
    for _ in range(3):
        vb.mul_(yh_np, B.T, xh_np)
    torch.cuda.synchronize()
    e2e_steps = max(10, min(K, 100))
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        vb.mul_(yh_np, B.T, xh_np)  # synchronous on return
    e2e_s = time.perf_counter() - t0
    result_checksum = float(yh_np.sum())
    from vbc_b200 import _lib
    h2d_elems = B.get_option(_lib.OPT_E2E_UPLOAD_ELEMS)

    # max over ranks
    rep_t = torch.tensor(rep_ms + [ms_kernel, e2e_s, float(h2d_elems), ms_plain_graph or 0.0], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(rep_t, op=dist.ReduceOp.MAX)
        nn = torch.tensor([float(nnz_local)], device="cuda", dtype=torch.float64)
        dist.all_reduce(nn)
        nnz_total = float(nn.item())
        ws = torch.tensor([float(wait_stats[k]) for k in ("wait_ns", "waits_that_spun", "longest_wait_ns")] if wait_stats else [0.0, 0.0, 0.0],
                          device="cuda", dtype=torch.float64)
        wl = [torch.zeros_like(ws) for _ in range(world)]
        dist.all_gather(wl, ws)
        wait_table = [{"rank": r, "wait_us_total_lanes": float(v[0]) / 1e3, "waits_that_spun": int(v[1]), "longest_wait_us": float(v[2]) / 1e3} for r, v in enumerate(wl)]
    else:
        nnz_total = float(nnz_local)
        wait_table = None
    vals = [float(v) for v in rep_t.tolist()]
    rep_ms_max, ms_kernel, e2e_s, h2d_elems, ms_plain_graph = vals[:replays], vals[replays], vals[replays + 1], vals[replays + 2], vals[replays + 3]

    c4 = None
    if world > 1 and peer is not None and not args.no_configs4:
        peer.close()
        peer = None
        B.close()
        del op
        torch.cuda.empty_cache()
        peak0, _ = measured_peak()
        try:
            c4 = configs4_slab(rank, world, dev, max(10, min(K, 50)), peak0)
        except Exception as e:  # keep the headline line even if the big slab does not fit / fails
            c4 = {"error": repr(e)[:500]}

    if rank == 0:
        ms_sorted = sorted(rep_ms_max)
        ms_total = ms_sorted[len(ms_sorted) // 2]   # median replay of K steps (max over ranks per replay)
        ms_step = ms_total / K
        value = 2.0 * nnz_total / (ms_step * 1e-3) / 1e9
        peak, peak_src = measured_peak()
        vec_bytes = 16 * N_LOCAL   # x entries the slab gathers from (own slice; the halo is a fraction of a percent) + y
        alg_bytes = adj_bytes + vec_bytes
        # one launch per step at every N (N > 1: the fused multiply + exchange kernel), so the average launch duration of the
        # dominant kernel is the step time itself; --sync-mode 1 / --exchange nccl add launches and are labelled below
        ms_roof = ms_step
        achieved = alg_bytes / (ms_roof * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tp) and world == 1:
            try:
                traffic = json.load(open(tp)).get("k_spmv_adj_dram_bytes_per_launch")
            except Exception:
                traffic = None
        kernel_name = "k_spmv_adj<double,8,DESC_BLOCKS>" if world == 1 else (
            "k_spmv_adj_halo<double,8,DESC_BLOCKS> (multiply + exchange, one launch per step)" if peer_mode(args) else "k_spmv_adj<double,8,DESC_BLOCKS> + NCCL all-gather")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(world, nnz_local),
            "replays": replays, "replays_ms_per_step": [m / K for m in rep_ms_max], "ms_per_step_min": ms_sorted[0] / K,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": ("ncu --set full capture of this kernel, profiles/ncu_traffic.json" if traffic else None),
                         "peak_source": peak_src, "kernel": kernel_name,
                         "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": ms_roof,
                         "plain_kernel_ms_graph": (ms_plain_graph if world > 1 else ms_step), "plain_kernel_ms_event_pairs_mean": ms_kernel, "plain_kernel_ms_min": ms_kernel_min, "plain_kernel_ms_median": ms_kernel_med,
                         "frac_of_nominal_8TBs": achieved / 8000.0,
                         "reference_format_bytes": ref_bytes + vec_bytes},
            "e2e": {"value": 2.0 * nnz_total * e2e_steps / e2e_s / 1e9, "unit": UNIT,
                    "h2d_bytes_per_step": int(h2d_elems * 8), "d2h_bytes_per_step": int(yh_np.nbytes),
                    "steps": e2e_steps, "ms_per_step": 1e3 * e2e_s / e2e_steps, "result_checksum": result_checksum,
                    "note": "per rank: pinned host x (the window of x this rank's stripes gather from, uploaded in pieces) -> kernel chunks -> y ranges copied back while later chunks run"},
            "gpu_launches": int(gpu_launches), "clocks": clk.summary(),
            "pack_seconds": pack_s, "cost_imbalance": imbalance, "alpha": alpha_step,
            "exchange": (None if world == 1 else {"halo": "fused into the multiply kernel (NVLink peer stores + per-step flags); each y segment goes to the ranks whose stripes read it",
                                                  "peer": "fused into the multiply kernel (NVLink peer stores + per-step flags); every y segment goes to every rank (all-gather)",
                                                  "nccl": "torch.distributed all_gather_into_tensor after the multiply"}[args.exchange]),
            "exchange_mode": (args.exchange if world > 1 else None),
            "exchange_sync": (None if world == 1 or not peer_mode(args) else
                              {0: "in-kernel: boundary stripes wait for the neighbours' previous step, the last one publishes this step",
                               1: "separate flag kernel (signal + wait) after every multiply"}[args.sync_mode]),
            "exchange_overhead_us": (None if world == 1 else 1e3 * (ms_step - ms_plain_graph)),
            "exchange_overhead_note": (None if world == 1 else "step minus the plain single-GPU kernel on the same slab, both as K launches in one CUDA graph (max over ranks)"),
            "exchange_wait_per_rank": wait_table,
            "exchange_parity": exchange_parity,
            "configs4": c4,
        }
        if world == 1 and not args.no_cpu_baseline:
            oracle, H = cpu_reference_setup(A, pi, phi)
            threads = host_threads()
            ts, y_cpu = time_cpu_reference(oracle, H, xg, 10, 2, threads)
            # the other rows of BASELINE.md's CPU plan, a few samples each: adjoint on 1 thread, the (serial) forward
            # multiply and the (serial) CSC TrSpMV! of the reference
            def _best(fn, reps=3):
                ts_ = []
                for _ in range(reps):
                    t0_ = time.perf_counter(); fn(); ts_.append(time.perf_counter() - t0_)
                return 2.0 * nnz_local / min(ts_) / 1e9
            ybuf, mbuf = np.empty(H.n), np.empty(H.m)
            xn_ = synth.vector(H.n, 2)
            extra = {"adjoint_1_thread": _best(lambda: oracle.mul(H, xg, trans=True, y=ybuf, nthreads=1)),
                     "forward_serial": _best(lambda: oracle.mul(H, xn_, trans=False, y=mbuf)),
                     "csc_trspmv_serial": _best(lambda: oracle.csc_trspmv(A.m, A.n, A.colptr, A.rowval, A.nzval, xg, y=ybuf))}
            oracle.set_static_schedule(True)   # context only: the same loops without the reference's shared counter
            ts_static, _ = time_cpu_reference(oracle, H, xg, 5, 1, threads)
            oracle.set_static_schedule(False)
            y_gpu = vb.mul_(np.empty(N_LOCAL), B.T, xg)
            err = float(np.max(np.abs(y_gpu - y_cpu) / np.maximum(np.abs(y_cpu), 1e-300)))
            line["cpu_baseline"] = {"value": 2.0 * nnz_local * len(ts) / sum(ts) / 1e9, "unit": UNIT, "cores": threads,
                                    "kind": "port",
                                    "sample": f"full configs[1] matrix, 10 adjoint multiplies after 2 warm-ups, all {threads} host threads "
                                              f"(OpenMP dynamic,1 over stripes); C restatement of the reference CPU path",
                                    "min_ms": 1e3 * min(ts), "max_rel_err_gpu_vs_cpu": err,
                                    "value_with_static_schedule": 2.0 * nnz_local * len(ts_static) / sum(ts_static) / 1e9,
                                    "other_rows_gflops": extra}
        if world == 1 and not args.no_extra:
            B.close()
            del A
            try:
                line["extra"] = {"configs": extra_configs(peak)}
            except Exception as e:
                line["extra"] = {"error": repr(e)[:500]}
        if world > 1:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        if world > 1:
            os.dup2(2, 1)  # NCCL also speaks while the communicator is destroyed: keep that off stdout too
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def peer_mode(args):
    return args.exchange in ("peer", "halo")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--replays", type=int, default=5, help="timed regions of K steps each (the reported step time is the median replay)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="N = 1: skip the sub-records of the other BASELINE configs")
    ap.add_argument("--no-configs4", action="store_true", help="N > 1: skip the BASELINE configs[4] sub-record")
    ap.add_argument("--sync-mode", type=int, default=0, choices=[0, 1],
                    help="fused exchange: 0 = flags inside the multiply kernel (one launch per step); 1 = a separate flag kernel "
                         "(signal + wait) after every multiply (comparator)")
    ap.add_argument("--exchange", default="halo", choices=["halo", "peer", "nccl"],
                    help="N > 1, how x_{t+1} reaches the ranks: 'halo' (default) = exchange fused into the multiply kernel through "
                         "NVLink peer stores, each y segment sent to exactly the ranks whose stripes read it; 'peer' = same kernel, "
                         "every segment sent to every rank (full replication, a fused all-gather); 'nccl' = multiply, then "
                         "torch.distributed all_gather_into_tensor")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under torch.distributed.run
        import subprocess
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    if world > 1:
        bind_to_gpu_cpus(local_rank)  # N = 1 keeps every core: the cpu_baseline leg uses them all
    run_ours(args, rank, world, local_rank)


def bind_to_gpu_cpus(local_rank):
    """Pin this rank to the CPU cores NVML reports as local to its GPU (before any pinned host buffer is allocated: first touch
    then places the e2e leg's host vectors on the GPU's own NUMA node, so that N ranks do not share one memory controller for
    their PCIe copies).  Best effort: any failure leaves the affinity alone."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:
        pass


if __name__ == "__main__":
    main()
