#!/usr/bin/env python
"""bench.py -- VBC SpMV throughput on B200 (BASELINE.json metric: GFLOP/s + achieved HBM GB/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one adjoint multiply y <- A' x (`mul!(y, B', x)`, the reference's benchmarked kernel,
bin/test_table.jl:122, costs.jl:224) over the synthetic matrix of BASELINE.json configs[1]:
2D-VBC, Float64, m = n = 1 000 000, U = W = 4, 13-block FEM-like band, nnz ≈ 52 M.  At N > 1
(torchrun, one rank per GPU) every rank owns a 1M-column slab of an (N·1M)² matrix (weak scaling)
and each step ends with the all-gather of the y slices into everyone's x over NVLink.

value  = useful GFLOP/s (2·nnz per multiply), whole job, inputs resident in HBM, CUDA events.
e2e    = same metric through the public API with HOST vectors (H2D x + kernel + D2H y per step).
roofline.achieved = algorithmic bytes per launch / average launch time.
`--impl reference` times the CPU restatement of the reference's `mul!(y, B', x)` (oracle/, all
host threads) on the same matrix -- Julia is not installed, see DESIGN.md.
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
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.002)

    def __enter__(self):
        if self.nv:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *a):
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
    line = {
        "impl": "reference", "metric": METRIC, "value": gflops, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(ts), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(1, A.nnz),
        "cpu_baseline": {"value": gflops, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"full configs[1] matrix (nnz={A.nnz}), {args.steps} adjoint multiplies, OpenMP dynamic,1 over stripes, "
                                   f"min {2.0 * A.nnz / min(ts) / 1e9:.2f} GFLOP/s; C restatement of the reference CPU path (no Julia in image)"},
        "e2e": {"value": gflops, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(world, nnz_local):
    return {"workload": "configs[1]: 2D-VBC adjoint SpMV y=A'x, Float64, U=W=4, 13-block FEM-like band (S=63), "
                        f"n={N_LOCAL} per GPU, nnz={nnz_local} per GPU",
            "n_global": N_LOCAL * world, "partition": "EquiChunker(4) rows and columns",
            "index_types": "Ti=Int64 canonical arrays; kernel reads 16-B stripe meta + Int32 block descriptors",
            "l2": "inputs (431 MB/GPU) larger than L2 (126 MB); no explicit flush",
            "timing": "K steps captured in one CUDA graph, replayed once between two CUDA events; max over ranks "
                      "(--exchange nccl: plain launch loop between the events)",
            "parallelism": f"row-block partition over {world} GPU(s)" + (", x exchange per step (see 'exchange')" if world > 1 else "")}


def run_ours(args, rank, world, local_rank):
    import torch
    import vbc_b200 as vb
    from vbc_b200 import dist as vdist
    from vbc_b200 import synth

    torch.cuda.set_device(local_rank)
    dev = local_rank
    if world > 1:
        import torch.distributed as dist
        # keep NCCL's version banner off stdout (one JSON line only): it is printed at NCCL_DEBUG >= VERSION
        if "VBC_NCCL_DEBUG" in os.environ:
            os.environ["NCCL_DEBUG"] = os.environ["VBC_NCCL_DEBUG"]
        else:
            os.environ.pop("NCCL_DEBUG", None)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    u = w = 4
    n_glob = N_LOCAL * world
    L_glob = n_glob // w
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
        rows_read = (A.rowval.astype(np.int64) - 1) if args.exchange == "halo" else None
        ranges = vdist.stripe_read_ranges(A, phi, pad=u - 1) if args.exchange == "halo" else None
        peer = vdist.PeerExchangeOperator(B, layout, rank, world, dev, alpha=alpha_step, rows_read=rows_read,
                                          fused_sync=args.sync_mode, stripe_ranges=ranges)
        peer.set_x(xg)
        dist.barrier()

    def step():
        if peer is not None:
            peer.step()          # multiply with the all-gather fused into its epilogue + flag kernel
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
    # so they are captured like any other work) and replayed once: the timed region then holds
    # exactly K steps with no host launch latency in between.
    use_graph = not (world > 1 and peer is None)  # NCCL collectives are left out of graph capture (plain loop)
    side = torch.cuda.Stream()
    launches_before = B.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if use_graph:
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            for _ in range(args.steps):
                step()
        flag_launches = 0 if peer is None else {0: 1, 1: 0, 2: (2 if peer.interior[1] > peer.interior[0] else 1), 3: 1}[args.sync_mode]
        gpu_launches = B.launch_count() - launches_before + args.steps * flag_launches  # + flag kernel(s) per step
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        with ClockSampler(dev) as clk:
            with torch.cuda.stream(side):
                ev0.record()
                graph.replay()
                ev1.record()
            torch.cuda.synchronize()
    else:
        with ClockSampler(dev) as clk:
            ev0.record()
            for _ in range(args.steps):
                step()
            ev1.record()
            torch.cuda.synchronize()
        gpu_launches = B.launch_count() - launches_before
    if world > 1:
        dist.barrier()
    ms_total = ev0.elapsed_time(ev1)
    if peer is not None and peer.timed_out():
        raise RuntimeError("peer flag wait timed out")
    # parity of the distributed iteration: both exchange paths must hold the same x after the run
    x_check = None
    if world > 1:
        xs = peer.x_global() if peer is not None else op.x_global()
        x_check = float(np.abs(xs).sum())
    # duration of the dominant kernel alone (no collective): one event pair per launch, so host launch
    # latency between launches does not count
    for _ in range(3):
        op.local_multiply()
    torch.cuda.synchronize()
    pairs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a, b in pairs:
        a.record()
        op.local_multiply()
        b.record()
    torch.cuda.synchronize()
    kt = sorted(a.elapsed_time(b) for a, b in pairs)
    ms_kernel = sum(kt) / len(kt)
    ms_kernel_min, ms_kernel_med = kt[0], kt[len(kt) // 2]

    # ---- e2e: the public API with host vectors (pinned), H2D + kernel + D2H every step
    xh = torch.from_numpy(layout.scatter(xg)).pin_memory()
    yh = torch.empty(N_LOCAL, dtype=torch.float64).pin_memory()
    xh_np, yh_np = xh.numpy(), yh.numpy()
    for _ in range(3):
        vb.mul_(yh_np, B.T, xh_np)
    torch.cuda.synchronize()
    e2e_steps = max(10, min(args.steps, 100))
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        vb.mul_(yh_np, B.T, xh_np)  # synchronous on return
    e2e_s = time.perf_counter() - t0
    result_checksum = float(yh_np.sum())

    # max over ranks
    if world > 1:
        t = torch.tensor([ms_total, ms_kernel, e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, ms_kernel, e2e_s = (float(v) for v in t.tolist())
        nn = torch.tensor([float(nnz_local)], device="cuda", dtype=torch.float64)
        dist.all_reduce(nn)
        nnz_total = float(nn.item())
    else:
        nnz_total = float(nnz_local)

    if rank == 0:
        ms_step = ms_total / args.steps
        value = 2.0 * nnz_total / (ms_step * 1e-3) / 1e9
        peak, peak_src = measured_peak()
        vec_bytes = 8 * (layout.padded_len if world > 1 else N_LOCAL) + 8 * N_LOCAL
        alg_bytes = adj_bytes + vec_bytes
        # N = 1: the timed region holds nothing but K launches of this kernel, so its average launch duration
        # is the step time itself; N > 1: the step also holds the exchange, so the kernel is timed on its own
        # (one event pair per launch of the same kernel without peer stores)
        ms_roof = ms_step if world == 1 else ms_kernel
        achieved = alg_bytes / (ms_roof * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("k_spmv_adj_dram_bytes_per_launch")
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(world, nnz_local),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "kernel": "k_spmv_adj<double,8,DESC_BLOCKS,false>",
                         "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": ms_roof, "kernel_ms_event_pairs_mean": ms_kernel,
                         "kernel_ms_min": ms_kernel_min, "kernel_ms_median": ms_kernel_med,
                         "frac_of_nominal_8TBs": achieved / 8000.0,
                         "reference_format_bytes": ref_bytes + vec_bytes},
            "e2e": {"value": 2.0 * nnz_total * e2e_steps / e2e_s / 1e9, "unit": UNIT,
                    "h2d_bytes_per_step": int(xh_np.nbytes), "d2h_bytes_per_step": int(yh_np.nbytes),
                    "steps": e2e_steps, "ms_per_step": 1e3 * e2e_s / e2e_steps, "result_checksum": result_checksum},
            "gpu_launches": int(gpu_launches), "clocks": clk.summary(),
            "pack_seconds": pack_s, "cost_imbalance": imbalance, "alpha": alpha_step,
            "exchange": (None if world == 1 else {"halo": "fused into the multiply kernel (NVLink peer stores); each y segment goes to the ranks whose stripes read it",
                                                  "peer": "fused into the multiply kernel (NVLink peer stores); every y segment goes to every rank (all-gather)",
                                                  "nccl": "torch.distributed all_gather_into_tensor after the multiply"}[args.exchange]),
            "exchange_mode": (args.exchange if world > 1 else None), "x_abs_sum_after_run": x_check,
            "exchange_sent_fraction": (peer.sent_fraction if peer is not None else None),
            "exchange_flag_neighbors_rank0": (peer.neighbors if peer is not None else None),
            "exchange_sync": (None if peer is None else {0: "flag kernel after the multiply", 1: f"in-kernel, stripes {list(peer.interior)} run before the wait",
                                                               2: f"split launches, stripes {list(peer.interior)} run before the wait",
                                                               3: "plain multiply into the own buffer, then one push-the-read-chunks + flag kernel"}[args.sync_mode]),
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
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sync-mode", type=int, default=0, choices=[0, 1, 2, 3],
                    help="peer/halo flag exchange: 0 = one flag kernel (signal+wait) after the multiply; 1 = (removed: same as 0); "
                         "2 = split launches [interior stripes][wait][other stripes][signal] (halo only); 3 = plain multiply kernel into the own "
                         "buffer, then one kernel that pushes the chunks other ranks read and exchanges flags (halo only; experimental)")
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
               "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.abspath(__file__),
               "--gpus", str(args.gpus), "--steps", str(args.steps), "--warmup", str(args.warmup), "--exchange", args.exchange]
        sys.exit(subprocess.call(cmd))
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
