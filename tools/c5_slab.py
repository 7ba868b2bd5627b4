#!/usr/bin/env python
"""BASELINE.json configs[4] shape per rank: row-partitioned 2D-VBC SpMV, Float32 / Int32, U = W = 4, 10 blocks per
stripe, n_local columns per GPU (full config: n = 50M over 8 GPUs = 6.25M per GPU), x exchanged by the fused
sparsity-aware peer stores.  Launch under torchrun; prints one JSON line on rank 0.

    python -m torch.distributed.run --nproc-per-node N tools/c5_slab.py [n_local] [steps]
"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import vbc_b200 as vb
from vbc_b200 import dist as vdist, synth

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
n_local = int(sys.argv[1]) if len(sys.argv) > 1 else 6_250_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
torch.cuda.set_device(local)
os.environ["NCCL_DEBUG"] = "WARN"
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
u = w = 4
n = n_local * world
K = L = n // 4
Lloc = n_local // 4
offs = [0, 1, -1, 2, -2, 57, -57, 58, -58, 3249]
t0 = time.perf_counter()
A, pi, phi = synth.banded_blocks(K, L, u, w, offs, dtype=np.float32, ti=np.int32, stripes=(rank * Lloc, (rank + 1) * Lloc))
t_gen = time.perf_counter() - t0
layout = vdist.PaddedLayout(np.arange(world + 1, dtype=np.int64) * n_local)
t0 = time.perf_counter()
B = vb.SparseMatrixVBC[u, w](A, pi, phi, device=local)
t_pack = time.perf_counter() - t0
nnz_local = A.nnz
ref_b, adj_b, _ = B.format_bytes()
rows_read = (A.rowval.astype(np.int64) - 1) if world > 1 else None
peer = vdist.PeerExchangeOperator(B, layout, rank, world, local, alpha=1.0 / 160.0, rows_read=rows_read)
del A
xg = synth.vector(n, 1, dtype=np.float32)
peer.set_x(xg)
if world > 1:
    dist.barrier()
for _ in range(5):
    peer.step()
torch.cuda.synchronize()
side, g = torch.cuda.Stream(), torch.cuda.CUDAGraph()
with torch.cuda.graph(g, stream=side):
    for _ in range(steps):
        peer.step()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.cuda.stream(side):
    e0.record(); g.replay(); e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / steps
t = torch.tensor([ms, float(nnz_local)], device="cuda", dtype=torch.float64)
if world > 1:
    mx = t.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    sm = t.clone(); dist.all_reduce(sm)
    ms, nnz_total = float(mx[0]), float(sm[1])
else:
    nnz_total = float(nnz_local)
assert not peer.timed_out()
if rank == 0:
    print(json.dumps({"workload": "configs[4] per-rank shape: 2D-VBC Float32/Int32, U=W=4, 10 blocks/stripe", "n_gpus": world, "n_local": n_local,
                      "n_global": n, "nnz_total": nnz_total, "ms_per_step": ms, "gflops": 2.0 * nnz_total / (ms * 1e-3) / 1e9,
                      "per_gpu_algorithmic_GBps": (adj_b + 4 * (layout.padded_len + n_local)) / (ms * 1e-3) / 1e9,
                      "gen_seconds": t_gen, "pack_seconds": t_pack, "exchange": "fused peer stores, sparsity-aware",
                      "sent_fraction": peer.sent_fraction}), flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
