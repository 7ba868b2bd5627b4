"""Worker of tests/test_gpu_peer.py::test_two_processes_cuda_ipc_halo_exchange (launched under torch.distributed.run).

Every process owns one slab of a banded operator, exchanges CUDA IPC handles over torch.distributed (gloo: works with or
without one GPU per process) and runs the fused multiply + halo exchange of libvbc; rank 0 prints one JSON line.
VBC_TEST_SHARE_GPU=1: all processes use cuda:0 (a one-GPU box)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    import vbc_b200 as vb
    from vbc_b200 import dist as vdist
    from vbc_b200 import synth

    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = 0 if os.environ.get("VBC_TEST_SHARE_GPU") else int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(dev)
    dist.init_process_group("gloo")
    n, u, w, S, steps, alpha = 40_000, 4, 4, 11, 6, 0.05
    L = n // w
    b = (np.arange(world + 1) * L) // world
    b[1:-1] += 9
    layout = vdist.PaddedLayout(b * w)
    Ar, _, phir = synth.config_c2(n=n, S=S, stripes=(int(b[rank]), int(b[rank + 1])))
    Ar = vdist.remap_rows_to_padded(Ar, layout, u)
    B = vb.SparseMatrixVBC[u, w](Ar, vdist.padded_row_partition(layout, u, np.int64), phir, device=dev)
    x0 = synth.vector(n, 5)

    peer = vdist.PeerExchangeOperator(B, layout, rank, world, dev, alpha=alpha, halo=True)
    peer.set_x(x0)
    dist.barrier()
    lockstep = bool(os.environ.get("VBC_TEST_SHARE_GPU"))
    if lockstep:  # the processes share one GPU: no kernel may wait for another one -> no in-kernel flags, the host is the barrier
        for _ in range(steps):
            peer.step(barrier=0)
            torch.cuda.synchronize()
            dist.barrier()
    else:
        for _ in range(steps):
            peer.step()
        peer.finish()
    torch.cuda.synchronize()
    timed_out = peer.timed_out()
    x_peer = peer.x_global()
    stats = peer.wait_stats()

    # the unfused path: multiply, then an all-gather over torch.distributed (staged through the host for gloo)
    y = torch.zeros(layout.S, dtype=torch.float64, device="cuda")
    x = torch.from_numpy(layout.scatter(x0)).cuda()
    for _ in range(steps):
        vb.mul_(y[: B.n], B.T, x, alpha, False)
        parts = [torch.zeros(layout.S, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(parts, y.cpu())
        x = torch.cat(parts).cuda()
    x_ag = layout.gather(x.cpu().numpy())

    if rank == 0:
        A, _, _ = synth.config_c2(n=n, S=S)
        Sg = A.to_scipy()
        ref = x0.copy()
        for _ in range(steps):
            ref = alpha * (Sg.T @ ref)
        err = float(np.max(np.abs(x_peer - ref) / np.maximum(np.abs(ref), 1e-300)))
        d = float(np.max(np.abs(x_peer - x_ag)))
        print(json.dumps({"ok": bool(d == 0.0 and err < 1e-12 and not timed_out), "max_abs_diff_vs_allgather_path": d,
                          "max_rel_err_vs_scipy": err, "timed_out": bool(timed_out), "interior": list(peer.interior),
                          "neighbors": peer.neighbors, "sent_fraction": peer.sent_fraction, "wait_stats": stats, "lockstep": lockstep}), flush=True)
    dist.barrier()
    peer.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
