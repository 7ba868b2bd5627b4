"""GPU tests of the row-partitioned multiply with the x exchange fused into the kernel (csrc/peer.cu, k_spmv_adj_halo):
sparsity-aware masks, neighbour-only flags, device-derived interior ranges, peer stores and the in-kernel wait / signal.

Kernels that wait for one another must never share a GPU (nothing guarantees they run at the same time), so:
  * several "ranks" on ONE GPU run in lockstep -- every rank's step is launched without in-kernel flags (barrier = 0) and
    the host synchronises between iterations; masks, interiors, staging and peer stores are exactly the product's;
  * the in-kernel wait / signal code runs on one GPU as a single kernel whose neighbour flags were preset by the host;
  * with >= 2 GPUs the same tests use one rank per device and the real flags (barrier = 3), and two real PROCESSES
    exchange CUDA IPC handles -- the path bench.py --gpus N uses (which also carries its own exchange_parity check).

Oracle: the iterate of the global operator, x_{t+1} = alpha * A' x_t, from scipy on the host CSC matrix and from the
single-GPU kernel (bit-exact: every stripe runs the same body in both)."""
import ctypes
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import vbc_b200 as vb
from conftest import ROOT
from vbc_b200 import _lib, synth
from vbc_b200 import dist as vdist

pytestmark = pytest.mark.gpu


def _fake_allgather(table):
    """allgather stand-in for one process holding every rank's contribution"""
    return lambda mine: np.stack(table)


def _run_ranks_one_process(n, S, P, steps, alpha, shift_bounds, Tv=np.float64, halo=True):
    import torch
    u = w = 4
    L = n // w
    A, pi, phi = synth.config_c2(n=n, S=S, dtype=Tv)
    Sg = A.to_scipy().astype(np.float64)
    b = (np.arange(P + 1) * L) // P
    b[1:-1] += shift_bounds  # unequal slices: the padded layout is exercised
    layout = vdist.PaddedLayout(b * w)
    Lh = _lib.lib()
    vt = _lib.VBC_F64 if Tv == np.float64 else _lib.VBC_F32
    mats, peers, slabs = [], [], []
    for r in range(P):
        Ar, _, phir = synth.config_c2(n=n, S=S, stripes=(int(b[r]), int(b[r + 1])), dtype=Tv)
        Ar = vdist.remap_rows_to_padded(Ar, layout, u)
        slabs.append(Ar)
        mats.append(vb.SparseMatrixVBC[u, w](Ar, vdist.padded_row_partition(layout, u, np.int64), phir))
        h = ctypes.c_void_p()
        _lib.check(Lh.vbc_peer_create(ctypes.byref(h), vt, layout.padded_len, r, P, 0, None))
        peers.append(h)
    ptrs = (ctypes.c_void_p * (P * 3))()
    p = ctypes.c_void_p()
    for r in range(P):
        for k in range(3):
            _lib.check(Lh.vbc_peer_buffer(peers[r], k, ctypes.byref(p)))
            ptrs[r * 3 + k] = p.value
    for r in range(P):
        _lib.check(Lh.vbc_peer_connect_local(peers[r], ptrs))
    interiors = []
    if halo:
        chunk_shift = 5
        needs = [mats[r].read_chunks(chunk_shift) for r in range(P)]
        for r in range(P):
            # the device-derived read set equals the one computed from the slab's CSC rows (a 2D block reads its whole row part)
            rr = slabs[r].rowval.astype(np.int64) - 1
            ref_need = np.zeros(len(needs[r]), dtype=np.uint8)
            ref_need[(rr // u * u) >> chunk_shift] = 1
            ref_need[(rr // u * u + u - 1) >> chunk_shift] = 1
            assert np.array_equal(needs[r], ref_need)
        for r in range(P):
            mask, nbr = vdist.halo_mask(needs[r], layout, r, P, mats[r].n, chunk_shift, _fake_allgather(needs))
            _lib.check(Lh.vbc_peer_set_mask(peers[r], mask.ctypes.data_as(ctypes.c_void_p), len(mask), chunk_shift))
            _lib.check(Lh.vbc_peer_set_neighbors(peers[r], nbr))
            i0, i1 = ctypes.c_int64(), ctypes.c_int64()
            _lib.check(Lh.vbc_peer_auto_interior(peers[r], mats[r]._h, r * layout.S, ctypes.byref(i0), ctypes.byref(i1)))
            interiors.append((i0.value, i1.value))
    x0 = synth.vector(n, 3, dtype=Tv)
    xp = layout.scatter(x0)
    rt = ctypes.CDLL("libcudart.so")
    for r in range(P):
        _lib.check(Lh.vbc_peer_buffer(peers[r], 0, ctypes.byref(p)))
        rt.cudaMemcpy(ctypes.c_void_p(p.value), ctypes.c_void_p(xp.ctypes.data), ctypes.c_size_t(xp.nbytes), 1)
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream() for _ in range(P)]
    for r in range(P):
        _lib.check(Lh.vbc_set_stream(mats[r]._h, ctypes.c_void_p(streams[r].cuda_stream)))
    x_ref = x0.astype(np.float64)
    for it in range(steps):  # lockstep: all ranks share this GPU, so no kernel may wait for another one (barrier = 0)
        for r in range(P):
            _lib.check(Lh.vbc_peer_spmv_step(peers[r], mats[r]._h, alpha, r * layout.S, 0))
        torch.cuda.synchronize()
        x_ref = alpha * (Sg.T @ x_ref)
    torch.cuda.synchronize()
    outs = []
    for r in range(P):
        cur, to = ctypes.c_int(), ctypes.c_int()
        _lib.check(Lh.vbc_peer_current(peers[r], ctypes.byref(cur)))
        assert cur.value == steps % 2
        _lib.check(Lh.vbc_peer_status(peers[r], ctypes.byref(to)))
        assert to.value == 0, f"rank {r}: a flag wait timed out"
        _lib.check(Lh.vbc_peer_buffer(peers[r], cur.value, ctypes.byref(p)))
        out = np.empty(layout.padded_len, dtype=Tv)
        rt.cudaMemcpy(ctypes.c_void_p(out.ctypes.data), ctypes.c_void_p(p.value), ctypes.c_size_t(out.nbytes), 2)
        outs.append(out)
    for h in peers:
        Lh.vbc_peer_destroy(h)
    return layout, outs, x_ref, interiors, b


@pytest.mark.parametrize("P", [2, 3])
def test_halo_exchange_ranks_on_one_gpu(P):
    """Sparsity-aware exchange (masks from the packed slabs), device-derived interiors, staged peer stores."""
    n, S, steps = 24_000, 9, 5
    layout, outs, x_ref, interiors, b = _run_ranks_one_process(n, S, P, steps, 0.05, 7)
    tol = 1e-12
    own = np.concatenate([outs[r][r * layout.S: r * layout.S + layout.lens[r]] for r in range(P)])
    assert np.allclose(own, x_ref, rtol=tol, atol=1e-300)
    for r in range(P):
        i0, i1 = interiors[r]
        Lr = int(b[r + 1] - b[r])
        assert 0 <= i0 < i1 <= Lr and (i1 - i0) > Lr // 2, (r, interiors[r])  # a banded slab is mostly interior
        assert (i0 > 0) == (r > 0) and (i1 < Lr) == (r < P - 1)                 # boundary only towards a neighbour
    # each rank's x is complete exactly where its stripes read it: one more multiply from any rank's buffer agrees
    # with the global iterate on that rank's slice (checked through the halo it received)
    A, pi, phi = synth.config_c2(n=n, S=S)
    Sg = A.to_scipy()
    for r in range(P):
        xr = layout.gather(outs[r])
        c0, c1 = int(b[r]) * 4, int(b[r + 1]) * 4
        y_loc = (Sg.T[c0:c1] @ xr)
        assert np.allclose(y_loc, (Sg.T @ x_ref)[c0:c1], rtol=1e-11, atol=1e-300), f"rank {r} holds a stale halo"


def test_full_replication_and_float32():
    """No mask: every stripe is a boundary stripe (claimed runs, everything sent everywhere) -- a fused all-gather."""
    n, S, P, steps = 16_000, 7, 2, 4
    layout, outs, x_ref, _, _ = _run_ranks_one_process(n, S, P, steps, 0.05, 5, Tv=np.float32, halo=False)
    for r in range(P):
        assert np.allclose(layout.gather(outs[r]).astype(np.float64), x_ref, rtol=2e-5, atol=0), f"rank {r}"


def test_fused_step_refuses_stripes_wider_than_its_staging_row():
    """the staged flush holds 32 columns per stripe; a host-packed W = 48 matrix must be refused, not corrupted."""
    n = 96
    W = 48
    phi = np.array([1, 49, 97], dtype=np.int64)
    pos = np.array([1, 2, 3], dtype=np.int64)
    idx = np.array([1, 2], dtype=np.int64)
    ofs = np.array([1, 49, 97], dtype=np.int64)
    val = np.arange(96, dtype=np.float64)
    B = vb.SparseMatrix1DVBC.from_packed(W, n, n, phi, pos, idx, ofs, val)
    x = np.arange(n, dtype=np.float64)
    y = vb.mul_(np.empty(n), B.T, x)  # the plain kernel takes wide stripes
    assert np.allclose(y[:48], val[:48] * x[0]) and np.allclose(y[48:], val[48:] * x[1])
    Lh = _lib.lib()
    h = ctypes.c_void_p()
    _lib.check(Lh.vbc_peer_create(ctypes.byref(h), _lib.VBC_F64, n, 0, 1, 0, None))
    rc = Lh.vbc_peer_spmv_step(h, B._h, 1.0, 0, 0)
    assert rc == _lib.VBC_ELIMIT and b"32 columns" in Lh.vbc_last_error()
    Lh.vbc_peer_destroy(h)


def test_in_kernel_wait_and_signal_with_preset_flags():
    """The flag code of the fused kernel on ONE GPU without any kernel waiting for another: rank 1 of 3; the host presets
    the neighbours' flags in this rank's flag block (so every wait passes at once) and the 'peers' are scratch buffers of
    this process -- the step must publish epoch + 1 into slot 1 of both neighbours' flag blocks and push the halo."""
    import torch
    n, u, w, P, me = 24_000, 4, 4, 3, 1
    L = n // w
    b = (np.arange(P + 1) * L) // P
    layout = vdist.PaddedLayout(b * w)
    A, pi, phi = synth.config_c2(n=n, S=9)
    Sg = A.to_scipy()
    Ar, _, phir = synth.config_c2(n=n, S=9, stripes=(int(b[me]), int(b[me + 1])))
    Ar = vdist.remap_rows_to_padded(Ar, layout, u)
    B = vb.SparseMatrixVBC[u, w](Ar, vdist.padded_row_partition(layout, u, np.int64), phir)
    Lh = _lib.lib()
    h = ctypes.c_void_p()
    _lib.check(Lh.vbc_peer_create(ctypes.byref(h), _lib.VBC_F64, layout.padded_len, me, P, 0, None))
    scratch = [[torch.zeros(layout.padded_len, dtype=torch.float64, device="cuda") for _ in range(2)] + [torch.zeros(8, dtype=torch.int64, device="cuda")] for _ in range(P)]
    ptrs = (ctypes.c_void_p * (P * 3))()
    p = ctypes.c_void_p()
    for r in range(P):
        for k in range(3):
            if r == me:
                _lib.check(Lh.vbc_peer_buffer(h, k, ctypes.byref(p)))
                ptrs[r * 3 + k] = p.value
            else:
                ptrs[r * 3 + k] = scratch[r][k].data_ptr()
    _lib.check(Lh.vbc_peer_connect_local(h, ptrs))
    # the read sets of the neighbours: what a banded slab reads = its own columns widened by the band
    need_me = B.read_chunks(5)
    needs = []
    for r in range(P):
        if r == me:
            needs.append(need_me)
        else:
            Aq, _, phiq = synth.config_c2(n=n, S=9, stripes=(int(b[r]), int(b[r + 1])))
            Aq = vdist.remap_rows_to_padded(Aq, layout, u)
            Bq = vb.SparseMatrixVBC[u, w](Aq, vdist.padded_row_partition(layout, u, np.int64), phiq)
            needs.append(Bq.read_chunks(5))
            Bq.close()
    mask, nbr = vdist.halo_mask(need_me, layout, me, P, B.n, 5, _fake_allgather(needs))
    assert nbr == 0b101
    _lib.check(Lh.vbc_peer_set_mask(h, mask.ctypes.data_as(ctypes.c_void_p), len(mask), 5))
    _lib.check(Lh.vbc_peer_set_neighbors(h, nbr))
    i0, i1 = ctypes.c_int64(), ctypes.c_int64()
    _lib.check(Lh.vbc_peer_auto_interior(h, B._h, me * layout.S, ctypes.byref(i0), ctypes.byref(i1)))
    assert 0 < i0.value < i1.value < B.L
    # x into the own current buffer; the neighbours' flags preset far ahead
    x0 = synth.vector(n, 3)
    rt = ctypes.CDLL("libcudart.so")
    _lib.check(Lh.vbc_peer_buffer(h, 0, ctypes.byref(p)))
    xp = layout.scatter(x0)
    rt.cudaMemcpy(ctypes.c_void_p(p.value), ctypes.c_void_p(xp.ctypes.data), ctypes.c_size_t(xp.nbytes), 1)
    _lib.check(Lh.vbc_peer_buffer(h, 2, ctypes.byref(p)))
    big = np.full(8, 1 << 40, dtype=np.int64)
    rt.cudaMemcpy(ctypes.c_void_p(p.value), ctypes.c_void_p(big.ctypes.data), ctypes.c_size_t(64), 1)
    y = 0.05 * (Sg.T @ x0)
    c0 = int(b[me]) * w
    for step in range(3):
        _lib.check(Lh.vbc_peer_spmv_step(h, B._h, 0.05, me * layout.S, 3))
        torch.cuda.synchronize()
        for r in (0, 2):
            assert int(scratch[r][2][me]) == step + 1, "the step was not published to the neighbour"
        if step == 0:
            # step 1 wrote buffer 1 of every destination: the halo pushed to a neighbour is this rank's y on the columns it reads
            for r, dest in ((0, 2), (2, 1)):  # destination index i = (r - me) % P
                got = scratch[r][1].cpu().numpy()
                cols = np.flatnonzero((np.repeat(mask, 32)[: B.n] >> dest) & 1)
                assert len(cols) > 0
                assert np.allclose(got[me * layout.S + cols], y[c0 + cols], rtol=1e-12)
                untouched = np.ones(layout.padded_len, dtype=bool)
                untouched[me * layout.S + cols] = False
                assert np.all(got[untouched] == 0.0)  # nothing else was sent
    to = ctypes.c_int()
    _lib.check(Lh.vbc_peer_status(h, ctypes.byref(to)))
    assert to.value == 0
    st = (ctypes.c_uint64 * 4)()
    _lib.check(Lh.vbc_peer_wait_stats(h, st, 0))
    assert st[0] == 3 and st[2] == 0  # three steps published, no wait had to spin
    Lh.vbc_peer_destroy(h)


def _spawn(world, extra_env=None, args=()):
    env = dict(os.environ)
    env.update(extra_env or {})
    env["PYTHONPATH"] = ROOT + os.pathsep + env.get("PYTHONPATH", "")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29547", os.path.join(ROOT, "tests", "peer_worker.py"), *args]
    return subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600)


def test_two_processes_cuda_ipc_halo_exchange():
    """vbc_peer_connect with real IPC handles between two processes, the halo mode bench.py runs by default: the distributed
    iterate equals the all-gather path (RowPartitionedOperator over torch.distributed) exactly and the host CSC iterate to
    rounding.  Two GPUs: one rank per device, in-kernel flags.  One GPU: both processes share it, so the steps run in
    lockstep (no in-kernel flags, a host barrier between iterations) -- CUDA IPC works between processes on one device."""
    import torch
    share = torch.cuda.device_count() < 2
    r = _spawn(2, extra_env={"VBC_TEST_SHARE_GPU": "1"} if share else None)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    res = json.loads(line)
    assert res["ok"], res
    assert res["max_abs_diff_vs_allgather_path"] == 0.0
    assert res["max_rel_err_vs_scipy"] < 1e-12
    assert res["timed_out"] is False
    assert res["lockstep"] == share


@pytest.mark.parametrize("kind", ["2d", "1d", "2d_f32_i32"])
def test_single_process_multi_device_api(kind):
    """vbc_dist_*: one process, P ranks (distinct GPUs when the box has them -- in-kernel flags, one CUDA graph per device --
    else the same GPU listed P times, which the library runs in lockstep): cost-balanced split, slabs packed per device,
    fused exchange -- the iterate equals the host CSC iterate, for even and odd iteration counts, 1D and 2D, unequal slices."""
    import torch
    ngpu = torch.cuda.device_count()
    n, S = 24_000, 9
    tv, ti = (np.float32, np.int32) if kind.endswith("i32") else (np.float64, np.int64)
    A, pi, phi = synth.config_c2(n=n, S=S, dtype=tv, ti=ti)
    # make the cost profile uneven: thin out the left third
    keep = np.ones(A.nnz, dtype=bool)
    cp = A.colptr.astype(np.int64) - 1
    rng = np.random.default_rng(1)
    for j in range(0, n // 3, 4):
        b, e = cp[j], cp[j + 4]
        if rng.random() < 0.7:
            keep[b:e] = (A.rowval[b:e] - 1) // 4 == j // 4  # keep only the diagonal block of this stripe
    counts = np.add.reduceat(keep.astype(np.int64), cp[:-1])
    A = vb.SparseMatrixCSC(n, n, np.concatenate([[1], 1 + np.cumsum(counts)]).astype(ti), A.rowval[keep], A.nzval[keep])
    Sg = A.to_scipy().astype(np.float64)
    x0 = synth.vector(n, 2, dtype=tv)
    tol = 1e-12 if tv == np.float64 else 2e-5
    for P in (1, 2, 3):
        devices = [r % max(ngpu, 1) for r in range(P)]
        op = vdist.DistributedOperator(A, phi, None if kind == "1d" else pi, U=4, W=4, ngpus=P, devices=devices)
        assert op.stripe_bounds[0] == 0 and op.stripe_bounds[-1] == len(phi) and all(np.diff(op.stripe_bounds) > 0)
        if P > 1:
            assert max(op.cost_per_gpu) < 1.25 * (sum(op.cost_per_gpu) / P)        # balanced by the memory model ...
            assert op.stripe_bounds[1] > len(phi) // P                                # ... not by stripe count: the thin third is on rank 0
            assert all(i1 > i0 for i0, i1 in op.interior)
        op.set_x(x0)
        ref = x0.astype(np.float64)
        for iters in (4, 1, 3):
            ms = op.iterate(iters, 0.05)
            assert ms > 0
            for _ in range(iters):
                ref = 0.05 * (Sg.T @ ref)
            x = op.x().astype(np.float64)
            assert np.allclose(x, ref, rtol=tol * 10, atol=1e-300 if tv == np.float64 else 1e-12), (kind, P, iters)
        op.close()


def test_single_process_nccl_comparator():
    """VBC_EXCH_NCCL: multiply + ncclAllGather through the dlopen'ed libnccl.  Needs two distinct GPUs; on a one-GPU box the
    duplicate device must come back as a clean VBC_ENCCL error, not a crash."""
    import torch
    n = 16_000
    A, pi, phi = synth.config_c2(n=n, S=7)
    if torch.cuda.device_count() < 2:
        with pytest.raises(_lib.VBCError) as ei:
            vdist.DistributedOperator(A, phi, pi, U=4, W=4, ngpus=2, devices=[0, 0], exchange="nccl")
        assert ei.value.code == _lib.VBC_ENCCL
        return
    Sg = A.to_scipy()
    x0 = synth.vector(n, 2)
    a = vdist.DistributedOperator(A, phi, pi, U=4, W=4, ngpus=2, exchange="nccl")
    b = vdist.DistributedOperator(A, phi, pi, U=4, W=4, ngpus=2, exchange="fused")
    a.set_x(x0); b.set_x(x0)
    a.iterate(5, 0.05); b.iterate(5, 0.05)
    ref = x0.copy()
    for _ in range(5):
        ref = 0.05 * (Sg.T @ ref)
    assert np.array_equal(a.x(), b.x())
    assert np.allclose(a.x(), ref, rtol=1e-11)
    a.close(); b.close()
