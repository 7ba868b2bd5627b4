"""Row-block partitioning of the adjoint multiply across the GPUs of one box (north_star (e)).

In the adjoint ("row block") orientation stripe l exclusively owns y[Φ.spl[l] : Φ.spl[l+1]) and
reads all of x (multiply_1DVBC.jl:172-175, :114-116), so stripes shard with no data-path
collective inside one multiply.  An iterated square operator x_{t+1} <- A' x_t needs one exchange
per iteration: every rank's y slice must reach every rank's x -- an all-gather over NVLink.

  * stripes are split into P contiguous ranges at the P-quantiles of the prefix sum of the
    reference's own memory cost model (costs.jl:10 1D, :140 2D) -- bytes streamed per rank;
  * slices have unequal lengths, so x lives in PADDED coordinates: rank r's slice occupies
    [r*S, r*S + len_r) with S = max_r len_r.  The row indices of each rank's slab are remapped
    to these coordinates on the host before the pack kernel runs (they are only gather
    indices), which makes the exchange a plain equal-size NCCL all-gather
    (torch.distributed.all_gather_into_tensor) straight from the kernel's y into everyone's x,
    with no staging copy.

One process per GPU (torchrun); torch.distributed is plumbing only -- the multiply is libvbc's.
"""
from __future__ import annotations

import numpy as np

from .partition import SparseMatrixCSC, SplitPartition


def stripe_costs(phi_widths, units_per_stripe, vals_per_stripe, tv_size, ti_size):
    """Reference memory model per stripe: 3|Ti| + units*|Ti| + values*|Tv| (costs.jl:10 with
    units = rows; costs.jl:140 with units = blocks, values = sum u*w)."""
    del phi_widths
    return 3 * ti_size + np.asarray(units_per_stripe, dtype=np.int64) * ti_size + np.asarray(vals_per_stripe, dtype=np.int64) * tv_size


def split_by_cost(cost, P):
    """Boundaries b[0..P] of P contiguous stripe ranges with (nearly) equal total cost."""
    cost = np.asarray(cost, dtype=np.int64)
    L = len(cost)
    pre = np.concatenate([[0], np.cumsum(cost)])
    total = pre[-1]
    b = [0]
    for r in range(1, P):
        target = total * r / P
        k = int(np.searchsorted(pre, target, side="left"))
        # pick the boundary closest to the target, keep ranges monotone
        if k > 0 and abs(pre[k - 1] - target) <= abs(pre[min(k, L)] - target):
            k -= 1
        b.append(min(max(k, b[-1]), L))
    b.append(L)
    return np.array(b, dtype=np.int64)


class PaddedLayout:
    """Map between global vector indices and the padded, per-rank-slice coordinates."""

    def __init__(self, col_bounds):
        self.col_bounds = np.asarray(col_bounds, dtype=np.int64)  # P+1 global element boundaries (0-based)
        self.P = len(self.col_bounds) - 1
        self.lens = np.diff(self.col_bounds)
        self.S = int(self.lens.max()) if self.P else 0
        self.padded_len = self.S * self.P

    def to_padded(self, i0):
        """global 0-based index array -> padded index array"""
        r = np.searchsorted(self.col_bounds, i0, side="right") - 1
        return r * self.S + (i0 - self.col_bounds[r])

    def scatter(self, x_global):
        xp = np.zeros(self.padded_len, dtype=x_global.dtype)
        for r in range(self.P):
            xp[r * self.S: r * self.S + self.lens[r]] = x_global[self.col_bounds[r]:self.col_bounds[r + 1]]
        return xp

    def gather(self, x_padded):
        return np.concatenate([x_padded[r * self.S: r * self.S + self.lens[r]] for r in range(self.P)])


def remap_rows_to_padded(A: SparseMatrixCSC, layout: PaddedLayout, u=1):
    """Slab CSC (global rows) -> slab CSC whose rows live in padded coordinates.  Row parts of
    height u stay contiguous because slice boundaries are multiples of the part height."""
    rows0 = A.rowval.astype(np.int64) - 1
    new_rows = layout.to_padded(rows0) + 1
    return SparseMatrixCSC(layout.padded_len, A.n, A.colptr, new_rows.astype(A.colptr.dtype), A.nzval)


def padded_row_partition(layout: PaddedLayout, u, ti):
    """Π over the padded row space: Equi(u) inside every rank slice (slice starts are part-aligned)."""
    assert layout.S % u == 0 and np.all(layout.col_bounds[:-1] % u == 0)
    return SplitPartition(np.arange(1, layout.padded_len + 2, u, dtype=ti))


def stripe_read_ranges(A: SparseMatrixCSC, phi: SplitPartition, pad=0):
    """Per stripe: smallest and largest 0-based row its columns store (-> x entries it gathers), widened by
    `pad` rows (a 2D block reads its whole row part).  Empty stripes get (big, -1)."""
    cp = A.colptr.astype(np.int64) - 1
    n = A.n
    big = np.int64(np.iinfo(np.int64).max // 4)
    nonempty = cp[1:] > cp[:-1]
    cmin = np.full(n, big, dtype=np.int64)
    cmax = np.full(n, -1, dtype=np.int64)
    cmin[nonempty] = A.rowval[cp[:-1][nonempty]].astype(np.int64) - 1
    cmax[nonempty] = A.rowval[cp[1:][nonempty] - 1].astype(np.int64) - 1
    spl = phi.spl.astype(np.int64) - 1
    L = len(spl) - 1
    rmin = np.full(L, big, dtype=np.int64)
    rmax = np.full(L, -1, dtype=np.int64)
    wide = spl[1:] > spl[:-1]
    if n > 0 and wide.any():
        starts = spl[:-1][wide]
        rmin[wide] = np.minimum.reduceat(cmin, starts)[: wide.sum()]
        rmax[wide] = np.maximum.reduceat(cmax, starts)[: wide.sum()]
        # reduceat runs to the next start; stripes are contiguous and cover all columns, so that is exact
    has = rmax >= 0
    rmin[has] -= pad
    rmax[has] += pad
    return rmin, rmax


def longest_true_run(flags):
    """[i0, i1) of the longest run of True in a boolean array ((0, 0) if none)."""
    f = np.concatenate([[False], np.asarray(flags, dtype=bool), [False]])
    d = np.diff(f.astype(np.int8))
    starts, ends = np.flatnonzero(d == 1), np.flatnonzero(d == -1)
    if len(starts) == 0:
        return 0, 0
    k = int(np.argmax(ends - starts))
    return int(starts[k]), int(ends[k])


def stripe_memory_costs(A: SparseMatrixCSC, phi: SplitPartition, pi: SplitPartition = None, tv_size=8, ti_size=8):
    """Exact per-stripe cost of the packed format under the reference's memory model (costs.jl:10 1D, :140 2D), from
    the CSC structure alone (before any packing): 3|Ti| + units*|Ti| + stored_rows*w*|Tv|, with units = distinct rows
    (1D) or distinct row parts (2D) of the stripe and stored_rows = the rows they cover."""
    from .partition import window_weight_sums
    spl = phi.spl.astype(np.int64) - 1
    widths = np.diff(spl)
    W = int(widths.max()) if len(widths) else 1
    if pi is None:
        D = window_weight_sums(A, W)
        units = rows = D[spl[1:], np.maximum(widths, 1) - 1, 0]
    else:
        heights = np.diff(pi.spl).astype(np.float64)
        D = window_weight_sums(A, W, pi, np.stack([np.ones(len(heights)), heights]))
        units = D[spl[1:], np.maximum(widths, 1) - 1, 0]
        rows = D[spl[1:], np.maximum(widths, 1) - 1, 1]
    units = np.where(widths > 0, units, 0.0)
    rows = np.where(widths > 0, rows, 0.0)
    return (3 * ti_size + units * ti_size + rows * widths * tv_size).astype(np.int64)


def distribute(A: SparseMatrixCSC, phi: SplitPartition, pi: SplitPartition = None, world=1, rank=0):
    """Row-block partition of the adjoint multiply of a SQUARE operator over `world` ranks, balanced by the
    reference's memory model: -> (local slab CSC with rows in padded coordinates, local Π (or None), local Φ,
    PaddedLayout, stripe bounds).  Rank r owns the stripes [b[r], b[r+1]) and the matching slice of the vectors.
    For 2D the rank boundaries must fall on row-part boundaries of Π (they do when Π and Φ share those split points)."""
    if A.m != A.n:
        raise ValueError("the iterated row-partitioned multiply needs a square operator")
    ti = A.colptr.dtype.type
    cost = stripe_memory_costs(A, phi, pi, A.nzval.dtype.itemsize, A.colptr.dtype.itemsize)
    b = split_by_cost(cost, world)
    spl = phi.spl.astype(np.int64) - 1
    col_bounds = spl[b]
    layout = PaddedLayout(col_bounds)
    if pi is not None:
        ps = pi.spl.astype(np.int64) - 1
        if not np.all(np.isin(col_bounds, ps)):
            raise ValueError("rank boundaries do not fall on row-part boundaries of Π")
    c0, c1 = int(col_bounds[rank]), int(col_bounds[rank + 1])
    cp = A.colptr.astype(np.int64) - 1
    lo, hi = cp[c0], cp[c1]
    slab = SparseMatrixCSC(A.m, c1 - c0, (cp[c0:c1 + 1] - lo + 1).astype(ti), A.rowval[lo:hi], A.nzval[lo:hi])
    slab = remap_rows_to_padded(slab, layout)
    phi_loc = SplitPartition((spl[b[rank]:b[rank + 1] + 1] - c0 + 1).astype(ti))
    pi_loc = None
    if pi is not None:  # Π in padded coordinates: every rank slice keeps its own parts, the padding becomes extra (empty) parts
        pieces = []
        for r in range(world):
            inside = ps[(ps >= col_bounds[r]) & (ps <= col_bounds[r + 1])] - col_bounds[r] + r * layout.S
            pieces.append(inside)
        flat = np.unique(np.concatenate(pieces + [np.array([layout.padded_len])]))
        # split padding gaps so that no part is taller than the tallest real part
        umax = int(np.diff(ps).max())
        out = [flat[0]]
        for v in flat[1:]:
            while v - out[-1] > umax:
                out.append(out[-1] + umax)
            out.append(v)
        pi_loc = SplitPartition((np.array(out, dtype=np.int64) + 1).astype(ti))
    return slab, pi_loc, phi_loc, layout, b


def _allgather_u8(arr):
    """all_gather of a uint8 numpy array over torch.distributed (any backend) -> [rank, ...]"""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(arr))
    if dist.get_backend() == "nccl":
        t = t.cuda()
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return np.stack([a.cpu().numpy() for a in out])


def halo_mask(need, layout: PaddedLayout, rank, world, n_local, chunk_shift, allgather):
    """Sparsity-aware replication plan of one rank (`vbc_peer_set_mask` / `vbc_peer_set_neighbors`).

    need[c] = 1 where this rank's stripes gather from chunk c of the padded x (`B.read_chunks(chunk_shift)`, computed
    on the device from the packed matrix); allgather(need_u8) -> [rank, chunk].
    Returns (mask, nbr): mask[c] bit i set <=> destination i (rank (rank+i) % world) gathers from column chunk c of this
    rank's slab; nbr bit r set <=> this rank sends to or receives from rank r (symmetric: every rank derives both
    directions from the same gathered table)."""
    C = 1 << chunk_shift
    need = np.ascontiguousarray(need, dtype=np.uint8)
    need_all = allgather(need)  # [rank, global chunk]
    y_offset = rank * layout.S
    nl = (n_local + C - 1) // C
    lo = (y_offset + np.arange(nl, dtype=np.int64) * C) >> chunk_shift
    hi = np.minimum(y_offset + (np.arange(nl, dtype=np.int64) + 1) * C - 1, layout.padded_len - 1) >> chunk_shift
    mask = np.ones(nl, dtype=np.uint8)  # bit 0: this rank always keeps its own slice
    for i in range(1, world):
        r = (rank + i) % world
        needed = need_all[r][lo] | need_all[r][hi]
        mask |= (needed.astype(np.uint8) << i).astype(np.uint8)
    S_ = layout.S
    nbr = 0
    for r in range(world):
        if r == rank:
            continue
        g0, g1 = (r * S_) >> chunk_shift, min((r + 1) * S_ - 1, layout.padded_len - 1) >> chunk_shift
        m0, m1 = (rank * S_) >> chunk_shift, min((rank + 1) * S_ - 1, layout.padded_len - 1) >> chunk_shift
        i_read_r = bool(need_all[rank][g0:g1 + 1].any())   # I gather from r's slice -> r sends to me
        r_reads_me = bool(need_all[r][m0:m1 + 1].any())    # r gathers from my slice -> I send to r
        if i_read_r or r_reads_me:
            nbr |= 1 << r
    return mask, nbr


class RowPartitionedOperator:
    """One rank's share of the iterated adjoint multiply x <- alpha * A' x with the x exchange done by
    a torch.distributed all-gather (NCCL on GPUs; gloo in the CPU tests).

    local_mul(y_slice, x) : fills this rank's y slice from the full (padded) x -- on a GPU
                            `lambda y, x: mul_(y, B.T, x, alpha)`, in CPU tests the oracle
    layout                : PaddedLayout of the global vector
    """

    def __init__(self, local_mul, layout: PaddedLayout, rank, world, dtype, device="cuda"):
        import torch
        self.local_mul, self.layout, self.rank, self.world = local_mul, layout, rank, world
        self.x = torch.zeros(layout.padded_len, dtype=dtype, device=device)
        self.y_pad = torch.zeros(layout.S, dtype=dtype, device=device)
        self.n_local = int(layout.lens[rank])

    def set_x(self, x_global_np):
        import torch
        self.x.copy_(torch.from_numpy(self.layout.scatter(x_global_np)))

    def local_multiply(self):
        self.local_mul(self.y_pad[: self.n_local], self.x)

    def exchange(self):
        import torch.distributed as dist
        if self.world == 1:
            self.x[: self.layout.S].copy_(self.y_pad)
            return
        try:
            dist.all_gather_into_tensor(self.x, self.y_pad)
        except (RuntimeError, NotImplementedError):  # backends without the flat variant
            dist.all_gather(list(self.x.view(self.world, self.layout.S).unbind(0)), self.y_pad)

    def step(self):
        self.local_multiply()
        self.exchange()

    def x_global(self):
        return self.layout.gather(self.x.cpu().numpy())


class PeerExchangeOperator:
    """Same iteration with the exchange FUSED into the multiply (libvbc `vbc_peer_*`): one kernel per
    iteration runs the interior stripes like the single-GPU kernel, while its boundary stripes store their
    results into the next-x buffers of the ranks that read them through CUDA-IPC peer mappings (NVLink)
    and publish a per-step flag.  x is double buffered inside libvbc; handles are exchanged once with
    torch.distributed (any backend)."""

    def __init__(self, B, layout: PaddedLayout, rank, world, device, alpha=1.0, halo=True, chunk_shift=7, sync_mode=0):
        """halo=True (every rank must agree): replication is sparsity-aware -- each rank publishes which chunks of x
        its stripes gather from (`vbc_read_chunks`, from the packed matrix on the device) and a y segment is stored
        only into the ranks that read it (`vbc_peer_set_mask`); halo=False: x is fully replicated (a fused all-gather).
        sync_mode 0: flags inside the multiply kernel (default); 1: a separate flag kernel (signal + wait)
        after every multiply -- the comparator the fused form is measured against."""
        import ctypes

        import torch
        import torch.distributed as dist

        from . import _lib
        self.B, self.layout, self.rank, self.world, self.alpha = B, layout, rank, world, float(alpha)
        self.sync_mode = int(sync_mode)
        self._h = ctypes.c_void_p()
        L = _lib.lib()
        vt = _lib.VBC_F64 if B.Tv == np.dtype(np.float64) else _lib.VBC_F32
        mine = (ctypes.c_char * (_lib.PEER_HANDLES * _lib.IPC_HANDLE_BYTES))()
        _lib.check(L.vbc_peer_create(ctypes.byref(self._h), vt, layout.padded_len, rank, world, int(device), mine))
        if world > 1:
            t = torch.tensor(list(bytes(mine)), dtype=torch.uint8)
            if dist.get_backend() == "nccl":
                t = t.cuda()
            allh = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(allh, t)
            blob = b"".join(bytes(h.cpu().tolist()) for h in allh)
            _lib.check(L.vbc_peer_connect(self._h, blob))
        self.y_offset = rank * layout.S
        self.Tv = B.Tv
        self._lib = _lib
        self.halo = False
        self.neighbors = [r for r in range(world) if r != rank]
        self.sent_fraction = 1.0
        if halo and world > 1:
            mask, nbr = halo_mask(B.read_chunks(chunk_shift), layout, rank, world, B.n, chunk_shift, _allgather_u8)
            _lib.check(L.vbc_peer_set_mask(self._h, mask.ctypes.data_as(ctypes.c_void_p), len(mask), chunk_shift))
            _lib.check(L.vbc_peer_set_neighbors(self._h, nbr))
            self.neighbors = [r for r in range(world) if (nbr >> r) & 1]
            self.halo = True
            bits = np.unpackbits(mask[:, None], axis=1).sum()
            self.sent_fraction = float(bits) / float(len(mask) * world)
            self._mask, self._chunk_shift = mask, chunk_shift
        # interior stripes (gather only from the own slice, feed only this rank): derived on the device from the packed matrix
        i0, i1 = ctypes.c_int64(), ctypes.c_int64()
        _lib.check(L.vbc_peer_auto_interior(self._h, B._h, self.y_offset, ctypes.byref(i0), ctypes.byref(i1)))
        self.interior = (int(i0.value), int(i1.value))

    def _buf_ptr(self, k):
        import ctypes
        p = ctypes.c_void_p()
        self._lib.check(self._lib.lib().vbc_peer_buffer(self._h, k, ctypes.byref(p)))
        return p.value

    def current(self):
        import ctypes
        c = ctypes.c_int()
        self._lib.check(self._lib.lib().vbc_peer_current(self._h, ctypes.byref(c)))
        return c.value

    def _as_tensor(self, k):
        """torch view of x buffer k (memory owned by libvbc)."""
        import torch

        class _Wrap:
            pass
        w = _Wrap()
        w.__cuda_array_interface__ = {"shape": (self.layout.padded_len,), "typestr": "<f8" if self.Tv == np.dtype(np.float64) else "<f4",
                                      "data": (self._buf_ptr(k), False), "version": 2}
        return torch.as_tensor(w, device="cuda")

    def set_x(self, x_global_np):
        import torch
        self._as_tensor(self.current()).copy_(torch.from_numpy(self.layout.scatter(x_global_np)))
        torch.cuda.synchronize()

    def step(self, barrier=3):
        self.B._use_torch_stream()
        L = self._lib.lib()
        if self.sync_mode == 1 and barrier == 3:
            import ctypes

            import torch
            self._lib.check(L.vbc_peer_spmv_step(self._h, self.B._h, self.alpha, self.y_offset, 0))
            self._lib.check(L.vbc_peer_barrier(self._h, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), 3))
            return
        self._lib.check(L.vbc_peer_spmv_step(self._h, self.B._h, self.alpha, self.y_offset, barrier))

    def finish(self):
        """After the last step: wait (on the current stream) until the neighbours' final halos have landed."""
        import ctypes

        import torch
        if self.world > 1 and self.sync_mode == 0:
            self._lib.check(self._lib.lib().vbc_peer_barrier(self._h, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), 2))

    def wait_stats(self, reset=False):
        """dict(steps, wait_ns, waits_that_spun, longest_wait_ns) of the in-kernel flag waits (vbc_peer_wait_stats)."""
        import ctypes
        st = (ctypes.c_uint64 * 4)()
        self._lib.check(self._lib.lib().vbc_peer_wait_stats(self._h, st, 1 if reset else 0))
        return {"steps": int(st[0]), "wait_ns": int(st[1]), "waits_that_spun": int(st[2]), "longest_wait_ns": int(st[3])}

    def timed_out(self):
        import ctypes
        c = ctypes.c_int()
        self._lib.check(self._lib.lib().vbc_peer_status(self._h, ctypes.byref(c)))
        return bool(c.value)

    def x_global(self):
        """The full vector, assembled from every rank's OWN slice (always current, with or without a mask)."""
        import torch
        import torch.distributed as dist
        torch.cuda.synchronize()
        cur = self._as_tensor(self.current())
        mine = cur[self.y_offset: self.y_offset + self.layout.S].clone()
        if self.world == 1:
            return self.layout.gather(mine.cpu().numpy())
        if dist.get_backend() != "nccl":
            mine = mine.cpu()
        parts = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(parts, mine)
        return self.layout.gather(torch.cat(parts).cpu().numpy())

    def close(self):
        if self._h.value:
            self._lib.lib().vbc_peer_destroy(self._h)
            import ctypes
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DistributedOperator:
    """`vbc_dist_*`: the iterated adjoint multiply x <- alpha * A' x of a square operator over several GPUs driven by THIS
    process (no torch.distributed): stripes split by the reference's memory cost model, one packed slab per device, the
    exchange fused into the multiply (exchange="fused") or multiply + ncclAllGather (exchange="nccl", the comparator)."""

    def __init__(self, A: SparseMatrixCSC, phi: SplitPartition, pi: SplitPartition = None, U=1, W=None, ngpus=1, devices=None, exchange="fused"):
        import ctypes

        from . import _lib
        if A.m != A.n:
            raise ValueError("the iterated row-partitioned multiply needs a square operator")
        self._lib, self._h = _lib, ctypes.c_void_p()
        ti = A.colptr.dtype
        vt = _lib.VBC_F64 if A.nzval.dtype == np.dtype(np.float64) else _lib.VBC_F32
        it = _lib.VBC_I64 if ti == np.dtype(np.int64) else _lib.VBC_I32
        phi = phi.astype(ti)
        pi = None if pi is None else pi.astype(ti)
        W = int(np.diff(phi.spl).max()) if W is None else int(W)
        dev = None if devices is None else (ctypes.c_int * ngpus)(*devices)
        vp = lambda a: None if a is None else ctypes.c_void_p(a.ctypes.data)
        _lib.check(_lib.lib().vbc_dist_create(ctypes.byref(self._h), int(ngpus), dev, vt, it, A.n, int(U), W, vp(A.colptr), vp(A.rowval), vp(A.nzval),
                                             vp(None if pi is None else pi.spl), 0 if pi is None else len(pi), vp(phi.spl), len(phi),
                                             _lib.EXCH_FUSED if exchange == "fused" else _lib.EXCH_NCCL))
        self.n, self.P, self.Tv = A.n, int(ngpus), A.nzval.dtype
        S = ctypes.c_int64()
        b = (ctypes.c_int64 * (self.P + 1))()
        c = (ctypes.c_int64 * self.P)()
        inter = (ctypes.c_int64 * (2 * self.P))()
        _lib.check(_lib.lib().vbc_dist_info(self._h, None, ctypes.byref(S), b, c, inter))
        self.slice_len, self.stripe_bounds, self.cost_per_gpu = S.value, list(b), list(c)
        self.interior = [(inter[2 * r], inter[2 * r + 1]) for r in range(self.P)]

    def set_x(self, x):
        import ctypes
        x = np.ascontiguousarray(x, dtype=self.Tv)
        assert len(x) == self.n
        self._lib.check(self._lib.lib().vbc_dist_set_x(self._h, ctypes.c_void_p(x.ctypes.data)))

    def iterate(self, iters, alpha=1.0):
        """-> device milliseconds per iteration (maximum over the devices)"""
        import ctypes
        ms = ctypes.c_double()
        self._lib.check(self._lib.lib().vbc_dist_spmv_iter(self._h, int(iters), float(alpha), ctypes.byref(ms)))
        return ms.value

    def x(self):
        import ctypes
        out = np.empty(self.n, dtype=self.Tv)
        self._lib.check(self._lib.lib().vbc_dist_gather_x(self._h, ctypes.c_void_p(out.ctypes.data)))
        return out

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            import ctypes
            self._lib.lib().vbc_dist_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
