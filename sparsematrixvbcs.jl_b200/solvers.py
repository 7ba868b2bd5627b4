"""Device-resident iterations around the adjoint multiply.

The reference benchmarks `mul!(y, B', x)` in isolation (bin/test_table.jl:80); in use it sits inside an
iteration x_{t+1} <- f(B' x_t).  These drivers keep every vector on the GPU, issue the multiply through
`mul_` (libvbc kernels on torch's current stream) and the handful of vector updates through torch, and
never read a scalar back inside the loop: step sizes stay 0-dim device tensors.  With `graph=True` one
whole iteration is captured in a CUDA graph and replayed, so an iteration costs one graph launch.

Convergence is looked at every `check_every` iterations (one 8-byte device-to-host read).
"""
from __future__ import annotations

import numpy as np

from .matrix import Adjoint, _CuVBC, mul_


def _operator(A):
    B = A.parent if isinstance(A, Adjoint) else A
    if not isinstance(B, _CuVBC):
        raise TypeError("expected a SparseMatrix1DVBC / SparseMatrixVBC or its adjoint")
    if A.shape[0] != A.shape[1]:
        raise ValueError("the iteration needs a square operator")
    return B


def _torch_dtype(B):
    import torch
    return torch.float64 if B.Tv == np.dtype(np.float64) else torch.float32


class _Replay:
    """Runs `body` eagerly, or captures it once (after a warm-up call on a side stream) and replays the graph."""

    def __init__(self, body, graph):
        import torch
        self.body, self.g = body, None
        if graph:
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                body()  # warm-up: lazy allocations and first-use setup happen outside the capture
            torch.cuda.current_stream().wait_stream(s)
            self.g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.g):
                body()
            self.warmup_calls = 1  # the warm-up call advanced the iteration once (the capture itself executes nothing)
        else:
            self.warmup_calls = 0

    def __call__(self):
        if self.g is not None:
            self.g.replay()
        else:
            self.body()


def power_iteration(A, x0=None, iters=100, tol=0.0, check_every=10, graph=False):
    """Dominant eigenpair of a square operator by x <- op(A) x / ||op(A) x||.  Pass `B.T` to iterate on the
    benchmarked row-block kernel.  -> (eigenvalue estimate (Rayleigh quotient), x as a device tensor, iterations run).

    With graph=True the captured body runs during warm-up too, so the returned iteration count includes those."""
    import torch
    B = _operator(A)
    n = A.shape[0]
    dt = _torch_dtype(B)
    x = torch.ones(n, dtype=dt, device="cuda") if x0 is None else torch.as_tensor(x0, dtype=dt).cuda().clone()
    x /= torch.linalg.vector_norm(x)
    y = torch.empty_like(x)
    lam = torch.zeros((), dtype=dt, device="cuda")
    dlam = torch.ones((), dtype=dt, device="cuda")

    def body():
        mul_(y, A, x)
        new = torch.dot(x, y)                       # Rayleigh quotient of the current unit vector
        dlam.copy_(torch.abs(new - lam))
        lam.copy_(new)
        torch.div(y, torch.linalg.vector_norm(y), out=x)

    step = _Replay(body, graph)
    done = step.warmup_calls
    while done < iters:
        step()
        done += 1
        if tol > 0 and done % check_every == 0:
            if float(dlam) <= tol * abs(float(lam)):
                break
    return float(lam), x, done


def cg(A, b, x0=None, iters=1000, rtol=1e-10, check_every=10, graph=False):
    """Conjugate gradients for op(A) x = b with op(A) symmetric positive definite; every product is one
    adjoint (or forward) multiply.  -> (x device tensor, iterations run, final relative residual).

    The loop body never synchronises: alpha = rs / (p . Ap) and beta = rs' / rs are 0-dim device tensors."""
    import torch
    B = _operator(A)
    dt = _torch_dtype(B)
    b = torch.as_tensor(b, dtype=dt).cuda()
    x = torch.zeros_like(b) if x0 is None else torch.as_tensor(x0, dtype=dt).cuda().clone()
    r = torch.empty_like(b)
    Ap = torch.empty_like(b)
    mul_(Ap, A, x)
    torch.sub(b, Ap, out=r)
    p = r.clone()
    rs = torch.dot(r, r)
    bnorm = float(torch.linalg.vector_norm(b))
    if bnorm == 0.0:
        return torch.zeros_like(b), 0, 0.0

    def body():
        mul_(Ap, A, p)
        alpha = rs / torch.dot(p, Ap)
        x.add_(p * alpha)
        r.sub_(Ap * alpha)
        rs_new = torch.dot(r, r)
        p.mul_(rs_new / rs).add_(r)
        rs.copy_(rs_new)

    res = float(torch.sqrt(rs)) / bnorm
    if res <= rtol:
        return x, 0, res
    step = _Replay(body, graph)
    done = step.warmup_calls
    while done < iters:
        step()
        done += 1
        if done % check_every == 0:
            res = float(torch.sqrt(rs)) / bnorm
            if res <= rtol:
                break
    res = float(torch.sqrt(rs)) / bnorm
    return x, done, res
