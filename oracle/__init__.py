"""oracle -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes front end of ``oracle/liboracle_vbc.so`` (built from ``vbc_oracle.c``), the CPU
restatement of SparseMatrixVBCs.jl's pack + multiply path.  See the header of
``vbc_oracle.c`` for the file:line map and the parity status ("pinned on the reference's
fixtures + one-hot invariants; packed arrays unpinned by reference outputs: no Julia here").

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this package.  The product package never does.

All index arrays are 1-based and Ti-typed exactly as the reference stores them.
"""
from __future__ import annotations

import ctypes
import hashlib
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_DIR, "liboracle_vbc.so")
_STAMP = os.path.join(_DIR, ".liboracle_vbc.stamp")


def _cpu_tag() -> str:
    """-march=native output is only valid on the CPU it was built on: key the build on it."""
    flags = ""
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    flags = line
                    break
    except OSError:
        pass
    src = b""
    for name in ("vbc_oracle.c", "vbc_oracle_body.inc", "Makefile"):
        with open(os.path.join(_DIR, name), "rb") as f:
            src += f.read()
    return hashlib.sha256(flags.encode() + src).hexdigest()


def build(force: bool = False) -> str:
    tag = _cpu_tag()
    have = None
    if os.path.exists(_STAMP) and os.path.exists(_SO):
        with open(_STAMP) as f:
            have = f.read().strip()
    if force or have != tag:
        subprocess.run(["make", "-C", _DIR, "-B", "liboracle_vbc.so"], check=True,
                       stdout=subprocess.DEVNULL)
        with open(_STAMP, "w") as f:
            f.write(tag)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.vbc_oracle_max_threads.restype = ctypes.c_int
    return _lib


def max_threads() -> int:
    return int(lib().vbc_oracle_max_threads())


def set_static_schedule(on: bool):
    """False (default): the reference's one-stripe-per-grab self-scheduling; True: OpenMP static schedule
    (not the reference's discipline -- only for stating what the same loops reach without the shared counter)."""
    lib().vbc_oracle_set_static_schedule(ctypes.c_int(1 if on else 0))


_TV = {np.dtype(np.float64): "f64", np.dtype(np.float32): "f32", np.dtype(np.int32): "s32", np.dtype(np.int64): "s64"}
_TI = {np.dtype(np.int64): "i64", np.dtype(np.int32): "i32"}


def _fn(name, tv, ti):
    f = getattr(lib(), f"{name}_{_TV[np.dtype(tv)]}_{_TI[np.dtype(ti)]}")
    f.restype = ctypes.c_int64
    return f


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _i64(v):
    return ctypes.c_int64(int(v))


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


class OracleError(Exception):
    """code 1: w > W assertion, 2: u > U assertion, 3: DimensionMismatch, -1: out of memory."""

    def __init__(self, code):
        super().__init__({1: "AssertionError: w <= W", 2: "AssertionError: u <= U",
                          3: "DimensionMismatch", -1: "out of memory"}.get(code, str(code)))
        self.code = code


def _check(rc):
    if rc != 0:
        raise OracleError(int(rc))


@dataclass
class Host1DVBC:
    """Field-for-field image of `SparseMatrix1DVBC{W,Tv,Ti}` (SparseMatrixVBCs.jl:36-43)."""
    m: int
    n: int
    W: int
    spl: np.ndarray  # Φ.spl, L+1
    pos: np.ndarray
    idx: np.ndarray
    ofs: np.ndarray
    val: np.ndarray  # ofs[L]-1 values (no SIMD tail pad)

    @property
    def L(self):
        return len(self.spl) - 1


@dataclass
class Host2DVBC:
    """Field-for-field image of `SparseMatrixVBC{U,W,Tv,Ti}` (SparseMatrixVBCs.jl:62-70)."""
    m: int
    n: int
    U: int
    W: int
    pi_spl: np.ndarray  # Π.spl, K+1
    spl: np.ndarray  # Φ.spl, L+1
    pos: np.ndarray
    idx: np.ndarray
    ofs: np.ndarray
    val: np.ndarray

    @property
    def L(self):
        return len(self.spl) - 1

    @property
    def K(self):
        return len(self.pi_spl) - 1


def pack_1d(m, n, colptr, rowval, nzval, spl, W, strict=False) -> Host1DVBC:
    ti, tv = colptr.dtype, nzval.dtype
    colptr, rowval, nzval, spl = _c(colptr, ti), _c(rowval, ti), _c(nzval, tv), _c(spl, ti)
    L = len(spl) - 1
    pos = np.empty(L + 1, dtype=ti)
    ofs = np.empty(L + 1, dtype=ti)
    if strict:
        _check(_fn("vbc1d_strict_count", tv, ti)(_i64(n), _p(colptr), _i64(L), _p(spl), _p(pos), _p(ofs)))
    else:
        _check(_fn("vbc1d_pack_count", tv, ti)(_i64(m), _i64(n), _p(colptr), _p(rowval), _i64(L),
                                               _p(spl), _p(pos), _p(ofs)))
    idx = np.empty(int(pos[-1]) - 1, dtype=ti)
    val = np.empty(int(ofs[-1]) - 1, dtype=tv)
    if strict:
        _check(_fn("vbc1d_strict_fill", tv, ti)(_i64(n), _i64(W), _p(colptr), _p(rowval), _p(nzval),
                                                _i64(L), _p(spl), _p(pos), _p(ofs), _p(idx), _p(val)))
    else:
        _check(_fn("vbc1d_pack_fill", tv, ti)(_i64(m), _i64(n), _i64(W), _p(colptr), _p(rowval),
                                              _p(nzval), _i64(L), _p(spl), _p(pos), _p(ofs),
                                              _p(idx), _p(val)))
    return Host1DVBC(int(m), int(n), int(W), spl, pos, idx, ofs, val)


def pack_2d(m, n, colptr, rowval, nzval, pi_spl, spl, U, W) -> Host2DVBC:
    ti, tv = colptr.dtype, nzval.dtype
    colptr, rowval, nzval = _c(colptr, ti), _c(rowval, ti), _c(nzval, tv)
    pi_spl, spl = _c(pi_spl, ti), _c(spl, ti)
    K, L = len(pi_spl) - 1, len(spl) - 1
    pos = np.empty(L + 1, dtype=ti)
    ofs = np.empty(L + 1, dtype=ti)
    _check(_fn("vbc2d_pack_count", tv, ti)(_i64(m), _i64(n), _p(colptr), _p(rowval), _i64(K),
                                           _p(pi_spl), _i64(L), _p(spl), _p(pos), _p(ofs)))
    idx = np.empty(int(pos[-1]) - 1, dtype=ti)
    val = np.empty(int(ofs[-1]) - 1, dtype=tv)
    _check(_fn("vbc2d_pack_fill", tv, ti)(_i64(m), _i64(n), _i64(U), _i64(W), _p(colptr), _p(rowval),
                                          _p(nzval), _i64(K), _p(pi_spl), _i64(L), _p(spl), _p(pos),
                                          _p(ofs), _p(idx), _p(val)))
    return Host2DVBC(int(m), int(n), int(U), int(W), pi_spl, spl, pos, idx, ofs, val)


def mul(B, x, trans=False, alpha=True, beta=False, y=None, nthreads=1):
    """`mul!(y, B, x, α, β)` / `mul!(y, B', x, α, β)` exactly as the reference computes them
    (α ignored; adjoint ignores β -- SURVEY.md R6)."""
    tv, ti = B.val.dtype, B.pos.dtype
    x = _c(x, tv)
    ylen = B.n if trans else B.m
    if y is None:
        y = np.zeros(ylen, dtype=tv)
    assert y.dtype == tv and y.flags.c_contiguous
    a, b = ctypes.c_double(float(alpha)), ctypes.c_double(float(beta))
    if isinstance(B, Host1DVBC):
        args = [_i64(B.m), _i64(B.n), _i64(B.L), _p(B.spl), _p(B.pos), _p(B.idx), _p(B.ofs), _p(B.val),
                _p(y), _i64(len(y)), _p(x), _i64(len(x)), a, b]
        if trans:
            _check(_fn("vbc1d_mul_adj", tv, ti)(*args, ctypes.c_int(nthreads)))
        else:
            _check(_fn("vbc1d_mul_fwd", tv, ti)(*args))
    else:
        args = [_i64(B.m), _i64(B.n), _i64(B.K), _p(B.pi_spl), _i64(B.L), _p(B.spl), _p(B.pos),
                _p(B.idx), _p(B.ofs), _p(B.val), _p(y), _i64(len(y)), _p(x), _i64(len(x)), a, b]
        if trans:
            _check(_fn("vbc2d_mul_adj", tv, ti)(*args, ctypes.c_int(nthreads)))
        else:
            _check(_fn("vbc2d_mul_fwd", tv, ti)(*args))
    return y


def csc_trspmv(m, n, colptr, rowval, nzval, x, y=None):
    """`TrSpMV!(y, A, x)`: y = A' x on CSC arrays (TrSpMV.jl:1-20)."""
    tv, ti = nzval.dtype, colptr.dtype
    x = _c(x, tv)
    if y is None:
        y = np.zeros(n, dtype=tv)
    _check(_fn("csc_trspmv", tv, ti)(_i64(m), _i64(n), _p(_c(colptr, ti)), _p(_c(rowval, ti)),
                                     _p(_c(nzval, tv)), _p(y), _i64(len(y)), _p(x), _i64(len(x))))
    return y


def csc_spmv(m, n, colptr, rowval, nzval, x, y=None):
    """y = A x on CSC arrays (the SparseArrays side of runtests.jl:36)."""
    tv, ti = nzval.dtype, colptr.dtype
    x = _c(x, tv)
    if y is None:
        y = np.zeros(m, dtype=tv)
    _check(_fn("csc_spmv", tv, ti)(_i64(m), _i64(n), _p(_c(colptr, ti)), _p(_c(rowval, ti)),
                                   _p(_c(nzval, tv)), _p(y), _i64(len(y)), _p(x), _i64(len(x))))
    return y


def memory_cost(B):
    """Per-stripe bytes of the packed format under the reference's memory models
    (costs.jl:10 for 1D, :140 for 2D).  Returns (cost[L], row_part_term)."""
    tv, ti = B.val.dtype, B.pos.dtype
    cost = np.empty(B.L, dtype=np.int64)
    if isinstance(B, Host1DVBC):
        extra = _fn("vbc1d_memory_cost", tv, ti)(_i64(B.L), _p(B.spl), _p(B.pos), _p(cost))
    else:
        extra = _fn("vbc2d_memory_cost", tv, ti)(_i64(B.K), _i64(B.L), _p(B.spl), _p(B.pos),
                                                 _p(B.ofs), _p(cost))
    return cost, int(extra)
