"""Host-side mirror of the reference's operator interface for the blocked multiply path, bound
to libvbc.so (CUDA, sm_100a).  Names, argument meaning and error behaviour follow the Julia API:

    Julia (reference)                                   here
    ------------------------------------------------    -------------------------------------------
    SparseMatrix1DVBC{W}(A, Φ | method)                 SparseMatrix1DVBC[W](A, Φ | method)
      constructors_1DVBC.jl:4, :9, :94
    SparseMatrixVBC{U,W}(A, Π, Φ | method)              SparseMatrixVBC[U, W](A, Π, Φ | method)
      constructors_VBC.jl:10, :15
    size(B)   SparseMatrixVBCs.jl:55, :84               B.shape / size(B)
    mul!(y, B, x, α, β)   multiply_1DVBC.jl:9, VBC :3   mul_(y, B, x, α, β)
    mul!(y, B', x, α, β)  multiply_1DVBC.jl:85, VBC :89 mul_(y, B.T, x, α, β)   (B.T == B.H == adjoint(B))
    B * x, B' * x   multiply_1DVBC.jl:182-183           B @ x, B.T @ x
    TrSpMV!(y, A, x)  TrSpMV.jl:1                       TrSpMV_(y, A, x)

The device types are the "CuVBC" of the north star: the packed arrays live in HBM, produced by the
CSC->VBC pack kernels from the host partition; there is no CPU implementation behind them.

Vectors may be numpy arrays (host: copied in/out, call is synchronous) or torch CUDA tensors
(device: enqueued on torch's current stream).
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from ._lib import ArgumentError, DimensionMismatch, check
from .partition import SparseMatrixCSC, SplitPartition, pack_plaid, pack_stripe

# Float32 / Float64 run the tuned kernels; Int32 / Int64 (test/runtests.jl:16) the exact wrapping-arithmetic kernels of
# csrc/inttypes.cu.  Bool (test/runtests.jl:15) is stored widened to Int32 on the device and narrowed at this boundary.
_VT = {np.dtype(np.float32): _lib.VBC_F32, np.dtype(np.float64): _lib.VBC_F64,
       np.dtype(np.int32): _lib.VBC_INT32, np.dtype(np.int64): _lib.VBC_INT64}
_BOOL = np.dtype(np.bool_)
_I32 = np.dtype(np.int32)


class InexactError(ValueError):
    """Julia's InexactError: an integer result that does not fit the element type (a Bool sum above 1)."""


def _torch_dtype(tv):
    import torch
    return {np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64,
            np.dtype(np.int32): torch.int32, np.dtype(np.int64): torch.int64}[np.dtype(tv)]


def _widen_bool(A):
    """Bool matrix -> (Int32 matrix, True); anything else -> (A, False)."""
    if A.nzval.dtype == _BOOL:
        return SparseMatrixCSC(A.m, A.n, A.colptr, A.rowval, A.nzval.astype(np.int32)), True
    return A, False
_IT = {np.dtype(np.int32): _lib.VBC_I32, np.dtype(np.int64): _lib.VBC_I64}
_VT_INV = {v: k for k, v in _VT.items()}
_IT_INV = {v: k for k, v in _IT.items()}


def _vp(a):
    return None if a is None else ctypes.c_void_p(a.ctypes.data)


def _is_torch(t):
    return type(t).__module__.startswith("torch")


def _is_f64(v):
    if _is_torch(v):
        import torch
        return v.dtype == torch.float64
    return np.asarray(v).dtype == np.dtype(np.float64)


def _vec_args(handle_dtype, v, name):
    """-> (pointer, length, on_device, keepalive)"""
    if _is_torch(v):
        import torch
        want = _torch_dtype(handle_dtype)
        if v.dtype != want or v.dim() != 1 or not v.is_contiguous():
            raise TypeError(f"{name} must be a contiguous 1-D {want} tensor")
        if not v.is_cuda:
            raise TypeError(f"{name}: torch tensors must live on the GPU (pass numpy arrays for host vectors)")
        return ctypes.c_void_p(v.data_ptr()), v.numel(), 1, v
    a = np.asarray(v)
    if a.dtype != handle_dtype or a.ndim != 1 or not a.flags.c_contiguous:
        raise TypeError(f"{name} must be a contiguous 1-D {handle_dtype} array (eltype(y) == eltype(x) == Tv)")
    return _vp(a), a.shape[0], 0, a


class Adjoint:
    """`B'` / `transpose(B)` (LinearAlgebra.Adjoint / Transpose wrapper)."""

    def __init__(self, parent):
        self.parent = parent

    @property
    def shape(self):
        m, n = self.parent.shape
        return (n, m)

    @property
    def T(self):
        return self.parent

    H = T

    def __matmul__(self, x):
        return _matvec(self, x)


class _CuVBC:
    """Common part of the device matrix types (an opaque vbc_mat* plus the shape)."""

    ndim_vbc = 0

    def __init__(self):
        self._h = ctypes.c_void_p()
        self._stream_set = None

    # -- lifetime -------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            _lib.lib().vbc_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):  # finalizer -> vbc_destroy
        try:
            self.close()
        except Exception:
            pass

    # -- introspection --------------------------------------------------------------------------
    def _query(self):
        L = _lib.lib()
        m, n, K, Ls = (ctypes.c_int64() for _ in range(4))
        U, W, nd, vt, it = (ctypes.c_int() for _ in range(5))
        check(L.vbc_shape(self._h, m, n, K, Ls, U, W, nd, vt, it))
        self.m, self.n, self.K, self.L = m.value, n.value, K.value, Ls.value
        self.U, self.W = U.value, W.value
        self.Tv, self.Ti = _VT_INV[vt.value], _IT_INV[it.value]
        self.eltype = _BOOL if getattr(self, "_bool", False) else self.Tv  # eltype(A) of the host matrix (Bool is held as Int32)
        nidx, nval = ctypes.c_int64(), ctypes.c_int64()
        check(L.vbc_sizes(self._h, nidx, nval))
        self.nidx, self.nval = nidx.value, nval.value

    @property
    def shape(self):
        return (self.m, self.n)

    @property
    def T(self):
        return Adjoint(self)

    H = T

    def adjoint(self):
        return Adjoint(self)

    def download(self):
        """-> dict(pos, idx, ofs, val): the reference struct fields (1-based, Ti / Tv typed)."""
        pos = np.empty(self.L + 1, dtype=self.Ti)
        ofs = np.empty(self.L + 1, dtype=self.Ti)
        idx = np.empty(self.nidx, dtype=self.Ti)
        val = np.empty(self.nval, dtype=self.Tv)
        check(_lib.lib().vbc_download(self._h, _vp(pos), _vp(idx), _vp(ofs), _vp(val)))
        if getattr(self, "_bool", False):
            val = val.astype(np.bool_)
        return dict(pos=pos, idx=idx, ofs=ofs, val=val)

    def format_bytes(self):
        """(reference accounting of bin/test_table.jl:78/:120, bytes read by an adjoint multiply,
        bytes read by a forward multiply) -- matrix part only."""
        b = (ctypes.c_int64 * 3)()
        check(_lib.lib().vbc_format_bytes(self._h, b))
        return int(b[0]), int(b[1]), int(b[2])

    def memory_cost(self):
        """Per-stripe cost under the reference's memory model (costs.jl:10 / :140)."""
        cost = np.empty(self.L, dtype=np.int64)
        row_term = ctypes.c_int64()
        check(_lib.lib().vbc_memory_cost(self._h, _vp(cost), row_term))
        return cost, row_term.value

    def read_chunks(self, chunk_shift):
        """need[c] = 1 where an adjoint multiply gathers from x chunk c (vbc_read_chunks)."""
        n = (self.m + (1 << chunk_shift) - 1) >> chunk_shift
        need = np.zeros(max(n, 1), dtype=np.uint8)
        check(_lib.lib().vbc_read_chunks(self._h, int(chunk_shift), _vp(need), len(need)))
        return need[:n]

    def set_option(self, option, value):
        check(_lib.lib().vbc_set_option(self._h, int(option), int(value)))

    def get_option(self, option):
        v = ctypes.c_int64()
        check(_lib.lib().vbc_get_option(self._h, int(option), v))
        return v.value

    def launch_count(self):
        c = ctypes.c_int64()
        check(_lib.lib().vbc_launch_count(self._h, c))
        return c.value

    def sync(self):
        check(_lib.lib().vbc_sync(self._h))

    def _use_torch_stream(self):
        import torch
        s = torch.cuda.current_stream().cuda_stream
        if s != self._stream_set:
            check(_lib.lib().vbc_set_stream(self._h, ctypes.c_void_p(s)))
            self._stream_set = s

    def __matmul__(self, x):
        return _matvec(self, x)


def _colptr_types(A: SparseMatrixCSC):
    tv, ti = A.nzval.dtype, A.colptr.dtype
    if tv not in _VT:
        raise TypeError(f"CuVBC supports Tv in (Float32, Float64, Int32, Int64, Bool); got {tv} -- convert with A.astype(...)")
    return _VT[tv], _IT[ti]


class _ParamMeta(type):
    """`SparseMatrix1DVBC[W]` / `SparseMatrixVBC[U, W]` stand in for Julia's type parameters."""

    def __getitem__(cls, params):
        if not isinstance(params, tuple):
            params = (params,)
        if len(params) != cls._nparams:
            raise TypeError(f"{cls.__name__} takes {cls._nparams} parameter(s)")
        for p in params:
            if not isinstance(p, (int, np.integer)) or isinstance(p, bool):
                raise ArgumentError(f"{'W' if cls._nparams == 1 else 'U, W'} must be an Int")  # SparseMatrixVBCs.jl:49, :76-77
        return type(f"{cls.__name__}{list(params)}", (cls,), {"_params": tuple(int(p) for p in params)})


class SparseMatrix1DVBC(_CuVBC, metaclass=_ParamMeta):
    """Device-resident `SparseMatrix1DVBC{W,Tv,Ti}` (SparseMatrixVBCs.jl:36-53)."""

    _nparams = 1
    _params = None
    ndim_vbc = 1

    def __init__(self, A: SparseMatrixCSC, method_or_phi=None, device=0):
        super().__init__()
        if self._params is None:
            raise TypeError("write SparseMatrix1DVBC[W](A, Φ)")
        (W,) = self._params
        if method_or_phi is None:
            raise TypeError("the default partitioner (DynamicTotalChunker, constructors_1DVBC.jl:1-2) lives in "
                            "ChainPartitioners, which is not vendored: pass Φ or a chunker from .partition")
        phi = method_or_phi if isinstance(method_or_phi, SplitPartition) else pack_stripe(A, method_or_phi)
        A, self._bool = _widen_bool(A)
        vt, it = _colptr_types(A)
        phi = phi.astype(A.colptr.dtype)
        self.Phi = phi
        check(_lib.lib().vbc_pack_csc(ctypes.byref(self._h), vt, it, A.m, A.n, 0, W, _vp(A.colptr), _vp(A.rowval),
                                      _vp(A.nzval), None, 0, _vp(phi.spl), len(phi), int(device)))
        self._query()

    @classmethod
    def from_device_csc(cls, W, m, n, colptr, rowval, nzval, phi_spl, device=0):
        """Pack from a CSC matrix whose arrays (and Φ.spl) are torch CUDA tensors (vbc_pack_csc_dev): the
        path for matrices generated on the device."""
        return _from_device_csc(cls[W] if cls._params is None else cls, 0, W, m, n, colptr, rowval, nzval, None, phi_spl, device)

    @classmethod
    def from_packed(cls, W, m, n, phi_spl, pos, idx, ofs, val, device=0):
        """Adopt the fields of a host `SparseMatrix1DVBC` packed by the reference (vbc_upload)."""
        self = object.__new__(cls[W] if cls._params is None else cls)
        _CuVBC.__init__(self)
        ti, tv = np.dtype(pos.dtype), np.dtype(val.dtype)
        arrs = [np.ascontiguousarray(a, dtype=ti) for a in (phi_spl, pos, idx, ofs)]
        self._bool = tv == _BOOL
        if self._bool:
            val, tv = np.asarray(val).astype(np.int32), _I32
        val = np.ascontiguousarray(val)
        self.Phi = SplitPartition(arrs[0])
        check(_lib.lib().vbc_upload(ctypes.byref(self._h), _VT[tv], _IT[ti], m, n, 0, W, None, 0, _vp(arrs[0]),
                                    len(arrs[0]) - 1, _vp(arrs[1]), _vp(arrs[2]), _vp(arrs[3]), _vp(val), int(device)))
        self._query()
        return self


class SparseMatrixVBC(_CuVBC, metaclass=_ParamMeta):
    """Device-resident `SparseMatrixVBC{U,W,Tv,Ti}` (SparseMatrixVBCs.jl:62-82)."""

    _nparams = 2
    _params = None
    ndim_vbc = 2

    def __init__(self, A: SparseMatrixCSC, pi_or_method=None, phi=None, device=0):
        super().__init__()
        if self._params is None:
            raise TypeError("write SparseMatrixVBC[U, W](A, Π, Φ)")
        U, W = self._params
        if pi_or_method is None:
            raise TypeError("the default packer (constructors_VBC.jl:1-8) lives in ChainPartitioners, which is "
                            "not vendored: pass Π, Φ or an AlternatingPacker from .partition")
        if isinstance(pi_or_method, SplitPartition):
            if not isinstance(phi, SplitPartition):
                raise TypeError("SparseMatrixVBC[U, W](A, Π, Φ) needs both partitions")
            pi = pi_or_method
        else:
            pi, phi = pack_plaid(A, pi_or_method)
        A, self._bool = _widen_bool(A)
        vt, it = _colptr_types(A)
        pi, phi = pi.astype(A.colptr.dtype), phi.astype(A.colptr.dtype)
        self.Pi, self.Phi = pi, phi
        check(_lib.lib().vbc_pack_csc(ctypes.byref(self._h), vt, it, A.m, A.n, U, W, _vp(A.colptr), _vp(A.rowval),
                                      _vp(A.nzval), _vp(pi.spl), len(pi), _vp(phi.spl), len(phi), int(device)))
        self._query()

    @classmethod
    def from_device_csc(cls, U, W, m, n, colptr, rowval, nzval, pi_spl, phi_spl, device=0):
        """2D pack from device-resident CSC arrays and partitions (vbc_pack_csc_dev)."""
        return _from_device_csc(cls[U, W] if cls._params is None else cls, U, W, m, n, colptr, rowval, nzval, pi_spl, phi_spl, device)

    @classmethod
    def from_packed(cls, U, W, m, n, pi_spl, phi_spl, pos, idx, ofs, val, device=0):
        self = object.__new__(cls[U, W] if cls._params is None else cls)
        _CuVBC.__init__(self)
        ti, tv = np.dtype(pos.dtype), np.dtype(val.dtype)
        arrs = [np.ascontiguousarray(a, dtype=ti) for a in (pi_spl, phi_spl, pos, idx, ofs)]
        self._bool = tv == _BOOL
        if self._bool:
            val, tv = np.asarray(val).astype(np.int32), _I32
        val = np.ascontiguousarray(val)
        self.Pi, self.Phi = SplitPartition(arrs[0]), SplitPartition(arrs[1])
        check(_lib.lib().vbc_upload(ctypes.byref(self._h), _VT[tv], _IT[ti], m, n, U, W, _vp(arrs[0]), len(arrs[0]) - 1,
                                    _vp(arrs[1]), len(arrs[1]) - 1, _vp(arrs[2]), _vp(arrs[3]), _vp(arrs[4]), _vp(val),
                                    int(device)))
        self._query()
        return self


def _from_device_csc(cls, U, W, m, n, colptr, rowval, nzval, pi_spl, phi_spl, device):
    import torch
    it = {torch.int32: _lib.VBC_I32, torch.int64: _lib.VBC_I64}[colptr.dtype]
    vts = {torch.float32: _lib.VBC_F32, torch.float64: _lib.VBC_F64, torch.int32: _lib.VBC_INT32, torch.int64: _lib.VBC_INT64}
    if nzval.dtype not in vts:
        raise TypeError(f"from_device_csc supports nzval in (float32, float64, int32, int64); got {nzval.dtype} (widen Bool with nzval.int())")
    vt = vts[nzval.dtype]
    tens = [colptr, rowval, nzval, phi_spl] + ([pi_spl] if pi_spl is not None else [])
    for t in tens:
        if not (t.is_cuda and t.is_contiguous()):
            raise TypeError("from_device_csc needs contiguous CUDA tensors")
    if rowval.dtype != colptr.dtype or phi_spl.dtype != colptr.dtype or (pi_spl is not None and pi_spl.dtype != colptr.dtype):
        raise TypeError("colptr, rowval and the partitions must share one index type Ti")
    torch.cuda.synchronize()
    self = object.__new__(cls)
    _CuVBC.__init__(self)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    check(_lib.lib().vbc_pack_csc_dev(ctypes.byref(self._h), vt, it, int(m), int(n), int(U), int(W), p(colptr), p(rowval), p(nzval),
                                      (p(pi_spl) if pi_spl is not None else None), (pi_spl.numel() - 1 if pi_spl is not None else 0),
                                      p(phi_spl), phi_spl.numel() - 1, int(device)))
    self.Phi = SplitPartition(phi_spl.cpu().numpy())
    if pi_spl is not None:
        self.Pi = SplitPartition(pi_spl.cpu().numpy())
    self._query()
    return self


CuVBC1D = SparseMatrix1DVBC
CuVBC2D = SparseMatrixVBC


def size(B):
    return B.shape


def adjoint(B):
    return B.T


def _panel_args(handle_dtype, v, name):
    """2-D panel -> (pointer, rows, k, ld, layout, on_device, keepalive); layout 0 = row-major (C order),
    1 = column-major (Fortran order, what a Julia Matrix is)."""
    if _is_torch(v):
        import torch
        want = _torch_dtype(handle_dtype)
        if v.dtype != want or v.dim() != 2 or not v.is_cuda:
            raise TypeError(f"{name} must be a 2-D {want} CUDA tensor")
        rows, k = v.shape
        s0, s1 = v.stride()
        if s1 == 1 and s0 >= max(k, 1):
            return ctypes.c_void_p(v.data_ptr()), rows, k, s0, 0, 1, v
        if s0 == 1 and s1 >= max(rows, 1):
            return ctypes.c_void_p(v.data_ptr()), rows, k, s1, 1, 1, v
        raise TypeError(f"{name} must be row- or column-contiguous")
    a = np.asarray(v)
    if a.dtype != handle_dtype or a.ndim != 2:
        raise TypeError(f"{name} must be a 2-D {handle_dtype} array")
    rows, k = a.shape
    if a.flags.c_contiguous:
        return _vp(a), rows, k, max(k, 1), 0, 0, a
    if a.flags.f_contiguous:
        return _vp(a), rows, k, max(rows, 1), 1, 0, a
    raise TypeError(f"{name} must be C- or Fortran-contiguous")


def _mul_panel(Y, B, trans, X, alpha, beta):
    xp, xr, xk, ldx, xl, xdev, _kx = _panel_args(B.Tv, X, "X")
    yp, yr, yk, ldy, yl, ydev, _ky = _panel_args(B.Tv, Y, "Y")
    if xdev != ydev:
        raise TypeError("X and Y must both be host arrays or both be device tensors")
    if xl != yl:
        raise TypeError("X and Y must have the same memory layout (both row-major or both column-major)")
    need_x, need_y = (B.m, B.n) if trans else (B.n, B.m)
    if xk != yk or xr != need_x or yr != need_y:
        raise DimensionMismatch(f"op(A) is {need_y} x {need_x}, X is {xr} x {xk}, Y is {yr} x {yk}")
    if xdev:
        B._use_torch_stream()
    check(_lib.lib().vbc_spmm(B._h, 1 if trans else 0, xk, float(alpha), xp, ldx, float(beta), yp, ldy, xl, xdev))
    return Y


def mul_(y, A, x, alpha=True, beta=False):
    """`LinearAlgebra.mul!(y, A, x, α, β)`: y <- α op(A) x + β y, in place; returns y.  x and y are
    vectors, or 2-D panels of k right-hand sides (SpMM).

    BLAS semantics.  (The reference ignores α everywhere and β in the adjoint -- SURVEY.md R6 --
    but its tests and benchmarks only ever pass (true, false), where both agree.)"""
    trans = isinstance(A, Adjoint)
    B = A.parent if trans else A
    if not isinstance(B, _CuVBC):
        raise TypeError("mul_ expects a SparseMatrix1DVBC / SparseMatrixVBC or its adjoint")
    if getattr(x, "ndim", 1) == 2:
        return _mul_panel(y, B, trans, x, alpha, beta)
    if getattr(B, "_bool", False) and not _is_torch(y) and np.asarray(y).dtype == _BOOL:
        # Bool matrix, Bool vectors (test/runtests.jl:15, :27-53): computed in Int32 on the device; a result outside {0, 1} is the
        # InexactError Julia raises when it stores an Int sum into a Bool vector
        yi = np.asarray(y).astype(np.int32)
        mul_(yi, A, np.ascontiguousarray(np.asarray(x), dtype=np.int32), int(alpha), int(beta))
        if yi.size and (yi.min() < 0 or yi.max() > 1):
            raise InexactError(f"Bool({int(yi.max() if yi.max() > 1 else yi.min())})")
        y[...] = yi.astype(np.bool_)
        return y
    if B.Tv == np.dtype(np.float32) and _is_f64(y):
        # eltype(y) wider than Tv: values and x are converted to eltype(y) before multiplying
        # (multiply_1DVBC.jl:23/27, :102) -- Float64 accumulation over the Float32 matrix
        f64 = np.dtype(np.float64)
        if not _is_f64(x):
            x = x.double() if _is_torch(x) else np.asarray(x, dtype=f64)
        xp, xlen, xdev, _kx = _vec_args(f64, x, "x")
        yp, ylen, ydev, _ky = _vec_args(f64, y, "y")
        if xdev != ydev:
            raise TypeError("x and y must both be host arrays or both be device tensors")
        if xdev:
            B._use_torch_stream()
        check(_lib.lib().vbc_spmv_mixed(B._h, 1 if trans else 0, float(alpha), xp, xlen, float(beta), yp, ylen, _lib.VBC_F64, xdev))
        return y
    xp, xlen, xdev, _kx = _vec_args(B.Tv, x, "x")
    yp, ylen, ydev, _ky = _vec_args(B.Tv, y, "y")
    if xdev != ydev:
        raise TypeError("x and y must both be host arrays or both be device tensors")
    if xdev:
        B._use_torch_stream()
    check(_lib.lib().vbc_spmv(B._h, 1 if trans else 0, float(alpha), xp, xlen, float(beta), yp, ylen, xdev))
    return y


def _matvec(A, x):
    """`A * x` (multiply_1DVBC.jl:182-183): allocates the result, calls mul!(…, true, false)."""
    rows = A.shape[0]
    if _is_torch(x):
        import torch
        shape = (rows,) if x.dim() == 1 else (rows, x.shape[1])
        y = torch.empty(shape, dtype=x.dtype, device=x.device)
        if x.dim() == 2 and x.stride(0) == 1 and x.shape[1] > 1:  # column-major in, column-major out
            y = torch.empty((x.shape[1], rows), dtype=x.dtype, device=x.device).t()
    else:
        x = np.asarray(x)
        shape = (rows,) if x.ndim == 1 else (rows, x.shape[1])
        y = np.empty(shape, dtype=x.dtype, order="F" if (x.ndim == 2 and x.flags.f_contiguous and not x.flags.c_contiguous) else "C")
    return mul_(y, A, x, True, False)


def trsv_analyse(A):
    """Level schedule for `ldiv_lower_` on `A` (an Adjoint of a device VBC matrix); returns #levels."""
    if not isinstance(A, Adjoint) or not isinstance(A.parent, _CuVBC):
        raise TypeError("the triangular solve works on the adjoint orientation: pass B.T")
    n = ctypes.c_int()
    check(_lib.lib().vbc_trsv_analyse(A.parent._h, ctypes.byref(n)))
    return n.value


def ldiv_lower_(x, A, b):
    """x <- LowerTriangular(B') \\ b  with A = B.T.  EXTENSION (no reference counterpart; the reference's
    "TrSpMV" is the transposed multiply): blocked, level-scheduled forward substitution on the row
    blocks of B'.  Entries of B' above its diagonal are ignored."""
    if not isinstance(A, Adjoint) or not isinstance(A.parent, _CuVBC):
        raise TypeError("the triangular solve works on the adjoint orientation: pass B.T")
    B = A.parent
    bp, blen, bdev, _kb = _vec_args(B.Tv, b, "b")
    xp, xlen, xdev, _kx = _vec_args(B.Tv, x, "x")
    if bdev != xdev:
        raise TypeError("b and x must both be host arrays or both be device tensors")
    if blen != xlen:
        raise DimensionMismatch(f"b has {blen} entries, x has {xlen}")
    if xdev:
        B._use_torch_stream()
    check(_lib.lib().vbc_trsv_lower(B._h, bp, xp, xlen, xdev))
    return x


class CuSparseMatrixCSC:
    """Device copy of a `SparseMatrixCSC` for the `TrSpMV!` comparator."""

    def __init__(self, A: SparseMatrixCSC, device=0):
        self._h = ctypes.c_void_p()
        A, self._bool = _widen_bool(A)
        vt, it = _colptr_types(A)
        self.m, self.n, self.Tv = A.m, A.n, A.nzval.dtype
        self._stream_set = None
        check(_lib.lib().vbc_csc_upload(ctypes.byref(self._h), vt, it, A.m, A.n, _vp(A.colptr), _vp(A.rowval),
                                        _vp(A.nzval), int(device)))

    @property
    def shape(self):
        return (self.m, self.n)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            _lib.lib().vbc_csc_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def TrSpMV_(y, A, x):
    """`TrSpMV!(y, A::SparseMatrixCSC, x)`: y = A' x (TrSpMV.jl:1-20).  A may be a host
    SparseMatrixCSC (uploaded for the call) or a CuSparseMatrixCSC."""
    own = None
    if isinstance(A, SparseMatrixCSC):
        own = A = CuSparseMatrixCSC(A)
    try:
        if A._bool and not _is_torch(y) and np.asarray(y).dtype == _BOOL:
            yi = np.zeros(len(y), dtype=np.int32)
            TrSpMV_(yi, A, np.ascontiguousarray(np.asarray(x), dtype=np.int32))
            if yi.size and (yi.min() < 0 or yi.max() > 1):
                raise InexactError(f"Bool({int(yi.max() if yi.max() > 1 else yi.min())})")
            y[...] = yi.astype(np.bool_)
            return y
        xp, xlen, xdev, _kx = _vec_args(A.Tv, x, "x")
        yp, ylen, ydev, _ky = _vec_args(A.Tv, y, "y")
        if xdev != ydev:
            raise TypeError("x and y must both be host arrays or both be device tensors")
        if xdev:
            import torch
            s = torch.cuda.current_stream().cuda_stream
            if s != A._stream_set:
                check(_lib.lib().vbc_csc_set_stream(A._h, ctypes.c_void_p(s)))
                A._stream_set = s
        check(_lib.lib().vbc_csc_trspmv(A._h, xp, xlen, yp, ylen, xdev))
    finally:
        if own is not None:
            own.close()
    return y
