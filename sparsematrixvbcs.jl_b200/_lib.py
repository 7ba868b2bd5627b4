"""ctypes binding of libvbc.so (include/vbc.h).  There is no CPU fallback: if the CUDA library
is missing or fails to load, importing the device types raises."""
from __future__ import annotations

import ctypes
import os

_DIR = os.path.dirname(os.path.abspath(__file__))
# VBC_LIBRARY: another build of the same library (kernel-variant experiments, tools/build_variants.sh); default = the in-tree one
LIB_PATH = os.environ.get("VBC_LIBRARY") or os.path.join(_DIR, "libvbc.so")

VBC_F32, VBC_F64, VBC_INT32, VBC_INT64 = 0, 1, 2, 3
VBC_I32, VBC_I64 = 0, 1
VBC_OK, VBC_EDIM, VBC_EARG, VBC_ELIMIT, VBC_ECUDA, VBC_ENCCL, VBC_ENOMEM = range(7)
OPT_ADJ_GROUP, OPT_FWD_GROUP, OPT_GRID_MULT, OPT_PARITY_MODE, OPT_FWD_MODE, OPT_SPMM_SIMT, OPT_E2E_PIPELINE, OPT_E2E_UPLOAD_ELEMS, OPT_E2E_GRAPH = 1, 2, 3, 4, 5, 6, 7, 8, 9

# every symbol include/vbc.h declares (tests/test_abi_symbols.py checks header <-> this list <-> .so)
SYMBOLS = [
    "vbc_last_error", "vbc_version", "vbc_device_count",
    "vbc_pack_csc", "vbc_pack_csc_dev", "vbc_upload", "vbc_destroy",
    "vbc_shape", "vbc_sizes", "vbc_download", "vbc_format_bytes", "vbc_memory_cost",
    "vbc_spmv", "vbc_spmv_mixed", "vbc_spmm", "vbc_trsv_analyse", "vbc_trsv_levels", "vbc_trsv_lower",
    "vbc_csc_upload", "vbc_csc_trspmv", "vbc_csc_destroy",
    "vbc_set_stream", "vbc_csc_set_stream", "vbc_sync", "vbc_set_option", "vbc_get_option",
    "vbc_launch_count", "vbc_dp_chunk", "vbc_overlap_chunk",
    "vbc_peer_create", "vbc_peer_connect", "vbc_peer_connect_local", "vbc_peer_buffer", "vbc_peer_current",
    "vbc_peer_spmv_step", "vbc_peer_set_mask", "vbc_peer_set_interior", "vbc_peer_auto_interior", "vbc_peer_set_neighbors", "vbc_peer_barrier", "vbc_peer_status",
    "vbc_peer_wait_stats", "vbc_peer_destroy",
    "vbc_gen_banded_csc", "vbc_gen_free", "vbc_read_chunks", "vbc_peer_get_interior",
    "vbc_dist_create", "vbc_dist_info", "vbc_dist_set_x", "vbc_dist_spmv_iter", "vbc_dist_gather_x", "vbc_dist_destroy",
]
EXCH_FUSED, EXCH_NCCL = 0, 1

IPC_HANDLE_BYTES, PEER_HANDLES, MAX_PEERS = 64, 3, 8


class VBCError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libvbc error {code}: {msg}")
        self.code = code


class DimensionMismatch(ValueError):
    """Julia's `DimensionMismatch` (multiply_1DVBC.jl:44-45 etc.)."""


class ArgumentError(ValueError):
    """Julia's `ArgumentError` (SparseMatrixVBCs.jl:45-50, :72-79)."""


_lib = None


def build_hint():
    return ("libvbc.so not found/loadable at %s -- build it with "
            "`python -c 'import __graft_entry__ as g; g.build()'` or "
            "`make -C sparsematrixvbcs.jl_b200/csrc`.  There is no CPU fallback." % LIB_PATH)


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(build_hint())
    try:
        L = ctypes.CDLL(LIB_PATH)
    except OSError as e:  # pragma: no cover
        raise ImportError(build_hint() + f" ({e})")
    c_i64, c_int, c_vp, c_dbl = ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_double
    pp = ctypes.POINTER(c_vp)
    L.vbc_last_error.restype = ctypes.c_char_p
    L.vbc_last_error.argtypes = []
    L.vbc_version.restype = c_int
    L.vbc_device_count.argtypes = [ctypes.POINTER(c_int)]
    pack_args = [pp, c_int, c_int, c_i64, c_i64, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_i64, c_int]
    L.vbc_pack_csc.argtypes = pack_args
    L.vbc_pack_csc_dev.argtypes = pack_args
    L.vbc_upload.argtypes = [pp, c_int, c_int, c_i64, c_i64, c_int, c_int, c_vp, c_i64, c_vp, c_i64,
                             c_vp, c_vp, c_vp, c_vp, c_int]
    L.vbc_destroy.argtypes = [c_vp]
    L.vbc_destroy.restype = None
    pi64, pint = ctypes.POINTER(c_i64), ctypes.POINTER(c_int)
    L.vbc_shape.argtypes = [c_vp, pi64, pi64, pi64, pi64, pint, pint, pint, pint, pint]
    L.vbc_sizes.argtypes = [c_vp, pi64, pi64]
    L.vbc_download.argtypes = [c_vp, c_vp, c_vp, c_vp, c_vp]
    L.vbc_format_bytes.argtypes = [c_vp, pi64]
    L.vbc_memory_cost.argtypes = [c_vp, c_vp, pi64]
    L.vbc_spmv.argtypes = [c_vp, c_int, c_dbl, c_vp, c_i64, c_dbl, c_vp, c_i64, c_int]
    L.vbc_spmv_mixed.argtypes = [c_vp, c_int, c_dbl, c_vp, c_i64, c_dbl, c_vp, c_i64, c_int, c_int]
    L.vbc_spmm.argtypes = [c_vp, c_int, c_i64, c_dbl, c_vp, c_i64, c_dbl, c_vp, c_i64, c_int, c_int]
    L.vbc_trsv_analyse.argtypes = [c_vp, pint]
    L.vbc_trsv_levels.argtypes = [c_vp, pint]
    L.vbc_trsv_lower.argtypes = [c_vp, c_vp, c_vp, c_i64, c_int]
    L.vbc_csc_upload.argtypes = [pp, c_int, c_int, c_i64, c_i64, c_vp, c_vp, c_vp, c_int]
    L.vbc_csc_trspmv.argtypes = [c_vp, c_vp, c_i64, c_vp, c_i64, c_int]
    L.vbc_csc_destroy.argtypes = [c_vp]
    L.vbc_csc_destroy.restype = None
    L.vbc_set_stream.argtypes = [c_vp, c_vp]
    L.vbc_csc_set_stream.argtypes = [c_vp, c_vp]
    L.vbc_sync.argtypes = [c_vp]
    L.vbc_set_option.argtypes = [c_vp, c_int, c_i64]
    L.vbc_get_option.argtypes = [c_vp, c_int, pi64]
    L.vbc_launch_count.argtypes = [c_vp, pi64]
    L.vbc_dp_chunk.argtypes = [c_i64, c_int, c_vp, c_vp, pi64]
    L.vbc_overlap_chunk.argtypes = [c_i64, c_vp, c_vp, c_dbl, c_int, c_vp, pi64]
    L.vbc_peer_create.argtypes = [pp, c_int, c_i64, c_int, c_int, c_int, c_vp]
    L.vbc_peer_connect.argtypes = [c_vp, c_vp]
    L.vbc_peer_connect_local.argtypes = [c_vp, ctypes.POINTER(c_vp)]
    L.vbc_peer_buffer.argtypes = [c_vp, c_int, pp]
    L.vbc_peer_current.argtypes = [c_vp, pint]
    L.vbc_peer_spmv_step.argtypes = [c_vp, c_vp, c_dbl, c_i64, c_int]
    L.vbc_peer_set_mask.argtypes = [c_vp, c_vp, c_i64, c_int]
    L.vbc_peer_set_interior.argtypes = [c_vp, c_i64, c_i64]
    L.vbc_peer_auto_interior.argtypes = [c_vp, c_vp, c_i64, pi64, pi64]
    L.vbc_peer_wait_stats.argtypes = [c_vp, ctypes.POINTER(ctypes.c_uint64), c_int]
    L.vbc_peer_set_neighbors.argtypes = [c_vp, ctypes.c_uint]
    L.vbc_peer_barrier.argtypes = [c_vp, c_vp, c_int]
    L.vbc_peer_status.argtypes = [c_vp, pint]
    L.vbc_peer_destroy.argtypes = [c_vp]
    L.vbc_gen_banded_csc.argtypes = [c_int, c_int, c_i64, c_i64, c_int, c_int, c_vp, c_int, c_i64, c_i64, ctypes.c_uint64, c_dbl,
                                     pp, pp, pp, pi64, c_int]
    L.vbc_gen_free.argtypes = [c_vp, c_vp, c_vp, c_int]
    L.vbc_read_chunks.argtypes = [c_vp, c_int, c_vp, c_i64]
    L.vbc_peer_get_interior.argtypes = [c_vp, pi64, pi64]
    L.vbc_dist_create.argtypes = [pp, c_int, pint, c_int, c_int, c_i64, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_i64, c_int]
    L.vbc_dist_info.argtypes = [c_vp, pint, pi64, pi64, pi64, pi64]
    L.vbc_dist_set_x.argtypes = [c_vp, c_vp]
    L.vbc_dist_spmv_iter.argtypes = [c_vp, c_int, c_dbl, ctypes.POINTER(c_dbl)]
    L.vbc_dist_gather_x.argtypes = [c_vp, c_vp]
    L.vbc_dist_destroy.argtypes = [c_vp]
    L.vbc_dist_destroy.restype = None
    L.vbc_peer_destroy.restype = None
    for name in SYMBOLS:
        f = getattr(L, name)
        if name not in ("vbc_last_error", "vbc_destroy", "vbc_csc_destroy", "vbc_peer_destroy", "vbc_dist_destroy"):
            f.restype = c_int
    _lib = L
    return L


def check(rc):
    """Map a vbc_status to the exception the reference would raise."""
    if rc == VBC_OK:
        return
    msg = lib().vbc_last_error().decode("utf-8", "replace")
    if rc == VBC_EDIM:
        raise DimensionMismatch(msg)
    if rc == VBC_EARG:
        raise ArgumentError(msg)
    if rc == VBC_ELIMIT and msg.startswith("AssertionError"):
        raise AssertionError(msg)
    if rc == VBC_ENOMEM:
        raise MemoryError(msg)
    raise VBCError(rc, msg)
