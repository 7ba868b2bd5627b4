"""CPU-only: libvbc.so loads and exports exactly the symbols include/vbc.h declares; the host
layer fails loudly (no CPU fallback) when the CUDA library is missing."""
import ctypes
import os
import re
import subprocess

import pytest

import vbc_b200 as vb
from conftest import ROOT
from vbc_b200 import _lib


def header_functions():
    src = open(os.path.join(ROOT, "include", "vbc.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vbc_[a-z_0-9]+)\s*\(", src)))


def test_header_matches_binding_list():
    assert header_functions() == sorted(_lib.SYMBOLS)


def test_julia_binding_covers_the_header():
    """Every entry point of include/vbc.h is `ccall`ed by julia/CuVBC.jl (the binding a maintainer of the reference adds; Julia is not
    in this image, so the file is checked textually), except the ones it lists as deliberately unbound."""
    jl = open(os.path.join(ROOT, "sparsematrixvbcs.jl_b200", "julia", "CuVBC.jl")).read()
    bound = set(re.findall(r"ccall\(\(:(vbc_[a-z_0-9]+), libvbc\)", jl))
    unbound = {"vbc_dp_chunk", "vbc_overlap_chunk", "vbc_gen_banded_csc", "vbc_gen_free", "vbc_peer_connect_local"}
    assert bound <= set(header_functions())
    assert set(header_functions()) - bound == unbound
    for name in unbound:
        assert name in jl  # named in the "not bound on purpose" note


def test_library_exports_every_declared_symbol():
    assert os.path.exists(vb.LIB_PATH), _lib.build_hint()
    out = subprocess.run(["nm", "-D", "--defined-only", vb.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (vbc_[a-z_0-9]+)$", out, flags=re.M))
    assert exported == set(header_functions())
    L = _lib.lib()
    for name in _lib.SYMBOLS:
        assert hasattr(L, name)
    assert L.vbc_version() >= 100
    assert isinstance(L.vbc_last_error(), bytes)


def test_library_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "--list-elf", vb.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_argument_errors_without_gpu():
    """Argument validation happens before any CUDA call, so it is testable on a CPU-only box."""
    L = _lib.lib()
    h = ctypes.c_void_p()
    import numpy as np
    one = np.array([1], dtype=np.int64)
    p = ctypes.c_void_p(one.ctypes.data)
    # ArgumentError: m must be >= 0  (SparseMatrixVBCs.jl:47)
    rc = L.vbc_pack_csc(ctypes.byref(h), _lib.VBC_F64, _lib.VBC_I64, -1, 0, 0, 4, p, p, p, None, 0, p, 0, 0)
    assert rc == _lib.VBC_EARG and b"rows" in L.vbc_last_error()
    # ArgumentError: W must be > 0  (SparseMatrixVBCs.jl:50)
    rc = L.vbc_pack_csc(ctypes.byref(h), _lib.VBC_F64, _lib.VBC_I64, 0, 0, 0, 0, p, p, p, None, 0, p, 0, 0)
    assert rc == _lib.VBC_EARG and b"W must be > 0" in L.vbc_last_error()
    # ArgumentError: U must be > 0  (SparseMatrixVBCs.jl:78) -- 2D because pi_spl != NULL
    rc = L.vbc_pack_csc(ctypes.byref(h), _lib.VBC_F64, _lib.VBC_I64, 0, 0, 0, 4, p, p, p, p, 0, p, 0, 0)
    assert rc == _lib.VBC_EARG and b"U must be > 0" in L.vbc_last_error()
    with pytest.raises(vb.ArgumentError):
        _lib.check(rc)
    assert L.vbc_spmv(None, 0, 1.0, None, 0, 0.0, None, 0, 0) == _lib.VBC_EARG
    with pytest.raises(vb.ArgumentError):
        vb.SparseMatrix1DVBC[4.0]  # "W must be an Int"  SparseMatrixVBCs.jl:49


def test_element_type_enumerants_and_host_type_rules():
    """include/vbc.h's vbc_dtype values are the ones the Python and Julia bindings pass; an unknown element type is refused
    before any CUDA call; Bool is widened to Int32 at the host boundary; other element types are a TypeError, not a reinterpretation."""
    import re
    import numpy as np
    from vbc_b200 import matrix
    hdr = open(os.path.join(ROOT, "include", "vbc.h")).read()
    m = re.search(r"enum vbc_dtype \{([^}]*)\}", hdr)
    vals = {k.strip(): int(v) for k, v in (kv.split("=") for kv in m.group(1).split(","))}
    assert vals == {"VBC_F32": _lib.VBC_F32, "VBC_F64": _lib.VBC_F64, "VBC_INT32": _lib.VBC_INT32, "VBC_INT64": _lib.VBC_INT64}
    jl = open(os.path.join(ROOT, "sparsematrixvbcs.jl_b200", "julia", "CuVBC.jl")).read()
    assert "const VBC_F32, VBC_F64, VBC_INT32, VBC_INT64 = Cint(0), Cint(1), Cint(2), Cint(3)" in jl
    L = _lib.lib()
    h = ctypes.c_void_p()
    one = np.array([1], dtype=np.int64)
    p = ctypes.c_void_p(one.ctypes.data)
    rc = L.vbc_pack_csc(ctypes.byref(h), 7, _lib.VBC_I64, 0, 0, 0, 4, p, p, p, None, 0, p, 0, 0)
    assert rc == _lib.VBC_EARG and b"vt must be" in L.vbc_last_error() and not h.value
    A = vb.SparseMatrixCSC(2, 2, np.array([1, 2, 3]), np.array([1, 2]), np.array([True, True]))
    W, was_bool = matrix._widen_bool(A)
    assert was_bool and W.nzval.dtype == np.int32 and W.nzval.tolist() == [1, 1] and W.colptr is not None
    assert matrix._widen_bool(W) == (W, False)
    assert matrix._colptr_types(W) == (_lib.VBC_INT32, _lib.VBC_I64)
    with pytest.raises(TypeError, match="Int32, Int64, Bool"):
        matrix._colptr_types(vb.SparseMatrixCSC(2, 2, np.array([1, 2, 3]), np.array([1, 2]), np.array([1, 1], dtype=np.int8)))
    assert issubclass(vb.InexactError, ValueError)


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", os.path.join(ROOT, "does_not_exist", "libvbc.so"))
    with pytest.raises(ImportError, match="no CPU fallback"):
        _lib.lib()


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under the product package may reference it."""
    pkg = os.path.join(ROOT, "sparsematrixvbcs.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".jl")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", text, flags=re.M), f
                assert "liboracle" not in text, f


def build_c_client(tmp_path):
    """gcc -std=c99 build of tests/c_abi/abi_smoke.c against include/vbc.h and libvbc.so -> (exe, env)."""
    exe = str(tmp_path / "abi_smoke")
    libdir = os.path.dirname(vb.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "c_abi", "abi_smoke.c"), "-o", exe, "-L", libdir, "-lvbc",
                    "-Wl,-rpath," + libdir, "-Wl,--allow-shlib-undefined"], check=True)
    # the CUDA runtime the library was linked against (pip-installed next to torch, not on the default loader path)
    ldd = subprocess.run(["ldd", vb.LIB_PATH], capture_output=True, text=True).stdout
    dirs = {os.path.dirname(m) for m in re.findall(r"=> (\S+libcudart\S*)", ldd)}
    try:
        import nvidia.cuda_runtime
        dirs.add(os.path.join(os.path.dirname(nvidia.cuda_runtime.__file__), "lib"))
    except ImportError:
        pass
    dirs.add("/usr/local/cuda/lib64")
    env = dict(os.environ, LD_LIBRARY_PATH=":".join(sorted(dirs) + [os.environ.get("LD_LIBRARY_PATH", "")]))
    return exe, env


def test_header_is_c99_and_a_plain_c_client_links(tmp_path):
    """The drop-in boundary is a C ABI: a C99 translation unit includes the header, links the library and gets the
    documented status codes from argument validation (no GPU needed)."""
    exe, env = build_c_client(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True, env=env)
    assert out.returncode == 0 and "abi_smoke ok" in out.stdout, out.stdout + out.stderr


@pytest.mark.gpu
def test_plain_c_client_packs_and_multiplies_on_the_gpu(tmp_path):
    exe, env = build_c_client(tmp_path)
    out = subprocess.run([exe, "gpu"], capture_output=True, text=True, env=env)
    assert out.returncode == 0 and "abi_smoke ok" in out.stdout, out.stdout + out.stderr
