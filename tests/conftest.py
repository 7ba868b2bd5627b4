import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # built artefacts are git-ignored: a fresh checkout has no libvbc.so / liboracle_vbc.so yet
    lib = os.path.join(ROOT, "sparsematrixvbcs.jl_b200", "libvbc.so")
    if not os.path.exists(lib):
        import __graft_entry__
        __graft_entry__.build()


def pytest_collection_modifyitems(config, items):
    # GPU tests are skipped (not failed) where no device is visible, e.g. in the build container.
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


SIZES = [1, 2, 3, 4, 5, 7, 8, 9, 15, 16, 17]  # test/runtests.jl:14


def load_fixtures():
    """The six literal matrices of the reference's test/matrices.jl:4-9 (tests/golden/make_golden.py)."""
    import vbc_b200 as vb
    fx = np.load(os.path.join(ROOT, "tests", "golden", "fixtures.npz"))
    out = {}
    for name in fx["names"]:
        name = str(name)
        out[name] = vb.SparseMatrixCSC(int(fx[f"{name}/m"]), int(fx[f"{name}/n"]), fx[f"{name}/colptr"],
                                       fx[f"{name}/rowval"], fx[f"{name}/nzval"])
    return out


def sprand(m, n, density, rng, kind="f64"):
    """Stand-in for Julia's sprand(m, n, 0.2) / sprand(Bool, ...) / sprand(Int32, ...) of runtests.jl:14-16
    (Julia's RNG stream is not reproducible outside Julia; only the distribution shape is)."""
    import scipy.sparse as sp
    import vbc_b200 as vb
    mask = rng.random((m, n)) < density
    if kind == "f64":
        vals = rng.random((m, n))
    elif kind == "bool":
        vals = np.ones((m, n))
    else:  # Int32-like: integer-valued floats (exact in f64/f32 for the one-hot products)
        vals = rng.integers(-2 ** 20, 2 ** 20, size=(m, n)).astype(np.float64)
        vals[vals == 0] = 1.0
    return vb.SparseMatrixCSC.from_scipy(sp.csc_matrix(np.where(mask, vals, 0.0)))


def sprand_typed(m, n, density, rng, dtype, ti=np.int64):
    """sprand(Bool, m, n, p) / sprand(Int32, m, n, p) of runtests.jl:15-16 with the element type kept: Bool entries are true,
    integer entries span the whole range of the type (so products and sums wrap)."""
    import vbc_b200 as vb
    dtype = np.dtype(dtype)
    mask = rng.random((m, n)) < density
    cols, rows = np.nonzero(mask.T)  # column-major order: rows ascend inside a column
    colptr = np.concatenate([[1], 1 + np.cumsum(mask.sum(axis=0))]).astype(ti)
    if dtype == np.bool_:
        vals = np.ones(len(rows), dtype=np.bool_)
    else:
        info = np.iinfo(dtype)
        vals = rng.integers(info.min, info.max, size=len(rows), dtype=dtype, endpoint=True)
        vals[vals == 0] = 1
    return vb.SparseMatrixCSC(m, n, colptr, (rows + 1).astype(ti), vals)


@pytest.fixture(scope="session")
def fixtures():
    return load_fixtures()
