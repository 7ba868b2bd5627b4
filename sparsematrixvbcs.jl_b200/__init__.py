"""sparsematrixvbcs.jl_b200 -- B200-native blocked sparse multiply behind the API of SparseMatrixVBCs.jl.

Import as `vbc_b200` (the directory name contains a dot, so the repo root ships `vbc_b200.py`,
a two-line importlib shim that loads this package under that name).

Only the hot path lives here: csrc/ (CUDA kernels + the C ABI of include/vbc.h -> libvbc.so), the
host mirror of the reference's operator interface (matrix.py), the host-side inputs of the pack
kernel (partition.py), synthetic workloads (synth.py) and the row-partitioned multi-GPU driver
(dist.py).
"""
from ._lib import ArgumentError, DimensionMismatch, VBCError, LIB_PATH  # noqa: F401
from . import partition  # noqa: F401
from .partition import (AlternatingPacker, DynamicTotalChunker, EquiChunker, OverlapChunker, RandomChunker, SparseMatrixCSC,  # noqa: F401
                        SplitPartition, StrictChunker, pack_plaid, pack_stripe, permutedims)
from .matrix import (Adjoint, CuSparseMatrixCSC, CuVBC1D, CuVBC2D, InexactError, SparseMatrix1DVBC,  # noqa: F401
                     SparseMatrixVBC, TrSpMV_, adjoint, ldiv_lower_, mul_, size, trsv_analyse)
from . import costs, solvers, synth  # noqa: F401
from .costs import (model_SparseMatrix1DVBC_blocks, model_SparseMatrix1DVBC_memory,  # noqa: F401
                    model_SparseMatrix1DVBC_TrSpMV_time, model_SparseMatrixVBC_blocks,
                    model_SparseMatrixVBC_memory, model_SparseMatrixVBC_TrSpMV_time, total_value)

__all__ = [
    "SparseMatrix1DVBC", "SparseMatrixVBC", "CuVBC1D", "CuVBC2D", "CuSparseMatrixCSC", "Adjoint",
    "mul_", "TrSpMV_", "adjoint", "size", "ldiv_lower_", "trsv_analyse",
    "SparseMatrixCSC", "SplitPartition", "EquiChunker", "StrictChunker", "RandomChunker",
    "AlternatingPacker", "DynamicTotalChunker", "OverlapChunker", "permutedims", "pack_stripe", "pack_plaid",
    "DimensionMismatch", "ArgumentError", "InexactError", "VBCError", "synth", "costs", "solvers",
    "model_SparseMatrix1DVBC_blocks", "model_SparseMatrix1DVBC_memory", "model_SparseMatrix1DVBC_TrSpMV_time",
    "model_SparseMatrixVBC_blocks", "model_SparseMatrixVBC_memory", "model_SparseMatrixVBC_TrSpMV_time", "total_value",
]
