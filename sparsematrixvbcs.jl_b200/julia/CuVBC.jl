# CuVBC.jl -- the reference-side binding: a device matrix type for SparseMatrixVBCs.jl whose
# constructors and `mul!` methods `ccall` libvbc.so (include/vbc.h).  Drop this file into the
# reference's src/ and `include("CuVBC.jl")` after constructors_VBC.jl (SparseMatrixVBCs.jl:90).
#
# NOT EXECUTED IN THIS REPO: the build image has no Julia toolchain.  The same C entry points are
# exercised by the Python ctypes mirror (sparsematrixvbcs.jl_b200/matrix.py) in tests/ and bench.py.
#
# Surface kept (reference file:line -> method here)
#   SparseMatrix1DVBC{W}(A, Φ)            constructors_1DVBC.jl:9   -> CuVBC{W}(A, Φ)        (device pack kernel)
#   SparseMatrixVBC{U,W}(A, Π, Φ)         constructors_VBC.jl:15    -> CuVBC{U,W}(A, Π, Φ)
#   SparseMatrix1DVBC{W}(A, method)       constructors_1DVBC.jl:4   -> CuVBC{W}(A, method)   (partitioner stays host-side)
#   SparseMatrixVBC{U,W}(A, method)       constructors_VBC.jl:10    -> CuVBC{U,W}(A, method)
#   (host-packed struct)                                             -> CuVBC(B::SparseMatrix1DVBC / SparseMatrixVBC)
#   Base.size                             SparseMatrixVBCs.jl:55,:84
#   LinearAlgebra.mul!(y, B, x, α, β)     multiply_1DVBC.jl:9, multiply_VBC.jl:3
#   LinearAlgebra.mul!(y, B', x, α, β)    multiply_1DVBC.jl:85, multiply_VBC.jl:89
#   Base.:*                               multiply_1DVBC.jl:182-183, multiply_VBC.jl:194-195
#   TrSpMV!(y, A, x)                      TrSpMV.jl:1 -> TrSpMV!(y, ::CuCSC, x)
#   model_*_memory (per-stripe bytes)     costs.jl:10, :140 -> memory_cost(::CuVBC), format_bytes(::CuVBC)
# Added for device-resident use (no reference counterpart; the reference's vectors are host Arrays):
#   CuVec{Tv}                             a device vector handle (raw CUdeviceptr + length), `mul!` methods on it enqueue only
#   CuVBCDist                             the row-partitioned iteration over several GPUs, one process (vbc_dist_*)
#   CuVBCPeer                             the same iteration with one Julia process per GPU (vbc_peer_*; handles exchanged by the caller's transport)
#   CuVBC{U,W}(m, n, ::CuVec...)          pack from a CSC matrix that is already on the device (vbc_pack_csc_dev)
#   set_option! / get_option              kernel options (VBC_OPT_*)

using LinearAlgebra
using SparseArrays
using ChainPartitioners

const libvbc = get(ENV, "LIBVBC", "libvbc.so")

const VBC_F32, VBC_F64, VBC_INT32, VBC_INT64 = Cint(0), Cint(1), Cint(2), Cint(3)
const VBC_I32, VBC_I64 = Cint(0), Cint(1)
vbc_vt(::Type{Float32}) = VBC_F32
vbc_vt(::Type{Float64}) = VBC_F64
vbc_vt(::Type{Int32}) = VBC_INT32   # integer element types (test/runtests.jl:16): wrapping arithmetic on the device, exact
vbc_vt(::Type{Int64}) = VBC_INT64
vbc_it(::Type{Int32}) = VBC_I32
vbc_it(::Type{Int64}) = VBC_I64

struct VBCError <: Exception
    code::Int
    msg::String
end

# vbc_status -> the exception the reference throws at the same place
function vbc_check(rc::Cint)
    rc == 0 && return nothing
    msg = unsafe_string(ccall((:vbc_last_error, libvbc), Cstring, ()))
    rc == 1 && throw(DimensionMismatch(msg))                      # multiply_1DVBC.jl:44-45 ...
    rc == 2 && throw(ArgumentError(msg))                          # SparseMatrixVBCs.jl:45-50
    rc == 3 && startswith(msg, "AssertionError") && throw(AssertionError(msg))  # constructors_1DVBC.jl:46
    rc == 6 && throw(OutOfMemoryError())
    throw(VBCError(rc, msg))
end

"""
    CuVBC{U, W, Tv, Ti} <: AbstractSparseMatrix{Tv, Ti}

Device-resident VBC matrix.  `U == 0` marks the 1D format (`SparseMatrix1DVBC{W}`).
"""
mutable struct CuVBC{U, W, Tv, Ti<:Integer} <: AbstractSparseMatrix{Tv, Ti}
    handle::Ptr{Cvoid}
    m::Int
    n::Int
    Π::Union{Nothing, SplitPartition{Ti}}
    Φ::SplitPartition{Ti}
    function CuVBC{U, W, Tv, Ti}(handle, m, n, Π, Φ) where {U, W, Tv, Ti}
        A = new{U, W, Tv, Ti}(handle, m, n, Π, Φ)
        finalizer(a -> (ccall((:vbc_destroy, libvbc), Cvoid, (Ptr{Cvoid},), a.handle); a.handle = C_NULL), A)
        return A
    end
end

Base.size(A::CuVBC) = (A.m, A.n)

# ---- constructors: host CSC + host partition -> device pack kernels -------------------------------
function CuVBC{W}(A::SparseMatrixCSC{Tv, Ti}, Φ::SplitPartition{Ti}; device::Integer = 0) where {W, Tv, Ti}
    h = Ref{Ptr{Cvoid}}(C_NULL)
    (m, n) = size(A)
    GC.@preserve A Φ begin
        vbc_check(ccall((:vbc_pack_csc, libvbc), Cint,
            (Ref{Ptr{Cvoid}}, Cint, Cint, Int64, Int64, Cint, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Cint),
            h, vbc_vt(Tv), vbc_it(Ti), m, n, 0, W, A.colptr, A.rowval, A.nzval, C_NULL, 0, Φ.spl, length(Φ), device))
    end
    return CuVBC{0, W, Tv, Ti}(h[], m, n, nothing, Φ)
end

function CuVBC{U, W}(A::SparseMatrixCSC{Tv, Ti}, Π::SplitPartition{Ti}, Φ::SplitPartition{Ti}; device::Integer = 0) where {U, W, Tv, Ti}
    h = Ref{Ptr{Cvoid}}(C_NULL)
    (m, n) = size(A)
    GC.@preserve A Π Φ begin
        vbc_check(ccall((:vbc_pack_csc, libvbc), Cint,
            (Ref{Ptr{Cvoid}}, Cint, Cint, Int64, Int64, Cint, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Cint),
            h, vbc_vt(Tv), vbc_it(Ti), m, n, U, W, A.colptr, A.rowval, A.nzval, Π.spl, length(Π), Φ.spl, length(Φ), device))
    end
    return CuVBC{U, W, Tv, Ti}(h[], m, n, Π, Φ)
end

# partitioner front doors: the partition is computed on the host exactly as in the reference
CuVBC{W}(A::SparseMatrixCSC, method; kw...) where {W} = CuVBC{W}(A, pack_stripe(A, method); kw...)
function CuVBC{U, W}(A::SparseMatrixCSC, method; kw...) where {U, W}
    Π, Φ = pack_plaid(A, method)
    return CuVBC{U, W}(A, convert(SplitPartition, Π), convert(SplitPartition, Φ); kw...)
end

# adopt a matrix already packed on the host by the reference
function CuVBC(B::SparseMatrix1DVBC{W, Tv, Ti}; device::Integer = 0) where {W, Tv, Ti}
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve B begin
        vbc_check(ccall((:vbc_upload, libvbc), Cint,
            (Ref{Ptr{Cvoid}}, Cint, Cint, Int64, Int64, Cint, Cint, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint),
            h, vbc_vt(Tv), vbc_it(Ti), B.m, B.n, 0, W, C_NULL, 0, B.Φ.spl, length(B.Φ), B.pos, B.idx, B.ofs, B.val, device))
    end
    return CuVBC{0, W, Tv, Ti}(h[], B.m, B.n, nothing, B.Φ)
end

function CuVBC(B::SparseMatrixVBC{U, W, Tv, Ti}; device::Integer = 0) where {U, W, Tv, Ti}
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve B begin
        vbc_check(ccall((:vbc_upload, libvbc), Cint,
            (Ref{Ptr{Cvoid}}, Cint, Cint, Int64, Int64, Cint, Cint, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint),
            h, vbc_vt(Tv), vbc_it(Ti), B.m, B.n, U, W, B.Π.spl, length(B.Π), B.Φ.spl, length(B.Φ), B.pos, B.idx, B.ofs, B.val, device))
    end
    return CuVBC{U, W, Tv, Ti}(h[], B.m, B.n, B.Π, B.Φ)
end

# download the packed arrays (bit-exact check against the host constructors)
function Base.collect(A::CuVBC{U, W, Tv, Ti}) where {U, W, Tv, Ti}
    nidx, nval = Ref{Int64}(0), Ref{Int64}(0)
    vbc_check(ccall((:vbc_sizes, libvbc), Cint, (Ptr{Cvoid}, Ref{Int64}, Ref{Int64}), A.handle, nidx, nval))
    L = length(A.Φ)
    pos, ofs = Vector{Ti}(undef, L + 1), Vector{Ti}(undef, L + 1)
    idx, val = Vector{Ti}(undef, nidx[]), Vector{Tv}(undef, nval[])
    vbc_check(ccall((:vbc_download, libvbc), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}), A.handle, pos, idx, ofs, val))
    return (pos = pos, idx = idx, ofs = ofs, val = val)
end

# ---- multiply ----------------------------------------------------------------------------------------
function _cuvbc_mul!(y::StridedVector{Tv}, A::CuVBC{U, W, Tv}, x::StridedVector{Tv}, α::Number, β::Number, trans::Bool) where {U, W, Tv}
    GC.@preserve x y begin
        vbc_check(ccall((:vbc_spmv, libvbc), Cint,
            (Ptr{Cvoid}, Cint, Cdouble, Ptr{Cvoid}, Int64, Cdouble, Ptr{Cvoid}, Int64, Cint),
            A.handle, trans, Float64(α), x, length(x), Float64(β), y, length(y), 0))
    end
    return y
end

# Bool matrices (test/runtests.jl:15) live on the device widened to Int32: `CuVBC{W}(A::SparseMatrixCSC{Bool}, Φ)` packs
# `SparseMatrixCSC{Int32}(A)`; Bool vectors are widened for the call and the Int32 result is converted back, which throws the
# same InexactError Julia raises when a sum above 1 is stored into a Bool vector.
function _cuvbc_mul!(y::StridedVector{Bool}, A::CuVBC{U, W, Int32}, x::StridedVector{Bool}, α::Number, β::Number, trans::Bool) where {U, W}
    yw = convert(Vector{Int32}, y)
    _cuvbc_mul!(yw, A, convert(Vector{Int32}, x), α, β, trans)
    y .= yw                                                     # InexactError: Bool(2)
    return y
end
CuVBC{W}(A::SparseMatrixCSC{Bool, Ti}, Φ::SplitPartition{Ti}; kw...) where {W, Ti} = CuVBC{W}(SparseMatrixCSC{Int32, Ti}(A), Φ; kw...)
CuVBC{U, W}(A::SparseMatrixCSC{Bool, Ti}, Π::SplitPartition{Ti}, Φ::SplitPartition{Ti}; kw...) where {U, W, Ti} = CuVBC{U, W}(SparseMatrixCSC{Int32, Ti}(A), Π, Φ; kw...)

# eltype(y) wider than the stored values: the reference converts values and x to eltype(y) before multiplying
# (multiply_1DVBC.jl:23/27/34, :102) -> Float64 accumulation over a Float32 matrix (csrc/mixed.cu)
function _cuvbc_mul!(y::StridedVector{Float64}, A::CuVBC{U, W, Float32}, x::StridedVector{Tx}, α::Number, β::Number, trans::Bool) where {U, W, Tx <: Union{Float32, Float64}}
    xw = Tx === Float64 ? x : convert(Vector{Float64}, x)
    GC.@preserve xw y begin
        vbc_check(ccall((:vbc_spmv_mixed, libvbc), Cint,
            (Ptr{Cvoid}, Cint, Cdouble, Ptr{Cvoid}, Int64, Cdouble, Ptr{Cvoid}, Int64, Cint, Cint),
            A.handle, trans, Float64(α), xw, length(xw), Float64(β), y, length(y), 1 #= VBC_F64 =#, 0))
    end
    return y
end

LinearAlgebra.mul!(y::StridedVector, A::CuVBC, x::StridedVector, α::Number, β::Number) =
    _cuvbc_mul!(y, A, x, α, β, false)
LinearAlgebra.mul!(y::StridedVector, adjA::Union{Adjoint{<:Any, <:CuVBC}, Transpose{<:Any, <:CuVBC}}, x::StridedVector, α::Number, β::Number) =
    _cuvbc_mul!(y, adjA.parent, x, α, β, true)

Base.:*(A::Union{CuVBC, Adjoint{<:Any, <:CuVBC}, Transpose{<:Any, <:CuVBC}}, x::StridedVector{Tx}) where {Tx} =
    (T = Base.promote_op(LinearAlgebra.matprod, eltype(A), Tx); mul!(similar(x, T, size(A, 1)), A, x, true, false))

# ---- k right-hand sides: the `*(A, B::DenseMatrix)` of multiply_1DVBC.jl:184-185 / multiply_VBC.jl:196-197, which the
# reference declares but cannot execute.  A Julia Matrix is column-major: layout = 1, ld = stride(X, 2).
function _cuvbc_mul!(Y::StridedMatrix{Tv}, A::CuVBC{U, W, Tv}, X::StridedMatrix{Tv}, α::Number, β::Number, trans::Bool) where {U, W, Tv}
    size(X, 2) == size(Y, 2) || throw(DimensionMismatch())
    GC.@preserve X Y begin
        vbc_check(ccall((:vbc_spmm, libvbc), Cint,
            (Ptr{Cvoid}, Cint, Int64, Cdouble, Ptr{Cvoid}, Int64, Cdouble, Ptr{Cvoid}, Int64, Cint, Cint),
            A.handle, trans, size(X, 2), Float64(α), X, stride(X, 2), Float64(β), Y, stride(Y, 2), 1, 0))
    end
    return Y
end
LinearAlgebra.mul!(Y::StridedMatrix, A::CuVBC, X::StridedMatrix, α::Number, β::Number) = _cuvbc_mul!(Y, A, X, α, β, false)
LinearAlgebra.mul!(Y::StridedMatrix, adjA::Union{Adjoint{<:Any, <:CuVBC}, Transpose{<:Any, <:CuVBC}}, X::StridedMatrix, α::Number, β::Number) =
    _cuvbc_mul!(Y, adjA.parent, X, α, β, true)
Base.:*(A::Union{CuVBC, Adjoint{<:Any, <:CuVBC}, Transpose{<:Any, <:CuVBC}}, X::StridedMatrix{Tx}) where {Tx} =
    (T = Base.promote_op(LinearAlgebra.matprod, eltype(A), Tx); mul!(similar(X, T, (size(A, 1), size(X, 2))), A, X, true, false))

# ---- extension: x = LowerTriangular(B') \ b (the reference has no solve; BASELINE.json north_star (d))
function ldiv_lower!(x::StridedVector{Tv}, adjA::Union{Adjoint{<:Any, <:CuVBC{U, W, Tv}}, Transpose{<:Any, <:CuVBC{U, W, Tv}}}, b::StridedVector{Tv}) where {U, W, Tv}
    length(x) == length(b) || throw(DimensionMismatch())
    GC.@preserve x b begin
        vbc_check(ccall((:vbc_trsv_lower, libvbc), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Cint),
            adjA.parent.handle, b, x, length(x), 0))
    end
    return x
end

# ---- CSC comparator ----------------------------------------------------------------------------------
mutable struct CuCSC{Tv, Ti}
    handle::Ptr{Cvoid}
    m::Int
    n::Int
    function CuCSC(A::SparseMatrixCSC{Tv, Ti}; device::Integer = 0) where {Tv, Ti}
        h = Ref{Ptr{Cvoid}}(C_NULL)
        GC.@preserve A begin
            vbc_check(ccall((:vbc_csc_upload, libvbc), Cint,
                (Ref{Ptr{Cvoid}}, Cint, Cint, Int64, Int64, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Cint),
                h, vbc_vt(Tv), vbc_it(Ti), size(A, 1), size(A, 2), A.colptr, A.rowval, A.nzval, device))
        end
        C = new{Tv, Ti}(h[], size(A, 1), size(A, 2))
        finalizer(c -> (ccall((:vbc_csc_destroy, libvbc), Cvoid, (Ptr{Cvoid},), c.handle); c.handle = C_NULL), C)
        return C
    end
end

function TrSpMV!(y::Vector{Tv}, A::CuCSC{Tv}, x::Vector{Tv}) where {Tv}
    GC.@preserve x y begin
        vbc_check(ccall((:vbc_csc_trspmv, libvbc), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Cint),
            A.handle, x, length(x), y, length(y), 0))
    end
    return y
end


# ---- device-resident vectors ---------------------------------------------------------------------------
# The reference's `mul!` takes any StridedVector on the host (multiply_1DVBC.jl:9, :85).  For iterative use the vectors
# should stay on the device: CuVec wraps a raw device pointer (from CUDA.jl's `pointer(::CuArray)`, or from cuda_malloc
# below) and the `mul!` methods on it pass on_device = 1, i.e. they only enqueue on the matrix' stream -- call `sync(A)`.
struct CuVec{Tv}
    ptr::Ptr{Cvoid}     # device address on the matrix' device
    len::Int
end
Base.length(v::CuVec) = v.len
Base.eltype(::CuVec{Tv}) where {Tv} = Tv

const libcudart = get(ENV, "LIBCUDART", "libcudart.so")
function cuda_malloc(::Type{Tv}, n::Integer) where {Tv}
    p = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:cudaMalloc, libcudart), Cint, (Ref{Ptr{Cvoid}}, Csize_t), p, n * sizeof(Tv))
    rc == 0 || throw(OutOfMemoryError())
    return CuVec{Tv}(p[], n)
end
cuda_free(v::CuVec) = ccall((:cudaFree, libcudart), Cint, (Ptr{Cvoid},), v.ptr)
function Base.copyto!(dst::CuVec{Tv}, src::Vector{Tv}) where {Tv}
    length(src) == dst.len || throw(DimensionMismatch())
    GC.@preserve src ccall((:cudaMemcpy, libcudart), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t, Cint), dst.ptr, src, sizeof(src), 1)
    return dst
end
function Base.copyto!(dst::Vector{Tv}, src::CuVec{Tv}) where {Tv}
    length(dst) == src.len || throw(DimensionMismatch())
    GC.@preserve dst ccall((:cudaMemcpy, libcudart), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Csize_t, Cint), dst, src.ptr, sizeof(dst), 2)
    return dst
end

function _cuvbc_mul!(y::CuVec{Tv}, A::CuVBC{U, W, Tv}, x::CuVec{Tv}, α::Number, β::Number, trans::Bool) where {U, W, Tv}
    vbc_check(ccall((:vbc_spmv, libvbc), Cint,
        (Ptr{Cvoid}, Cint, Cdouble, Ptr{Cvoid}, Int64, Cdouble, Ptr{Cvoid}, Int64, Cint),
        A.handle, trans, Float64(α), x.ptr, x.len, Float64(β), y.ptr, y.len, 1))   # on_device = 1: enqueued, not awaited
    return y
end
LinearAlgebra.mul!(y::CuVec, A::CuVBC, x::CuVec, α::Number = true, β::Number = false) = _cuvbc_mul!(y, A, x, α, β, false)
LinearAlgebra.mul!(y::CuVec, adjA::Union{Adjoint{<:Any, <:CuVBC}, Transpose{<:Any, <:CuVBC}}, x::CuVec, α::Number = true, β::Number = false) =
    _cuvbc_mul!(y, adjA.parent, x, α, β, true)
sync(A::CuVBC) = vbc_check(ccall((:vbc_sync, libvbc), Cint, (Ptr{Cvoid},), A.handle))
set_stream!(A::CuVBC, stream::Ptr{Cvoid}) = vbc_check(ccall((:vbc_set_stream, libvbc), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), A.handle, stream))

# ---- cost models and options ---------------------------------------------------------------------------
# model_SparseMatrix1DVBC_memory / model_SparseMatrixVBC_memory evaluated per stripe on the packed matrix (costs.jl:10, :140):
# the balance weight of the multi-GPU split.  Returns (cost per stripe, row term = K * sizeof(Ti) for 2D).
function memory_cost(A::CuVBC)
    cost = Vector{Int64}(undef, length(A.Φ))
    row_term = Ref{Int64}(0)
    vbc_check(ccall((:vbc_memory_cost, libvbc), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ref{Int64}), A.handle, cost, row_term))
    return cost, row_term[]
end
# (reference accounting of bin/test_table.jl:78/:120, bytes one adjoint multiply reads, bytes one forward multiply reads)
function format_bytes(A::CuVBC)
    b = Vector{Int64}(undef, 3)
    vbc_check(ccall((:vbc_format_bytes, libvbc), Cint, (Ptr{Cvoid}, Ptr{Int64}), A.handle, b))
    return (reference = b[1], adjoint = b[2], forward = b[3])
end
const VBC_OPT_ADJ_GROUP, VBC_OPT_FWD_GROUP, VBC_OPT_GRID_MULT, VBC_OPT_PARITY_MODE = 1, 2, 3, 4
const VBC_OPT_FWD_MODE, VBC_OPT_SPMM_SIMT, VBC_OPT_E2E_PIPELINE, VBC_OPT_E2E_UPLOAD_ELEMS, VBC_OPT_E2E_GRAPH = 5, 6, 7, 8, 9
set_option!(A::CuVBC, opt::Integer, value::Integer) =
    vbc_check(ccall((:vbc_set_option, libvbc), Cint, (Ptr{Cvoid}, Cint, Int64), A.handle, opt, value))
function get_option(A::CuVBC, opt::Integer)
    v = Ref{Int64}(0)
    vbc_check(ccall((:vbc_get_option, libvbc), Cint, (Ptr{Cvoid}, Cint, Ref{Int64}), A.handle, opt, v))
    return v[]
end
launch_count(A::CuVBC) = (c = Ref{Int64}(0); vbc_check(ccall((:vbc_launch_count, libvbc), Cint, (Ptr{Cvoid}, Ref{Int64}), A.handle, c)); c[])

# ---- multi-GPU: the row-partitioned iteration x <- α A' x over the GPUs of one box, driven by this process -----------------
# (north_star (e).  The loop it scales is the @threads stripe loop of multiply_1DVBC.jl:169-177 / multiply_VBC.jl:182-189:
# stripes are independent row blocks of A'; here contiguous ranges of them live on different GPUs, balanced by memory_cost.)
const VBC_EXCH_FUSED, VBC_EXCH_NCCL = Cint(0), Cint(1)
mutable struct CuVBCDist{Tv, Ti}
    handle::Ptr{Cvoid}
    n::Int
    ngpus::Int
    function CuVBCDist(A::SparseMatrixCSC{Tv, Ti}, Φ::SplitPartition{Ti}; Π::Union{Nothing, SplitPartition{Ti}} = nothing, U::Integer = 0, W::Integer,
                       ngpus::Integer, devices::Union{Nothing, Vector{Cint}} = nothing, exchange::Cint = VBC_EXCH_FUSED) where {Tv, Ti}
        size(A, 1) == size(A, 2) || throw(DimensionMismatch("the iterated operator must be square"))
        h = Ref{Ptr{Cvoid}}(C_NULL)
        GC.@preserve A Φ Π devices begin
            vbc_check(ccall((:vbc_dist_create, libvbc), Cint,
                (Ref{Ptr{Cvoid}}, Cint, Ptr{Cint}, Cint, Cint, Int64, Cint, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Cint),
                h, ngpus, devices === nothing ? C_NULL : pointer(devices), vbc_vt(Tv), vbc_it(Ti), size(A, 2), U, W, A.colptr, A.rowval, A.nzval,
                Π === nothing ? C_NULL : pointer(Π.spl), Π === nothing ? 0 : length(Π), Φ.spl, length(Φ), exchange))
        end
        D = new{Tv, Ti}(h[], size(A, 2), ngpus)
        finalizer(d -> (ccall((:vbc_dist_destroy, libvbc), Cvoid, (Ptr{Cvoid},), d.handle); d.handle = C_NULL), D)
        return D
    end
end
set_x!(D::CuVBCDist{Tv}, x::Vector{Tv}) where {Tv} =
    (GC.@preserve x vbc_check(ccall((:vbc_dist_set_x, libvbc), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), D.handle, x)); D)
# `iters` iterations x <- α A' x, device-resident; returns the device milliseconds per iteration (maximum over the GPUs)
function iterate!(D::CuVBCDist, iters::Integer; α::Number = 1.0)
    ms = Ref{Cdouble}(0.0)
    vbc_check(ccall((:vbc_dist_spmv_iter, libvbc), Cint, (Ptr{Cvoid}, Cint, Cdouble, Ref{Cdouble}), D.handle, iters, Float64(α), ms))
    return ms[]
end
function gather_x(D::CuVBCDist{Tv}) where {Tv}
    x = Vector{Tv}(undef, D.n)
    GC.@preserve x vbc_check(ccall((:vbc_dist_gather_x, libvbc), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), D.handle, x))
    return x
end
# stripe ranges, padded slice length, bytes per GPU under the memory model, interior stripe range of every rank
function partition_info(D::CuVBCDist)
    S = Ref{Int64}(0)
    bounds, cost, interior = Vector{Int64}(undef, D.ngpus + 1), Vector{Int64}(undef, D.ngpus), Vector{Int64}(undef, 2 * D.ngpus)
    vbc_check(ccall((:vbc_dist_info, libvbc), Cint, (Ptr{Cvoid}, Ptr{Cint}, Ref{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}), D.handle, C_NULL, S, bounds, cost, interior))
    return (slice_len = S[], stripe_bounds = bounds, cost_per_gpu = cost, interior = reshape(interior, 2, :))
end

# ---- multi-GPU, one Julia process per GPU (e.g. under MPI.jl): the vbc_peer_* front end of the same fused kernel ---------
# Each rank packs its own stripes (`CuVBC{U, W}(A_slab, Π, Φ_slab; device)`), creates a CuVBCPeer, the ranks exchange the IPC
# handles with whatever transport they have (`MPI.Allgather(handles(P), comm)`), connect, and iterate.  `allgather` below is
# that transport: a function `Vector{UInt8} -> Vector{UInt8}` returning the concatenation over ranks in rank order.
const VBC_IPC_HANDLE_BYTES, VBC_PEER_HANDLES = 64, 3
mutable struct CuVBCPeer{Tv}
    handle::Ptr{Cvoid}
    rank::Int
    nranks::Int
    xlen::Int
    my_handles::Vector{UInt8}
    function CuVBCPeer{Tv}(xlen::Integer, rank::Integer, nranks::Integer; device::Integer = rank) where {Tv}
        h = Ref{Ptr{Cvoid}}(C_NULL)
        hs = zeros(UInt8, VBC_IPC_HANDLE_BYTES * VBC_PEER_HANDLES)
        GC.@preserve hs vbc_check(ccall((:vbc_peer_create, libvbc), Cint, (Ref{Ptr{Cvoid}}, Cint, Int64, Cint, Cint, Cint, Ptr{Cvoid}),
                                        h, vbc_vt(Tv), xlen, rank, nranks, device, hs))
        P = new{Tv}(h[], rank, nranks, xlen, hs)
        finalizer(p -> (ccall((:vbc_peer_destroy, libvbc), Cvoid, (Ptr{Cvoid},), p.handle); p.handle = C_NULL), P)
        return P
    end
end
handles(P::CuVBCPeer) = P.my_handles
function connect!(P::CuVBCPeer, allgather::Function)
    all = allgather(P.my_handles)
    length(all) == P.nranks * length(P.my_handles) || throw(ArgumentError("allgather must return the handles of all $(P.nranks) ranks"))
    GC.@preserve all vbc_check(ccall((:vbc_peer_connect, libvbc), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), P.handle, all))
    return P
end
# device pointer of x buffer k (0 or 1) as a CuVec, and which of the two holds the current x
function buffer(P::CuVBCPeer{Tv}, k::Integer) where {Tv}
    p = Ref{Ptr{Cvoid}}(C_NULL)
    vbc_check(ccall((:vbc_peer_buffer, libvbc), Cint, (Ptr{Cvoid}, Cint, Ref{Ptr{Cvoid}}), P.handle, k, p))
    return CuVec{Tv}(p[], P.xlen)
end
current(P::CuVBCPeer) = (c = Ref{Cint}(0); vbc_check(ccall((:vbc_peer_current, libvbc), Cint, (Ptr{Cvoid}, Ref{Cint}), P.handle, c)); Int(c[]))
# sparsity-aware replication: mask[c >> chunk_shift] has bit i set when destination (rank + i) % nranks reads that chunk of this
# rank's columns (`read_chunks` of the other ranks' matrices, gathered by the caller); neighbours = bit set of the ranks to synchronise with
set_mask!(P::CuVBCPeer, mask::Vector{UInt8}, chunk_shift::Integer) =
    (GC.@preserve mask vbc_check(ccall((:vbc_peer_set_mask, libvbc), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Cint), P.handle, mask, length(mask), chunk_shift)); P)
set_neighbors!(P::CuVBCPeer, mask::Integer) = (vbc_check(ccall((:vbc_peer_set_neighbors, libvbc), Cint, (Ptr{Cvoid}, Cuint), P.handle, mask)); P)
function auto_interior!(P::CuVBCPeer, A::CuVBC, y_offset::Integer)
    i0, i1 = Ref{Int64}(0), Ref{Int64}(0)
    vbc_check(ccall((:vbc_peer_auto_interior, libvbc), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ref{Int64}, Ref{Int64}), P.handle, A.handle, y_offset, i0, i1))
    return (i0[], i1[])
end
function read_chunks(A::CuVBC, chunk_shift::Integer)
    n = (A.m + (1 << chunk_shift) - 1) >> chunk_shift
    need = zeros(UInt8, max(n, 1))
    GC.@preserve need vbc_check(ccall((:vbc_read_chunks, libvbc), Cint, (Ptr{Cvoid}, Cint, Ptr{UInt8}, Int64), A.handle, chunk_shift, need, length(need)))
    return need[1:n]
end
# one iteration x_next[y_offset .+ (1:size(A, 2))] = α A' x_cur on this rank's stripes, stored on every rank that reads it: ONE kernel launch
# on A's stream, the exchange fused in (barrier: 1 signal | 2 wait | 3 both).  Nothing is awaited on the host.
step!(P::CuVBCPeer, A::CuVBC, y_offset::Integer; α::Number = 1.0, barrier::Integer = 3) =
    (vbc_check(ccall((:vbc_peer_spmv_step, libvbc), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cdouble, Int64, Cint), P.handle, A.handle, Float64(α), y_offset, barrier)); P)
barrier!(P::CuVBCPeer, stream::Ptr{Cvoid} = C_NULL; mode::Integer = 3) =
    (vbc_check(ccall((:vbc_peer_barrier, libvbc), Cint, (Ptr{Cvoid}, Ptr{Cvoid}, Cint), P.handle, stream, mode)); P)
function status(P::CuVBCPeer)   # true when no in-kernel flag wait has timed out (call after `sync(A)`)
    t = Ref{Cint}(0)
    vbc_check(ccall((:vbc_peer_status, libvbc), Cint, (Ptr{Cvoid}, Ref{Cint}), P.handle, t))
    return t[] == 0
end
function wait_stats(P::CuVBCPeer; reset::Bool = false)
    s = zeros(UInt64, 4)
    GC.@preserve s vbc_check(ccall((:vbc_peer_wait_stats, libvbc), Cint, (Ptr{Cvoid}, Ptr{UInt64}, Cint), P.handle, s, reset))
    return (steps = s[1], wait_ns = s[2], waits_that_spun = s[3], longest_wait_ns = s[4])
end

# ---- the rest of the surface -------------------------------------------------------------------------------------------------
version() = Int(ccall((:vbc_version, libvbc), Cint, ()))
device_count() = (c = Ref{Cint}(0); vbc_check(ccall((:vbc_device_count, libvbc), Cint, (Ref{Cint},), c)); Int(c[]))
# (m, n, K, L, U, W, ndim, vt, it) as the device holds them
function shape(A::CuVBC)
    m, n, K, L = Ref{Int64}(0), Ref{Int64}(0), Ref{Int64}(0), Ref{Int64}(0)
    U, W, nd, vt, it = Ref{Cint}(0), Ref{Cint}(0), Ref{Cint}(0), Ref{Cint}(0), Ref{Cint}(0)
    vbc_check(ccall((:vbc_shape, libvbc), Cint, (Ptr{Cvoid}, Ref{Int64}, Ref{Int64}, Ref{Int64}, Ref{Int64}, Ref{Cint}, Ref{Cint}, Ref{Cint}, Ref{Cint}, Ref{Cint}),
                    A.handle, m, n, K, L, U, W, nd, vt, it))
    return (m = m[], n = n[], K = K[], L = L[], U = Int(U[]), W = Int(W[]), ndim = Int(nd[]), vt = vt[], it = it[])
end
# level schedule of the triangular solve (extension): built once, reused by every `ldiv_lower!`
trsv_analyse!(A::CuVBC) = (l = Ref{Cint}(0); vbc_check(ccall((:vbc_trsv_analyse, libvbc), Cint, (Ptr{Cvoid}, Ref{Cint}), A.handle, l)); Int(l[]))
trsv_levels(A::CuVBC) = (l = Ref{Cint}(0); vbc_check(ccall((:vbc_trsv_levels, libvbc), Cint, (Ptr{Cvoid}, Ref{Cint}), A.handle, l)); Int(l[]))
set_stream!(A::CuCSC, stream::Ptr{Cvoid}) = vbc_check(ccall((:vbc_csc_set_stream, libvbc), Cint, (Ptr{Cvoid}, Ptr{Cvoid}), A.handle, stream))
# pack from a CSC matrix that already lives on the device (colptr / rowval / Φ.spl / Π.spl as CuVec{Ti}, nzval as CuVec{Tv}); the host
# copies of the partitions are kept for `A.Π` / `A.Φ`
function CuVBC{U, W}(m::Integer, n::Integer, colptr::CuVec{Ti}, rowval::CuVec{Ti}, nzval::CuVec{Tv}, Π::Union{Nothing, SplitPartition{Ti}}, dΠ::Union{Nothing, CuVec{Ti}},
                     Φ::SplitPartition{Ti}, dΦ::CuVec{Ti}; device::Integer = 0) where {U, W, Tv, Ti}
    h = Ref{Ptr{Cvoid}}(C_NULL)
    vbc_check(ccall((:vbc_pack_csc_dev, libvbc), Cint,
        (Ref{Ptr{Cvoid}}, Cint, Cint, Int64, Int64, Cint, Cint, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Ptr{Cvoid}, Int64, Cint),
        h, vbc_vt(Tv), vbc_it(Ti), m, n, U, W, colptr.ptr, rowval.ptr, nzval.ptr,
        dΠ === nothing ? C_NULL : dΠ.ptr, Π === nothing ? 0 : length(Π), dΦ.ptr, length(Φ), device))
    return CuVBC{U, W, Tv, Ti}(h[], m, n, Π, Φ)
end
set_interior!(P::CuVBCPeer, i0::Integer, i1::Integer) = (vbc_check(ccall((:vbc_peer_set_interior, libvbc), Cint, (Ptr{Cvoid}, Int64, Int64), P.handle, i0, i1)); P)
function get_interior(P::CuVBCPeer)
    i0, i1 = Ref{Int64}(0), Ref{Int64}(0)
    vbc_check(ccall((:vbc_peer_get_interior, libvbc), Cint, (Ptr{Cvoid}, Ref{Int64}, Ref{Int64}), P.handle, i0, i1))
    return (i0[], i1[])
end
# Not bound on purpose: vbc_dp_chunk / vbc_overlap_chunk (ChainPartitioners does this work on the Julia side), vbc_gen_banded_csc /
# vbc_gen_free (the benchmark's device-side matrix generator), vbc_peer_connect_local (ranks sharing one process: vbc_dist_* does that).
