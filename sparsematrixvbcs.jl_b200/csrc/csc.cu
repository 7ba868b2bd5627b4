// csc.cu -- `TrSpMV!(y, A::SparseMatrixCSC, x)`: y = A' x on the CSC arrays themselves.
// Replaces /root/reference/src/TrSpMV.jl:1-20 (the "reference" row of bin/test_table.jl:40).
// Column i of A is row i of A': G lanes walk colptr[i]..colptr[i+1] with coalesced loads,
// gather x[rowval[q]], and finish with a shuffle reduction (CSR-vector shape).
#include "common.cuh"

#include <type_traits>

namespace vbc {

// a * b + c: one FMA for the floating-point types; the integer element types are computed on the unsigned type of their width
// (Julia's wrapping arithmetic without signed-overflow undefined behaviour)
template <typename T> __device__ __forceinline__ T mad_any(const T a, const T b, const T c)
{
    if constexpr (std::is_floating_point<T>::value) return fma(a, b, c);
    else return a * b + c;
}

template <typename Tv, typename Ti, int G>
__global__ void __launch_bounds__(256) k_csc_trspmv(const Ti *__restrict__ colptr, const Ti *__restrict__ rowval,
                                                     const Tv *__restrict__ nzval, const Tv *__restrict__ x,
                                                     Tv *__restrict__ y, const int64_t n)
{
    const int lane = threadIdx.x % G;
    unsigned gmask = 0xffffffffu;
    if constexpr (G < 32) gmask = ((1u << G) - 1u) << (((threadIdx.x & 31) / G) * G);
    const int64_t ngroups = ((int64_t)gridDim.x * blockDim.x) / G;
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G; i < n; i += ngroups) {
        const int64_t b = (int64_t)__ldg(colptr + i) - 1, e = (int64_t)__ldg(colptr + i + 1) - 1;
        Tv tmp = (Tv)0;
        int64_t q = b + lane;
        for (; q + 3 * G < e; q += 4 * G) {
            Tv v[4];
            Ti r[4];
#pragma unroll
            for (int k = 0; k < 4; k++) { v[k] = __ldcs(nzval + q + k * G); r[k] = __ldcs(rowval + q + k * G); }
#pragma unroll
            for (int k = 0; k < 4; k++) tmp = mad_any(v[k], __ldg(x + (int64_t)r[k] - 1), tmp);
        }
        for (; q < e; q += G) tmp = mad_any(__ldcs(nzval + q), __ldg(x + (int64_t)__ldcs(rowval + q) - 1), tmp);
#pragma unroll
        for (int d = 1; d < G; d <<= 1) tmp += __shfl_xor_sync(gmask, tmp, d, G);
        if (lane == 0) y[i] = tmp; // TrSpMV.jl:16  (plain store, no alpha/beta)
    }
}

template <typename Tv, typename Ti, int G>
static int launch_g(vbc_csc *A, const Tv *x, Tv *y)
{
    int occ = 0;
    VBC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_csc_trspmv<Tv, Ti, G>, 256, 0));
    if (occ < 1) occ = 1;
    int64_t grid = (int64_t)A->sm_count * occ;
    const int64_t need = (A->n * G + 255) / 256;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    k_csc_trspmv<Tv, Ti, G><<<(unsigned)grid, 256, 0, A->stream>>>((const Ti *)A->d_colptr, (const Ti *)A->d_rowval, (const Tv *)A->d_nzval, x, y, A->n);
    A->launches++;
    VBC_CUDA(cudaGetLastError());
    return VBC_OK;
}

template <typename Tv, typename Ti>
static int launch_t(vbc_csc *A, const void *x, void *y)
{
    if (A->n == 0) return VBC_OK;
    const double avg = (double)A->nnz / (double)A->n;
    if (avg >= 96.0) return launch_g<Tv, Ti, 32>(A, (const Tv *)x, (Tv *)y);
    if (avg >= 12.0) return launch_g<Tv, Ti, 8>(A, (const Tv *)x, (Tv *)y);
    return launch_g<Tv, Ti, 2>(A, (const Tv *)x, (Tv *)y);
}

template <typename Ti> static int launch_ti(vbc_csc *A, const void *d_x, void *d_y)
{
    switch (A->vt) {
    case VBC_F64: return launch_t<double, Ti>(A, d_x, d_y);
    case VBC_F32: return launch_t<float, Ti>(A, d_x, d_y);
    case VBC_INT64: return launch_t<unsigned long long, Ti>(A, d_x, d_y);
    default: return launch_t<unsigned, Ti>(A, d_x, d_y);
    }
}

int launch_csc_trspmv(vbc_csc *A, const void *d_x, void *d_y)
{
    return A->it == VBC_I64 ? launch_ti<int64_t>(A, d_x, d_y) : launch_ti<int32_t>(A, d_x, d_y);
}

} // namespace vbc
