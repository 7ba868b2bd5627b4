// inttypes.cu -- multiplies of matrices with integer element types (Int32, Int64; the host layers widen Bool to Int32).
//
// The reference's methods are generic over `Tv <: SIMD.VecTypes` and its test suite packs and multiplies `Bool` and `Int32`
// matrices next to the floating-point ones (test/runtests.jl:15-16, one-hot products compared with `==`, :29-53, :63-87).
// Julia's fixed-width integer arithmetic wraps, and wrapping addition is associative: any summation order gives the
// reference's bits, so the forward multiply may scatter with integer atomics and stay exact.  All arithmetic here is done on
// the unsigned type of the same width (same bits as two's-complement wrapping, no signed-overflow undefined behaviour).
// The kernels read the compact layout of spmv.cu (StripeMeta + desc); pack.cu moves values as bits and needs no integer
// instantiation of its own (the 4- and 8-byte ones serve).  The adjoint has the lane mapping and 16-byte loads of the floating-point
// kernels (8 lanes per stripe); the forward multiply is a plain scatter.
#include "walk.cuh"
#include <limits>

namespace vbc {

namespace {

constexpr int IG = 8; // lanes per stripe

template <int MODE> __device__ __forceinline__ int int_x_index(const int *__restrict__ desc, const int pos0, const int r, const int u0)
{
    if constexpr (MODE == DESC_ROWS) return __ldg(desc + pos0 + r);
    else return __ldg(desc + pos0 + r / u0) + r % u0;
}

// EPV elements of a stripe's slab in one 16- / 8- / 4-byte load (streaming: every value is read once)
template <typename T, int EPV> __device__ __forceinline__ void ld_vec(const T *p, T (&v)[EPV])
{
    if constexpr (sizeof(T) * EPV == 16) {
        const uint4 q = __ldcs(reinterpret_cast<const uint4 *>(p));
        if constexpr (sizeof(T) == 4) { v[0] = (T)q.x; v[1] = (T)q.y; v[2] = (T)q.z; v[3] = (T)q.w; }
        else { v[0] = (T)(((unsigned long long)q.y << 32) | q.x); v[1] = (T)(((unsigned long long)q.w << 32) | q.z); }
    } else if constexpr (sizeof(T) * EPV == 8 && sizeof(T) == 4) {
        const uint2 q = __ldcs(reinterpret_cast<const uint2 *>(p));
        v[0] = (T)q.x; v[1] = (T)q.y;
    } else {
        static_assert(EPV == 1, "vector shapes: 4 x 4 B, 2 x 4 B, 2 x 8 B, or one element");
        v[0] = __ldcs(p);
    }
}

// y[j + c] = alpha * sum_r val[ofs + r w + c] * x[i_r] (+ beta y[j + c])      multiply_1DVBC.jl:98-118, multiply_VBC.jl:99-135
// The lane mapping of the floating-point adjoint kernels (spmv.cu, mixed.cu): the IG lanes of a group read consecutive EPV-element
// vectors of the stripe's slab, lane v holds column-vector v mod cpr and rows r0, r0 + rps, ...; the loads of four row-steps are issued
// before the first product; the lanes that share a column-vector are summed with the strided shuffle tree (any order is exact here).
template <typename T, int MODE, int EPV>
__device__ __forceinline__ void int_adj_stripe(const StripeMeta a, const int w, const int R, const int lane, const unsigned gmask,
                                               const int *__restrict__ desc, const T *__restrict__ val, const T *__restrict__ x,
                                               T *__restrict__ y, const int u0, const int log2u, const T alpha, const T beta)
{
    const int cpr = w / EPV, rps = small_div(IG, cpr), r0 = small_div(lane, cpr), c = lane - r0 * cpr;
    const bool active = lane < rps * cpr;
    T acc[EPV];
#pragma unroll
    for (int e = 0; e < EPV; e++) acc[e] = (T)0;
    const T *vp = val + a.ofs + (long long)r0 * w + c * EPV;
    const int vstride = rps * w;
    RowWalk<MODE> walk;
    walk.init(desc, a.pos, r0, rps, u0, log2u);
    constexpr int UNR = 4;
    for (int r = active ? r0 : R; r < R; r += UNR * rps) {
        T v[UNR][EPV];
        int xi[UNR];
        T xv[UNR];
        bool ok[UNR];
#pragma unroll
        for (int k = 0; k < UNR; k++) {
            ok[k] = r + k * rps < R;
#pragma unroll
            for (int e = 0; e < EPV; e++) v[k][e] = (T)0;
            if (ok[k]) ld_vec<T, EPV>(vp, v[k]);
            vp += vstride;
        }
#pragma unroll
        for (int k = 0; k < UNR; k++) xi[k] = walk.next_if(ok[k]);
#pragma unroll
        for (int k = 0; k < UNR; k++) xv[k] = ok[k] ? __ldg(x + xi[k]) : (T)0;
#pragma unroll
        for (int k = 0; k < UNR; k++)
#pragma unroll
            for (int e = 0; e < EPV; e++) acc[e] += v[k][e] * xv[k];
    }
    __syncwarp(gmask);
    for (int d = cpr; d < IG; d <<= 1)
#pragma unroll
        for (int e = 0; e < EPV; e++) {
            const T t = __shfl_down_sync(gmask, acc[e], d, IG);
            if (lane + d < IG) acc[e] += t;
        }
    if (lane < cpr) {
        T *yp = y + a.col + lane * EPV;
#pragma unroll
        for (int e = 0; e < EPV; e++) yp[e] = (beta == (T)0) ? alpha * acc[e] : alpha * acc[e] + beta * yp[e];
    }
}

template <typename T, int MODE>
__global__ void __launch_bounds__(256) k_int_adj(const StripeMeta *__restrict__ meta, const int *__restrict__ desc, const T *__restrict__ val,
                                                 const T *__restrict__ x, T *__restrict__ y, const int L, const int u0, const int log2u, const T alpha, const T beta)
{
    constexpr int VE = 16 / (int)sizeof(T); // elements per 16-byte vector
    const int lane = threadIdx.x % IG;
    const unsigned gmask = ((1u << IG) - 1u) << (((threadIdx.x & 31) / IG) * IG);
    const long long groups = (long long)gridDim.x * (blockDim.x / IG);
    for (long long l = (long long)blockIdx.x * (blockDim.x / IG) + threadIdx.x / IG; l < L; l += groups) {
        const StripeMeta a = ld_meta(meta + l), b = ld_meta(meta + l + 1);
        const int w = b.col - a.col;
        if (w <= 0) continue; // uniform over the group
        const int rows = (MODE == DESC_ROWS) ? (b.pos - a.pos) : (int)((b.ofs - a.ofs) / w);
        if (w % VE == 0 && a.ofs % VE == 0 && w / VE <= IG) int_adj_stripe<T, MODE, VE>(a, w, rows, lane, gmask, desc, val, x, y, u0, log2u, alpha, beta);
        else if (VE == 4 && (w & 1) == 0 && (a.ofs & 1) == 0 && (w >> 1) <= IG) int_adj_stripe<T, MODE, (VE == 4 ? 2 : 1)>(a, w, rows, lane, gmask, desc, val, x, y, u0, log2u, alpha, beta);
        else if (w <= IG) int_adj_stripe<T, MODE, 1>(a, w, rows, lane, gmask, desc, val, x, y, u0, log2u, alpha, beta);
        else { // wider than the group: one lane per column, IG columns per pass
            const T *vbase = val + a.ofs;
            for (int c = lane; c < w; c += IG) {
                T acc = (T)0;
                for (int r = 0; r < rows; r++) acc += vbase[(long long)r * w + c] * __ldg(x + int_x_index<MODE>(desc, a.pos, r, u0));
                T *yp = y + a.col + c;
                *yp = (beta == (T)0) ? alpha * acc : alpha * acc + beta * *yp;
            }
        }
    }
}

// y[i_r] += alpha * sum_c val[ofs + r w + c] * x[j + c]      multiply_1DVBC.jl:26-36, multiply_VBC.jl:40-45
template <typename T, int MODE>
__global__ void __launch_bounds__(256) k_int_fwd(const StripeMeta *__restrict__ meta, const int *__restrict__ desc, const T *__restrict__ val,
                                                 const T *__restrict__ x, T *__restrict__ y, const int L, const int u0, const T alpha)
{
    const int lane = threadIdx.x % IG;
    const long long groups = (long long)gridDim.x * (blockDim.x / IG);
    for (long long l = (long long)blockIdx.x * (blockDim.x / IG) + threadIdx.x / IG; l < L; l += groups) {
        const StripeMeta a = ld_meta(meta + l), b = ld_meta(meta + l + 1);
        const int w = b.col - a.col;
        if (w <= 0) continue;
        const int rows = (int)((b.ofs - a.ofs) / w);
        for (int r = lane; r < rows; r += IG) {
            const T *vp = val + a.ofs + (long long)r * w;
            T s = (T)0;
            for (int c = 0; c < w; c++) s += vp[c] * __ldg(x + a.col + c);
            atomicAdd(y + int_x_index<MODE>(desc, a.pos, r, u0), alpha * s); // wrapping adds commute: exact in any order
        }
    }
}

template <typename T> __global__ void __launch_bounds__(256) k_int_scale(T *__restrict__ y, const int64_t len, const T beta)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (int64_t)gridDim.x * blockDim.x)
        y[i] = (beta == (T)0) ? (T)0 : beta * y[i];
}

template <typename T>
int launch_int_t(vbc_mat *A, int trans, T alpha, const T *x, T beta, T *y)
{
    if (!trans) { // forward through the transposed copy when the matrix has one (the adjoint kernel on it, no atomics), like mixed.cu
        VBC_TRY(ensure_tindex(A));
        vbc_mat *At = tindex_copy(A);
        if (At != nullptr) { At->stream = A->stream; const int64_t before = At->launches; const int rc = launch_int_t<T>(At, 1, alpha, x, beta, y); A->launches += At->launches - before; return rc; }
    }
    const int64_t ylen = trans ? A->n : A->m;
    const int per_block = 256 / IG;
    int64_t grid = (A->L + per_block - 1) / per_block;
    const int64_t cap = (int64_t)A->sm_count * 32;
    if (grid > cap) grid = cap;
    if (!trans && ylen > 0 && beta != (T)1) {
        int64_t g = (ylen + 255) / 256;
        if (g > cap) g = cap;
        k_int_scale<T><<<(unsigned)g, 256, 0, A->stream>>>(y, ylen, beta);
        A->launches++;
    }
    int log2u = -1;
    if (A->u0 > 0 && !(A->u0 & (A->u0 - 1))) { log2u = 0; while ((1 << log2u) < A->u0) log2u++; }
    if (A->L > 0 && grid > 0) {
        const bool rows = A->desc_mode == DESC_ROWS;
        const T *val = (const T *)A->d_val;
        if (trans) {
            if (rows) k_int_adj<T, DESC_ROWS><<<(unsigned)grid, 256, 0, A->stream>>>(A->d_meta, A->d_desc, val, x, y, (int)A->L, A->u0, log2u, alpha, beta);
            else k_int_adj<T, DESC_BLOCKS><<<(unsigned)grid, 256, 0, A->stream>>>(A->d_meta, A->d_desc, val, x, y, (int)A->L, A->u0, log2u, alpha, beta);
        } else {
            if (rows) k_int_fwd<T, DESC_ROWS><<<(unsigned)grid, 256, 0, A->stream>>>(A->d_meta, A->d_desc, val, x, y, (int)A->L, A->u0, alpha);
            else k_int_fwd<T, DESC_BLOCKS><<<(unsigned)grid, 256, 0, A->stream>>>(A->d_meta, A->d_desc, val, x, y, (int)A->L, A->u0, alpha);
        }
        A->launches++;
    }
    VBC_CUDA(cudaGetLastError());
    return VBC_OK;
}

// `convert(eltype(y), alpha)` (multiply_1DVBC.jl:48): a scalar that is not an integer of the element type's range is an InexactError
template <typename S> bool exact_int(const double v, S *out)
{
    const double lo = (double)std::numeric_limits<S>::min(); // -2^31 / -2^63, exact
    if (!(v >= lo) || !(v < -lo)) return false;              // also rejects NaN
    const long long t = (long long)v;
    if ((double)t != v) return false;
    *out = (S)t;
    return true;
}

} // namespace

int launch_spmv_int(vbc_mat *A, int trans, double alpha, const void *d_x, double beta, void *d_y)
{
    if (A->vt == VBC_INT64) {
        long long a = 0, b = 0;
        if (!exact_int<long long>(alpha, &a) || !exact_int<long long>(beta, &b)) VBC_FAIL(VBC_EARG, "InexactError: alpha = %g / beta = %g are not Int64 values", alpha, beta);
        return launch_int_t<unsigned long long>(A, trans, (unsigned long long)a, (const unsigned long long *)d_x, (unsigned long long)b, (unsigned long long *)d_y);
    }
    int a = 0, b = 0;
    if (!exact_int<int>(alpha, &a) || !exact_int<int>(beta, &b)) VBC_FAIL(VBC_EARG, "InexactError: alpha = %g / beta = %g are not Int32 values", alpha, beta);
    return launch_int_t<unsigned>(A, trans, (unsigned)a, (const unsigned *)d_x, (unsigned)b, (unsigned *)d_y);
}

} // namespace vbc
