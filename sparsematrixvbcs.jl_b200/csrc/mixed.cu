// mixed.cu -- multiplies whose vectors are wider than the stored values (Float32 matrix, Float64 x and y).
//
// The reference converts the stored values AND x to eltype(y) before multiplying
// (src/multiply_1DVBC.jl:23/27/34 forward, :102 adjoint; src/multiply_VBC.jl:40-45, :131), so a Float32
// matrix applied to Float64 vectors accumulates in Float64.  These kernels keep that rule: every product
// is Tu(val) * Tu(x) with Tu the vector type, summed in ascending stored-row order like the reference's loops.
// They read the same compact layout as spmv.cu (StripeMeta + desc) in a separate pair of kernels (the tuned same-type
// kernels stay untouched): the adjoint with the lane mapping of the same-type kernel, the forward multiply through the
// transposed copy when the matrix has one (else a scatter with atomics).
#include "walk.cuh"

namespace vbc {

namespace {

constexpr int MG = 8; // lanes per stripe

template <int MODE> __device__ __forceinline__ int x_index(const int *__restrict__ desc, const int pos0, const int r, const int u0)
{
    if constexpr (MODE == DESC_ROWS) return __ldg(desc + pos0 + r);
    else return __ldg(desc + pos0 + r / u0) + r % u0;
}

// y[j + c] = alpha * sum_r Tu(val[ofs + r w + c]) * x[i_r]  (+ beta y)     multiply_1DVBC.jl:98-118, multiply_VBC.jl:99-135
// Same lane mapping as the same-type adjoint kernel (spmv.cu): the MG lanes of a group read consecutive EPV-element vectors of
// the stripe's slab (16 / 8 / 4 bytes, by the slab's alignment), lane v always holds column-vector v mod cpr and rows r0, r0 + rps, ...;
// loads of four row-steps are issued before the first FMA; every product is Tu(val) * x in the vector type; the lanes that
// share a column-vector are summed with the strided shuffle tree.  Stripes with more vectors per row than lanes: one lane per column.
template <typename Tm, typename Tu, int MODE, int EPV>
__device__ __forceinline__ void mixed_adj_stripe(const StripeMeta a, const int w, const int R, const int lane, const unsigned gmask,
                                                 const int *__restrict__ desc, const Tm *__restrict__ val, const Tu *__restrict__ x,
                                                 Tu *__restrict__ y, const int u0, const int log2u, const Tu alpha, const Tu beta)
{
    const int cpr = w / EPV, rps = small_div(MG, cpr), r0 = small_div(lane, cpr), c = lane - r0 * cpr;
    const bool active = lane < rps * cpr;
    Tu acc[EPV];
#pragma unroll
    for (int e = 0; e < EPV; e++) acc[e] = (Tu)0;
    const Tm *vp = val + a.ofs + (long long)r0 * w + c * EPV;
    const int vstride = rps * w;
    RowWalk<MODE> walk;
    walk.init(desc, a.pos, r0, rps, u0, log2u);
    constexpr int UNR = 4;
    for (int r = active ? r0 : R; r < R; r += UNR * rps) {
        Tm v[UNR][EPV];
        int xi[UNR];
        Tu xv[UNR];
        bool ok[UNR];
#pragma unroll
        for (int k = 0; k < UNR; k++) {
            ok[k] = r + k * rps < R;
#pragma unroll
            for (int e = 0; e < EPV; e++) v[k][e] = (Tm)0;
            if (ok[k]) {
                if constexpr (EPV == 4) { const float4 q = __ldcs(reinterpret_cast<const float4 *>(vp)); v[k][0] = q.x; v[k][1] = q.y; v[k][2] = q.z; v[k][3] = q.w; }
                else if constexpr (EPV == 2) { const float2 q = __ldcs(reinterpret_cast<const float2 *>(vp)); v[k][0] = q.x; v[k][1] = q.y; }
                else v[k][0] = __ldcs(vp);
            }
            vp += vstride;
        }
#pragma unroll
        for (int k = 0; k < UNR; k++) xi[k] = walk.next_if(ok[k]);
#pragma unroll
        for (int k = 0; k < UNR; k++) xv[k] = ok[k] ? __ldg(x + xi[k]) : (Tu)0;
#pragma unroll
        for (int k = 0; k < UNR; k++)
#pragma unroll
            for (int e = 0; e < EPV; e++) acc[e] = fma((Tu)v[k][e], xv[k], acc[e]);
    }
    __syncwarp(gmask);
    for (int d = cpr; d < MG; d <<= 1)
#pragma unroll
        for (int e = 0; e < EPV; e++) {
            const Tu t = __shfl_down_sync(gmask, acc[e], d, MG);
            if (lane + d < MG) acc[e] += t;
        }
    if (lane < cpr) {
        Tu *yp = y + a.col + lane * EPV;
#pragma unroll
        for (int e = 0; e < EPV; e++) yp[e] = (beta == (Tu)0) ? alpha * acc[e] : alpha * acc[e] + beta * yp[e];
    }
}

template <typename Tm, typename Tu, int MODE>
__global__ void __launch_bounds__(256, 4) k_mixed_adj(const StripeMeta *__restrict__ meta, const int *__restrict__ desc, const Tm *__restrict__ val,
                                                   const Tu *__restrict__ x, Tu *__restrict__ y, const int L, const int u0, const int log2u,
                                                   const Tu alpha, const Tu beta)
{
    static_assert(sizeof(Tm) == 4, "the vector widths below are those of a 4-byte matrix type");
    const int lane = threadIdx.x % MG;
    const unsigned gmask = ((1u << MG) - 1u) << (((threadIdx.x & 31) / MG) * MG);
    const long long groups = (long long)gridDim.x * (blockDim.x / MG);
    for (long long l = (long long)blockIdx.x * (blockDim.x / MG) + threadIdx.x / MG; l < L; l += groups) {
        const StripeMeta a = ld_meta(meta + l), b = ld_meta(meta + l + 1);
        const int w = b.col - a.col;
        if (w <= 0) continue;
        const int rows = (MODE == DESC_ROWS) ? (b.pos - a.pos) : (int)((b.ofs - a.ofs) / w);
        if ((w & 3) == 0 && (a.ofs & 3) == 0 && (w >> 2) <= MG) mixed_adj_stripe<Tm, Tu, MODE, 4>(a, w, rows, lane, gmask, desc, val, x, y, u0, log2u, alpha, beta);
        else if ((w & 1) == 0 && (a.ofs & 1) == 0 && (w >> 1) <= MG) mixed_adj_stripe<Tm, Tu, MODE, 2>(a, w, rows, lane, gmask, desc, val, x, y, u0, log2u, alpha, beta);
        else if (w <= MG) mixed_adj_stripe<Tm, Tu, MODE, 1>(a, w, rows, lane, gmask, desc, val, x, y, u0, log2u, alpha, beta);
        else {
            for (int c = lane; c < w; c += MG) {
                Tu acc = (Tu)0;
                const Tm *vp = val + a.ofs + c;
                for (int r = 0; r < rows; r++) acc += (Tu)vp[(long long)r * w] * x[x_index<MODE>(desc, a.pos, r, u0)];
                Tu *yp = y + a.col + c;
                *yp = (beta == (Tu)0) ? alpha * acc : alpha * acc + beta * *yp;
            }
        }
    }
}

// y[i_r] += alpha * sum_c Tu(val[ofs + r w + c]) * x[j + c]      multiply_1DVBC.jl:26-36, multiply_VBC.jl:40-45
template <typename Tm, typename Tu, int MODE>
__global__ void __launch_bounds__(256) k_mixed_fwd(const StripeMeta *__restrict__ meta, const int *__restrict__ desc, const Tm *__restrict__ val,
                                                   const Tu *__restrict__ x, Tu *__restrict__ y, const int L, const int u0, const Tu alpha)
{
    const int lane = threadIdx.x % MG;
    const long long groups = (long long)gridDim.x * (blockDim.x / MG);
    for (long long l = (long long)blockIdx.x * (blockDim.x / MG) + threadIdx.x / MG; l < L; l += groups) {
        const StripeMeta a = ld_meta(meta + l), b = ld_meta(meta + l + 1);
        const int w = b.col - a.col;
        if (w <= 0) continue;
        const int rows = (int)((b.ofs - a.ofs) / w);
        for (int r = lane; r < rows; r += MG) {
            const Tm *vp = val + a.ofs + (long long)r * w;
            Tu s = (Tu)0;
            for (int c = 0; c < w; c++) s += (Tu)vp[c] * x[a.col + c];
            atomicAdd(y + x_index<MODE>(desc, a.pos, r, u0), alpha * s);
        }
    }
}

template <typename Tu> __global__ void __launch_bounds__(256) k_mixed_scale(Tu *__restrict__ y, const int64_t len, const Tu beta)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (int64_t)gridDim.x * blockDim.x)
        y[i] = (beta == (Tu)0) ? (Tu)0 : beta * y[i];
}

template <typename Tm, typename Tu>
int launch_mixed_t(vbc_mat *A, int trans, Tu alpha, const Tu *x, Tu beta, Tu *y)
{
    if (!trans) { // forward through the transposed copy when the matrix has one: the adjoint kernel on it, no atomics
        VBC_TRY(ensure_tindex(A));
        vbc_mat *At = tindex_copy(A);
        if (At != nullptr) { At->stream = A->stream; const int64_t before = At->launches; const int rc = launch_mixed_t<Tm, Tu>(At, 1, alpha, x, beta, y); A->launches += At->launches - before; return rc; }
    }
    const int64_t ylen = trans ? A->n : A->m;
    int log2u = -1;
    if (A->u0 > 0 && !(A->u0 & (A->u0 - 1))) { log2u = 0; while ((1 << log2u) < A->u0) log2u++; }
    const int per_block = 256 / MG;
    int64_t grid = (A->L + per_block - 1) / per_block;
    const int64_t cap = (int64_t)A->sm_count * 32;
    if (grid > cap) grid = cap;
    if (!trans && ylen > 0 && beta != (Tu)1) {
        int64_t g = (ylen + 255) / 256;
        if (g > cap) g = cap;
        k_mixed_scale<Tu><<<(unsigned)g, 256, 0, A->stream>>>(y, ylen, beta);
        A->launches++;
    }
    if (A->L > 0 && grid > 0) {
        const bool rows = A->desc_mode == DESC_ROWS;
        const Tm *val = (const Tm *)A->d_val;
        if (trans) {
            if (rows) k_mixed_adj<Tm, Tu, DESC_ROWS><<<(unsigned)grid, 256, 0, A->stream>>>(A->d_meta, A->d_desc, val, x, y, (int)A->L, A->u0, log2u, alpha, beta);
            else k_mixed_adj<Tm, Tu, DESC_BLOCKS><<<(unsigned)grid, 256, 0, A->stream>>>(A->d_meta, A->d_desc, val, x, y, (int)A->L, A->u0, log2u, alpha, beta);
        } else {
            if (rows) k_mixed_fwd<Tm, Tu, DESC_ROWS><<<(unsigned)grid, 256, 0, A->stream>>>(A->d_meta, A->d_desc, val, x, y, (int)A->L, A->u0, alpha);
            else k_mixed_fwd<Tm, Tu, DESC_BLOCKS><<<(unsigned)grid, 256, 0, A->stream>>>(A->d_meta, A->d_desc, val, x, y, (int)A->L, A->u0, alpha);
        }
        A->launches++;
    }
    VBC_CUDA(cudaGetLastError());
    return VBC_OK;
}

} // namespace

int launch_spmv_mixed(vbc_mat *A, int trans, double alpha, const void *d_x, double beta, void *d_y)
{
    // the only widening pair of the two device value types
    return launch_mixed_t<float, double>(A, trans, alpha, (const double *)d_x, beta, (double *)d_y);
}

} // namespace vbc
