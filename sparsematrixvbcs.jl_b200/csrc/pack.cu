// pack.cu -- CSC + partition -> VBC pack kernels, and the canonical -> compact layout pass.
//
// Device restatement of
//   SparseMatrix1DVBC{W}(A, Φ)      /root/reference/src/constructors_1DVBC.jl:9-92
//   SparseMatrixVBC{U,W}(A, Π, Φ)   /root/reference/src/constructors_VBC.jl:15-133
// The outputs pos/idx/ofs/val are bit-identical to the reference's (integer arithmetic and
// value copies only).  The reference is two serial sweeps over the stripes; here
//   (1) one thread per stripe merges its <= W sorted columns and COUNTS distinct rows (1D) or
//       distinct row parts and their heights (2D)              [:22-32 / VBC :31-47]
//   (2) exclusive scans give pos and ofs                        [:23,:31 / VBC :42-43]
//   (3) the same merge WRITES idx                                [:84 / VBC :125]
//   (4) val is zero-filled and every stored nonzero is scattered to its slot, found by a
//       binary search of its row (part) in the stripe's idx segment -- equivalent to the
//       reference's in-order emission with `zero(Tv)` fill       [:66-87 / VBC :94-128]
#include <algorithm>
#include <stdlib.h>

#include "common.cuh"
#include "scan.cuh"

namespace vbc {

constexpr int WMAX = 32; // widest stripe the merge kernels hold in per-thread state

// error flag values written by kernels
enum { PERR_W = 1, PERR_U = 2, PERR_SPL = 3, PERR_WMAX = 4, PERR_CSC = 5 };

template <typename Ti>
__global__ void k_check_spl(const Ti *__restrict__ spl, int64_t P, int64_t dim, int limit, int errcode, int *err)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k == 0) {
        if ((int64_t)spl[0] != 1 || (int64_t)spl[P] != dim + 1) atomicMax(err, PERR_SPL);
    }
    if (k < P) {
        const int64_t d = (int64_t)spl[k + 1] - (int64_t)spl[k];
        if (d < 0) atomicMax(err, PERR_SPL);
        else if (d > limit) atomicMax(err, errcode);
    }
}

// CSC validation (the reference trusts SparseMatrixCSC's invariants; here a malformed index would send the merge, the
// value scatter and later every multiply out of bounds): colptr[1] == 1, colptr nondecreasing and ending at nnz + 1,
// 1 <= rowval <= m, rows strictly ascending within a column.  8 lanes per column.
template <typename Ti>
__global__ void __launch_bounds__(256) k_check_csc(const Ti *__restrict__ colptr, const Ti *__restrict__ rowval, int64_t n, int64_t m, int64_t nnz, int *err)
{
    const int64_t j = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const int lane = threadIdx.x & 7;
    if (j >= n) return;
    const int64_t b = (int64_t)colptr[j] - 1, e = (int64_t)colptr[j + 1] - 1;
    if ((j == 0 && b != 0) || b < 0 || e < b || e > nnz) { if (lane == 0) atomicMax(err, PERR_CSC); return; }
    bool bad = false;
    for (int64_t t = b + lane; t < e; t += 8) {
        const int64_t r = (int64_t)rowval[t];
        if (r < 1 || r > m || (t > b && (int64_t)rowval[t - 1] >= r)) bad = true;
    }
    if (bad) atomicMax(err, PERR_CSC);
}

// asg[i] = 0-based part of row i  (`convert(MapPartition, Π).asg`, constructors_VBC.jl:22)
template <typename Ti>
__global__ void k_build_map(const Ti *__restrict__ spl, int64_t P, int64_t dim, int *__restrict__ asg)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= dim) return;
    const Ti key = (Ti)(i + 1);
    int64_t lo = 0, hi = P; // largest k with spl[k] <= key
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (spl[mid] <= key) lo = mid; else hi = mid;
    }
    asg[i] = (int)lo;
}

// One thread per stripe: w-way merge of the stripe's sorted columns.
//   DIM2 = false: units are rows           (constructors_1DVBC.jl:26-30 count, :57-87 emit)
//   DIM2 = true : units are row parts      (constructors_VBC.jl:37-46 count, :85-128 emit)
// WRITE = false: cnt[l] = #units, nvals[l] = #values of the stripe.   WRITE = true: idx written.
template <typename Ti, bool DIM2, bool WRITE>
__global__ void __launch_bounds__(128) k_merge_stripes(const Ti *__restrict__ colptr, const Ti *__restrict__ rowval,
                                                       const Ti *__restrict__ phi_spl, int64_t L,
                                                       const Ti *__restrict__ pi_spl, const int *__restrict__ asg,
                                                       long long *__restrict__ cnt, long long *__restrict__ nvals,
                                                       const Ti *__restrict__ pos, Ti *__restrict__ idx, int *err)
{
    const int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= L) return;
    const int64_t j0 = (int64_t)phi_spl[l] - 1;
    const int w = (int)((int64_t)phi_spl[l + 1] - 1 - j0);
    if (w > WMAX) {
        atomicMax(err, PERR_WMAX);
        if (!WRITE) { cnt[l] = 0; nvals[l] = 0; }
        return;
    }
    int64_t out = WRITE ? (int64_t)pos[l] - 1 : 0;
    if (w == 1 && !DIM2) { // :47-55 -- the column itself
        const int64_t b = (int64_t)colptr[j0] - 1, e = (int64_t)colptr[j0 + 1] - 1;
        if (!WRITE) { cnt[l] = e - b; nvals[l] = e - b; }
        else for (int64_t q = b; q < e; q++) idx[out++] = rowval[q];
        return;
    }
    int64_t q[WMAX], e[WMAX];
    long long key[WMAX]; // head unit of each column, LLONG_MAX when exhausted
    const long long SENT = 0x7fffffffffffffffLL;
    long long cur = SENT;
    for (int dj = 0; dj < w; dj++) {
        q[dj] = (int64_t)colptr[j0 + dj] - 1;
        e[dj] = (int64_t)colptr[j0 + dj + 1] - 1;
        long long kk = SENT;
        if (q[dj] < e[dj]) {
            const long long r = (long long)rowval[q[dj]] - 1;
            kk = DIM2 ? (long long)asg[r] : r;
        }
        key[dj] = kk;
        cur = kk < cur ? kk : cur;
    }
    long long units = 0, rows = 0;
    while (cur != SENT) {
        long long nxt = SENT;
        for (int dj = 0; dj < w; dj++) {
            long long kk = key[dj];
            if (kk == cur) { // advance this column past the current unit
                int64_t qq = q[dj] + 1;
                kk = SENT;
                while (qq < e[dj]) {
                    const long long r = (long long)rowval[qq] - 1;
                    const long long k2 = DIM2 ? (long long)asg[r] : r;
                    if (k2 != cur) { kk = k2; break; }
                    qq++; // only reachable in 2D (several rows of one part) or on duplicate rows
                }
                q[dj] = qq;
                key[dj] = kk;
            }
            nxt = kk < nxt ? kk : nxt;
        }
        if (WRITE) idx[out++] = (Ti)(cur + 1);
        units++;
        if (DIM2) rows += (long long)pi_spl[cur + 1] - (long long)pi_spl[cur];
        cur = nxt;
    }
    if (!WRITE) {
        cnt[l] = units;
        nvals[l] = (DIM2 ? rows : units) * (long long)w;
    }
}

// longest CSC segment of a stripe (entries of all its columns): sizes the shared-memory staging of the warp merge
template <typename Ti>
__global__ void __launch_bounds__(256) k_max_seglen(const Ti *__restrict__ colptr, const Ti *__restrict__ phi_spl, int64_t L, unsigned long long *__restrict__ out)
{
    unsigned long long mx = 0;
    for (int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; l < L; l += (int64_t)gridDim.x * blockDim.x) {
        const unsigned long long len = (unsigned long long)((int64_t)colptr[(int64_t)phi_spl[l + 1] - 1] - (int64_t)colptr[(int64_t)phi_spl[l] - 1]);
        mx = len > mx ? len : mx;
    }
    for (int d = 16; d > 0; d >>= 1) { const unsigned long long o = __shfl_xor_sync(0xffffffffu, mx, d); mx = o > mx ? o : mx; }
    if ((threadIdx.x & 31) == 0 && mx > 0) atomicMax(out, mx);
}

// One WARP per stripe (the same merge, with memory-level parallelism): the stripe's CSC segment -- its columns are adjacent
// in rowval -- is loaded coalesced and mapped to units (rows / row parts) by all lanes into shared memory; lane dj then owns
// column dj's head, the next unit is a warp-wide minimum (REDUX), the lanes whose head equals it advance, and the sorted
// distinct units collect in a second shared list from which the counts (2D: the heights, in parallel) or idx (coalesced
// stores) are produced.  One thread per stripe walking global memory spent ~2 us per entry on two dependent loads.
template <typename Ti, bool DIM2, bool WRITE>
__global__ void __launch_bounds__(256) k_merge_stripes_warp(const Ti *__restrict__ colptr, const Ti *__restrict__ rowval,
                                                            const Ti *__restrict__ phi_spl, int64_t L,
                                                            const Ti *__restrict__ pi_spl, const int *__restrict__ asg,
                                                            long long *__restrict__ cnt, long long *__restrict__ nvals,
                                                            const Ti *__restrict__ pos, Ti *__restrict__ idx, const int cap, int *err,
                                                            int *__restrict__ keep = nullptr)
{
    // keep (count pass only): the merged units of stripe l are parked at keep[b, b + n), b = the stripe's first CSC entry (n <= its
    // entries, so stripes never overlap); the write pass is then a plain copy (k_copy_units) instead of a second merge
    extern __shared__ int merge_smem[]; // per warp: units[cap], outs[cap]
    const int lane = threadIdx.x & 31;
    int *units = merge_smem + (size_t)(threadIdx.x >> 5) * 2 * cap, *outs = units + cap;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t l = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; l < L; l += nwarps) {
        const int64_t j0 = (int64_t)phi_spl[l] - 1;
        const int w = (int)((int64_t)phi_spl[l + 1] - 1 - j0);
        if (w > WMAX) {
            if (lane == 0) { atomicMax(err, PERR_WMAX); if (!WRITE) { cnt[l] = 0; nvals[l] = 0; } }
            continue;
        }
        const int64_t b = (int64_t)colptr[j0] - 1;
        // lane dj: [h, end) = column dj's entries, relative to the segment
        int h = 0, end = 0;
        if (lane < w) { h = (int)((int64_t)colptr[j0 + lane] - 1 - b); end = (int)((int64_t)colptr[j0 + lane + 1] - 1 - b); }
        const int len = __shfl_sync(0xffffffffu, end, w > 0 ? w - 1 : 0);
        for (int tb = 0; tb < len; tb += 32 * 8) { // eight independent row loads in flight per lane, then their (dependent) part lookups
            long long r[8];
#pragma unroll
            for (int k = 0; k < 8; k++) { const int t = tb + k * 32 + lane; r[k] = t < len ? (long long)rowval[b + t] - 1 : -1; }
            int uu[8];
#pragma unroll
            for (int k = 0; k < 8; k++) uu[k] = r[k] < 0 ? 0 : (DIM2 ? asg[r[k]] : (int)r[k]);
#pragma unroll
            for (int k = 0; k < 8; k++) { const int t = tb + k * 32 + lane; if (t < len) units[t] = uu[k]; }
        }
        __syncwarp();
        int n = 0;
        if (w == 1 && !DIM2) { // :47-55 -- the column itself
            n = len;
            if (WRITE) { const int64_t out = (int64_t)pos[l] - 1; for (int t = lane; t < len; t += 32) idx[out + t] = (Ti)(units[t] + 1); }
        } else {
            const unsigned SENT = 0x7fffffffu;
            unsigned key = h < end ? (unsigned)units[h] : SENT;
            for (;;) {
                const unsigned cur = __reduce_min_sync(0xffffffffu, key);
                if (cur == SENT) break;
                if (key == cur) { // advance this column past the current unit (2D: several rows of one part)
                    do { h++; } while (h < end && (unsigned)units[h] == cur);
                    key = h < end ? (unsigned)units[h] : SENT;
                }
                if (lane == 0) outs[n] = (int)cur;
                n++;
            }
            __syncwarp();
            if (WRITE) { const int64_t out = (int64_t)pos[l] - 1; for (int t = lane; t < n; t += 32) idx[out + t] = (Ti)(outs[t] + 1); }
        }
        if (!WRITE) {
            long long rows = 0;
            if (DIM2) {
                for (int t = lane; t < n; t += 32) { const int u = (w == 1 && !DIM2) ? units[t] : outs[t]; rows += (long long)pi_spl[u + 1] - (long long)pi_spl[u]; }
                for (int d = 16; d > 0; d >>= 1) rows += __shfl_xor_sync(0xffffffffu, rows, d);
            }
            if (lane == 0) { cnt[l] = n; nvals[l] = (DIM2 ? rows : (long long)n) * (long long)w; }
            if (keep != nullptr) { const int *src = (w == 1 && !DIM2) ? units : outs; for (int t = lane; t < n; t += 32) keep[b + t] = src[t]; }
        }
        __syncwarp();
    }
}

// idx[pos[l] - 1 + t] = keep[first CSC entry of stripe l + t] + 1: the write pass when the count pass kept its merged units
template <typename Ti>
__global__ void __launch_bounds__(256) k_copy_units(const Ti *__restrict__ colptr, const Ti *__restrict__ phi_spl, int64_t L, const Ti *__restrict__ pos,
                                                    const int *__restrict__ keep, Ti *__restrict__ idx)
{
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t l = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; l < L; l += nwarps) {
        const int64_t b = (int64_t)colptr[(int64_t)phi_spl[l] - 1] - 1, out = (int64_t)pos[l] - 1;
        const int n = (int)((int64_t)pos[l + 1] - 1 - out);
        for (int t = lane; t < n; t += 32) idx[out + t] = (Ti)(keep[b + t] + 1);
    }
}

// col2stripe[j] = 0-based stripe of column j
template <typename Ti>
__global__ void k_col_to_stripe(const Ti *__restrict__ spl, int64_t L, int *__restrict__ c2s)
{
    const int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= L) return;
    for (int64_t j = (int64_t)spl[l] - 1; j < (int64_t)spl[l + 1] - 1; j++) c2s[j] = (int)l;
}

// Value scatter: 8 lanes per column walk its nonzeros; slot = binary search in the stripe's idx.
template <typename Ti, typename Tv, bool DIM2>
__global__ void __launch_bounds__(256) k_scatter_values(const Ti *__restrict__ colptr, const Ti *__restrict__ rowval,
                                                        const Tv *__restrict__ nzval, int64_t n,
                                                        const Ti *__restrict__ phi_spl, const int *__restrict__ c2s,
                                                        const Ti *__restrict__ pi_spl, const int *__restrict__ asg,
                                                        const int *__restrict__ brow, const Ti *__restrict__ pos,
                                                        const Ti *__restrict__ idx, const Ti *__restrict__ ofs,
                                                        Tv *__restrict__ val)
{
    constexpr int G = 8;
    const int64_t j = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const int lane = threadIdx.x % G;
    if (j >= n) return;
    const int l = c2s[j];
    const int64_t j0 = (int64_t)phi_spl[l] - 1;
    const int64_t w = (int64_t)phi_spl[l + 1] - 1 - j0;
    const int64_t dj = j - j0;
    const int64_t p0 = (int64_t)pos[l] - 1, p1 = (int64_t)pos[l + 1] - 1;
    const int64_t o0 = (int64_t)ofs[l] - 1;
    const int64_t b = (int64_t)colptr[j] - 1, e = (int64_t)colptr[j + 1] - 1;
    for (int64_t t = b + lane; t < e; t += G) {
        const int64_t r = (int64_t)rowval[t] - 1;
        const Ti key = DIM2 ? (Ti)(asg[r] + 1) : (Ti)(r + 1);
        int64_t lo = p0, hi = p1; // first slot with idx >= key (it is present by construction)
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            if (idx[mid] < key) lo = mid + 1; else hi = mid;
        }
        int64_t row;
        if (DIM2) row = (int64_t)brow[lo] + (r - ((int64_t)pi_spl[key - 1] - 1));
        else row = lo - p0;
        val[o0 + row * w + dj] = nzval[t];
    }
}

// ---- canonical -> compact layout ------------------------------------------------------------
template <typename Ti>
__global__ void k_desc_rows_1d(const Ti *__restrict__ idx, int64_t nidx, int *__restrict__ desc)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nidx) desc[t] = (int)((int64_t)idx[t] - 1);
}

// 2D, one thread per stripe: desc[Q] = first x index of block Q, brow[Q] = rows before it in the stripe
template <typename Ti>
__global__ void k_desc_blocks_2d(const Ti *__restrict__ pos, const Ti *__restrict__ idx, const Ti *__restrict__ pi_spl,
                                 int64_t L, int *__restrict__ desc, int *__restrict__ brow, long long *__restrict__ rows_out)
{
    const int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= L) return;
    int run = 0;
    for (int64_t Q = (int64_t)pos[l] - 1; Q < (int64_t)pos[l + 1] - 1; Q++) {
        const int64_t k = (int64_t)idx[Q] - 1;
        const int64_t i = (int64_t)pi_spl[k] - 1;
        desc[Q] = (int)i;
        brow[Q] = run;
        run += (int)((int64_t)pi_spl[k + 1] - 1 - i);
    }
    if (rows_out) rows_out[l] = run;
}

// 2D with non-uniform part heights: expand blocks to one descriptor per stored row
template <typename Ti>
__global__ void k_expand_rows_2d(const Ti *__restrict__ pos, const Ti *__restrict__ idx, const Ti *__restrict__ pi_spl,
                                 int64_t L, const int *__restrict__ rowbase, int *__restrict__ rowx)
{
    const int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= L) return;
    int64_t out = rowbase[l];
    for (int64_t Q = (int64_t)pos[l] - 1; Q < (int64_t)pos[l + 1] - 1; Q++) {
        const int64_t k = (int64_t)idx[Q] - 1;
        const int64_t i = (int64_t)pi_spl[k] - 1;
        const int64_t u = (int64_t)pi_spl[k + 1] - 1 - i;
        for (int64_t di = 0; di < u; di++) rowx[out++] = (int)(i + di);
    }
}

template <typename Ti, typename Tp>
__global__ void k_meta(const Ti *__restrict__ ofs, const Tp *__restrict__ first_desc, int sub1, const Ti *__restrict__ spl, int64_t L, StripeMeta *__restrict__ meta)
{
    const int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (l > L) return;
    StripeMeta s;
    s.ofs = (long long)ofs[l] - 1;
    s.pos = (int)((long long)first_desc[l] - sub1);
    s.col = (int)((long long)spl[l] - 1);
    meta[l] = s;
}

template <typename Ti>
__global__ void k_memory_cost(const Ti *__restrict__ spl, const Ti *__restrict__ pos, const Ti *__restrict__ ofs, int64_t L, int ndim, int tv_size, long long *__restrict__ cost)
{
    const int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= L) return;
    const long long ti = sizeof(Ti);
    const long long units = (long long)pos[l + 1] - (long long)pos[l];
    const long long nv = (long long)ofs[l + 1] - (long long)ofs[l];
    // costs.jl:10  3|Ti| + rows*(|Ti| + w|Tv|)   /   costs.jl:140  3|Ti| + sum_blocks(|Ti| + u*w*|Tv|)
    (void)spl; (void)ndim;
    cost[l] = 3 * ti + units * ti + nv * (long long)tv_size;
}

static inline unsigned nblk(int64_t n, int t);

// Kernel-body class of a stripe: must mirror the per-stripe dispatch of k_spmv_adj / k_spmv_fwd (spmv.cu):
// elements per load (from width and slab alignment) and vectors per row.
__device__ __forceinline__ int stripe_class(const StripeMeta a, const StripeMeta b, const int VE, const int rows_mode, const long long long_vals)
{
    const int w = b.col - a.col;
    if (w <= 0) return 0;
    if (b.ofs - a.ofs > long_vals) return 255; // a LONG stripe (a dense column group): last in the order, multiplied by one CTA each (k_spmv_adj_long)
    int epv_code, cpr;
    if ((w % VE) == 0 && (a.ofs % VE) == 0) { epv_code = 0; cpr = w / VE; }
    else if (VE == 4 && (w % 2) == 0 && (a.ofs % 2) == 0) { epv_code = 1; cpr = w / 2; }
    else { epv_code = 2; cpr = w; }
    return (1 + epv_code * 64 + (cpr > 62 ? 62 : cpr)) & 255;
}

// hist[0..255]: stripes per class; hist[256], hist[257]: min and max stripe width
__global__ void k_class_hist(const StripeMeta *__restrict__ meta, int64_t L, int VE, int rows_mode, long long long_vals, unsigned *__restrict__ hist)
{
    __shared__ unsigned sh[256];
    sh[threadIdx.x] = 0;
    __syncthreads();
    unsigned wmin = 0xffffffffu, wmax = 0;
    for (int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; l < L; l += (int64_t)gridDim.x * blockDim.x) {
        atomicAdd(&sh[stripe_class(meta[l], meta[l + 1], VE, rows_mode, long_vals) & 255], 1u);
        const unsigned w = (unsigned)(meta[l + 1].col - meta[l].col);
        wmin = w < wmin ? w : wmin;
        wmax = w > wmax ? w : wmax;
    }
    if (wmin != 0xffffffffu) { atomicMin(&hist[256], wmin); atomicMax(&hist[257], wmax); }
    __syncthreads();
    if (sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], sh[threadIdx.x]);
}

// order[cursor[class]++] = l.  Stripes are visited in ascending blocks, so neighbours of one class stay close.
__global__ void k_class_scatter(const StripeMeta *__restrict__ meta, int64_t L, int VE, int rows_mode, long long long_vals, unsigned *__restrict__ cursor, int *__restrict__ order)
{
    const int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= L) return;
    const int c = stripe_class(meta[l], meta[l + 1], VE, rows_mode, long_vals) & 255;
    order[atomicAdd(&cursor[c], 1u)] = (int)l;
}

// Group the stripes by kernel-body class when the matrix mixes several (variable widths): a warp then runs
// one body instead of serialising up to four.
int build_class_order(vbc_mat *A)
{
    const int64_t L = A->L;
    A->nclasses = 1;
    A->w_uniform = 0;
    A->n_long = 0;
    if (L < 1) return VBC_OK;
    // a stripe is LONG when it holds more than 2 K values and more than 16 times the average: one group of lanes would
    // still be streaming it long after every other stripe is done (dense column groups; dense ROWS in the transposed copy)
    long long long_vals = 16 * (A->nval / L + 1);
    if (long_vals < 2048) long_vals = 2048;
    { const char *e = getenv("VBC_LONG_STRIPE_VALUES"); if (e && atoll(e) > 0) long_vals = atoll(e); }
    cudaStream_t st = A->stream;
    const int VE = 16 / (int)vt_size(A->vt);
    unsigned *d_hist = nullptr;
    VBC_CUDA(cudaMalloc(&d_hist, 258 * sizeof(unsigned)));
    unsigned h[258];
    cudaError_t e = cudaMemsetAsync(d_hist, 0, 258 * sizeof(unsigned), st);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_hist + 256, 0xff, sizeof(unsigned), st); // running minimum
    if (e == cudaSuccess) {
        int64_t g = (L + 255) / 256;
        if (g > 2048) g = 2048;
        k_class_hist<<<(unsigned)g, 256, 0, st>>>(A->d_meta, L, VE, A->desc_mode == DESC_ROWS ? 1 : 0, long_vals, d_hist);
        A->launches++;
        e = cudaMemcpyAsync(h, d_hist, sizeof(h), cudaMemcpyDeviceToHost, st);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { cudaFree(d_hist); VBC_FAIL(VBC_ECUDA, "class histogram: %s", cudaGetErrorString(e)); }
    if (h[256] == h[257] && h[257] > 0) A->w_uniform = (int)h[257];
    A->has_unaligned = 0;
    A->opt_no_flat = getenv("VBC_NO_FLAT") != nullptr; // experiments: per-element bodies for unaligned stripes, as before round 2
    for (int c = 65; c < 192; c++) if (h[c]) A->has_unaligned = 1; // class codes 1 + epv_code * 64 + cpr with epv_code 1, 2: stripes that are not 16-byte aligned
    int ncls = 0;
    unsigned run = 0, cur[256];
    for (int c = 0; c < 256; c++) { cur[c] = run; run += h[c]; if (h[c] && c != 255) ncls++; }
    A->nclasses = ncls > 0 ? ncls : 1;
    A->n_long = (int64_t)h[255];
    if (ncls <= 1 && A->n_long == 0) { cudaFree(d_hist); return VBC_OK; }
    e = cudaMalloc(&A->d_order, sizeof(int) * (size_t)L);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_hist, cur, sizeof(cur), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
        k_class_scatter<<<nblk(L, 256), 256, 0, st>>>(A->d_meta, L, VE, A->desc_mode == DESC_ROWS ? 1 : 0, long_vals, d_hist, A->d_order);
        A->launches++;
        e = cudaStreamSynchronize(st);
    }
    cudaFree(d_hist);
    if (e != cudaSuccess) VBC_FAIL(VBC_ECUDA, "class order: %s", cudaGetErrorString(e));
    return VBC_OK;
}

static inline unsigned nblk(int64_t n, int t) { return (unsigned)((n + t - 1) / t > 0 ? (n + t - 1) / t : 1); }

// ---------------------------------------------------------------------------------------------
template <typename Ti>
static int finalize_t(vbc_mat *A, const void *h_pi_spl_v)
{
    cudaStream_t st = A->stream;
    const Ti *pos = (const Ti *)A->d_pos, *idx = (const Ti *)A->d_idx, *ofs = (const Ti *)A->d_ofs;
    const Ti *phi = (const Ti *)A->d_phi_spl, *pi = (const Ti *)A->d_pi_spl;
    const int64_t L = A->L;
    if (A->nidx >= (1LL << 31) || A->m >= (1LL << 31) || A->n >= (1LL << 31))
        VBC_FAIL(VBC_ELIMIT, "device layout uses 32-bit row/column fields: nidx=%lld m=%lld n=%lld", (long long)A->nidx, (long long)A->m, (long long)A->n);
    VBC_CUDA(cudaMalloc(&A->d_meta, sizeof(StripeMeta) * (size_t)(L + 1)));
    if (A->ndim == 1) {
        A->desc_mode = DESC_ROWS;
        A->ndesc = A->nidx;
        VBC_CUDA(cudaMalloc(&A->d_desc, sizeof(int) * (size_t)(A->ndesc > 0 ? A->ndesc : 1)));
        if (A->nidx > 0) { k_desc_rows_1d<Ti><<<nblk(A->nidx, 256), 256, 0, st>>>(idx, A->nidx, A->d_desc); A->launches++; }
        k_meta<Ti, Ti><<<nblk(L + 1, 256), 256, 0, st>>>(ofs, pos, 1, phi, L, A->d_meta); A->launches++;
        VBC_CUDA(cudaGetLastError());
        return VBC_OK;
    }
    // 2D: uniform part heights?  (all parts u0 high, the last one possibly shorter)
    const Ti *hp = (const Ti *)h_pi_spl_v;
    bool uniform = true;
    int64_t u0 = A->K > 0 ? (int64_t)hp[1] - (int64_t)hp[0] : 1;
    for (int64_t k = 0; k < A->K && uniform; k++) {
        const int64_t u = (int64_t)hp[k + 1] - (int64_t)hp[k];
        if (k + 1 < A->K ? (u != u0) : (u > u0 || u < 1)) uniform = false;
    }
    if (u0 < 1) uniform = false;
    A->u0 = uniform ? (int)u0 : 1;
    const int64_t nb = A->nidx;
    VBC_CUDA(cudaMalloc(&A->d_brow, sizeof(int) * (size_t)(nb > 0 ? nb : 1)));
    struct Scratch { // freed on every exit path
        int *bdesc = nullptr, *rowbase = nullptr;
        long long *rows = nullptr, *tmp = nullptr;
        ~Scratch() { cudaFree(bdesc); cudaFree(rowbase); cudaFree(rows); cudaFree(tmp); }
    } sc;
    int *&d_bdesc = sc.bdesc;
    long long *&d_rows = sc.rows;
    VBC_CUDA(cudaMalloc(&d_bdesc, sizeof(int) * (size_t)(nb > 0 ? nb : 1)));
    if (!uniform) VBC_CUDA(cudaMalloc(&d_rows, sizeof(long long) * (size_t)(L > 0 ? L : 1)));
    if (L > 0) { k_desc_blocks_2d<Ti><<<nblk(L, 128), 128, 0, st>>>(pos, idx, pi, L, d_bdesc, A->d_brow, d_rows); A->launches++; }
    VBC_CUDA(cudaGetLastError());
    if (uniform) {
        A->desc_mode = DESC_BLOCKS;
        A->d_desc = d_bdesc;
        d_bdesc = nullptr; // ownership moved to the matrix
        A->ndesc = nb;
        k_meta<Ti, Ti><<<nblk(L + 1, 256), 256, 0, st>>>(ofs, pos, 1, phi, L, A->d_meta); A->launches++;
        VBC_CUDA(cudaGetLastError());
        return VBC_OK;
    }
    // expanded rows
    A->desc_mode = DESC_ROWS;
    int *&d_rowbase = sc.rowbase;
    long long *&d_tmp = sc.tmp;
    long long total = 0;
    VBC_CUDA(cudaMalloc(&d_rowbase, sizeof(int) * (size_t)(L + 1)));
    VBC_CUDA(cudaMalloc(&d_tmp, sizeof(long long) * (size_t)scan_tmp_elems(L)));
    int rc = exclusive_scan<int>(d_rows, d_rowbase, L, 0, d_tmp, &total, st, &A->launches);
    if (rc == VBC_OK && total >= (1LL << 31)) { set_error("expanded row count %lld exceeds 32-bit descriptor space", total); rc = VBC_ELIMIT; }
    if (rc == VBC_OK) {
        A->ndesc = total;
        if (cudaMalloc(&A->d_desc, sizeof(int) * (size_t)(total > 0 ? total : 1)) != cudaSuccess) { set_error("cudaMalloc(rowx)"); rc = VBC_ENOMEM; }
    }
    if (rc == VBC_OK) {
        if (L > 0) { k_expand_rows_2d<Ti><<<nblk(L, 128), 128, 0, st>>>(pos, idx, pi, L, d_rowbase, A->d_desc); A->launches++; }
        k_meta<Ti, int><<<nblk(L + 1, 256), 256, 0, st>>>(ofs, d_rowbase, 0, phi, L, A->d_meta); A->launches++;
        if (cudaStreamSynchronize(st) != cudaSuccess || cudaGetLastError() != cudaSuccess) { set_error("finalize (expanded rows) kernels failed"); rc = VBC_ECUDA; }
    }
    return rc;
}

int finalize_layout(vbc_mat *A, const void *h_pi_spl)
{
    VBC_TRY(A->it == VBC_I64 ? finalize_t<int64_t>(A, h_pi_spl) : finalize_t<int32_t>(A, h_pi_spl));
    return build_class_order(A);
}

template <typename Ti, typename Tv>
static int pack_t(vbc_mat *A, const Ti *colptr, const Ti *rowval, const Tv *nzval)
{
    cudaStream_t st = A->stream;
    const int64_t L = A->L, K = A->K, m = A->m, n = A->n;
    const bool d2 = A->ndim == 2;
    const Ti *phi = (const Ti *)A->d_phi_spl, *pi = (const Ti *)A->d_pi_spl;

    struct Tmp {
        int *err = nullptr, *asg = nullptr, *c2s = nullptr, *keep = nullptr;
        long long *cnt = nullptr, *nv = nullptr, *scan = nullptr;
        ~Tmp() { cudaFree(err); cudaFree(asg); cudaFree(c2s); cudaFree(keep); cudaFree(cnt); cudaFree(nv); cudaFree(scan); }
    } t;
    VBC_CUDA(cudaMalloc(&t.err, sizeof(int)));
    VBC_CUDA(cudaMemsetAsync(t.err, 0, sizeof(int), st));
    // partition validation; `@assert w <= W` (constructors_1DVBC.jl:46, VBC :65), `@assert u <= U` (VBC :58-60)
    k_check_spl<Ti><<<nblk(L + 1, 256), 256, 0, st>>>(phi, L, n, A->W, PERR_W, t.err); A->launches++;
    if (d2) { k_check_spl<Ti><<<nblk(K + 1, 256), 256, 0, st>>>(pi, K, m, A->U, PERR_U, t.err); A->launches++; }
    int herr = 0;
    int64_t nnz_in = 0; // entries of the input CSC
    VBC_CUDA(cudaMemcpyAsync(&herr, t.err, sizeof(int), cudaMemcpyDeviceToHost, st));
    VBC_CUDA(cudaStreamSynchronize(st));
    if (herr == 0 && n > 0) { // CSC arrays (the partitions are valid, so the kernels below index them safely)
        int64_t nnz_h = 0;
        { Ti last; VBC_CUDA(cudaMemcpy(&last, colptr + n, sizeof(Ti), cudaMemcpyDeviceToHost)); nnz_h = (int64_t)last - 1; }
        if (nnz_h < 0) VBC_FAIL(VBC_EARG, "ArgumentError: colptr[end] < 1");
        nnz_in = nnz_h;
        k_check_csc<Ti><<<nblk(n * 8, 256), 256, 0, st>>>(colptr, rowval, n, m, nnz_h, t.err); A->launches++;
        VBC_CUDA(cudaMemcpyAsync(&herr, t.err, sizeof(int), cudaMemcpyDeviceToHost, st));
        VBC_CUDA(cudaStreamSynchronize(st));
        if (herr == PERR_CSC) VBC_FAIL(VBC_EARG, "ArgumentError: malformed SparseMatrixCSC (colptr must start at 1 and be nondecreasing, rowval must lie in 1:%lld and ascend strictly within each column)", (long long)m);
    }
    if (herr == PERR_SPL) VBC_FAIL(VBC_EARG, "partition is not a SplitPartition of the matrix dimension (spl[1]==1, nondecreasing, spl[end]==dim+1)");
    if (herr == PERR_W) VBC_FAIL(VBC_ELIMIT, "AssertionError: w <= W (a stripe is wider than W=%d)", A->W);
    if (herr == PERR_U) VBC_FAIL(VBC_ELIMIT, "AssertionError: u <= U (a row part is taller than U=%d)", A->U);

    VBC_CUDA(cudaMalloc(&t.cnt, sizeof(long long) * (size_t)(L > 0 ? L : 1)));
    VBC_CUDA(cudaMalloc(&t.nv, sizeof(long long) * (size_t)(L > 0 ? L : 1)));
    VBC_CUDA(cudaMalloc(&t.scan, sizeof(long long) * (size_t)scan_tmp_elems(L)));
    if (d2) {
        VBC_CUDA(cudaMalloc(&t.asg, sizeof(int) * (size_t)(m > 0 ? m : 1)));
        if (m > 0) { k_build_map<Ti><<<nblk(m, 256), 256, 0, st>>>(pi, K, m, t.asg); A->launches++; }
    }
    // (1) count.  Warp-per-stripe merge staged in shared memory when the longest stripe segment fits (2 x cap ints per warp),
    // else one thread per stripe straight from global memory.
    int cap = 0;
    if (L > 0) {
        unsigned long long *d_mx = nullptr, h_mx = 0;
        VBC_CUDA(cudaMalloc(&d_mx, sizeof(unsigned long long)));
        cudaError_t ce = cudaMemsetAsync(d_mx, 0, sizeof(unsigned long long), st);
        if (ce == cudaSuccess) { k_max_seglen<Ti><<<(unsigned)(nblk(L, 256) > 1024 ? 1024 : nblk(L, 256)), 256, 0, st>>>(colptr, phi, L, d_mx); A->launches++; ce = cudaMemcpyAsync(&h_mx, d_mx, sizeof(h_mx), cudaMemcpyDeviceToHost, st); }
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
        cudaFree(d_mx);
        if (ce != cudaSuccess) VBC_FAIL(VBC_ECUDA, "pack: %s", cudaGetErrorString(ce));
        for (int c = 256; c <= 2048; c *= 2)
            if (h_mx <= (unsigned long long)c) { cap = c; break; }
    }
    const size_t merge_smem_bytes = (size_t)8 * 2 * cap * sizeof(int);
    if (cap > 0 && merge_smem_bytes > 48 * 1024) {
        cudaError_t ce = d2 ? cudaFuncSetAttribute(k_merge_stripes_warp<Ti, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)merge_smem_bytes)
                            : cudaFuncSetAttribute(k_merge_stripes_warp<Ti, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)merge_smem_bytes);
        if (ce == cudaSuccess) ce = d2 ? cudaFuncSetAttribute(k_merge_stripes_warp<Ti, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)merge_smem_bytes)
                                        : cudaFuncSetAttribute(k_merge_stripes_warp<Ti, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)merge_smem_bytes);
        if (ce != cudaSuccess) { cudaGetLastError(); cap = 0; }
    }
    const unsigned warp_grid = (unsigned)std::min<int64_t>((L + 7) / 8 > 0 ? (L + 7) / 8 : 1, (int64_t)A->sm_count * 8);
    if (L > 0) {
        if (cap > 0) {
            // scratch for the merged units (4 bytes per CSC entry, freed with the other temporaries); without it the write pass merges again
            if (nnz_in > 0 && cudaMalloc(&t.keep, sizeof(int) * (size_t)nnz_in) != cudaSuccess) { cudaGetLastError(); t.keep = nullptr; }
            if (d2) k_merge_stripes_warp<Ti, true, false><<<warp_grid, 256, merge_smem_bytes, st>>>(colptr, rowval, phi, L, pi, t.asg, t.cnt, t.nv, nullptr, nullptr, cap, t.err, t.keep);
            else    k_merge_stripes_warp<Ti, false, false><<<warp_grid, 256, merge_smem_bytes, st>>>(colptr, rowval, phi, L, pi, t.asg, t.cnt, t.nv, nullptr, nullptr, cap, t.err, t.keep);
        } else {
            if (d2) k_merge_stripes<Ti, true, false><<<nblk(L, 128), 128, 0, st>>>(colptr, rowval, phi, L, pi, t.asg, t.cnt, t.nv, nullptr, nullptr, t.err);
            else    k_merge_stripes<Ti, false, false><<<nblk(L, 128), 128, 0, st>>>(colptr, rowval, phi, L, pi, t.asg, t.cnt, t.nv, nullptr, nullptr, t.err);
        }
        A->launches++;
    }
    VBC_CUDA(cudaGetLastError());
    // (2) scans -> pos, ofs (1-based)
    VBC_CUDA(cudaMalloc(&A->d_pos, sizeof(Ti) * (size_t)(L + 1)));
    VBC_CUDA(cudaMalloc(&A->d_ofs, sizeof(Ti) * (size_t)(L + 1)));
    long long nidx = 0, nval = 0;
    VBC_TRY(exclusive_scan<Ti>(t.cnt, (Ti *)A->d_pos, L, 1, t.scan, &nidx, st, &A->launches));
    VBC_TRY(exclusive_scan<Ti>(t.nv, (Ti *)A->d_ofs, L, 1, t.scan, &nval, st, &A->launches));
    VBC_CUDA(cudaMemcpyAsync(&herr, t.err, sizeof(int), cudaMemcpyDeviceToHost, st));
    VBC_CUDA(cudaStreamSynchronize(st));
    if (herr == PERR_WMAX) VBC_FAIL(VBC_ELIMIT, "a stripe is wider than the pack kernel's limit of %d columns", WMAX);
    if (sizeof(Ti) == 4 && (nval + 1 > 0x7fffffffLL || nidx + 1 > 0x7fffffffLL))
        VBC_FAIL(VBC_ELIMIT, "packed sizes overflow Ti=Int32 (nidx=%lld nval=%lld)", nidx, nval);
    A->nidx = nidx;
    A->nval = nval;
    // (3) idx
    const size_t pad = 64; // values; lets 128-bit loads touch the tail safely
    VBC_CUDA(cudaMalloc(&A->d_idx, sizeof(Ti) * (size_t)(nidx > 0 ? nidx : 1)));
    VBC_CUDA(cudaMalloc(&A->d_val, sizeof(Tv) * ((size_t)nval + pad)));
    VBC_CUDA(cudaMemsetAsync(A->d_val, 0, sizeof(Tv) * ((size_t)nval + pad), st));
    if (L > 0) {
        if (cap > 0 && t.keep != nullptr) {
            k_copy_units<Ti><<<warp_grid, 256, 0, st>>>(colptr, phi, L, (const Ti *)A->d_pos, t.keep, (Ti *)A->d_idx);
        } else if (cap > 0) {
            if (d2) k_merge_stripes_warp<Ti, true, true><<<warp_grid, 256, merge_smem_bytes, st>>>(colptr, rowval, phi, L, pi, t.asg, nullptr, nullptr, (const Ti *)A->d_pos, (Ti *)A->d_idx, cap, t.err);
            else    k_merge_stripes_warp<Ti, false, true><<<warp_grid, 256, merge_smem_bytes, st>>>(colptr, rowval, phi, L, pi, t.asg, nullptr, nullptr, (const Ti *)A->d_pos, (Ti *)A->d_idx, cap, t.err);
        } else {
            if (d2) k_merge_stripes<Ti, true, true><<<nblk(L, 128), 128, 0, st>>>(colptr, rowval, phi, L, pi, t.asg, nullptr, nullptr, (const Ti *)A->d_pos, (Ti *)A->d_idx, t.err);
            else    k_merge_stripes<Ti, false, true><<<nblk(L, 128), 128, 0, st>>>(colptr, rowval, phi, L, pi, t.asg, nullptr, nullptr, (const Ti *)A->d_pos, (Ti *)A->d_idx, t.err);
        }
        A->launches++;
    }
    VBC_CUDA(cudaGetLastError());
    // compact layout (also yields brow for the 2D scatter)
    void *h_pi = nullptr;
    if (d2) {
        h_pi = malloc(sizeof(Ti) * (size_t)(K + 1));
        if (!h_pi) VBC_FAIL(VBC_ENOMEM, "host malloc");
        cudaError_t ce = cudaMemcpyAsync(h_pi, pi, sizeof(Ti) * (size_t)(K + 1), cudaMemcpyDeviceToHost, st);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
        if (ce != cudaSuccess) { free(h_pi); VBC_FAIL(VBC_ECUDA, "copy of pi_spl to host failed: %s", cudaGetErrorString(ce)); }
    }
    int rc = finalize_layout(A, h_pi);
    free(h_pi);
    VBC_TRY(rc);
    // (4) values
    if (n > 0 && L > 0) {
        VBC_CUDA(cudaMalloc(&t.c2s, sizeof(int) * (size_t)n));
        k_col_to_stripe<Ti><<<nblk(L, 256), 256, 0, st>>>(phi, L, t.c2s); A->launches++;
        const unsigned g = nblk(n * 8, 256);
        if (d2) k_scatter_values<Ti, Tv, true><<<g, 256, 0, st>>>(colptr, rowval, nzval, n, phi, t.c2s, pi, t.asg, A->d_brow, (const Ti *)A->d_pos, (const Ti *)A->d_idx, (const Ti *)A->d_ofs, (Tv *)A->d_val);
        else    k_scatter_values<Ti, Tv, false><<<g, 256, 0, st>>>(colptr, rowval, nzval, n, phi, t.c2s, pi, t.asg, A->d_brow, (const Ti *)A->d_pos, (const Ti *)A->d_idx, (const Ti *)A->d_ofs, (Tv *)A->d_val);
        A->launches++;
    }
    VBC_CUDA(cudaGetLastError());
    VBC_CUDA(cudaStreamSynchronize(st));
    return VBC_OK;
}

int pack_from_device_csc(vbc_mat *A, const void *c, const void *r, const void *v)
{
    if (A->it == VBC_I64) {
        if (vt_size(A->vt) == 8) return pack_t<int64_t, double>(A, (const int64_t *)c, (const int64_t *)r, (const double *)v);
        return pack_t<int64_t, float>(A, (const int64_t *)c, (const int64_t *)r, (const float *)v);
    }
    if (vt_size(A->vt) == 8) return pack_t<int32_t, double>(A, (const int32_t *)c, (const int32_t *)r, (const double *)v);
    return pack_t<int32_t, float>(A, (const int32_t *)c, (const int32_t *)r, (const float *)v);
}

int memory_cost_device(const vbc_mat *A, int64_t *h_cost, int64_t *row_term)
{
    const int64_t L = A->L;
    if (row_term) *row_term = A->ndim == 2 ? A->K * (int64_t)it_size(A->it) : 0;
    if (L == 0) return VBC_OK;
    long long *d = nullptr;
    VBC_CUDA(cudaMalloc(&d, sizeof(long long) * (size_t)L));
    if (A->it == VBC_I64) k_memory_cost<int64_t><<<nblk(L, 256), 256, 0, A->stream>>>((const int64_t *)A->d_phi_spl, (const int64_t *)A->d_pos, (const int64_t *)A->d_ofs, L, A->ndim, (int)vt_size(A->vt), d);
    else k_memory_cost<int32_t><<<nblk(L, 256), 256, 0, A->stream>>>((const int32_t *)A->d_phi_spl, (const int32_t *)A->d_pos, (const int32_t *)A->d_ofs, L, A->ndim, (int)vt_size(A->vt), d);
    cudaError_t ce = cudaMemcpyAsync(h_cost, d, sizeof(long long) * (size_t)L, cudaMemcpyDeviceToHost, A->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(A->stream);
    cudaFree(d);
    if (ce != cudaSuccess) VBC_FAIL(VBC_ECUDA, "memory_cost: %s", cudaGetErrorString(ce));
    return VBC_OK;
}

} // namespace vbc
