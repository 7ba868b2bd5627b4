// fwdt.cu -- owner-computes forward multiply  y <- alpha * A x + beta * y  through a transposed unit index.
//
// The reference's forward `mul!` (multiply_1DVBC.jl:9-83, multiply_VBC.jl:3-87) is a serial scatter,
// y[idx[Q]] += ...; the straightforward GPU version of it (k_spmv_fwd, spmv.cu) needs one fp atomic per
// stored row and is bound by atomic throughput (3.3 TB/s on configs[1]).  Here the scatter is turned
// around at first use: a device-built index lists, for every destination row (1D / expanded rows) or row
// part (2D blocks), the units that write to it -- 16 B per unit {value offset, stripe column, stripe
// width}.  A group of SG lanes then owns one destination: it streams its units' values with the same
// 16-byte coalesced loads as the adjoint kernel, multiplies by the stripe's x segment, reduces with
// shuffles and stores y once with alpha/beta -- no atomics, no separate y <- beta*y pass, deterministic
// for a given handle.  Extra traffic: 16 B per unit (12.5 % for 4x4 Float64 blocks).
//
// Blocks mode keeps per-lane accumulators for a fixed (row-in-block, column-vector) and therefore needs
// one stripe width for the whole matrix; rows mode takes any widths.  Everything else falls back to the
// atomic kernel (VBC_OPT_FWD_MODE = 1 forces that).
#include <new>

#include "common.cuh"
#include "scan.cuh"
#include "walk.cuh"

namespace vbc {

struct __align__(16) TRec {
    long long vofs; // element offset of the unit's first value in val
    int col;        // first column of the unit's stripe
    int w;          // width of that stripe
};

__device__ __forceinline__ TRec ld_rec(const TRec *p)
{
    const int4 t = __ldcs(reinterpret_cast<const int4 *>(p));
    TRec r;
    r.vofs = (long long)(((unsigned long long)(unsigned)t.y << 32) | (unsigned)t.x);
    r.col = t.z;
    r.w = t.w;
    return r;
}

// destination key of unit t: rows mode -> the row; blocks mode -> the row part
template <int MODE> __device__ __forceinline__ int unit_key(const int *__restrict__ desc, const int t, const int u0, const int log2u)
{
    const int i = __ldg(desc + t);
    if (MODE == DESC_ROWS) return i;
    return log2u >= 0 ? (i >> log2u) : i / u0;
}

template <int MODE>
__global__ void __launch_bounds__(128) k_t_count(const StripeMeta *__restrict__ meta, const int *__restrict__ desc, const int L,
                                                  const int u0, const int log2u, unsigned long long *__restrict__ cnt)
{
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= L) return;
    const StripeMeta a = meta[l], b = meta[l + 1];
    if (b.col - a.col <= 0) return;
    for (int t = a.pos; t < b.pos; t++) atomicAdd(&cnt[unit_key<MODE>(desc, t, u0, log2u)], 1ull);
}

template <int MODE>
__global__ void __launch_bounds__(128) k_t_fill(const StripeMeta *__restrict__ meta, const int *__restrict__ desc, const int L,
                                                 const int u0, const int log2u, const int *__restrict__ tptr,
                                                 unsigned *__restrict__ cursor, TRec *__restrict__ rec)
{
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= L) return;
    const StripeMeta a = meta[l], b = meta[l + 1];
    const int w = b.col - a.col;
    if (w <= 0) return;
    const long long unit_vals = (MODE == DESC_ROWS) ? w : (long long)u0 * w; // only a stripe's last block can be shorter
    for (int t = a.pos; t < b.pos; t++) {
        const int key = unit_key<MODE>(desc, t, u0, log2u);
        TRec r;
        r.vofs = a.ofs + (long long)(t - a.pos) * unit_vals;
        r.col = a.col;
        r.w = w;
        rec[tptr[key] + atomicAdd(&cursor[key], 1u)] = r;
    }
}

// The atomic cursor leaves a key's units in arrival order; sorting them by stripe column (a handful per key) makes the
// index -- and the summation order of the owner-computes forward multiply -- the same on every run.
__global__ void __launch_bounds__(256) k_t_sort(const int *__restrict__ tptr, TRec *__restrict__ rec, const int nkeys)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nkeys) return;
    const int t0 = tptr[k], t1 = tptr[k + 1];
    for (int i = t0 + 1; i < t1; i++) { // insertion sort
        const TRec r = rec[i];
        int j = i - 1;
        while (j >= t0 && rec[j].col > r.col) { rec[j + 1] = rec[j]; j--; }
        rec[j + 1] = r;
    }
}

// Transposed copy of a matrix of uniform u0 x w0 blocks: row part k becomes "stripe" k (u columns wide: the part's rows)
// holding its blocks transposed, w0 x u row-major, in ascending stripe-column order.  The forward multiply y = A x is then
// the ADJOINT multiply of this copy -- the same streaming owner-computes kernel, at the same fraction of the HBM roofline
// as mul!(y, B', x) -- at the price of a second copy of the values.
template <typename Tv>
__global__ void __launch_bounds__(256) k_t_meta(const int *__restrict__ tptr, const int nkeys, const int u0, const int w0, const long long m,
                                                StripeMeta *__restrict__ meta2)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > nkeys) return;
    // every part before the last is u0 high, so the value offset is closed-form: blocks before * (w0 * u0)
    StripeMeta s;
    s.ofs = (long long)tptr[k] * w0 * u0;
    if (k == nkeys && nkeys > 0) { // the last part may be shorter: its blocks are w0 x u_last
        const long long u_last = m - (long long)(nkeys - 1) * u0;
        s.ofs = (long long)tptr[k - 1] * w0 * u0 + (long long)(tptr[k] - tptr[k - 1]) * w0 * u_last;
    }
    s.pos = tptr[k];
    const long long c = (long long)k * u0;
    s.col = (int)(c < m ? c : m);
    meta2[k] = s;
}

template <typename Tv>
__global__ void __launch_bounds__(256) k_t_copy(const int *__restrict__ tptr, const TRec *__restrict__ rec, const Tv *__restrict__ val, const int nkeys,
                                                const int u0, const int w0, const long long m, int *__restrict__ desc2, Tv *__restrict__ val2)
{
    // one warp per row part; lanes walk the elements of its transposed blocks, consecutive lanes -> consecutive outputs
    const int k = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (k >= nkeys) return;
    const int t0 = tptr[k], t1 = tptr[k + 1];
    const long long i0 = (long long)k * u0;
    const int u = (int)((m - i0) < u0 ? (m - i0) : u0);
    const long long obase = (long long)t0 * w0 * u0;
    const int per = w0 * u;
    for (int t = t0; t < t1; t++) {
        const TRec r = rec[t];
        if (lane == 0) desc2[t] = r.col;
        Tv *out = val2 + obase + (long long)(t - t0) * per;
        for (int e = lane; e < per; e += 32) {
            const int dj = e / u, di = e - dj * u; // transposed block: row dj (a column of A), column di (a row of the part)
            out[e] = val[r.vofs + (long long)di * w0 + dj];
        }
    }
}

// ---- rows mode: SG lanes per destination row, units are w-wide row segments ---------------------
template <typename Tv, int SG>
__global__ void __launch_bounds__(256) k_fwdt_rows(const int *__restrict__ tptr, const TRec *__restrict__ rec, const Tv *__restrict__ val,
                                                    const Tv *__restrict__ x, Tv *__restrict__ y, const int nkeys, const Tv alpha, const Tv beta)
{
    constexpr int VE = 16 / (int)sizeof(Tv);
    const int lane = threadIdx.x % SG;
    unsigned gmask = 0xffffffffu;
    if constexpr (SG < 32) gmask = ((1u << SG) - 1u) << (((threadIdx.x & 31) / SG) * SG);
    const int ngroups = (int)((gridDim.x * blockDim.x) / SG);
    for (int i = (int)((blockIdx.x * blockDim.x + threadIdx.x) / SG); i < nkeys; i += ngroups) {
        const int t0 = __ldg(tptr + i), t1 = __ldg(tptr + i + 1);
        Tv acc = (Tv)0;
        constexpr int UNR = 4;
        for (int t = t0; t < t1; t += UNR) {
            TRec r[UNR];
#pragma unroll
            for (int k = 0; k < UNR; k++) {
                r[k].w = 0;
                if (t + k < t1) r[k] = ld_rec(rec + t + k);
            }
#pragma unroll
            for (int k = 0; k < UNR; k++) {
                const int w = r[k].w;
                if (w == 0) continue;
                const Tv *vp = val + r[k].vofs;
                const Tv *xp = x + r[k].col;
                if ((w % VE) == 0 && (r[k].vofs % VE) == 0) {
                    for (int e = lane * VE; e < w; e += SG * VE) {
                        Tv v[VE];
                        if constexpr (VE == 2) { const double2 q = __ldcs(reinterpret_cast<const double2 *>(vp + e)); v[0] = (Tv)q.x; v[1] = (Tv)q.y; }
                        else { const float4 q = __ldcs(reinterpret_cast<const float4 *>(vp + e)); v[0] = (Tv)q.x; v[1] = (Tv)q.y; v[2] = (Tv)q.z; v[3] = (Tv)q.w; }
#pragma unroll
                        for (int j = 0; j < VE; j++) acc = fma(v[j], __ldg(xp + e + j), acc);
                    }
                } else {
                    for (int e = lane; e < w; e += SG) acc = fma(__ldcs(vp + e), __ldg(xp + e), acc);
                }
            }
        }
#pragma unroll
        for (int d = 1; d < SG; d <<= 1) acc += __shfl_xor_sync(gmask, acc, d, SG);
        if (lane == 0) y[i] = (beta == (Tv)0) ? alpha * acc : alpha * acc + beta * y[i];
    }
}

// ---- rows mode with the transposed copy: row i owns a contiguous run of its w0-wide segments (val2) and their first columns
// (desc2).  A 16-byte vector never leaves a segment (w0 is a multiple of the vector), so lane v multiplies one vector of values
// with one vector of x: a plain dot product per row, SG lanes per row, fully coalesced -- y[i] is stored once, no atomics.
template <typename Tv, int SG>
__global__ void __launch_bounds__(256) k_fwdc_rows(const StripeMeta *__restrict__ meta2, const int *__restrict__ desc2, const Tv *__restrict__ val2,
                                                    const Tv *__restrict__ x, Tv *__restrict__ y, const int nrows, const int w0, const int log2vps,
                                                    const Tv alpha, const Tv beta)
{
    constexpr int VE = 16 / (int)sizeof(Tv);
    const int lane = threadIdx.x % SG;
    unsigned gmask = 0xffffffffu;
    if constexpr (SG < 32) gmask = ((1u << SG) - 1u) << (((threadIdx.x & 31) / SG) * SG);
    const int ngroups = (int)((gridDim.x * blockDim.x) / SG);
    const int vps = w0 / VE; // vectors per segment
    for (int i = (int)((blockIdx.x * blockDim.x + threadIdx.x) / SG); i < nrows; i += ngroups) {
        const StripeMeta a = ld_meta(meta2 + i), b = ld_meta(meta2 + i + 1);
        const int nvec = (int)((b.ofs - a.ofs) / VE);
        const Tv *vp = val2 + a.ofs;
        const int *dp = desc2 + a.pos;
        Tv acc[VE];
#pragma unroll
        for (int e = 0; e < VE; e++) acc[e] = (Tv)0;
        constexpr int UNR = 4;
        for (int v = lane; v < nvec; v += UNR * SG) {
            Tv vv[UNR][VE], xx[UNR][VE];
            int col[UNR];
#pragma unroll
            for (int k = 0; k < UNR; k++) {
                const int vk = v + k * SG;
                const bool ok = vk < nvec;
                const int seg = log2vps >= 0 ? (vk >> log2vps) : vk / vps;
                col[k] = ok ? __ldg(dp + seg) + (vk - seg * vps) * VE : 0;
                if (ok) {
                    if constexpr (VE == 2) { const double2 q = __ldcs(reinterpret_cast<const double2 *>(vp + (long long)vk * VE)); vv[k][0] = (Tv)q.x; vv[k][1] = (Tv)q.y; }
                    else { const float4 q = __ldcs(reinterpret_cast<const float4 *>(vp + (long long)vk * VE)); vv[k][0] = (Tv)q.x; vv[k][1] = (Tv)q.y; vv[k][2] = (Tv)q.z; vv[k][3] = (Tv)q.w; }
                } else {
#pragma unroll
                    for (int e = 0; e < VE; e++) vv[k][e] = (Tv)0;
                }
            }
#pragma unroll
            for (int k = 0; k < UNR; k++) {
                const bool ok = v + k * SG < nvec;
                if (ok && (col[k] % VE) == 0) { // aligned stripe start: one 16-byte x load
                    if constexpr (VE == 2) { const double2 q = __ldg(reinterpret_cast<const double2 *>(x + col[k])); xx[k][0] = (Tv)q.x; xx[k][1] = (Tv)q.y; }
                    else { const float4 q = __ldg(reinterpret_cast<const float4 *>(x + col[k])); xx[k][0] = (Tv)q.x; xx[k][1] = (Tv)q.y; xx[k][2] = (Tv)q.z; xx[k][3] = (Tv)q.w; }
                } else {
#pragma unroll
                    for (int e = 0; e < VE; e++) xx[k][e] = ok ? __ldg(x + col[k] + e) : (Tv)0;
                }
            }
#pragma unroll
            for (int k = 0; k < UNR; k++)
#pragma unroll
                for (int e = 0; e < VE; e++) acc[e] = fma(vv[k][e], xx[k][e], acc[e]);
        }
        Tv sum = (Tv)0;
#pragma unroll
        for (int e = 0; e < VE; e++) sum += acc[e];
#pragma unroll
        for (int d = 1; d < SG; d <<= 1) sum += __shfl_xor_sync(gmask, sum, d, SG);
        if (lane == 0) y[i] = (beta == (Tv)0) ? alpha * sum : alpha * sum + beta * y[i];
    }
}

// ---- blocks mode: one stripe width w0 and part height u0 for the whole matrix; SG = lanes per part ---
// lane v of the group holds vector v of every block (row-in-block v / cpr, column-vector v % cpr).
template <typename Tv, int SG>
__global__ void __launch_bounds__(256) k_fwdt_blocks(const int *__restrict__ tptr, const TRec *__restrict__ rec, const Tv *__restrict__ val,
                                                      const Tv *__restrict__ x, Tv *__restrict__ y, const int nkeys, const int u0, const int w0,
                                                      const long long m, const Tv alpha, const Tv beta)
{
    constexpr int VE = 16 / (int)sizeof(Tv);
    const int lane = threadIdx.x % SG;
    unsigned gmask = 0xffffffffu;
    if constexpr (SG < 32) gmask = ((1u << SG) - 1u) << (((threadIdx.x & 31) / SG) * SG);
    const int ngroups = (int)((gridDim.x * blockDim.x) / SG);
    const int cpr = w0 / VE;          // vectors per block row
    const int nvec = u0 * cpr;        // vectors per full block
    for (int k = (int)((blockIdx.x * blockDim.x + threadIdx.x) / SG); k < nkeys; k += ngroups) {
        const int t0 = __ldg(tptr + k), t1 = __ldg(tptr + k + 1);
        const long long i0 = (long long)k * u0;
        const int u = (int)((m - i0) < u0 ? (m - i0) : u0); // the last part may be shorter
        for (int vb = 0; vb < nvec; vb += SG) { // SG >= nvec in the tuned cases: one pass
            const int v = vb + lane;
            const int di = small_div(v, cpr), cv = v - di * cpr;
            const bool mine = v < u * cpr;
            Tv acc[VE];
#pragma unroll
            for (int j = 0; j < VE; j++) acc[j] = (Tv)0;
            constexpr int UNR = 4;
            for (int t = t0; t < t1; t += UNR) {
                TRec r[UNR];
                Tv vv[UNR][VE], xx[UNR][VE];
#pragma unroll
                for (int q = 0; q < UNR; q++) {
                    r[q].w = 0;
                    if (t + q < t1) r[q] = ld_rec(rec + t + q);
                }
#pragma unroll
                for (int q = 0; q < UNR; q++) {
                    const bool ok = mine && r[q].w != 0;
#pragma unroll
                    for (int j = 0; j < VE; j++) { vv[q][j] = (Tv)0; xx[q][j] = (Tv)0; }
                    if (ok) {
                        const Tv *vp = val + r[q].vofs + (long long)v * VE;
                        if constexpr (VE == 2) { const double2 a2 = __ldcs(reinterpret_cast<const double2 *>(vp)); vv[q][0] = (Tv)a2.x; vv[q][1] = (Tv)a2.y; }
                        else { const float4 a4 = __ldcs(reinterpret_cast<const float4 *>(vp)); vv[q][0] = (Tv)a4.x; vv[q][1] = (Tv)a4.y; vv[q][2] = (Tv)a4.z; vv[q][3] = (Tv)a4.w; }
#pragma unroll
                        for (int j = 0; j < VE; j++) xx[q][j] = __ldg(x + r[q].col + cv * VE + j);
                    }
                }
#pragma unroll
                for (int q = 0; q < UNR; q++)
#pragma unroll
                    for (int j = 0; j < VE; j++) acc[j] = fma(vv[q][j], xx[q][j], acc[j]);
            }
            Tv s = (Tv)0;
#pragma unroll
            for (int j = 0; j < VE; j++) s += acc[j];
            // sum the cpr lanes of a block row (adjacent lanes)
            for (int d = 1; d < cpr; d <<= 1) {
                const Tv o = __shfl_down_sync(gmask, s, d, SG);
                if (cv + d < cpr) s += o;
            }
            if (mine && cv == 0) {
                Tv *yp = y + i0 + di;
                *yp = (beta == (Tv)0) ? alpha * s : alpha * s + beta * *yp;
            }
        }
    }
}

struct TIndex {
    int *d_tptr = nullptr;
    TRec *d_rec = nullptr;
    int nkeys = 0;
    int mode = 0; // DESC_ROWS / DESC_BLOCKS
    vbc_mat *At = nullptr; // transposed copy (uniform blocks): an internal handle whose adjoint multiply is this matrix' forward multiply
};

void destroy_tindex(TIndex *t)
{
    if (!t) return;
    cudaFree(t->d_tptr);
    cudaFree(t->d_rec);
    if (t->At) {
        cudaFree(t->At->d_meta); cudaFree(t->At->d_desc); cudaFree(t->At->d_val); cudaFree(t->At->d_order);
        delete t->At;
    }
    delete t;
}

static int ilog2x(int v)
{
    if (v <= 0 || (v & (v - 1))) return -1;
    int l = 0;
    while ((1 << l) < v) l++;
    return l;
}

template <int MODE>
static int build_tindex_mode(vbc_mat *A, TIndex *T)
{
    const int L = (int)A->L;
    const int64_t nkeys = MODE == DESC_ROWS ? A->m : (A->m + A->u0 - 1) / A->u0;
    T->nkeys = (int)nkeys;
    T->mode = MODE;
    cudaStream_t st = A->stream;
    unsigned long long *d_cnt = nullptr;
    long long *d_tmp = nullptr;
    unsigned *d_cur = nullptr;
    const size_t nk = (size_t)(nkeys > 0 ? nkeys : 1);
    VBC_CUDA(cudaMalloc(&d_cnt, sizeof(unsigned long long) * nk));
    cudaError_t e = cudaMalloc(&d_tmp, sizeof(long long) * (size_t)scan_tmp_elems(nkeys));
    if (e == cudaSuccess) e = cudaMalloc(&d_cur, sizeof(unsigned) * nk);
    if (e == cudaSuccess) e = cudaMalloc(&T->d_tptr, sizeof(int) * (nk + 1));
    if (e == cudaSuccess) e = cudaMalloc(&T->d_rec, sizeof(TRec) * (size_t)(A->ndesc > 0 ? A->ndesc : 1));
    if (e == cudaSuccess) e = cudaMemsetAsync(d_cnt, 0, sizeof(unsigned long long) * nk, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_cur, 0, sizeof(unsigned) * nk, st);
    int rc = VBC_OK;
    if (e != cudaSuccess) { set_error("transposed index allocation: %s", cudaGetErrorString(e)); rc = VBC_ENOMEM; }
    const int log2u = ilog2x(A->u0);
    const unsigned g = (unsigned)((L + 127) / 128 > 0 ? (L + 127) / 128 : 1);
    if (rc == VBC_OK && L > 0) { k_t_count<MODE><<<g, 128, 0, st>>>(A->d_meta, A->d_desc, L, A->u0, log2u, d_cnt); A->launches++; }
    long long total = 0;
    if (rc == VBC_OK) rc = exclusive_scan<int>((const long long *)d_cnt, T->d_tptr, nkeys, 0, d_tmp, &total, st, &A->launches);
    if (rc == VBC_OK && L > 0) { k_t_fill<MODE><<<g, 128, 0, st>>>(A->d_meta, A->d_desc, L, A->u0, log2u, T->d_tptr, d_cur, T->d_rec); A->launches++; }
    if (rc == VBC_OK && nkeys > 0) { k_t_sort<<<(unsigned)((nkeys + 255) / 256), 256, 0, st>>>(T->d_tptr, T->d_rec, (int)nkeys); A->launches++; }
    if (rc == VBC_OK && (cudaStreamSynchronize(st) != cudaSuccess || cudaGetLastError() != cudaSuccess)) { set_error("transposed index kernels failed"); rc = VBC_ECUDA; }
    cudaFree(d_cnt); cudaFree(d_tmp); cudaFree(d_cur);
    return rc;
}

// uniform blocks: materialise the transposed copy from the index (opt_fwd_atomic 0 = when device memory is plentiful, 3 = always)
template <typename Tv>
static int build_transposed_copy(vbc_mat *A, TIndex *T)
{
    if (A->w_uniform <= 0) return VBC_OK; // the copy's blocks are w0 high: one stripe width for the whole matrix
    const int uu = T->mode == DESC_BLOCKS ? A->u0 : 1; // rows of a unit: a u0-high block, or one stored row (1D / expanded rows)
    const size_t need = sizeof(Tv) * ((size_t)A->nval + 64) + 4 * (size_t)A->ndesc + sizeof(StripeMeta) * ((size_t)T->nkeys + 1);
    if (A->opt_fwd_atomic != 3) {
        size_t fr = 0, tot = 0;
        if (cudaMemGetInfo(&fr, &tot) != cudaSuccess || need * 4 > fr) { cudaGetLastError(); return VBC_OK; } // keep the second copy for when memory is plentiful
    }
    vbc_mat *At = new (std::nothrow) vbc_mat();
    if (!At) VBC_FAIL(VBC_ENOMEM, "host allocation failed");
    At->vt = A->vt; At->it = A->it; At->ndim = 2; At->device = A->device;
    At->m = A->n; At->n = A->m; At->K = A->L; At->L = T->nkeys; At->U = A->W; At->W = uu;
    At->nidx = A->ndesc; At->nval = A->nval; At->ndesc = A->ndesc;
    At->desc_mode = DESC_BLOCKS; At->u0 = A->w_uniform; At->w_uniform = uu;
    At->sm_count = A->sm_count; At->stream = A->stream;
    cudaError_t e = cudaMalloc(&At->d_meta, sizeof(StripeMeta) * ((size_t)T->nkeys + 1));
    if (e == cudaSuccess) e = cudaMalloc(&At->d_desc, sizeof(int) * (size_t)(A->ndesc > 0 ? A->ndesc : 1));
    if (e == cudaSuccess) e = cudaMalloc(&At->d_val, sizeof(Tv) * ((size_t)A->nval + 64));
    if (e == cudaSuccess) e = cudaMemsetAsync((char *)At->d_val + sizeof(Tv) * (size_t)A->nval, 0, sizeof(Tv) * 64, A->stream);
    if (e == cudaSuccess) {
        k_t_meta<Tv><<<(unsigned)((T->nkeys + 1 + 255) / 256), 256, 0, A->stream>>>(T->d_tptr, T->nkeys, uu, A->w_uniform, A->m, At->d_meta);
        if (T->nkeys > 0) k_t_copy<Tv><<<(unsigned)(((long long)T->nkeys * 32 + 255) / 256), 256, 0, A->stream>>>(T->d_tptr, T->d_rec, (const Tv *)A->d_val, T->nkeys, uu, A->w_uniform, A->m, At->d_desc, (Tv *)At->d_val);
        A->launches += 2;
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaStreamSynchronize(A->stream);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        cudaFree(At->d_meta); cudaFree(At->d_desc); cudaFree(At->d_val);
        delete At;
        if (A->opt_fwd_atomic == 3) VBC_FAIL(e == cudaErrorMemoryAllocation ? VBC_ENOMEM : VBC_ECUDA, "transposed copy: %s", cudaGetErrorString(e));
        return VBC_OK; // auto mode: fall back to the index
    }
    T->At = At;
    // the records are no longer needed: the copy's descriptors and meta carry the structure
    cudaFree(T->d_rec);
    T->d_rec = nullptr;
    return VBC_OK;
}

// ---- variable blocks (2D, parts of different heights): the transposed copy in ROWS mode ----------------------------------
// Row part k becomes stripe k of the copy (u_k columns wide: the part's rows); each of its blocks, taken in ascending
// stripe-column order, contributes w_l stored rows (one per column of A in that block) whose descriptor is the column index
// and whose u_k values are that column of the block.  y = A x is then the adjoint multiply of the copy -- the owner-computes
// streaming kernel with its flat-slab bodies for the odd widths, no atomics, fixed summation order -- instead of one
// floating-point atomic per stored row (0.35 of the HBM peak on the C2v matrix).
template <typename Ti>
__global__ void __launch_bounds__(128) k_tb_count(const Ti *__restrict__ pos, const Ti *__restrict__ idx, const int L, unsigned long long *__restrict__ cnt)
{
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= L) return;
    for (long long Q = (long long)pos[l] - 1; Q < (long long)pos[l + 1] - 1; Q++) atomicAdd(&cnt[(long long)idx[Q] - 1], 1ull);
}

template <typename Ti>
__global__ void __launch_bounds__(128) k_tb_fill(const Ti *__restrict__ pos, const Ti *__restrict__ idx, const Ti *__restrict__ ofs, const Ti *__restrict__ phi,
                                                  const int *__restrict__ brow, const int L, const int *__restrict__ tptr, unsigned *__restrict__ cursor,
                                                  TRec *__restrict__ rec)
{
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= L) return;
    const int j0 = (int)((long long)phi[l] - 1), w = (int)((long long)phi[l + 1] - (long long)phi[l]);
    if (w <= 0) return;
    const long long v0 = (long long)ofs[l] - 1;
    for (long long Q = (long long)pos[l] - 1; Q < (long long)pos[l + 1] - 1; Q++) {
        const long long k = (long long)idx[Q] - 1;
        TRec r;
        r.vofs = v0 + (long long)brow[Q] * w;
        r.col = j0;
        r.w = w;
        rec[tptr[k] + atomicAdd(&cursor[k], 1u)] = r;
    }
}

// stored rows and values of every stripe of the copy
template <typename Ti>
__global__ void __launch_bounds__(256) k_tb_rows(const int *__restrict__ tptr, const TRec *__restrict__ rec, const Ti *__restrict__ pi, const int K,
                                                  long long *__restrict__ rows, long long *__restrict__ vals)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    long long r = 0;
    for (int t = tptr[k]; t < tptr[k + 1]; t++) r += rec[t].w;
    rows[k] = r;
    vals[k] = r * ((long long)pi[k + 1] - (long long)pi[k]);
}

template <typename Ti>
__global__ void __launch_bounds__(256) k_tb_meta(const long long *__restrict__ ofs2, const int *__restrict__ pos2, const Ti *__restrict__ pi, const int K,
                                                  StripeMeta *__restrict__ meta2)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > K) return;
    StripeMeta s;
    s.ofs = ofs2[k]; s.pos = pos2[k]; s.col = (int)((long long)pi[k] - 1);
    meta2[k] = s;
}

template <typename Ti, typename Tv>
__global__ void __launch_bounds__(256) k_tb_copy(const int *__restrict__ tptr, const TRec *__restrict__ rec, const Tv *__restrict__ val, const Ti *__restrict__ pi,
                                                  const StripeMeta *__restrict__ meta2, const int K, int *__restrict__ desc2, Tv *__restrict__ val2)
{
    const int k = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (k >= K) return;
    const int u = (int)((long long)pi[k + 1] - (long long)pi[k]);
    const StripeMeta a = meta2[k];
    int row = 0;
    for (int t = tptr[k]; t < tptr[k + 1]; t++) {
        const TRec r = rec[t];
        for (int dj = lane; dj < r.w; dj += 32) desc2[a.pos + row + dj] = r.col + dj;
        Tv *out = val2 + a.ofs + (long long)row * u;
        const int per = r.w * u;
        for (int e = lane; e < per; e += 32) {
            const int dj = e / u, di = e - dj * u; // stored row dj of the copy (a column of A), column di (a row of the part)
            out[e] = val[r.vofs + (long long)di * r.w + dj];
        }
        row += r.w;
    }
}

template <typename Ti, typename Tv>
static int build_transposed_blocks(vbc_mat *A, TIndex *T)
{
    const int L = (int)A->L, K = (int)A->K;
    cudaStream_t st = A->stream;
    const size_t need = sizeof(Tv) * ((size_t)A->nval + 64) + 4 * (size_t)A->nval + sizeof(StripeMeta) * ((size_t)K + 1);
    if (A->opt_fwd_atomic != 3) {
        size_t fr = 0, tot = 0;
        if (cudaMemGetInfo(&fr, &tot) != cudaSuccess || need * 4 > fr) { cudaGetLastError(); return VBC_OK; }
    }
    T->nkeys = K; T->mode = DESC_ROWS;
    struct Scratch {
        unsigned long long *cnt = nullptr; long long *tmp = nullptr, *rows = nullptr, *vals = nullptr, *ofs2 = nullptr; unsigned *cur = nullptr; int *pos2 = nullptr;
        ~Scratch() { cudaFree(cnt); cudaFree(tmp); cudaFree(rows); cudaFree(vals); cudaFree(ofs2); cudaFree(cur); cudaFree(pos2); }
    } sc;
    const size_t nk = (size_t)(K > 0 ? K : 1);
    VBC_CUDA(cudaMalloc(&sc.cnt, 8 * nk));
    VBC_CUDA(cudaMalloc(&sc.tmp, 8 * (size_t)scan_tmp_elems(K)));
    VBC_CUDA(cudaMalloc(&sc.rows, 8 * nk));
    VBC_CUDA(cudaMalloc(&sc.vals, 8 * nk));
    VBC_CUDA(cudaMalloc(&sc.ofs2, 8 * (nk + 1)));
    VBC_CUDA(cudaMalloc(&sc.cur, 4 * nk));
    VBC_CUDA(cudaMalloc(&sc.pos2, 4 * (nk + 1)));
    VBC_CUDA(cudaMalloc(&T->d_tptr, 4 * (nk + 1)));
    VBC_CUDA(cudaMalloc(&T->d_rec, sizeof(TRec) * (size_t)(A->nidx > 0 ? A->nidx : 1)));
    VBC_CUDA(cudaMemsetAsync(sc.cnt, 0, 8 * nk, st));
    VBC_CUDA(cudaMemsetAsync(sc.cur, 0, 4 * nk, st));
    const Ti *pos = (const Ti *)A->d_pos, *idx = (const Ti *)A->d_idx, *ofs = (const Ti *)A->d_ofs, *phi = (const Ti *)A->d_phi_spl, *pi = (const Ti *)A->d_pi_spl;
    const unsigned g = (unsigned)((L + 127) / 128 > 0 ? (L + 127) / 128 : 1);
    long long total = 0, nrows = 0, nvals = 0;
    if (L > 0) { k_tb_count<Ti><<<g, 128, 0, st>>>(pos, idx, L, sc.cnt); A->launches++; }
    VBC_TRY(exclusive_scan<int>((const long long *)sc.cnt, T->d_tptr, K, 0, sc.tmp, &total, st, &A->launches));
    if (L > 0) { k_tb_fill<Ti><<<g, 128, 0, st>>>(pos, idx, ofs, phi, A->d_brow, L, T->d_tptr, sc.cur, T->d_rec); A->launches++; }
    if (K > 0) {
        k_t_sort<<<(unsigned)((K + 255) / 256), 256, 0, st>>>(T->d_tptr, T->d_rec, K);
        k_tb_rows<Ti><<<(unsigned)((K + 255) / 256), 256, 0, st>>>(T->d_tptr, T->d_rec, pi, K, sc.rows, sc.vals);
        A->launches += 2;
    }
    VBC_CUDA(cudaGetLastError());
    VBC_TRY(exclusive_scan<int>(sc.rows, sc.pos2, K, 0, sc.tmp, &nrows, st, &A->launches));
    VBC_TRY(exclusive_scan<long long>(sc.vals, sc.ofs2, K, 0, sc.tmp, &nvals, st, &A->launches));
    if (nrows >= (1LL << 31)) return VBC_OK; // 32-bit descriptor space: keep the scatter kernel
    vbc_mat *At = new (std::nothrow) vbc_mat();
    if (!At) VBC_FAIL(VBC_ENOMEM, "host allocation failed");
    At->vt = A->vt; At->it = A->it; At->ndim = 2; At->device = A->device;
    At->m = A->n; At->n = A->m; At->K = A->L; At->L = K; At->U = A->W; At->W = A->U;
    At->nidx = nrows; At->nval = nvals; At->ndesc = nrows;
    At->desc_mode = DESC_ROWS; At->u0 = 1;
    At->sm_count = A->sm_count; At->stream = st;
    cudaError_t e = cudaMalloc(&At->d_meta, sizeof(StripeMeta) * ((size_t)K + 1));
    if (e == cudaSuccess) e = cudaMalloc(&At->d_desc, sizeof(int) * (size_t)(nrows > 0 ? nrows : 1));
    if (e == cudaSuccess) e = cudaMalloc(&At->d_val, sizeof(Tv) * ((size_t)nvals + 64));
    if (e == cudaSuccess) e = cudaMemsetAsync((char *)At->d_val + sizeof(Tv) * (size_t)nvals, 0, sizeof(Tv) * 64, st);
    int rc = VBC_OK;
    if (e == cudaSuccess) {
        k_tb_meta<Ti><<<(unsigned)((K + 1 + 255) / 256), 256, 0, st>>>(sc.ofs2, sc.pos2, pi, K, At->d_meta);
        if (K > 0) k_tb_copy<Ti, Tv><<<(unsigned)(((long long)K * 32 + 255) / 256), 256, 0, st>>>(T->d_tptr, T->d_rec, (const Tv *)A->d_val, pi, At->d_meta, K, At->d_desc, (Tv *)At->d_val);
        A->launches += 2;
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e == cudaSuccess) rc = build_class_order(At);
    }
    if (e != cudaSuccess || rc != VBC_OK) {
        cudaGetLastError();
        cudaFree(At->d_meta); cudaFree(At->d_desc); cudaFree(At->d_val); cudaFree(At->d_order);
        delete At;
        if (A->opt_fwd_atomic == 3) VBC_FAIL(e == cudaErrorMemoryAllocation ? VBC_ENOMEM : VBC_ECUDA, "transposed copy: %s", cudaGetErrorString(e));
        return VBC_OK;
    }
    A->launches += At->launches;
    At->launches = 0;
    T->At = At;
    cudaFree(T->d_rec);
    T->d_rec = nullptr;
    return VBC_OK;
}

// eligible: rows mode always; blocks mode when every stripe has one width that is a multiple of the 16-byte vector
// (opt_fwd_atomic: 0 = auto, 1 = never, 2 = whenever possible).  Measured (profiles/r01_tuning.md): the index wins
// for uniform 2D blocks (113 vs 127 us on configs[1]) and loses to 32-lane atomics in rows mode (116 vs 110 us
// on the 1D w=8 matrix), so auto uses it for blocks mode only.
static bool tindex_eligible(const vbc_mat *A)
{
    if (A->L == 0 || A->m == 0 || A->opt_fwd_atomic == 1) return false;
    if (A->desc_mode == DESC_ROWS) return A->opt_fwd_atomic == 2 || A->w_uniform > 0 || A->ndim == 2; // one stripe width, or variable 2D blocks: a transposed copy is possible
    const int VE = 16 / (int)vt_size(A->vt);
    return A->w_uniform > 0 && (A->w_uniform % VE) == 0 && A->u0 * (A->w_uniform / VE) <= 32;
}

int ensure_tindex(vbc_mat *A)
{
    if (A->tindex || A->opt_fwd_no_copy || !tindex_eligible(A)) return VBC_OK;
    TIndex *T = new (std::nothrow) TIndex();
    if (!T) VBC_FAIL(VBC_ENOMEM, "host allocation failed");
    int rc;
    if (A->ndim == 2 && A->desc_mode == DESC_ROWS && A->opt_fwd_atomic != 2 && A->w_uniform <= 0) { // variable blocks: the copy is built from the canonical block arrays
        // the copies move values as bits: the 8- and 4-byte instantiations also serve the integer element types
        if (A->it == VBC_I64) rc = vt_size(A->vt) == 8 ? build_transposed_blocks<int64_t, double>(A, T) : build_transposed_blocks<int64_t, float>(A, T);
        else rc = vt_size(A->vt) == 8 ? build_transposed_blocks<int32_t, double>(A, T) : build_transposed_blocks<int32_t, float>(A, T);
    } else {
        rc = A->desc_mode == DESC_ROWS ? build_tindex_mode<DESC_ROWS>(A, T) : build_tindex_mode<DESC_BLOCKS>(A, T);
        if (rc == VBC_OK && A->opt_fwd_atomic != 2) rc = vt_size(A->vt) == 8 ? build_transposed_copy<double>(A, T) : build_transposed_copy<float>(A, T);
    }
    if (rc != VBC_OK) { destroy_tindex(T); return rc; }
    if (!T->At && A->desc_mode == DESC_ROWS && A->opt_fwd_atomic != 2) { destroy_tindex(T); A->opt_fwd_no_copy = 1; return VBC_OK; } // rows mode without the copy: atomics win over the index
    A->tindex = T;
    return VBC_OK;
}

// bytes one forward multiply reads besides the values: the copy's meta + descriptors, or the index
int64_t tindex_bytes(const vbc_mat *A)
{
    if (!A->tindex) return 0;
    if (A->tindex->At) return (int64_t)sizeof(StripeMeta) * ((int64_t)A->tindex->nkeys + 1) + 4 * A->tindex->At->ndesc;
    return (int64_t)sizeof(TRec) * A->ndesc + 4 * ((int64_t)A->tindex->nkeys + 1);
}

int tindex_kind(const vbc_mat *A) { return !A->tindex ? 0 : (A->tindex->At ? 2 : 1); }
vbc_mat *tindex_copy(const vbc_mat *A) { return A->tindex ? A->tindex->At : nullptr; }

template <typename Tv>
static int launch_fwdt_t(vbc_mat *A, Tv alpha, const Tv *x, Tv beta, Tv *y)
{
    const TIndex *T = A->tindex;
    const int VE = 16 / (int)sizeof(Tv);
    int64_t grid = (int64_t)A->sm_count * 6;
    auto clamp = [&](int sg) { const int64_t need = ((int64_t)T->nkeys * sg + 255) / 256; return (unsigned)(grid > need ? (need > 0 ? need : 1) : grid); };
    if (T->mode == DESC_ROWS) {
        // lanes per destination row ~ vectors of a typical row segment
        const double w_avg = A->ndesc > 0 ? (double)A->nval / (double)A->ndesc : 1.0;
        const int cpr = (int)((w_avg + VE - 1) / VE);
        if (cpr >= 4) k_fwdt_rows<Tv, 4><<<clamp(4), 256, 0, A->stream>>>(T->d_tptr, T->d_rec, (const Tv *)A->d_val, x, y, T->nkeys, alpha, beta);
        else if (cpr >= 2) k_fwdt_rows<Tv, 2><<<clamp(2), 256, 0, A->stream>>>(T->d_tptr, T->d_rec, (const Tv *)A->d_val, x, y, T->nkeys, alpha, beta);
        else k_fwdt_rows<Tv, 1><<<clamp(1), 256, 0, A->stream>>>(T->d_tptr, T->d_rec, (const Tv *)A->d_val, x, y, T->nkeys, alpha, beta);
    } else {
        const int nvec = A->u0 * (A->w_uniform / VE);
        if (nvec > 16) k_fwdt_blocks<Tv, 32><<<clamp(32), 256, 0, A->stream>>>(T->d_tptr, T->d_rec, (const Tv *)A->d_val, x, y, T->nkeys, A->u0, A->w_uniform, A->m, alpha, beta);
        else if (nvec > 8) k_fwdt_blocks<Tv, 16><<<clamp(16), 256, 0, A->stream>>>(T->d_tptr, T->d_rec, (const Tv *)A->d_val, x, y, T->nkeys, A->u0, A->w_uniform, A->m, alpha, beta);
        else if (nvec > 4) k_fwdt_blocks<Tv, 8><<<clamp(8), 256, 0, A->stream>>>(T->d_tptr, T->d_rec, (const Tv *)A->d_val, x, y, T->nkeys, A->u0, A->w_uniform, A->m, alpha, beta);
        else k_fwdt_blocks<Tv, 4><<<clamp(4), 256, 0, A->stream>>>(T->d_tptr, T->d_rec, (const Tv *)A->d_val, x, y, T->nkeys, A->u0, A->w_uniform, A->m, alpha, beta);
    }
    A->launches++;
    VBC_CUDA(cudaGetLastError());
    return VBC_OK;
}

int launch_fwdt(vbc_mat *A, double alpha, const void *x, double beta, void *y)
{
    if (A->tindex->At && A->tindex->mode == DESC_ROWS && A->w_uniform > 0 && (A->w_uniform % (16 / (int)vt_size(A->vt))) == 0) {
        // rows mode: the copy is row-contiguous segments -> a dot product per row (k_fwdc_rows)
        const vbc_mat *At = A->tindex->At;
        const int VE = 16 / (int)vt_size(A->vt), vps = A->w_uniform / VE;
        int64_t grid = (int64_t)A->sm_count * 8;
        const double vec_per_row = At->L > 0 ? (double)At->nval / VE / (double)At->L : 1.0;
        const int SG = vec_per_row >= 64 ? 16 : (vec_per_row >= 16 ? 8 : 4);
        const int64_t need = (At->L * SG + 255) / 256;
        if (grid > need) grid = need;
        if (grid < 1) grid = 1;
        const int l2 = ilog2x(vps);
#define FWDC(Tv, SGv) k_fwdc_rows<Tv, SGv><<<(unsigned)grid, 256, 0, A->stream>>>(At->d_meta, At->d_desc, (const Tv *)At->d_val, (const Tv *)x, (Tv *)y, (int)At->L, A->w_uniform, l2, (Tv)alpha, (Tv)beta)
        if (A->vt == VBC_F64) { if (SG == 16) FWDC(double, 16); else if (SG == 8) FWDC(double, 8); else FWDC(double, 4); }
        else { if (SG == 16) FWDC(float, 16); else if (SG == 8) FWDC(float, 8); else FWDC(float, 4); }
#undef FWDC
        A->launches++;
        VBC_CUDA(cudaGetLastError());
        return VBC_OK;
    }
    if (A->tindex->At) { // forward multiply of A = adjoint multiply of its transposed copy
        vbc_mat *At = A->tindex->At;
        At->stream = A->stream;
        At->opt_adj_group = A->opt_fwd_group == 8 || A->opt_fwd_group == 32 ? A->opt_fwd_group : 0;
        const int64_t before = At->launches;
        const int rc = launch_spmv(At, 1, alpha, x, beta, y);
        A->launches += At->launches - before;
        return rc;
    }
    return A->vt == VBC_F64 ? launch_fwdt_t<double>(A, alpha, (const double *)x, beta, (double *)y)
                            : launch_fwdt_t<float>(A, (float)alpha, (const float *)x, (float)beta, (float *)y);
}

} // namespace vbc
