// trsv.cu -- lower-triangular VBC solve  tril(A') x = b,  north_star item (d).
//
// EXTENSION WITHOUT A REFERENCE COUNTERPART: the reference's "TrSpMV" is the TRANSPOSED multiply
// (TrSpMV.jl:1-20); it has no triangular solve.  BASELINE.json asks for "a blocked
// triangular solve ... using level-scheduled row blocks", so this file provides one over the same
// device layout; its oracle is plain forward substitution on the CSC matrix (parity unpinned).
//
// Orientation: stripe l (columns j0..j0+w) of A is row block l of A'.  Row j0+dj of A' holds
// A[i, j0+dj] for the stripe's stored rows i; entries with i > j0+dj lie above the diagonal of A' and
// are ignored (BLAS trsv 'L' semantics).  Per stripe:
//     acc[dj]    = sum over stored rows i < j0 of val[r, dj] * x[i]          (needs earlier row blocks)
//     x[j0+dj]   = (b[j0+dj] - acc[dj] - sum_{d < dj} D[d][dj] * x[j0+d]) / D[dj][dj]
// with D[d][dj] = A[j0+d, j0+dj] the w x w diagonal block.
//
// Schedule: vbc_trsv_analyse computes level[l] = 1 + max level over the row blocks l depends on, orders the row blocks by
// level, and extracts every row block's w x w diagonal block (with reciprocal diagonal) into a dense array.  ONE cooperative
// persistent kernel walks that order, one warp per row block.  There are no completion flags: x itself carries the
// dependency.  Before a solve x is filled with a sentinel (a NaN with a payload no computation produces); a lane that needs
// x[i] polls x[i] (L2-coherent 8-byte loads, single-copy atomic) until it is no longer the sentinel; the producer just
// stores its results.  One L2 round trip per dependency level instead of store + fence + flag + poll + load, no epoch
// counter to keep in step with CUDA-graph replays, no column -> row-block table.  Dependencies always sit earlier in the
// order and all CTAs are co-resident (cooperative launch), hence the first unfinished row block can always run: no
// deadlock.  A wall-clock bound in the poll loop turns a lost dependency into an error instead of a hang.
#include <stdlib.h>

#include <algorithm>
#include <new>
#include <vector>

#include "common.cuh"
#include "walk.cuh"

struct vbc_trsv_plan {
    int *d_order = nullptr;   // row blocks sorted by (level, index)
    void *d_diag = nullptr;   // L x wmax x wmax: D[l][d][dj] = A[j0 + d, j0 + dj], the diagonal entries replaced by their reciprocals
    int *d_err = nullptr;
    int nlevels = 0;
    int wmax = 0;
};

namespace vbc {

constexpr int TRSV_WMAX = 32; // one lane per unknown of a row block

template <typename Tv> struct Sentinel;
template <> struct Sentinel<double> {
    static __device__ __forceinline__ double value() { return __longlong_as_double(0x7ff8dead0b200b20LL); }
    static __device__ __forceinline__ bool is(double v) { return __double_as_longlong(v) == 0x7ff8dead0b200b20LL; }
};
template <> struct Sentinel<float> {
    static __device__ __forceinline__ float value() { return __int_as_float(0x7fc0dead); }
    static __device__ __forceinline__ bool is(float v) { return __float_as_int(v) == 0x7fc0dead; }
};

// GPU-scope relaxed accesses for the x entries that carry the dependencies (L2 is their point of coherence; `volatile`
// would be system scope)
__device__ __forceinline__ double ld_x_gpu(const double *p) { double v; asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ float ld_x_gpu(const float *p) { float v; asm volatile("ld.relaxed.gpu.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_x_gpu(double *p, double v) { asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory"); }
__device__ __forceinline__ void st_x_gpu(float *p, float v) { asm volatile("st.relaxed.gpu.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory"); }

// Loads that must be ISSUED before the polling starts (values, right-hand side, diagonal block): `asm volatile` keeps them
// ahead of the (volatile) polling loads -- the compiler otherwise sinks them to their first use, after the wait, and puts a
// DRAM round trip on the dependency chain of every level.
__device__ __forceinline__ double ld_early(const double *p) { double v; asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p)); return v; }
__device__ __forceinline__ float ld_early(const float *p) { float v; asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p)); return v; }

template <typename Tv>
__global__ void __launch_bounds__(256) k_trsv_fill(Tv *__restrict__ x, const int64_t n, int *__restrict__ err)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) *err = 0; // every solve starts with a clean error flag
    const Tv s = Sentinel<Tv>::value();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) x[i] = s;
}

// D[l][d][dj] for the stored rows i = j0 + d of stripe l that fall inside its own column range; diagonal -> reciprocal
template <typename Tv, int MODE>
__global__ void __launch_bounds__(256) k_trsv_diag(const StripeMeta *__restrict__ meta, const int *__restrict__ desc, const Tv *__restrict__ val,
                                                    const int L, const int u0, const int log2u, const int wmax, Tv *__restrict__ D, int *__restrict__ err)
{
    const int l = (int)((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (l >= L) return;
    const StripeMeta a = ld_meta(meta + l), b = ld_meta(meta + l + 1);
    const int w = b.col - a.col, j0 = a.col;
    if (w <= 0) return;
    Tv *Dl = D + (size_t)l * wmax * wmax;
    for (int i = lane; i < wmax * wmax; i += 32) Dl[i] = (Tv)0;
    __syncwarp();
    const int R = (MODE == DESC_ROWS) ? (b.pos - a.pos) : (int)((b.ofs - a.ofs) / w);
    for (int r = lane; r < R; r += 32) {
        int xi;
        if (MODE == DESC_ROWS) xi = desc[a.pos + r];
        else xi = (log2u >= 0) ? desc[a.pos + (r >> log2u)] + (r & (u0 - 1)) : desc[a.pos + r / u0] + r % u0;
        if (xi >= j0 && xi < j0 + w)
            for (int dj = 0; dj < w; dj++) Dl[(xi - j0) * wmax + dj] = val[a.ofs + (long long)r * w + dj];
    }
    __syncwarp();
    if (lane < w) {
        const Tv d = Dl[lane * wmax + lane];
        if (d == (Tv)0) atomicExch(err, 2); // a zero (or missing) diagonal entry
        Dl[lane * wmax + lane] = (Tv)1 / d;
    }
}

template <typename Tv, int MODE>
__global__ void __launch_bounds__(256) k_trsv_lower(const StripeMeta *__restrict__ meta, const int *__restrict__ desc,
                                                     const Tv *__restrict__ val, const int *__restrict__ order, const Tv *__restrict__ D, const int wmax,
                                                     const Tv *__restrict__ bvec, Tv *x, const int L, const int u0, const int log2u,
                                                     int *__restrict__ err)
{
    const int lane = threadIdx.x & 31;
    const int nwarps = (int)((gridDim.x * blockDim.x) >> 5);
    int t = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    int lnext = t < L ? __ldg(order + t) : 0;
    for (; t < L; t += nwarps) {
        const int l = lnext;
        if (t + nwarps < L) lnext = __ldg(order + t + nwarps);
        const StripeMeta a = ld_meta(meta + l), b = ld_meta(meta + l + 1);
        const int w = b.col - a.col;
        if (w <= 0) continue;
        const int j0 = a.col;
        const int R = (MODE == DESC_ROWS) ? (b.pos - a.pos) : (int)((b.ofs - a.ofs) / w);
        // everything that does not depend on other row blocks first: right-hand side, the diagonal block's column of this lane
        const Tv rhs0 = lane < w ? ld_early(bvec + j0 + lane) : (Tv)0;
        const Tv *Dl = D + (size_t)l * wmax * wmax;
        // this lane's column of the diagonal block, fetched NOW: inside the substitution loop each load would sit on the
        // dependency chain of the whole solve (one memory round trip per row of the block and per level)
        constexpr int WR = 8;
        Tv dcol[WR];
#pragma unroll
        for (int d = 0; d < WR; d++) dcol[d] = (d < w && lane < w) ? ld_early(Dl + d * wmax + lane) : (Tv)0;
        const Tv dinv = lane < w ? ld_early(Dl + lane * wmax + lane) : (Tv)0;
        // element mapping: lane -> (row r0 + k*rps, column c)
        const int rps = 32 / w, c = lane % w, r0 = lane / w;
        const bool active = lane < rps * w;
        Tv acc = (Tv)0;
        if (active) {
            constexpr int TB = 4; // rows per lane in flight: descriptors and values are issued before the first poll
            RowWalk<MODE> walk;
            walk.init(desc, a.pos, r0, rps, u0, log2u);
            const Tv *vp = val + a.ofs + (long long)r0 * w + c;
            for (int r = r0; r < R; r += TB * rps) {
                int xi[TB];
                Tv v[TB], xv[TB];
#pragma unroll
                for (int k = 0; k < TB; k++) {
                    const bool ok = r + k * rps < R;
                    xi[k] = walk.next_if(ok);
                    v[k] = ok ? ld_early(vp) : (Tv)0;
                    vp += (long long)rps * w;
                    if (!ok || xi[k] >= j0) xi[k] = -1; // the diagonal block and everything above the diagonal of A' is not a dependency
                }
#pragma unroll
                for (int k = 0; k < TB; k++) xv[k] = xi[k] >= 0 ? ld_x_gpu(x + xi[k]) : (Tv)0; // written by another SM
                bool pend = false;
#pragma unroll
                for (int k = 0; k < TB; k++) pend = pend || (xi[k] >= 0 && Sentinel<Tv>::is(xv[k]));
                if (pend) { // not solved yet: poll all missing entries together (the clock is read once per 256 rounds only)
                    unsigned long long t0 = 0, t1;
                    unsigned polls = 0;
                    do {
#pragma unroll
                        for (int k = 0; k < TB; k++)
                            if (xi[k] >= 0 && Sentinel<Tv>::is(xv[k])) xv[k] = ld_x_gpu(x + xi[k]);
                        pend = false;
#pragma unroll
                        for (int k = 0; k < TB; k++) pend = pend || (xi[k] >= 0 && Sentinel<Tv>::is(xv[k]));
                        if ((++polls & 255u) == 0) {
                            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                            if (t0 == 0) t0 = t1;
                            if (t1 - t0 > 3000000000ull) { atomicExch(err, 1); break; }
                        }
                    } while (pend);
                }
#pragma unroll
                for (int k = 0; k < TB; k++) acc = fma(v[k], (xi[k] >= 0 && !Sentinel<Tv>::is(xv[k])) ? xv[k] : (Tv)0, acc);
            }
        }
        // The lanes leave the poll loop one by one.  Without an explicit reconvergence here ptxas reaches the first shuffle with
        // the warp still split and takes its divergent-warp path (BRA.DIV): 1.5 us per level on the dependency chain of the
        // Float64 solve (2.71 -> 1.2 us per level; the Float32 build happened to reconverge by itself).
        __syncwarp();
        for (int d = w; d < 32; d <<= 1) {
            const Tv tsum = __shfl_down_sync(0xffffffffu, acc, d);
            if (lane + d < 32) acc += tsum;
        }
        // forward substitution inside the w x w diagonal block: lane dj owns unknown j0 + dj; the diagonal holds reciprocals
        Tv rhs = rhs0 - acc;
        Tv xj = (Tv)0;
        if (w <= WR) {
#pragma unroll
            for (int d = 0; d < WR; d++) {
                if (d < w) { // warp-uniform
                    const Tv xd = __shfl_sync(0xffffffffu, rhs * dinv, d); // x[j0 + d]
                    if (lane == d) xj = xd;
                    if (lane > d) rhs = fma(-dcol[d], xd, rhs);
                }
            }
        } else {
            for (int d = 0; d < w; d++) {
                const Tv xd = __shfl_sync(0xffffffffu, rhs * dinv, d);
                if (lane == d) xj = xd;
                if (lane > d && lane < w) rhs -= __ldg(Dl + d * wmax + lane) * xd;
            }
        }
        if (lane < w) st_x_gpu(x + j0 + lane, xj); // the store IS the completion signal
    }
}

static void free_plan(vbc_trsv_plan *p)
{
    if (!p) return;
    cudaFree(p->d_order); cudaFree(p->d_diag); cudaFree(p->d_err);
    delete p;
}

static int ilog2u(int u0)
{
    int log2u = -1;
    if (u0 > 0 && !(u0 & (u0 - 1))) { log2u = 0; while ((1 << log2u) < u0) log2u++; }
    return log2u;
}

template <typename Tv, int MODE>
static int launch_trsv(vbc_mat *A, vbc_trsv_plan *P, const Tv *b, Tv *x)
{
    int occ = 0;
    VBC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_trsv_lower<Tv, MODE>, 256, 0));
    if (occ < 1) occ = 1;
    int64_t grid = (int64_t)A->sm_count * occ;
    const int64_t need = (A->L * 32 + 255) / 256;
    if (grid > need) grid = need;
    // Only about one level's worth of row blocks can make progress at a time; every other resident warp polls.  Enough warps
    // to hold `ahead` levels (so that the values and descriptors of the coming levels are already in registers), no more: thousands
    // of polling warps slow the few that are on the critical path.
    static int ahead = -1;
    if (ahead < 0) { const char *e = getenv("VBC_TRSV_LEVELS_AHEAD"); ahead = e ? atoi(e) : 8; if (ahead < 1) ahead = 1; }
    const int64_t width = P->nlevels > 0 ? (A->L + P->nlevels - 1) / P->nlevels : A->L;
    const int tpb = 256, wpc = 8;
    int64_t want = (width * ahead + wpc - 1) / wpc;
    if (want < A->sm_count / 2) want = A->sm_count / 2;
    if (grid > want) grid = want;
    if (grid < 1) grid = 1;
    int L = (int)A->L, u0 = A->u0, log2u = ilog2u(A->u0), wmax = P->wmax;
    int64_t fg = (A->n + 255) / 256;
    if (fg > (int64_t)A->sm_count * 8) fg = (int64_t)A->sm_count * 8;
    k_trsv_fill<Tv><<<(unsigned)fg, 256, 0, A->stream>>>(x, A->n, P->d_err); // x <- sentinel, error flag <- 0 (in stream order: graph replays stay correct)
    A->launches++;
    const StripeMeta *meta = A->d_meta;
    const int *desc = A->d_desc, *order = P->d_order;
    const Tv *val = (const Tv *)A->d_val, *D = (const Tv *)P->d_diag;
    int *err = P->d_err;
    void *args[] = {&meta, &desc, &val, &order, &D, &wmax, &b, &x, &L, &u0, &log2u, &err};
    VBC_CUDA(cudaLaunchCooperativeKernel((const void *)k_trsv_lower<Tv, MODE>, dim3((unsigned)grid), dim3((unsigned)tpb), args, 0, A->stream));
    A->launches++;
    return VBC_OK;
}

template <typename Tv>
static int build_diag(vbc_mat *A, vbc_trsv_plan *P)
{
    const int64_t L = A->L;
    const size_t elems = (size_t)(L > 0 ? L : 1) * P->wmax * P->wmax;
    VBC_CUDA(cudaMalloc(&P->d_diag, sizeof(Tv) * (elems > 0 ? elems : 1)));
    if (L == 0) return VBC_OK;
    const unsigned g = (unsigned)((L * 32 + 255) / 256);
    if (A->desc_mode == DESC_ROWS) k_trsv_diag<Tv, DESC_ROWS><<<g, 256, 0, A->stream>>>(A->d_meta, A->d_desc, (const Tv *)A->d_val, (int)L, A->u0, ilog2u(A->u0), P->wmax, (Tv *)P->d_diag, P->d_err);
    else k_trsv_diag<Tv, DESC_BLOCKS><<<g, 256, 0, A->stream>>>(A->d_meta, A->d_desc, (const Tv *)A->d_val, (int)L, A->u0, ilog2u(A->u0), P->wmax, (Tv *)P->d_diag, P->d_err);
    A->launches++;
    VBC_CUDA(cudaGetLastError());
    int herr = 0;
    VBC_CUDA(cudaMemcpyAsync(&herr, P->d_err, sizeof(int), cudaMemcpyDeviceToHost, A->stream));
    VBC_CUDA(cudaStreamSynchronize(A->stream));
    if (herr == 2) VBC_FAIL(VBC_EARG, "ArgumentError: tril(A') has a zero or missing diagonal entry (singular)");
    return VBC_OK;
}

} // namespace vbc

using namespace vbc;

extern "C" {

int vbc_trsv_analyse(vbc_mat *A, int *nlevels)
{
    if (!A) VBC_FAIL(VBC_EARG, "matrix handle is NULL");
    if (A->m != A->n) VBC_FAIL(VBC_EDIM, "DimensionMismatch: triangular solve needs a square matrix, got %lld x %lld", (long long)A->m, (long long)A->n);
    if (A->opt_parity) VBC_FAIL(VBC_EARG, "triangular solve needs the compact layout (parity mode is on)");
    if (!vt_is_float(A->vt)) VBC_FAIL(VBC_EARG, "ArgumentError: the triangular solve is offered for Float32 / Float64 matrices");
    DeviceGuard guard(A->device);
    if (!guard.ok) VBC_FAIL(VBC_ECUDA, "cudaSetDevice(%d) failed", A->device);
    const int64_t L = A->L, n = A->n;
    std::vector<StripeMeta> meta((size_t)L + 1);
    std::vector<int> desc((size_t)(A->ndesc > 0 ? A->ndesc : 1));
    VBC_CUDA(cudaStreamSynchronize(A->stream));
    VBC_CUDA(cudaMemcpy(meta.data(), A->d_meta, sizeof(StripeMeta) * (size_t)(L + 1), cudaMemcpyDeviceToHost));
    if (A->ndesc > 0) VBC_CUDA(cudaMemcpy(desc.data(), A->d_desc, sizeof(int) * (size_t)A->ndesc, cudaMemcpyDeviceToHost));
    std::vector<int> c2s((size_t)(n > 0 ? n : 1)), level((size_t)(L > 0 ? L : 1), 1);
    int wmax = 0;
    for (int64_t l = 0; l < L; l++) {
        const int w = meta[l + 1].col - meta[l].col;
        wmax = std::max(wmax, w);
        for (int j = meta[l].col; j < meta[l + 1].col; j++) c2s[j] = (int)l;
    }
    if (wmax > TRSV_WMAX) VBC_FAIL(VBC_ELIMIT, "triangular solve supports stripes up to %d columns wide (widest is %d)", TRSV_WMAX, wmax);
    if (wmax < 1) wmax = 1;
    int maxlevel = L > 0 ? 1 : 0;
    const bool blocks = A->desc_mode == DESC_BLOCKS;
    for (int64_t l = 0; l < L; l++) {
        const int j0 = meta[l].col;
        int lev = 1;
        for (int q = meta[l].pos; q < meta[l + 1].pos; q++) {
            const int i0 = desc[q];
            const int i1 = blocks ? (int)std::min<int64_t>((int64_t)i0 + A->u0, A->m) : i0 + 1; // rows [i0, i1)
            // rows below j0 are dependencies; visit each row of the unit
            for (int i = i0; i < i1 && i < j0; i++) lev = std::max(lev, level[c2s[i]] + 1);
        }
        level[l] = lev;
        maxlevel = std::max(maxlevel, lev);
    }
    // counting sort by level (stable: index order inside a level)
    std::vector<int> count((size_t)maxlevel + 2, 0), order((size_t)(L > 0 ? L : 1));
    for (int64_t l = 0; l < L; l++) count[level[l] + 1]++;
    for (int v = 1; v <= maxlevel + 1; v++) count[v] += count[v - 1];
    for (int64_t l = 0; l < L; l++) order[count[level[l]]++] = (int)l;

    vbc_trsv_plan *P = new (std::nothrow) vbc_trsv_plan();
    if (!P) VBC_FAIL(VBC_ENOMEM, "host allocation failed");
    P->nlevels = maxlevel;
    P->wmax = wmax;
    cudaError_t e = cudaMalloc(&P->d_order, sizeof(int) * (size_t)(L > 0 ? L : 1));
    if (e == cudaSuccess) e = cudaMalloc(&P->d_err, sizeof(int));
    if (e == cudaSuccess && L > 0) e = cudaMemcpy(P->d_order, order.data(), sizeof(int) * (size_t)L, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemset(P->d_err, 0, sizeof(int));
    if (e != cudaSuccess) { free_plan(P); VBC_FAIL(VBC_ECUDA, "vbc_trsv_analyse: %s", cudaGetErrorString(e)); }
    const int rc = A->vt == VBC_F64 ? build_diag<double>(A, P) : build_diag<float>(A, P);
    if (rc != VBC_OK) { free_plan(P); return rc; }
    free_plan(A->trsv);
    A->trsv = P;
    if (nlevels) *nlevels = maxlevel;
    return VBC_OK;
}

int vbc_trsv_lower(vbc_mat *A, const void *b, void *x, int64_t len, int on_device)
{
    if (!A) VBC_FAIL(VBC_EARG, "matrix handle is NULL");
    if (A->m != A->n || len != A->n) VBC_FAIL(VBC_EDIM, "DimensionMismatch: triangular solve with a %lld x %lld matrix and vectors of length %lld", (long long)A->m, (long long)A->n, (long long)len);
    if (A->opt_parity) VBC_FAIL(VBC_EARG, "triangular solve needs the compact layout (parity mode is on)");
    if (!vt_is_float(A->vt)) VBC_FAIL(VBC_EARG, "ArgumentError: the triangular solve is offered for Float32 / Float64 matrices");
    if (!A->trsv) VBC_TRY(vbc_trsv_analyse(A, nullptr));
    if (len == 0) return VBC_OK;
    if (!b || !x) VBC_FAIL(VBC_EARG, "NULL vector");
    if (b == x) VBC_FAIL(VBC_EARG, "ArgumentError: the triangular solve is not in place: x and b must be different vectors");
    DeviceGuard guard(A->device);
    if (!guard.ok) VBC_FAIL(VBC_ECUDA, "cudaSetDevice(%d) failed", A->device);
    const size_t tv = vt_size(A->vt);
    const void *db = b;
    void *dx = x;
    void *tb = nullptr, *tx = nullptr;
    if (!on_device) {
        VBC_CUDA(cudaMalloc(&tb, tv * (size_t)len));
        if (cudaMalloc(&tx, tv * (size_t)len) != cudaSuccess) { cudaFree(tb); VBC_FAIL(VBC_ENOMEM, "trsv staging allocation failed"); }
        cudaError_t e = cudaMemcpyAsync(tb, b, tv * (size_t)len, cudaMemcpyHostToDevice, A->stream);
        if (e != cudaSuccess) { cudaFree(tb); cudaFree(tx); VBC_FAIL(VBC_ECUDA, "trsv: %s", cudaGetErrorString(e)); }
        db = tb; dx = tx;
    }
    int rc;
    const bool rows = A->desc_mode == DESC_ROWS;
    if (A->vt == VBC_F64) rc = rows ? launch_trsv<double, DESC_ROWS>(A, A->trsv, (const double *)db, (double *)dx) : launch_trsv<double, DESC_BLOCKS>(A, A->trsv, (const double *)db, (double *)dx);
    else rc = rows ? launch_trsv<float, DESC_ROWS>(A, A->trsv, (const float *)db, (float *)dx) : launch_trsv<float, DESC_BLOCKS>(A, A->trsv, (const float *)db, (float *)dx);
    if (!on_device) {
        cudaError_t e = cudaSuccess;
        if (rc == VBC_OK) e = cudaMemcpyAsync(x, tx, tv * (size_t)len, cudaMemcpyDeviceToHost, A->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(A->stream);
        int herr = 0;
        if (e == cudaSuccess) e = cudaMemcpy(&herr, A->trsv->d_err, sizeof(int), cudaMemcpyDeviceToHost);
        cudaFree(tb); cudaFree(tx);
        if (rc == VBC_OK && e != cudaSuccess) VBC_FAIL(VBC_ECUDA, "vbc_trsv_lower: %s", cudaGetErrorString(e));
        if (rc == VBC_OK && herr) VBC_FAIL(VBC_ECUDA, "vbc_trsv_lower: a dependency wait timed out (is tril(A') really the schedule that was analysed?)");
    }
    return rc;
}

int vbc_trsv_levels(const vbc_mat *A, int *nlevels)
{
    if (!A || !nlevels) VBC_FAIL(VBC_EARG, "NULL argument");
    if (!A->trsv) VBC_FAIL(VBC_EARG, "vbc_trsv_analyse has not been called");
    *nlevels = A->trsv->nlevels;
    return VBC_OK;
}

} // extern "C"

namespace vbc {
void destroy_trsv_plan(vbc_trsv_plan *p) { free_plan(p); }
int trsv_error_flag(const vbc_mat *A, int *flag)
{
    *flag = 0;
    if (!A->trsv) return VBC_OK;
    VBC_CUDA(cudaMemcpy(flag, A->trsv->d_err, sizeof(int), cudaMemcpyDeviceToHost));
    return VBC_OK;
}
} // namespace vbc
