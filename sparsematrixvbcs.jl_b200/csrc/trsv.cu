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
// Schedule: vbc_trsv_analyse computes level[l] = 1 + max level over the row blocks l depends on,
// and orders the row blocks by level.  ONE cooperative persistent kernel walks that order, one warp
// per row block; a warp spins (acquire loads) on the completion flags of exactly the row blocks it
// gathers from, so there is no grid-wide barrier between levels.  Dependencies always sit earlier
// in the order, all CTAs are co-resident (cooperative launch), hence the first unfinished row block
// can always run: no deadlock.  A wall-clock bound in the spin loop turns a lost dependency into an
// error instead of a hang.
#include <algorithm>
#include <new>
#include <vector>

#include "common.cuh"
#include "walk.cuh"

struct vbc_trsv_plan {
    int *d_order = nullptr;   // row blocks sorted by (level, index)
    int *d_c2s = nullptr;     // column -> stripe
    unsigned *d_flags = nullptr; // per stripe: epoch of the last solve that finished it
    int *d_err = nullptr;
    unsigned epoch = 0;
    int nlevels = 0;
    int wmax = 0;
};

namespace vbc {

constexpr int TRSV_WMAX = 8;

__device__ __forceinline__ unsigned ld_acquire_gpu_u32(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu_u32(unsigned *p, unsigned v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <typename Tv, int MODE>
__global__ void __launch_bounds__(256) k_trsv_lower(const StripeMeta *__restrict__ meta, const int *__restrict__ desc,
                                                     const Tv *__restrict__ val, const int *__restrict__ order,
                                                     const int *__restrict__ c2s, unsigned *__restrict__ flags, const unsigned epoch,
                                                     const Tv *__restrict__ bvec, Tv *x, const int L, const int u0, const int log2u,
                                                     int *__restrict__ err)
{
    __shared__ Tv Dsm[8][TRSV_WMAX][TRSV_WMAX + 1]; // diagonal block of each warp's row block
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int nwarps = (int)((gridDim.x * blockDim.x) >> 5);
    for (int t = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5); t < L; t += nwarps) {
        const int l = order[t];
        const StripeMeta a = ld_meta(meta + l), b = ld_meta(meta + l + 1);
        const int w = b.col - a.col;
        if (w <= 0) { if (lane == 0) st_release_gpu_u32(flags + l, epoch); continue; }
        const int j0 = a.col;
        const int R = (MODE == DESC_ROWS) ? (b.pos - a.pos) : (int)((b.ofs - a.ofs) / w);
        for (int i = lane; i < TRSV_WMAX * TRSV_WMAX; i += 32) Dsm[wib][i / TRSV_WMAX][i % TRSV_WMAX] = (Tv)0;
        __syncwarp();
        // element mapping: lane -> (row r0 + k*rps, column c)
        const int rps = 32 / w, c = lane % w, r0 = lane / w;
        const bool active = lane < rps * w;
        Tv acc = (Tv)0;
        if (active) {
            // Everything that does not depend on other row blocks is issued first (descriptors, values, the
            // column->stripe lookups), in batches of TB rows per lane; only then does the lane wait on the
            // flags of the row blocks it gathers from, and only then does it read x.
            constexpr int TB = 4;
            RowWalk<MODE> walk;
            walk.init(desc, a.pos, r0, rps, u0, log2u);
            const Tv *vp = val + a.ofs + (long long)r0 * w + c;
            for (int r = r0; r < R; r += TB * rps) {
                int xi[TB], dep[TB];
                Tv v[TB];
#pragma unroll
                for (int k = 0; k < TB; k++) {
                    const bool ok = r + k * rps < R;
                    xi[k] = walk.next_if(ok);
                    v[k] = ok ? __ldcs(vp) : (Tv)0;
                    vp += (long long)rps * w;
                    if (!ok) xi[k] = -1;
                }
#pragma unroll
                for (int k = 0; k < TB; k++) {
                    dep[k] = -1;
                    if (xi[k] >= 0 && xi[k] < j0) dep[k] = __ldg(c2s + xi[k]);
                    else if (xi[k] >= j0 && xi[k] < j0 + w) Dsm[wib][xi[k] - j0][c] = v[k]; // diagonal block row
                }
#pragma unroll
                for (int k = 0; k < TB; k++) {
                    if (dep[k] < 0) continue;
                    if (ld_acquire_gpu_u32(flags + dep[k]) != epoch) {
                        unsigned long long t0, t1;
                        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
                        while (ld_acquire_gpu_u32(flags + dep[k]) != epoch) {
                            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                            if (t1 - t0 > 3000000000ull) { atomicExch(err, 1); break; }
                        }
                    }
                }
                Tv xv[TB];
#pragma unroll
                for (int k = 0; k < TB; k++) xv[k] = dep[k] >= 0 ? __ldcg(x + xi[k]) : (Tv)0; // .cg: written by another SM
#pragma unroll
                for (int k = 0; k < TB; k++) acc = fma(v[k], xv[k], acc);
            }
        }
        for (int d = w; d < 32; d <<= 1) {
            const Tv tsum = __shfl_down_sync(0xffffffffu, acc, d);
            if (lane + d < 32) acc += tsum;
        }
        __syncwarp();
        // forward substitution inside the w x w diagonal block: lane dj owns unknown j0 + dj
        Tv rhs = (lane < w) ? bvec[j0 + lane] - acc : (Tv)0;
        Tv xj = (Tv)0;
        for (int d = 0; d < w; d++) {
            const Tv diag = Dsm[wib][d][d];
            const Tv xd = __shfl_sync(0xffffffffu, rhs, d) / diag; // x[j0 + d]
            if (lane == d) xj = xd;
            if (lane > d && lane < w) rhs -= Dsm[wib][d][lane] * xd;
        }
        if (lane < w) x[j0 + lane] = xj;
        __syncwarp();
        if (lane == 0) {
            __threadfence();
            st_release_gpu_u32(flags + l, epoch);
        }
    }
}

static void free_plan(vbc_trsv_plan *p)
{
    if (!p) return;
    cudaFree(p->d_order); cudaFree(p->d_c2s); cudaFree(p->d_flags); cudaFree(p->d_err);
    delete p;
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
    if (grid < 1) grid = 1;
    int L = (int)A->L, u0 = A->u0, log2u = -1;
    if (u0 > 0 && !(u0 & (u0 - 1))) { log2u = 0; while ((1 << log2u) < u0) log2u++; }
    P->epoch++;
    const StripeMeta *meta = A->d_meta;
    const int *desc = A->d_desc, *order = P->d_order, *c2s = P->d_c2s;
    const Tv *val = (const Tv *)A->d_val;
    unsigned *flags = P->d_flags;
    unsigned epoch = P->epoch;
    int *err = P->d_err;
    void *args[] = {&meta, &desc, &val, &order, &c2s, &flags, &epoch, &b, &x, &L, &u0, &log2u, &err};
    VBC_CUDA(cudaLaunchCooperativeKernel((const void *)k_trsv_lower<Tv, MODE>, dim3((unsigned)grid), dim3(256), args, 0, A->stream));
    A->launches++;
    return VBC_OK;
}

} // namespace vbc

using namespace vbc;

extern "C" {

int vbc_trsv_analyse(vbc_mat *A, int *nlevels)
{
    if (!A) VBC_FAIL(VBC_EARG, "matrix handle is NULL");
    if (A->m != A->n) VBC_FAIL(VBC_EDIM, "DimensionMismatch: triangular solve needs a square matrix, got %lld x %lld", (long long)A->m, (long long)A->n);
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
    int maxlevel = L > 0 ? 1 : 0;
    const bool blocks = A->desc_mode == DESC_BLOCKS;
    for (int64_t l = 0; l < L; l++) {
        const int j0 = meta[l].col;
        int lev = 1;
        for (int q = meta[l].pos; q < meta[l + 1].pos; q++) {
            const int i0 = desc[q];
            const int i1 = blocks ? std::min<int64_t>(i0 + A->u0, A->m) : i0 + 1; // rows [i0, i1)
            // rows below j0 are dependencies; they sit in at most two stripes' worth of rows for a block,
            // so visit each row of the unit
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
    if (e == cudaSuccess) e = cudaMalloc(&P->d_c2s, sizeof(int) * (size_t)(n > 0 ? n : 1));
    if (e == cudaSuccess) e = cudaMalloc(&P->d_flags, sizeof(unsigned) * (size_t)(L > 0 ? L : 1));
    if (e == cudaSuccess) e = cudaMalloc(&P->d_err, sizeof(int));
    if (e == cudaSuccess && L > 0) e = cudaMemcpy(P->d_order, order.data(), sizeof(int) * (size_t)L, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && n > 0) e = cudaMemcpy(P->d_c2s, c2s.data(), sizeof(int) * (size_t)n, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemset(P->d_flags, 0, sizeof(unsigned) * (size_t)(L > 0 ? L : 1));
    if (e == cudaSuccess) e = cudaMemset(P->d_err, 0, sizeof(int));
    if (e != cudaSuccess) { free_plan(P); VBC_FAIL(VBC_ECUDA, "vbc_trsv_analyse: %s", cudaGetErrorString(e)); }
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
    if (!A->trsv) VBC_TRY(vbc_trsv_analyse(A, nullptr));
    if (len == 0) return VBC_OK;
    if (!b || !x) VBC_FAIL(VBC_EARG, "NULL vector");
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
