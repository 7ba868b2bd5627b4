// api.cu -- the extern "C" surface declared in include/vbc.h.
#include <stdarg.h>
#include <stdlib.h>
#include <new>

#include "common.cuh"

namespace vbc {

static thread_local char g_err[1024] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
const char *get_error() { return g_err; }

static int check_types(int vt, int it)
{
    if (!vt_is_valid(vt)) VBC_FAIL(VBC_EARG, "vt must be VBC_F32, VBC_F64, VBC_INT32 or VBC_INT64, got %d", vt);
    if (it != VBC_I32 && it != VBC_I64) VBC_FAIL(VBC_EARG, "it must be VBC_I32 or VBC_I64, got %d", it);
    return VBC_OK;
}

// ArgumentError checks of the struct constructors, SparseMatrixVBCs.jl:45-50 and :72-79
static int check_shape(int64_t m, int64_t n, int U, int W, bool d2, int64_t K, int64_t L)
{
    if (m < 0) VBC_FAIL(VBC_EARG, "ArgumentError: number of rows (m) must be >= 0, got %lld", (long long)m);
    if (n < 0) VBC_FAIL(VBC_EARG, "ArgumentError: number of columns (n) must be >= 0, got %lld", (long long)n);
    if (W <= 0) VBC_FAIL(VBC_EARG, "ArgumentError: W must be > 0");
    if (d2 && U <= 0) VBC_FAIL(VBC_EARG, "ArgumentError: U must be > 0");
    if (L < 0 || (d2 && K < 0)) VBC_FAIL(VBC_EARG, "partition lengths must be >= 0");
    return VBC_OK;
}

static int64_t rd(const void *a, int it, int64_t i) { return it == VBC_I64 ? ((const int64_t *)a)[i] : (int64_t)((const int32_t *)a)[i]; }

static int check_spl_host(const void *spl, int it, int64_t P, int64_t dim, int64_t limit, const char *what, int over_code, const char *over_msg)
{
    if (rd(spl, it, 0) != 1 || rd(spl, it, P) != dim + 1) VBC_FAIL(VBC_EARG, "%s is not a SplitPartition of 1:%lld (spl[1]=%lld, spl[end]=%lld)", what, (long long)dim, (long long)rd(spl, it, 0), (long long)rd(spl, it, P));
    for (int64_t k = 0; k < P; k++) {
        const int64_t d = rd(spl, it, k + 1) - rd(spl, it, k);
        if (d < 0) VBC_FAIL(VBC_EARG, "%s.spl is decreasing at %lld", what, (long long)k);
        if (d > limit) VBC_FAIL(over_code, "%s", over_msg);
    }
    return VBC_OK;
}

static vbc_mat *new_mat(int vt, int it, int64_t m, int64_t n, int U, int W, bool d2, int64_t K, int64_t L, int device)
{
    vbc_mat *A = new (std::nothrow) vbc_mat();
    if (!A) return nullptr;
    A->vt = vt; A->it = it; A->m = m; A->n = n; A->U = d2 ? U : 0; A->W = W; A->ndim = d2 ? 2 : 1;
    A->K = d2 ? K : 0; A->L = L; A->device = device;
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && sms > 0) A->sm_count = sms;
    return A;
}

static int copy_in(void **dst, const void *src, size_t bytes, cudaMemcpyKind kind, cudaStream_t st)
{
    VBC_CUDA(cudaMalloc(dst, bytes > 0 ? bytes : 1));
    if (bytes > 0) VBC_CUDA(cudaMemcpyAsync(*dst, src, bytes, kind, st));
    return VBC_OK;
}

static int pack_common(vbc_mat **out, int vt, int it, int64_t m, int64_t n, int U, int W, const void *colptr, const void *rowval,
                       const void *nzval, const void *pi_spl, int64_t K, const void *phi_spl, int64_t L, int device, bool on_dev)
{
    if (!out) VBC_FAIL(VBC_EARG, "out is NULL");
    *out = nullptr;
    VBC_TRY(check_types(vt, it));
    const bool d2 = pi_spl != nullptr;
    VBC_TRY(check_shape(m, n, U, W, d2, K, L));
    if (!colptr || !phi_spl || (!rowval && !on_dev)) VBC_FAIL(VBC_EARG, "NULL array argument");
    DeviceGuard guard(device);
    if (!guard.ok) VBC_FAIL(VBC_ECUDA, "cudaSetDevice(%d) failed", device);
    const size_t ti = it_size(it), tv = vt_size(vt);
    const cudaMemcpyKind kind = on_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;

    int64_t nnz = 0;
    if (on_dev) {
        char tmp[8] = {0};
        VBC_CUDA(cudaMemcpy(tmp, (const char *)colptr + ti * (size_t)n, ti, cudaMemcpyDeviceToHost));
        nnz = rd(tmp, it, 0) - 1;
    } else {
        if (rd(colptr, it, 0) != 1) VBC_FAIL(VBC_EARG, "colptr[1] must be 1");
        for (int64_t j = 0; j < n; j++)
            if (rd(colptr, it, j + 1) < rd(colptr, it, j)) VBC_FAIL(VBC_EARG, "colptr is decreasing at column %lld", (long long)(j + 1));
        nnz = rd(colptr, it, n) - 1;
    }
    if (nnz < 0) VBC_FAIL(VBC_EARG, "colptr[end] < 1");

    vbc_mat *A = new_mat(vt, it, m, n, U, W, d2, K, L, device);
    if (!A) VBC_FAIL(VBC_ENOMEM, "host allocation failed");
    int rc = VBC_OK;
    void *d_colptr = nullptr, *d_rowval = nullptr, *d_nzval = nullptr;
    do {
        if ((rc = copy_in(&A->d_phi_spl, phi_spl, ti * (size_t)(L + 1), kind, A->stream)) != VBC_OK) break;
        if (d2 && (rc = copy_in(&A->d_pi_spl, pi_spl, ti * (size_t)(K + 1), kind, A->stream)) != VBC_OK) break;
        if (on_dev) {
            d_colptr = const_cast<void *>(colptr); d_rowval = const_cast<void *>(rowval); d_nzval = const_cast<void *>(nzval);
        } else {
            if ((rc = copy_in(&d_colptr, colptr, ti * (size_t)(n + 1), kind, A->stream)) != VBC_OK) break;
            if ((rc = copy_in(&d_rowval, rowval, ti * (size_t)nnz, kind, A->stream)) != VBC_OK) break;
            if ((rc = copy_in(&d_nzval, nzval, tv * (size_t)nnz, kind, A->stream)) != VBC_OK) break;
        }
        rc = pack_from_device_csc(A, d_colptr, d_rowval, d_nzval);
    } while (0);
    if (!on_dev) { cudaFree(d_colptr); cudaFree(d_rowval); cudaFree(d_nzval); }
    if (rc != VBC_OK) { vbc_destroy(A); return rc; }
    *out = A;
    return VBC_OK;
}

static int ensure_vec(void **buf, int64_t *cap, int64_t len, size_t esz)
{
    if (*cap >= len && *buf) return VBC_OK;
    if (*buf) { cudaFree(*buf); *buf = nullptr; *cap = 0; }
    VBC_CUDA(cudaMalloc(buf, esz * (size_t)(len > 0 ? len : 1)));
    *cap = len;
    return VBC_OK;
}

} // namespace vbc

using namespace vbc;

// largest x index gathered by the descriptors [d0, d1) -> *out (atomicMax), for the chunked host-vector multiply
static __global__ void __launch_bounds__(256) k_desc_max(const int *__restrict__ desc, const long long d0, const long long d1, int *out, int *out_min)
{
    int mx = -1, mn = 0x7fffffff;
    for (long long i = d0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < d1; i += (long long)gridDim.x * blockDim.x) {
        const int v = desc[i];
        mx = max(mx, v);
        mn = min(mn, v);
    }
    for (int d = 16; d > 0; d >>= 1) {
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, d));
        mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, d));
    }
    if ((threadIdx.x & 31) == 0 && mx >= 0) { atomicMax(out, mx); atomicMin(out_min, mn); }
}

// chunk_xhi[c] = one past the largest x index the stripes of chunks 0..c gather from (cumulative, clamped to m)
static int prepare_x_ranges(vbc_mat *A)
{
    A->xhi_ready = -1;
    const int NC = A->nchunks;
    int *d_mx = nullptr;
    VBC_CUDA(cudaMalloc(&d_mx, sizeof(int) * 9)); // [0..7]: per-chunk max, [8]: global min
    int rc = VBC_OK;
    int h_mx[9];
    do {
        const int big = 0x7fffffff;
        if (cudaMemsetAsync(d_mx, 0xff, sizeof(int) * 8, A->stream) != cudaSuccess ||  // -1
            cudaMemcpyAsync(d_mx + 8, &big, sizeof(int), cudaMemcpyHostToDevice, A->stream) != cudaSuccess) { rc = VBC_ECUDA; break; }
        for (int c = 0; c < NC && rc == VBC_OK; c++) {
            StripeMeta m0, m1;
            if (cudaMemcpy(&m0, A->d_meta + A->chunk_l[c], sizeof(StripeMeta), cudaMemcpyDeviceToHost) != cudaSuccess ||
                cudaMemcpy(&m1, A->d_meta + A->chunk_l[c + 1], sizeof(StripeMeta), cudaMemcpyDeviceToHost) != cudaSuccess) { rc = VBC_ECUDA; break; }
            const long long n = (long long)m1.pos - m0.pos;
            if (n <= 0) continue;
            long long g = (n + 255) / 256;
            if (g > (long long)A->sm_count * 8) g = (long long)A->sm_count * 8;
            k_desc_max<<<(unsigned)g, 256, 0, A->stream>>>(A->d_desc, m0.pos, m1.pos, d_mx + c, d_mx + 8);
        }
        if (rc != VBC_OK) break;
        if (cudaMemcpyAsync(h_mx, d_mx, sizeof(int) * 9, cudaMemcpyDeviceToHost, A->stream) != cudaSuccess || cudaStreamSynchronize(A->stream) != cudaSuccess) { rc = VBC_ECUDA; break; }
    } while (0);
    cudaFree(d_mx);
    if (rc != VBC_OK) { cudaGetLastError(); set_error("x-range analysis of the chunked multiply failed"); return rc; }
    // a block descriptor is the FIRST x index of a block of up to u0 rows (DESC_BLOCKS); a row descriptor is the index itself
    const int64_t reach = (A->desc_mode == DESC_BLOCKS) ? A->u0 : 1;
    int64_t hi = 0;
    for (int c = 0; c < NC; c++) {
        if (h_mx[c] >= 0 && (int64_t)h_mx[c] + reach > hi) hi = (int64_t)h_mx[c] + reach;
        if (hi > A->m) hi = A->m;
        A->chunk_xhi[c] = hi;
    }
    A->x_lo = (h_mx[8] != 0x7fffffff && h_mx[8] > 0) ? h_mx[8] : 0; // nothing below the smallest gathered index is ever read
    if (A->x_lo > hi) A->x_lo = hi;
    A->xhi_ready = 1;
    return VBC_OK;
}

using namespace vbc;

extern "C" {

const char *vbc_last_error(void) { return get_error(); }
int vbc_version(void) { return 100; }

int vbc_device_count(int *count)
{
    if (!count) VBC_FAIL(VBC_EARG, "count is NULL");
    VBC_CUDA(cudaGetDeviceCount(count));
    return VBC_OK;
}

int vbc_pack_csc(vbc_mat **out, int vt, int it, int64_t m, int64_t n, int U, int W, const void *colptr, const void *rowval,
                 const void *nzval, const void *pi_spl, int64_t K, const void *phi_spl, int64_t L, int device)
{
    return pack_common(out, vt, it, m, n, U, W, colptr, rowval, nzval, pi_spl, K, phi_spl, L, device, false);
}

int vbc_pack_csc_dev(vbc_mat **out, int vt, int it, int64_t m, int64_t n, int U, int W, const void *colptr, const void *rowval,
                     const void *nzval, const void *pi_spl, int64_t K, const void *phi_spl, int64_t L, int device)
{
    return pack_common(out, vt, it, m, n, U, W, colptr, rowval, nzval, pi_spl, K, phi_spl, L, device, true);
}

int vbc_upload(vbc_mat **out, int vt, int it, int64_t m, int64_t n, int U, int W, const void *pi_spl, int64_t K, const void *phi_spl,
               int64_t L, const void *pos, const void *idx, const void *ofs, const void *val, int device)
{
    if (!out) VBC_FAIL(VBC_EARG, "out is NULL");
    *out = nullptr;
    VBC_TRY(check_types(vt, it));
    const bool d2 = pi_spl != nullptr;
    VBC_TRY(check_shape(m, n, U, W, d2, K, L));
    if (!phi_spl || !pos || !ofs) VBC_FAIL(VBC_EARG, "NULL array argument");
    VBC_TRY(check_spl_host(phi_spl, it, L, n, W, "Φ", VBC_ELIMIT, "AssertionError: w <= W"));
    if (d2) VBC_TRY(check_spl_host(pi_spl, it, K, m, U, "Π", VBC_ELIMIT, "AssertionError: u <= U"));
    if (rd(pos, it, 0) != 1 || rd(ofs, it, 0) != 1) VBC_FAIL(VBC_EARG, "pos[1] and ofs[1] must be 1");
    for (int64_t l = 0; l < L; l++)
        if (rd(pos, it, l + 1) < rd(pos, it, l) || rd(ofs, it, l + 1) < rd(ofs, it, l)) VBC_FAIL(VBC_EARG, "pos/ofs decreasing at stripe %lld", (long long)(l + 1));
    const int64_t nidx = rd(pos, it, L) - 1, nval = rd(ofs, it, L) - 1;
    if ((nidx > 0 && !idx) || (nval > 0 && !val)) VBC_FAIL(VBC_EARG, "NULL idx/val");
    // idx range check: a malformed idx would make the kernels gather out of bounds
    const int64_t idx_hi = d2 ? K : m;
    for (int64_t t = 0; t < nidx; t++)
        if (rd(idx, it, t) < 1 || rd(idx, it, t) > idx_hi) VBC_FAIL(VBC_EARG, "idx[%lld]=%lld out of 1:%lld", (long long)(t + 1), (long long)rd(idx, it, t), (long long)idx_hi);
    DeviceGuard guard(device);
    if (!guard.ok) VBC_FAIL(VBC_ECUDA, "cudaSetDevice(%d) failed", device);
    vbc_mat *A = new_mat(vt, it, m, n, U, W, d2, K, L, device);
    if (!A) VBC_FAIL(VBC_ENOMEM, "host allocation failed");
    A->nidx = nidx; A->nval = nval;
    const size_t ti = it_size(it), tv = vt_size(vt);
    int rc = VBC_OK;
    do {
        if ((rc = copy_in(&A->d_phi_spl, phi_spl, ti * (size_t)(L + 1), cudaMemcpyHostToDevice, A->stream)) != VBC_OK) break;
        if (d2 && (rc = copy_in(&A->d_pi_spl, pi_spl, ti * (size_t)(K + 1), cudaMemcpyHostToDevice, A->stream)) != VBC_OK) break;
        if ((rc = copy_in(&A->d_pos, pos, ti * (size_t)(L + 1), cudaMemcpyHostToDevice, A->stream)) != VBC_OK) break;
        if ((rc = copy_in(&A->d_ofs, ofs, ti * (size_t)(L + 1), cudaMemcpyHostToDevice, A->stream)) != VBC_OK) break;
        if ((rc = copy_in(&A->d_idx, idx, ti * (size_t)nidx, cudaMemcpyHostToDevice, A->stream)) != VBC_OK) break;
        const size_t pad = 64;
        if (cudaMalloc(&A->d_val, tv * ((size_t)nval + pad)) != cudaSuccess) { set_error("cudaMalloc(val)"); rc = VBC_ENOMEM; break; }
        if (cudaMemsetAsync((char *)A->d_val + tv * (size_t)nval, 0, tv * pad, A->stream) != cudaSuccess) { set_error("memset"); rc = VBC_ECUDA; break; }
        if (nval > 0 && cudaMemcpyAsync(A->d_val, val, tv * (size_t)nval, cudaMemcpyHostToDevice, A->stream) != cudaSuccess) { set_error("memcpy(val)"); rc = VBC_ECUDA; break; }
        rc = finalize_layout(A, pi_spl);
        if (rc == VBC_OK && cudaStreamSynchronize(A->stream) != cudaSuccess) { set_error("upload: sync failed: %s", cudaGetErrorString(cudaGetLastError())); rc = VBC_ECUDA; }
    } while (0);
    if (rc != VBC_OK) { vbc_destroy(A); return rc; }
    *out = A;
    return VBC_OK;
}

void vbc_destroy(vbc_mat *A)
{
    if (!A) return;
    DeviceGuard guard(A->device);
    cudaFree(A->d_pi_spl); cudaFree(A->d_phi_spl); cudaFree(A->d_pos); cudaFree(A->d_idx); cudaFree(A->d_ofs); cudaFree(A->d_val);
    cudaFree(A->d_meta); cudaFree(A->d_desc); cudaFree(A->d_brow); cudaFree(A->d_order); cudaFree(A->d_px); cudaFree(A->d_py); cudaFree(A->d_x); cudaFree(A->d_y);
    destroy_trsv_plan(A->trsv);
    destroy_tindex(A->tindex);
    if (A->e2e_exec) cudaGraphExecDestroy(A->e2e_exec);
    if (A->e2e_stream) cudaStreamDestroy(A->e2e_stream);
    for (int k = 0; k < 3; k++)
        if (A->e2e_ev[k]) cudaEventDestroy(A->e2e_ev[k]);
    if (A->copy_stream) cudaStreamDestroy(A->copy_stream);
    if (A->h2d_stream) cudaStreamDestroy(A->h2d_stream);
    for (int c = 0; c < 8; c++)
        if (A->h2d_ev[c]) cudaEventDestroy(A->h2d_ev[c]);
    for (int c = 0; c < 8; c++)
        if (A->chunk_ev[c]) cudaEventDestroy(A->chunk_ev[c]);
    delete A;
}

int vbc_shape(const vbc_mat *A, int64_t *m, int64_t *n, int64_t *K, int64_t *L, int *U, int *W, int *ndim, int *vt, int *it)
{
    if (!A) VBC_FAIL(VBC_EARG, "matrix handle is NULL");
    if (m) *m = A->m; if (n) *n = A->n; if (K) *K = A->K; if (L) *L = A->L;
    if (U) *U = A->U; if (W) *W = A->W; if (ndim) *ndim = A->ndim; if (vt) *vt = A->vt; if (it) *it = A->it;
    return VBC_OK;
}

int vbc_sizes(const vbc_mat *A, int64_t *nidx, int64_t *nval)
{
    if (!A) VBC_FAIL(VBC_EARG, "matrix handle is NULL");
    if (nidx) *nidx = A->nidx;
    if (nval) *nval = A->nval;
    return VBC_OK;
}

int vbc_download(const vbc_mat *A, void *pos, void *idx, void *ofs, void *val)
{
    if (!A) VBC_FAIL(VBC_EARG, "matrix handle is NULL");
    DeviceGuard guard(A->device);
    const size_t ti = it_size(A->it), tv = vt_size(A->vt);
    VBC_CUDA(cudaStreamSynchronize(A->stream));
    if (pos) VBC_CUDA(cudaMemcpy(pos, A->d_pos, ti * (size_t)(A->L + 1), cudaMemcpyDeviceToHost));
    if (ofs) VBC_CUDA(cudaMemcpy(ofs, A->d_ofs, ti * (size_t)(A->L + 1), cudaMemcpyDeviceToHost));
    if (idx && A->nidx > 0) VBC_CUDA(cudaMemcpy(idx, A->d_idx, ti * (size_t)A->nidx, cudaMemcpyDeviceToHost));
    if (val && A->nval > 0) VBC_CUDA(cudaMemcpy(val, A->d_val, tv * (size_t)A->nval, cudaMemcpyDeviceToHost));
    return VBC_OK;
}

int vbc_format_bytes(const vbc_mat *A, int64_t bytes[3])
{
    if (!A || !bytes) VBC_FAIL(VBC_EARG, "NULL argument");
    const int64_t ti = (int64_t)it_size(A->it), tv = (int64_t)vt_size(A->vt);
    bytes[0] = ti * (A->L + 1) * 3 + (A->ndim == 2 ? ti * (A->K + 1) : 0) + ti * A->nidx + tv * A->nval;
    if (A->opt_parity) { bytes[1] = bytes[0]; bytes[2] = bytes[0]; return VBC_OK; }
    bytes[1] = (int64_t)sizeof(StripeMeta) * (A->L + 1) + 4 * A->ndesc + tv * A->nval + (A->d_order ? 4 * A->L : 0);
    bytes[2] = (A->tindex && A->opt_fwd_atomic != 1) ? tv * A->nval + tindex_bytes(A) : bytes[1];
    return VBC_OK;
}

int vbc_memory_cost(const vbc_mat *A, int64_t *cost, int64_t *row_term)
{
    if (!A || (!cost && A->L > 0)) VBC_FAIL(VBC_EARG, "NULL argument");
    DeviceGuard guard(A->device);
    return memory_cost_device(A, cost, row_term);
}

// Host-vector adjoint multiply, chunked: enqueues (no synchronisation) the upload of x -- in pieces on h2d_stream when
// `pipeline`, each chunk of stripes waiting only for the rows it gathers from -- the chunk kernels on A->stream and the
// copies of the finished y ranges on copy_stream.  Used eagerly and under stream capture (the replayed graph below).
static int host_adjoint_enqueue(vbc_mat *A, double alpha, const void *x, int64_t xlen, double beta, void *y, int64_t ylen, bool pipeline, int64_t *uploaded)
{
    const size_t tv = vt_size(A->vt);
    if (!pipeline && xlen > 0) VBC_CUDA(cudaMemcpyAsync(A->d_x, x, tv * (size_t)xlen, cudaMemcpyHostToDevice, A->stream));
    if (beta != 0.0 && ylen > 0) VBC_CUDA(cudaMemcpyAsync(A->d_y, y, tv * (size_t)ylen, cudaMemcpyHostToDevice, A->stream));
    int rc = VBC_OK;
    int64_t xcopied = pipeline ? A->x_lo : 0;
    for (int c = 0; c < A->nchunks && rc == VBC_OK; c++) {
        if (pipeline) { // the piece of x this chunk still lacks, on the upload stream; the chunk's kernel waits for it
            const int64_t hi = A->chunk_xhi[c];
            cudaError_t e = cudaSuccess;
            if (hi > xcopied) { e = cudaMemcpyAsync((char *)A->d_x + tv * (size_t)xcopied, (const char *)x + tv * (size_t)xcopied, tv * (size_t)(hi - xcopied), cudaMemcpyHostToDevice, A->h2d_stream); xcopied = hi; }
            if (e == cudaSuccess) e = cudaEventRecord(A->h2d_ev[c], A->h2d_stream);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(A->stream, A->h2d_ev[c], 0);
            if (e != cudaSuccess) { set_error("pipelined upload failed: %s", cudaGetErrorString(e)); rc = VBC_ECUDA; break; }
        }
        A->range_l0 = A->chunk_l[c]; A->range_l1 = A->chunk_l[c + 1];
        rc = launch_spmv(A, 1, alpha, A->d_x, beta, A->d_y);
        if (rc == VBC_OK && cudaEventRecord(A->chunk_ev[c], A->stream) != cudaSuccess) { set_error("cudaEventRecord failed"); rc = VBC_ECUDA; }
    }
    A->range_l0 = A->range_l1 = -1;
    VBC_TRY(rc);
    for (int c = 0; c < A->nchunks; c++) {
        const int64_t c0 = A->chunk_col[c], c1 = A->chunk_col[c + 1];
        VBC_CUDA(cudaStreamWaitEvent(A->copy_stream, A->chunk_ev[c], 0));
        if (c1 > c0) VBC_CUDA(cudaMemcpyAsync((char *)y + tv * (size_t)c0, (char *)A->d_y + tv * (size_t)c0, tv * (size_t)(c1 - c0), cudaMemcpyDeviceToHost, A->copy_stream));
    }
    *uploaded = pipeline ? xcopied - A->x_lo : xlen;
    return VBC_OK;
}

static bool is_pinned_host(const void *p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
}

static void drop_e2e_graph(vbc_mat *A)
{
    if (A->e2e_exec) { cudaGraphExecDestroy(A->e2e_exec); A->e2e_exec = nullptr; }
    A->e2e_x = nullptr; A->e2e_y = nullptr; A->e2e_seen = 0;
}

// The same work as ONE replayed CUDA graph.  An iterative caller multiplies with the same pinned x and y again and again
// (`mul!(y, A', x)` in a loop): the second call with the same pointers and scalars captures the enqueue above -- copies,
// chunk kernels and their cross-stream dependencies -- into a graph on an internal stream, and every later call is a
// single cudaGraphLaunch instead of ~50 runtime calls, which on a PCIe-bound path is most of the host time.
static int host_adjoint_graph(vbc_mat *A, double alpha, const void *x, int64_t xlen, double beta, void *y, int64_t ylen, bool pipeline, bool *done)
{
    *done = false;
    if (!A->opt_e2e_graph) return VBC_OK;
    if (A->e2e_x != x || A->e2e_y != y || A->e2e_alpha != alpha || A->e2e_beta != beta || A->e2e_pipeline != (int)pipeline) {
        drop_e2e_graph(A);
        A->e2e_x = x; A->e2e_y = y; A->e2e_alpha = alpha; A->e2e_beta = beta; A->e2e_pipeline = (int)pipeline;
        A->e2e_seen = 1;
        return VBC_OK; // first sighting: run eagerly
    }
    if (!A->e2e_exec) {
        if (A->e2e_seen < 0) return VBC_OK; // capture failed before for these buffers: stay eager
        if (!is_pinned_host(x) || !is_pinned_host(y)) { A->e2e_seen = -1; return VBC_OK; }
        if (!A->e2e_stream && cudaStreamCreateWithFlags(&A->e2e_stream, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); A->e2e_seen = -1; return VBC_OK; }
        for (int k = 0; k < 3; k++)
            if (!A->e2e_ev[k] && cudaEventCreateWithFlags(&A->e2e_ev[k], cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); A->e2e_seen = -1; return VBC_OK; }
        cudaStream_t user = A->stream;
        VBC_CUDA(cudaStreamSynchronize(user)); // earlier device work of this handle is finished before the internal stream takes over
        A->stream = A->e2e_stream;
        cudaGraph_t g = nullptr;
        int64_t up = 0;
        cudaError_t e = cudaStreamBeginCapture(A->e2e_stream, cudaStreamCaptureModeThreadLocal);
        int rc = VBC_OK;
        if (e == cudaSuccess) {
            e = cudaEventRecord(A->e2e_ev[0], A->e2e_stream);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(A->copy_stream, A->e2e_ev[0], 0);
            if (e == cudaSuccess && A->h2d_stream) e = cudaStreamWaitEvent(A->h2d_stream, A->e2e_ev[0], 0);
            if (e == cudaSuccess) rc = host_adjoint_enqueue(A, alpha, x, xlen, beta, y, ylen, pipeline, &up);
            if (e == cudaSuccess && rc == VBC_OK) e = cudaEventRecord(A->e2e_ev[1], A->copy_stream);
            if (e == cudaSuccess && rc == VBC_OK) e = cudaStreamWaitEvent(A->e2e_stream, A->e2e_ev[1], 0);
            if (e == cudaSuccess && rc == VBC_OK && A->h2d_stream) { e = cudaEventRecord(A->e2e_ev[2], A->h2d_stream); if (e == cudaSuccess) e = cudaStreamWaitEvent(A->e2e_stream, A->e2e_ev[2], 0); }
            cudaError_t e2 = cudaStreamEndCapture(A->e2e_stream, &g);
            if (e == cudaSuccess) e = e2;
        }
        A->stream = user;
        if (e == cudaSuccess && rc == VBC_OK && g) e = cudaGraphInstantiate(&A->e2e_exec, g, 0);
        if (g) cudaGraphDestroy(g);
        if (e != cudaSuccess || rc != VBC_OK || !A->e2e_exec) { // capture is an optimisation: fall back to the eager path for good
            cudaGetLastError();
            A->e2e_exec = nullptr;
            A->e2e_seen = -1;
            return VBC_OK;
        }
        A->e2e_upload = up;
    }
    VBC_CUDA(cudaGraphLaunch(A->e2e_exec, A->e2e_stream));
    VBC_CUDA(cudaStreamSynchronize(A->e2e_stream));
    A->last_upload_elems = A->e2e_upload;
    *done = true;
    return VBC_OK;
}

extern "C" int vbc_spmv(vbc_mat *A, int trans, double alpha, const void *x, int64_t xlen, double beta, void *y, int64_t ylen, int on_device)
{
    if (!A) VBC_FAIL(VBC_EARG, "matrix handle is NULL");
    const int64_t need_x = trans ? A->m : A->n, need_y = trans ? A->n : A->m;
    if (xlen != need_x || ylen != need_y)
        VBC_FAIL(VBC_EDIM, "DimensionMismatch: op(A) is %lld x %lld, x has %lld, y has %lld", (long long)need_y, (long long)need_x, (long long)xlen, (long long)ylen);
    if ((!x && xlen > 0) || (!y && ylen > 0)) VBC_FAIL(VBC_EARG, "NULL vector");
    DeviceGuard guard(A->device);
    if (!guard.ok) VBC_FAIL(VBC_ECUDA, "cudaSetDevice(%d) failed", A->device);
    if (on_device) return launch_spmv(A, trans, alpha, x, beta, y);
    const size_t tv = vt_size(A->vt);
    VBC_TRY(ensure_vec(&A->d_x, &A->x_cap, xlen, tv));
    VBC_TRY(ensure_vec(&A->d_y, &A->y_cap, ylen, tv));
    const bool chunkable = trans && !A->opt_parity && A->d_order == nullptr && vt_is_float(A->vt);
    // adjoint with a large y: launch the stripes in chunks and copy each finished y range back on a second stream
    // while the next chunk computes (the D2H copy is as long as the whole kernel)
    if (chunkable && A->nchunks == 0) {
        A->nchunks = -1;
        int NC = 8;
        if (const char *e = getenv("VBC_E2E_CHUNKS")) { NC = atoi(e); if (NC > 8) NC = 8; }
        if (NC >= 2 && A->L >= 64 * NC && tv * (size_t)ylen >= (1u << 20)) {
            bool ok = cudaStreamCreateWithFlags(&A->copy_stream, cudaStreamNonBlocking) == cudaSuccess;
            for (int c = 0; c < NC && ok; c++) ok = cudaEventCreateWithFlags(&A->chunk_ev[c], cudaEventDisableTiming) == cudaSuccess;
            for (int c = 0; c <= NC && ok; c++) {
                A->chunk_l[c] = (int)(A->L * c / NC);
                StripeMeta mt;
                ok = cudaMemcpy(&mt, A->d_meta + A->chunk_l[c], sizeof(StripeMeta), cudaMemcpyDeviceToHost) == cudaSuccess;
                A->chunk_col[c] = mt.col;
            }
            if (ok) A->nchunks = NC;
            else cudaGetLastError();
        }
    }
    if (chunkable && A->nchunks > 0) {
        // pipelined upload: needs the x ranges of the chunks and a third stream; anything missing -> plain upload first
        bool pipeline = false;
        if (A->opt_e2e_pipeline && xlen > 0) {
            if (A->xhi_ready == 0) {
                bool ok = cudaStreamCreateWithFlags(&A->h2d_stream, cudaStreamNonBlocking) == cudaSuccess;
                for (int c = 0; c < A->nchunks && ok; c++) ok = cudaEventCreateWithFlags(&A->h2d_ev[c], cudaEventDisableTiming) == cudaSuccess;
                if (ok) ok = prepare_x_ranges(A) == VBC_OK;
                if (!ok) { cudaGetLastError(); A->xhi_ready = -1; }
            }
            pipeline = A->xhi_ready == 1;
        }
        bool done = false;
        VBC_TRY(host_adjoint_graph(A, alpha, x, xlen, beta, y, ylen, pipeline, &done));
        if (done) return VBC_OK;
        int64_t up = 0;
        VBC_TRY(host_adjoint_enqueue(A, alpha, x, xlen, beta, y, ylen, pipeline, &up));
        VBC_CUDA(cudaStreamSynchronize(A->copy_stream));
        VBC_CUDA(cudaStreamSynchronize(A->stream));
        if (pipeline) VBC_CUDA(cudaStreamSynchronize(A->h2d_stream));
        A->last_upload_elems = up;
        return VBC_OK;
    }
    if (xlen > 0) VBC_CUDA(cudaMemcpyAsync(A->d_x, x, tv * (size_t)xlen, cudaMemcpyHostToDevice, A->stream));
    if (beta != 0.0 && ylen > 0) VBC_CUDA(cudaMemcpyAsync(A->d_y, y, tv * (size_t)ylen, cudaMemcpyHostToDevice, A->stream));
    VBC_TRY(launch_spmv(A, trans, alpha, A->d_x, beta, A->d_y));
    if (ylen > 0) VBC_CUDA(cudaMemcpyAsync(y, A->d_y, tv * (size_t)ylen, cudaMemcpyDeviceToHost, A->stream));
    VBC_CUDA(cudaStreamSynchronize(A->stream));
    A->last_upload_elems = xlen;
    return VBC_OK;
}

extern "C" int vbc_spmv_mixed(vbc_mat *A, int trans, double alpha, const void *x, int64_t xlen, double beta, void *y, int64_t ylen, int vec_vt, int on_device)
{
    if (!A) VBC_FAIL(VBC_EARG, "matrix handle is NULL");
    if (vec_vt == A->vt && vt_is_valid(vec_vt)) return vbc_spmv(A, trans, alpha, x, xlen, beta, y, ylen, on_device);
    if (!vt_is_float(A->vt) || !vt_is_float(vec_vt)) VBC_FAIL(VBC_EARG, "ArgumentError: the mixed-type multiply pairs a Float32 matrix with Float64 vectors; integer element types multiply vectors of their own type");
    if (vec_vt == A->vt) return vbc_spmv(A, trans, alpha, x, xlen, beta, y, ylen, on_device);
    if (vec_vt == VBC_F32) VBC_FAIL(VBC_EARG, "ArgumentError: Float64 values with Float32 vectors (narrowing) is not offered");
    const int64_t need_x = trans ? A->m : A->n, need_y = trans ? A->n : A->m;
    if (xlen != need_x || ylen != need_y)
        VBC_FAIL(VBC_EDIM, "DimensionMismatch: op(A) is %lld x %lld, x has %lld, y has %lld", (long long)need_y, (long long)need_x, (long long)xlen, (long long)ylen);
    if ((!x && xlen > 0) || (!y && ylen > 0)) VBC_FAIL(VBC_EARG, "NULL vector");
    DeviceGuard guard(A->device);
    if (!guard.ok) VBC_FAIL(VBC_ECUDA, "cudaSetDevice(%d) failed", A->device);
    if (on_device) return launch_spmv_mixed(A, trans, alpha, x, beta, y);
    const size_t tu = vt_size(vec_vt);
    void *dx = nullptr, *dy = nullptr; // wider than the handle's staging vectors: call-local buffers
    int rc = VBC_OK;
    do {
        if (cudaMalloc(&dx, tu * (size_t)(xlen > 0 ? xlen : 1)) != cudaSuccess || cudaMalloc(&dy, tu * (size_t)(ylen > 0 ? ylen : 1)) != cudaSuccess) { cudaGetLastError(); set_error("cudaMalloc of the staging vectors failed"); rc = VBC_ENOMEM; break; }
        cudaError_t e = cudaSuccess;
        if (xlen > 0) e = cudaMemcpyAsync(dx, x, tu * (size_t)xlen, cudaMemcpyHostToDevice, A->stream);
        if (e == cudaSuccess && ylen > 0 && beta != 0.0) e = cudaMemcpyAsync(dy, y, tu * (size_t)ylen, cudaMemcpyHostToDevice, A->stream);
        if (e != cudaSuccess) { set_error("host-to-device copy failed: %s", cudaGetErrorString(e)); rc = VBC_ECUDA; break; }
        if ((rc = launch_spmv_mixed(A, trans, alpha, dx, beta, dy)) != VBC_OK) break;
        if (ylen > 0) e = cudaMemcpyAsync(y, dy, tu * (size_t)ylen, cudaMemcpyDeviceToHost, A->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(A->stream);
        if (e != cudaSuccess) { set_error("device-to-host copy failed: %s", cudaGetErrorString(e)); rc = VBC_ECUDA; }
    } while (0);
    cudaFree(dx); cudaFree(dy);
    return rc;
}

int vbc_csc_upload(vbc_csc **out, int vt, int it, int64_t m, int64_t n, const void *colptr, const void *rowval, const void *nzval, int device)
{
    if (!out) VBC_FAIL(VBC_EARG, "out is NULL");
    *out = nullptr;
    VBC_TRY(check_types(vt, it));
    if (m < 0 || n < 0) VBC_FAIL(VBC_EARG, "ArgumentError: m, n must be >= 0");
    if (!colptr) VBC_FAIL(VBC_EARG, "colptr is NULL");
    if (rd(colptr, it, 0) != 1) VBC_FAIL(VBC_EARG, "colptr[1] must be 1");
    for (int64_t j = 0; j < n; j++)
        if (rd(colptr, it, j + 1) < rd(colptr, it, j)) VBC_FAIL(VBC_EARG, "colptr is decreasing at column %lld", (long long)(j + 1));
    const int64_t nnz = rd(colptr, it, n) - 1;
    for (int64_t t = 0; t < nnz; t++)
        if (rd(rowval, it, t) < 1 || rd(rowval, it, t) > m) VBC_FAIL(VBC_EARG, "rowval[%lld] out of 1:%lld", (long long)(t + 1), (long long)m);
    DeviceGuard guard(device);
    if (!guard.ok) VBC_FAIL(VBC_ECUDA, "cudaSetDevice(%d) failed", device);
    vbc_csc *A = new (std::nothrow) vbc_csc();
    if (!A) VBC_FAIL(VBC_ENOMEM, "host allocation failed");
    A->vt = vt; A->it = it; A->m = m; A->n = n; A->nnz = nnz; A->device = device;
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && sms > 0) A->sm_count = sms;
    const size_t ti = it_size(it), tv = vt_size(vt);
    int rc = VBC_OK;
    do {
        if ((rc = copy_in(&A->d_colptr, colptr, ti * (size_t)(n + 1), cudaMemcpyHostToDevice, A->stream)) != VBC_OK) break;
        if ((rc = copy_in(&A->d_rowval, rowval, ti * (size_t)nnz, cudaMemcpyHostToDevice, A->stream)) != VBC_OK) break;
        if ((rc = copy_in(&A->d_nzval, nzval, tv * (size_t)nnz, cudaMemcpyHostToDevice, A->stream)) != VBC_OK) break;
        if (cudaMalloc(&A->d_x, tv * (size_t)(m > 0 ? m : 1)) != cudaSuccess || cudaMalloc(&A->d_y, tv * (size_t)(n > 0 ? n : 1)) != cudaSuccess) { set_error("cudaMalloc(x,y)"); rc = VBC_ENOMEM; break; }
        if (cudaStreamSynchronize(A->stream) != cudaSuccess) { set_error("csc upload sync failed"); rc = VBC_ECUDA; }
    } while (0);
    if (rc != VBC_OK) { vbc_csc_destroy(A); return rc; }
    *out = A;
    return VBC_OK;
}

int vbc_csc_trspmv(vbc_csc *A, const void *x, int64_t xlen, void *y, int64_t ylen, int on_device)
{
    if (!A) VBC_FAIL(VBC_EARG, "matrix handle is NULL");
    if (A->n != ylen || A->m != xlen) VBC_FAIL(VBC_EDIM, "DimensionMismatch: A' is %lld x %lld, x has %lld, y has %lld", (long long)A->n, (long long)A->m, (long long)xlen, (long long)ylen);
    DeviceGuard guard(A->device);
    if (!guard.ok) VBC_FAIL(VBC_ECUDA, "cudaSetDevice(%d) failed", A->device);
    if (on_device) return launch_csc_trspmv(A, x, y);
    const size_t tv = vt_size(A->vt);
    if (xlen > 0) VBC_CUDA(cudaMemcpyAsync(A->d_x, x, tv * (size_t)xlen, cudaMemcpyHostToDevice, A->stream));
    VBC_TRY(launch_csc_trspmv(A, A->d_x, A->d_y));
    if (ylen > 0) VBC_CUDA(cudaMemcpyAsync(y, A->d_y, tv * (size_t)ylen, cudaMemcpyDeviceToHost, A->stream));
    VBC_CUDA(cudaStreamSynchronize(A->stream));
    return VBC_OK;
}

void vbc_csc_destroy(vbc_csc *A)
{
    if (!A) return;
    DeviceGuard guard(A->device);
    cudaFree(A->d_colptr); cudaFree(A->d_rowval); cudaFree(A->d_nzval); cudaFree(A->d_x); cudaFree(A->d_y);
    delete A;
}

int vbc_set_stream(vbc_mat *A, void *s)
{
    if (!A) VBC_FAIL(VBC_EARG, "matrix handle is NULL");
    A->stream = (cudaStream_t)s;
    return VBC_OK;
}

int vbc_csc_set_stream(vbc_csc *A, void *s)
{
    if (!A) VBC_FAIL(VBC_EARG, "matrix handle is NULL");
    A->stream = (cudaStream_t)s;
    return VBC_OK;
}

int vbc_sync(vbc_mat *A)
{
    if (!A) VBC_FAIL(VBC_EARG, "matrix handle is NULL");
    DeviceGuard guard(A->device);
    VBC_CUDA(cudaStreamSynchronize(A->stream));
    int flag = 0;
    VBC_TRY(trsv_error_flag(A, &flag));
    if (flag) VBC_FAIL(VBC_ECUDA, "a triangular-solve dependency wait timed out");
    return VBC_OK;
}

int vbc_set_option(vbc_mat *A, int option, int64_t value)
{
    if (!A) VBC_FAIL(VBC_EARG, "matrix handle is NULL");
    switch (option) {
    case VBC_OPT_ADJ_GROUP:
    case VBC_OPT_FWD_GROUP:
        if (value != 0 && value != 4 && value != 8 && value != 16 && value != 32) VBC_FAIL(VBC_EARG, "group must be 0 (auto), 4, 8, 16 or 32");
        if (option == VBC_OPT_FWD_GROUP && (value == 16 || value == 4)) VBC_FAIL(VBC_EARG, "forward kernel groups: 0 (auto), 8 or 32");
        (option == VBC_OPT_ADJ_GROUP ? A->opt_adj_group : A->opt_fwd_group) = (int)value;
        return VBC_OK;
    case VBC_OPT_GRID_MULT:
        if (value < 0 || value > 32) VBC_FAIL(VBC_EARG, "grid multiplier must be in 0..32");
        A->opt_grid_mult = (int)value;
        return VBC_OK;
    case VBC_OPT_PARITY_MODE:
        A->opt_parity = value ? 1 : 0;
        return VBC_OK;
    case VBC_OPT_SPMM_SIMT:
        if (value < 0 || value > 3) VBC_FAIL(VBC_EARG, "SpMM kernel must be 0 (auto), 1 (SIMT), 2 (tensor tiles fed by per-lane loads) or 3 (TMA-fed tensor tiles)");
        A->opt_spmm_simt = (int)value;
        return VBC_OK;
    case VBC_OPT_E2E_PIPELINE:
        A->opt_e2e_pipeline = value ? 1 : 0;
        return VBC_OK;
    case VBC_OPT_E2E_GRAPH:
        A->opt_e2e_graph = value ? 1 : 0;
        if (!value) { DeviceGuard guard(A->device); drop_e2e_graph(A); }
        return VBC_OK;
    case VBC_OPT_E2E_UPLOAD_ELEMS:
        VBC_FAIL(VBC_EARG, "VBC_OPT_E2E_UPLOAD_ELEMS is read-only");
    case VBC_OPT_FWD_MODE:
        if (value < 0 || value > 3) VBC_FAIL(VBC_EARG, "forward mode must be 0 (auto), 1 (atomic scatter), 2 (transposed unit index) or 3 (transposed copy)");
        if ((int)value != A->opt_fwd_atomic && A->tindex) { DeviceGuard guard(A->device); cudaStreamSynchronize(A->stream); destroy_tindex(A->tindex); A->tindex = nullptr; } // rebuilt at the next forward multiply
        A->opt_fwd_atomic = (int)value;
        A->opt_fwd_no_copy = 0;
        return VBC_OK;
    }
    VBC_FAIL(VBC_EARG, "unknown option %d", option);
}

int vbc_get_option(const vbc_mat *A, int option, int64_t *value)
{
    if (!A || !value) VBC_FAIL(VBC_EARG, "NULL argument");
    switch (option) {
    case VBC_OPT_ADJ_GROUP: *value = A->opt_adj_group; return VBC_OK;
    case VBC_OPT_FWD_GROUP: *value = A->opt_fwd_group; return VBC_OK;
    case VBC_OPT_GRID_MULT: *value = A->opt_grid_mult; return VBC_OK;
    case VBC_OPT_PARITY_MODE: *value = A->opt_parity; return VBC_OK;
    case VBC_OPT_FWD_MODE: *value = A->opt_fwd_atomic; return VBC_OK;
    case VBC_OPT_SPMM_SIMT: *value = A->opt_spmm_simt; return VBC_OK;
    case VBC_OPT_E2E_PIPELINE: *value = A->opt_e2e_pipeline; return VBC_OK;
    case VBC_OPT_E2E_GRAPH: *value = A->e2e_exec ? 2 : A->opt_e2e_graph; return VBC_OK;
    case VBC_OPT_E2E_UPLOAD_ELEMS: *value = A->last_upload_elems; return VBC_OK;
    }
    VBC_FAIL(VBC_EARG, "unknown option %d", option);
}

int vbc_launch_count(const vbc_mat *A, int64_t *count)
{
    if (!A || !count) VBC_FAIL(VBC_EARG, "NULL argument");
    *count = A->launches;
    return VBC_OK;
}

} // extern "C"
