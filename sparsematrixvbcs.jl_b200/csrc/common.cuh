// common.cuh -- shared declarations of libvbc.so (device VBC matrix, error plumbing).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/vbc.h"

namespace vbc {

// ---- error plumbing -------------------------------------------------------------------------
void set_error(const char *fmt, ...);
const char *get_error();

#define VBC_CUDA(call)                                                                          \
    do {                                                                                        \
        cudaError_t e__ = (call);                                                               \
        if (e__ != cudaSuccess) {                                                               \
            vbc::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return (e__ == cudaErrorMemoryAllocation) ? VBC_ENOMEM : VBC_ECUDA;                 \
        }                                                                                       \
    } while (0)

#define VBC_TRY(expr)                                                                           \
    do {                                                                                        \
        int rc__ = (expr);                                                                      \
        if (rc__ != VBC_OK) return rc__;                                                        \
    } while (0)

#define VBC_FAIL(code, ...)                                                                     \
    do {                                                                                        \
        vbc::set_error(__VA_ARGS__);                                                            \
        return (code);                                                                          \
    } while (0)

static inline size_t vt_size(int vt) { return (vt == VBC_F64 || vt == VBC_INT64) ? 8 : 4; }
static inline bool vt_is_float(int vt) { return vt == VBC_F32 || vt == VBC_F64; }
static inline bool vt_is_valid(int vt) { return vt == VBC_F32 || vt == VBC_F64 || vt == VBC_INT32 || vt == VBC_INT64; }
static inline size_t it_size(int it) { return it == VBC_I64 ? 8 : 4; }

// ---- device layout --------------------------------------------------------------------------
// One entry per stripe boundary (L+1 entries), 0-based.  16 bytes so one 128-bit load fetches it.
struct __align__(16) StripeMeta {
    long long ofs; // element offset of the stripe's first value in `val`
    int pos;       // first descriptor of the stripe in `desc`
    int col;       // first column of the stripe (Φ.spl[l] - 1)
};

enum DescMode {
    DESC_ROWS = 0,   // desc[t] = 0-based x index of stored row t (1D; 2D with non-uniform part heights, expanded)
    DESC_BLOCKS = 1  // desc[Q] = 0-based first x index of block Q; every part has height u0 (the last may be shorter)
};

} // namespace vbc

struct vbc_trsv_plan; // trsv.cu
namespace vbc { struct TIndex; } // fwdt.cu

struct vbc_mat {
    int vt = VBC_F64, it = VBC_I64, ndim = 1, device = 0;
    int64_t m = 0, n = 0, K = 0, L = 0;
    int U = 0, W = 0;
    int64_t nidx = 0, nval = 0;
    // canonical arrays, device copies of the reference struct fields (1-based, Ti typed)
    void *d_pi_spl = nullptr, *d_phi_spl = nullptr, *d_pos = nullptr, *d_idx = nullptr, *d_ofs = nullptr;
    void *d_val = nullptr; // nval values + zero pad
    // compact layout read by the multiply kernels
    vbc::StripeMeta *d_meta = nullptr; // L+1
    int *d_desc = nullptr;             // ndesc
    int *d_brow = nullptr;             // 2D only: first (stripe-relative) expanded row of each block
    int *d_order = nullptr;            // stripes grouped by kernel body class (null: all stripes share one class)
    int nclasses = 1;
    int64_t n_long = 0;                // LONG stripes (dense column groups): the last n_long entries of d_order, multiplied by one CTA each
    int64_t ndesc = 0;
    int desc_mode = vbc::DESC_ROWS;
    int u0 = 1; // DESC_BLOCKS: uniform part height
    int w_uniform = 0; // >0 when every stripe has this width
    int has_unaligned = 0; // some stripe has an odd width or starts on an odd element (Float64: the flat-slab adjoint bodies apply)
    int opt_no_flat = 0;   // experiments: 1 = never use the flat-slab bodies
    // staging vectors for host-pointer multiplies
    void *d_x = nullptr, *d_y = nullptr;
    int64_t x_cap = 0, y_cap = 0;
    // row-major staging panels of the column-major SpMM path (grow-only, reused across calls)
    void *d_px = nullptr, *d_py = nullptr;
    int64_t px_cap = 0, py_cap = 0;
    cudaStream_t stream = nullptr;
    // options
    int opt_adj_group = 0, opt_fwd_group = 0, opt_grid_mult = 0, opt_parity = 0;
    int sm_count = 148;
    int64_t launches = 0;
    // host-vector adjoint multiplies: the stripes are launched in chunks so the D2H copy of a finished y range
    // overlaps the next chunk's kernel (second, non-blocking stream)
    int range_l0 = -1, range_l1 = -1;      // stripe range of the next adjoint launch (-1: all)
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t chunk_ev[8] = {};
    int chunk_l[9] = {};                   // stripe boundaries of the chunks
    int64_t chunk_col[9] = {};             // first column of each chunk
    int nchunks = 0;                       // 0: not prepared, -1: chunking not applicable
    // VBC_OPT_E2E_PIPELINE (default on): x is uploaded in pieces on its own stream; chunk c starts once x[0, chunk_xhi[c]) is there
    int opt_e2e_pipeline = 1;
    cudaStream_t h2d_stream = nullptr;
    cudaEvent_t h2d_ev[8] = {};
    int64_t chunk_xhi[8] = {};             // cumulative: one past the largest x index gathered by the stripes of chunks 0..c
    int xhi_ready = 0;                     // 0: not computed, 1: ready, -1: failed (pipeline off)
    int64_t x_lo = 0;                      // smallest x index any stripe gathers from
    int64_t last_upload_elems = 0;         // x elements the last host-vector multiply copied to the device
    // VBC_OPT_E2E_GRAPH (default on): repeated host-vector adjoint multiplies with the same pinned x / y replay one captured graph
    int opt_e2e_graph = 1;
    cudaGraphExec_t e2e_exec = nullptr;
    cudaStream_t e2e_stream = nullptr;
    cudaEvent_t e2e_ev[3] = {};
    const void *e2e_x = nullptr; void *e2e_y = nullptr;
    double e2e_alpha = 0.0, e2e_beta = 0.0;
    int e2e_pipeline = 0, e2e_seen = 0;    // seen: 1 after the first call with these buffers, -1 if they cannot be captured
    int64_t e2e_upload = 0;
    vbc_trsv_plan *trsv = nullptr; // level schedule of the triangular solve (vbc_trsv_analyse)
    vbc::TIndex *tindex = nullptr; // transposed unit index of the owner-computes forward multiply (built at first use)
    int opt_fwd_atomic = 0;        // VBC_OPT_FWD_MODE: 0 auto, 1 atomic scatter kernel, 2 transposed unit index, 3 transposed copy
    int opt_fwd_no_copy = 0;       // the transposed copy was tried and is not available (memory): do not retry at every multiply
    int opt_spmm_simt = 0;         // 1: Float64 adjoint SpMM on the SIMT (DFMA) kernel instead of the DMMA tiles
};

struct vbc_csc {
    int vt = VBC_F64, it = VBC_I64, device = 0;
    int64_t m = 0, n = 0, nnz = 0;
    void *d_colptr = nullptr, *d_rowval = nullptr, *d_nzval = nullptr;
    void *d_x = nullptr, *d_y = nullptr;
    cudaStream_t stream = nullptr;
    int sm_count = 148;
    int64_t launches = 0;
};

namespace vbc {

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; prev = -1; }
        if (ok && prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// pack.cu
int pack_from_device_csc(vbc_mat *A, const void *d_colptr, const void *d_rowval, const void *d_nzval);
int finalize_layout(vbc_mat *A, const void *h_pi_spl /* host copy, may be null for 1D */);
int build_class_order(vbc_mat *A); // stripes grouped by kernel-body class; sets w_uniform / has_unaligned
int memory_cost_device(const vbc_mat *A, int64_t *h_cost, int64_t *row_term);
// spmv.cu
int launch_spmv(vbc_mat *A, int trans, double alpha, const void *d_x, double beta, void *d_y);
struct HaloLaunch { // one step of the row-partitioned iteration (peer.cu -> spmv.cu, k_spmv_adj_halo)
    int n;                                     // destinations; dst[0] = own next-x buffer, dst[i] = rank (me + i) % n
    void *dst[VBC_MAX_PEERS];
    const unsigned char *d_mask; int chunk_shift;
    int i0, i1;                                // interior stripes [i0, i1)
    int me, nranks, do_wait, do_signal;
    unsigned nbr_mask;
    unsigned long long *flags[VBC_MAX_PEERS];
    unsigned long long *ctl;                   // [1] finished boundary warps, [2] epoch, [3..5] wait statistics
    int *timed_out;
};
int launch_spmv_adj_halo(vbc_mat *A, double alpha, const void *d_x, const HaloLaunch *hl);
// inttypes.cu (Int32 / Int64 element types, wrapping arithmetic)
int launch_spmv_int(vbc_mat *A, int trans, double alpha, const void *d_x, double beta, void *d_y);
// mixed.cu
int launch_spmv_mixed(vbc_mat *A, int trans, double alpha, const void *d_x, double beta, void *d_y);
// fwdt.cu
int ensure_tindex(vbc_mat *A);
int launch_fwdt(vbc_mat *A, double alpha, const void *x, double beta, void *y);
void destroy_tindex(TIndex *t);
int64_t tindex_bytes(const vbc_mat *A);
int tindex_kind(const vbc_mat *A); // 0 none, 1 transposed unit index, 2 transposed copy
vbc_mat *tindex_copy(const vbc_mat *A); // the transposed copy (forward multiply = adjoint multiply of it), or null
// trsv.cu
void destroy_trsv_plan(vbc_trsv_plan *p);
int trsv_error_flag(const vbc_mat *A, int *flag);
// csc.cu
int launch_csc_trspmv(vbc_csc *A, const void *d_x, void *d_y);

} // namespace vbc
