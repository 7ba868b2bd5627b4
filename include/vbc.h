/* include/vbc.h -- C ABI of libvbc.so: B200 (sm_100a) device implementation of the blocked
 * sparse multiply path of SparseMatrixVBCs.jl.
 *
 * This header is the drop-in boundary.  Each entry point names the reference interface it
 * replaces (paths relative to the reference checkout).  All arguments are plain pointers and
 * integers; no C++/torch types.  Index arrays crossing the boundary are 1-BASED and Ti-typed,
 * exactly as the reference's Julia structs store them, so "bit-exact packed arrays" is a
 * memcmp after vbc_download.
 *
 * Conventions
 *   - every function returns a vbc_status (0 = VBC_OK); vbc_last_error() gives the message of
 *     the calling thread's last failure.  The library never aborts or exits.
 *   - `vt` is a vbc_dtype (Tv, and also the element type of x and y), `it` a vbc_itype (Ti).
 *   - `on_device` = 0: x / y are host pointers, the call copies in/out and is synchronous on
 *     return.  `on_device` = 1: x / y are device pointers on the matrix' device, the call only
 *     enqueues work on the matrix' stream (vbc_set_stream, default: the legacy default
 *     stream) -- use vbc_sync() or your own events to wait.
 *   - a handle may be used from any thread, one call at a time; different handles are
 *     independent.
 */
#ifndef VBC_H
#define VBC_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vbc_mat vbc_mat; /* device-resident 1D-VBC or 2D-VBC matrix ("CuVBC")            */
typedef struct vbc_csc vbc_csc; /* device-resident CSC matrix (comparator for TrSpMV!)           */

enum vbc_dtype { VBC_F32 = 0, VBC_F64 = 1 };
enum vbc_itype { VBC_I32 = 0, VBC_I64 = 1 };

enum vbc_status {
    VBC_OK = 0,
    VBC_EDIM = 1,   /* DimensionMismatch: multiply_1DVBC.jl:44-45,:139-140; multiply_VBC.jl:51-52,:152-153; TrSpMV.jl:3-4 */
    VBC_EARG = 2,   /* ArgumentError: SparseMatrixVBCs.jl:45-50, :72-79 (m,n >= 0; U,W > 0) + malformed inputs   */
    VBC_ELIMIT = 3, /* AssertionError `w <= W` constructors_1DVBC.jl:46 / constructors_VBC.jl:65, `u <= U` :58-60;
                       also: sizes beyond the device layout's 32-bit fields                                        */
    VBC_ECUDA = 4,  /* a CUDA runtime call failed (message has the CUDA error string)                            */
    VBC_ENCCL = 5,  /* reserved for the multi-GPU layer                                                          */
    VBC_ENOMEM = 6  /* host or device allocation failed                                                          */
};

/* options for vbc_set_option */
enum vbc_option {
    VBC_OPT_ADJ_GROUP = 1,  /* lanes cooperating on one stripe in the adjoint kernel: 0 = auto, else 4|8|16|32 */
    VBC_OPT_FWD_GROUP = 2,  /* same for the forward (scatter) kernel                                         */
    VBC_OPT_GRID_MULT = 3,  /* CTAs per SM for the persistent grid-stride launch: 0 = auto                   */
    VBC_OPT_PARITY_MODE = 4, /* 1: multiply straight from the canonical Ti arrays (pos/idx/ofs/spl) with the
                                  generic kernel instead of the compact device layout                         */
    VBC_OPT_FWD_MODE = 5,    /* forward multiply: 0 = auto (owner-computes through a transposed unit index, built at first
                                  use, for uniform 2D blocks; atomic scatter kernel otherwise), 1 = always the atomic
                                  scatter kernel, 2 = the transposed index whenever the layout allows it              */
    VBC_OPT_SPMM_SIMT = 6,   /* Float64 adjoint SpMM: 0 = auto (FP64 tensor DMMA m8n8k4 tiles), 1 = the SIMT (DFMA) kernel,
                              * 2 = DMMA with scalar X loads (what 0 selects), 3 = DMMA with 256-bit X-row loads, 4 = DMMA tiles fed
                              * through shared memory by bulk copies, 5 = by cp.async (3-5: experiments, profiles/r01_spmm_ncu.md;
                              * they fall back to 2 when the panels are not suitably aligned) */
    VBC_OPT_E2E_PIPELINE = 7, /* host-vector adjoint multiplies: 1 = upload x in pieces, each chunk of stripes starting as soon as the
                              * x rows it gathers from have arrived (pays when the matrix is banded); 0 = upload x first (default).
                              * Only x[lo, hi) with lo / hi the smallest / largest index any stripe gathers from is uploaded at all.
                              * Experimental in round 1: not yet run on a GPU. */
    VBC_OPT_E2E_UPLOAD_ELEMS = 8 /* read-only: x elements the last host-vector multiply copied to the device */
};

const char *vbc_last_error(void);
int vbc_version(void);
int vbc_device_count(int *count);

/* ---- constructors ------------------------------------------------------------------------
 * vbc_pack_csc: host CSC + host partition(s) -> device pack kernels.
 *   1D (pi_spl == NULL): replaces `SparseMatrix1DVBC{W}(A::SparseMatrixCSC, Φ::SplitPartition)`
 *                        constructors_1DVBC.jl:9-92 (and the StrictChunker body :94-143, whose
 *                        output is identical on a strict partition).  U is ignored.
 *   2D:                  replaces `SparseMatrixVBC{U,W}(A, Π, Φ)` constructors_VBC.jl:15-133.
 *   colptr[n+1], rowval[nnz], nzval[nnz]: SparseMatrixCSC fields (1-based, rows ascending in a
 *   column, no duplicates).  pi_spl[K+1], phi_spl[L+1]: SplitPartition.spl (1-based, spl[0]=1,
 *   spl[K]=m+1 resp. spl[L]=n+1).
 * vbc_pack_csc_dev: same, but every array argument is already a device pointer on `device`. */
int vbc_pack_csc(vbc_mat **out, int vt, int it, int64_t m, int64_t n, int U, int W,
                 const void *colptr, const void *rowval, const void *nzval,
                 const void *pi_spl, int64_t K, const void *phi_spl, int64_t L, int device);
int vbc_pack_csc_dev(vbc_mat **out, int vt, int it, int64_t m, int64_t n, int U, int W,
                     const void *colptr, const void *rowval, const void *nzval,
                     const void *pi_spl, int64_t K, const void *phi_spl, int64_t L, int device);

/* vbc_upload: adopt arrays already packed by the reference on the host -- the fields of a
 * `SparseMatrix1DVBC` (SparseMatrixVBCs.jl:36-43; pi_spl == NULL) or `SparseMatrixVBC`
 * (:62-70).  val needs only its first ofs[L]-1 entries (the SIMD tail pad is not read). */
int vbc_upload(vbc_mat **out, int vt, int it, int64_t m, int64_t n, int U, int W,
               const void *pi_spl, int64_t K, const void *phi_spl, int64_t L,
               const void *pos, const void *idx, const void *ofs, const void *val, int device);

void vbc_destroy(vbc_mat *A);

/* ---- introspection (`Base.size`, SparseMatrixVBCs.jl:55/:84, and the struct fields) -------- */
int vbc_shape(const vbc_mat *A, int64_t *m, int64_t *n, int64_t *K, int64_t *L, int *U, int *W,
              int *ndim /* 1 or 2 */, int *vt, int *it);
/* nidx = pos[L]-1 (stored rows / blocks), nval = ofs[L]-1 (stored values) */
int vbc_sizes(const vbc_mat *A, int64_t *nidx, int64_t *nval);
/* copies pos[L+1], idx[nidx], ofs[L+1], val[nval] (1-based, Ti / Tv typed) to host; any may be NULL */
int vbc_download(const vbc_mat *A, void *pos, void *idx, void *ofs, void *val);
/* bytes: [0] the reference's format accounting `sizeof(Φ)+sizeof(Π)+sizeof(pos)+sizeof(idx)+sizeof(ofs)+sizeof(val)`
 *            (bin/test_table.jl:78, :120);
 *        [1] bytes of the device layout that one adjoint multiply reads (matrix part only);
 *        [2] same for one forward multiply. */
int vbc_format_bytes(const vbc_mat *A, int64_t bytes[3]);
/* per-stripe cost under the reference's memory model (costs.jl:10 1D, :140 2D), computed on the
 * device from the packed arrays: cost[L] (host pointer); *row_term = K*sizeof(Ti) for 2D, else 0. */
int vbc_memory_cost(const vbc_mat *A, int64_t *cost, int64_t *row_term);

/* ---- multiply ------------------------------------------------------------------------------
 * y <- alpha * op(A) * x + beta * y.
 *   trans = 0: op(A) = A   replaces `mul!(y, A, x, α, β)`  multiply_1DVBC.jl:9 / multiply_VBC.jl:3
 *   trans = 1: op(A) = A'  replaces `mul!(y, A', x, α, β)` multiply_1DVBC.jl:85 / multiply_VBC.jl:89
 * (BLAS semantics; coincides with the reference wherever its tests pin it, α=1, β=0.  The
 * reference ignores α and, in the adjoint, β -- documented in DESIGN.md, not reproduced.)
 * xlen / ylen are checked against the matrix shape -> VBC_EDIM. */
int vbc_spmv(vbc_mat *A, int trans, double alpha, const void *x, int64_t xlen, double beta,
             void *y, int64_t ylen, int on_device);

/* Same multiply with vectors of a wider type than the stored values: vec_vt is the type of x AND y.
 * The reference converts values and x to eltype(y) before multiplying (multiply_1DVBC.jl:23/27/34, :102;
 * multiply_VBC.jl:40-45, :131), so a Float32 matrix with Float64 vectors accumulates in Float64.
 * vec_vt == the matrix' value type: identical to vbc_spmv.  Float32 matrix + VBC_F64 vectors: dedicated
 * kernels (csrc/mixed.cu).  Float64 matrix + VBC_F32 vectors: VBC_EARG (narrowing is not offered). */
int vbc_spmv_mixed(vbc_mat *A, int trans, double alpha, const void *x, int64_t xlen, double beta,
                   void *y, int64_t ylen, int vec_vt, int on_device);

/* k right-hand sides: Y <- alpha * op(A) * X + beta * Y, X: cols(op(A)) x k, Y: rows(op(A)) x k.
 * Stands in for `*(A, B::DenseMatrix)` (multiply_1DVBC.jl:184-185, multiply_VBC.jl:196-197), which the
 * reference declares but cannot execute (no matrix `mul!` method) -- new functionality,
 * oracle = k independent vbc_spmv.  layout 0: row-major panels, element (i, c) at X[i*ldx + c]
 * (ldx >= k; the fast path); layout 1: column-major as Julia stores a Matrix, X[c*ldx + i]
 * (ldx >= rows; transposed on the device into row-major staging buffers and back). */
int vbc_spmm(vbc_mat *A, int trans, int64_t k, double alpha, const void *X, int64_t ldx, double beta,
             void *Y, int64_t ldy, int layout, int on_device);

/* Lower-triangular solve  tril(A') x = b  (A square).  EXTENSION: the reference has no triangular solve
 * ("TrSpMV" there is the transposed multiply); BASELINE.json's north_star (d) asks for a
 * blocked, level-scheduled one.  Row block l of A' is stripe l; entries above the diagonal of A' are
 * ignored (BLAS trsv 'L'), the diagonal must be stored and nonzero.  vbc_trsv_analyse builds the level
 * schedule (also done lazily by the first solve) and reports the number of levels; the solve is one
 * cooperative persistent kernel that walks the row blocks in level order and waits on per-row-block
 * flags.  Stripes may be at most 8 columns wide (VBC_ELIMIT).  b and x have n entries; x != b. */
int vbc_trsv_analyse(vbc_mat *A, int *nlevels);
int vbc_trsv_levels(const vbc_mat *A, int *nlevels);
int vbc_trsv_lower(vbc_mat *A, const void *b, void *x, int64_t len, int on_device);

/* CSC comparator: `TrSpMV!(y, A::SparseMatrixCSC, x)` TrSpMV.jl:1-20, y = A' x. */
int vbc_csc_upload(vbc_csc **out, int vt, int it, int64_t m, int64_t n, const void *colptr,
                   const void *rowval, const void *nzval, int device);
int vbc_csc_trspmv(vbc_csc *A, const void *x, int64_t xlen, void *y, int64_t ylen, int on_device);
void vbc_csc_destroy(vbc_csc *A);

/* ---- host-side helper of the partitioners (no device work) ---------------------------------
 * Minimum-total-cost contiguous partition with stripes at most W wide: cost[b*W + (w-1)] is the cost of
 * the stripe of columns [b-w, b) (0-based, b = 1..n).  Writes the 1-based SplitPartition.spl (at most n+1
 * entries) and its length L.  The recurrence behind ChainPartitioners' DynamicTotalChunker
 * (constructors_1DVBC.jl:1-2, test/runtests.jl:22-23); ChainPartitioners is not vendored, tie-breaking
 * parity is unpinned. */
int vbc_dp_chunk(int64_t n, int W, const double *cost, int64_t *spl, int64_t *L_out);
/* Greedy overlap chunker, a stand-in for ChainPartitioners' OverlapChunker(rho, w_max) (test/runtests.jl:21) under an
 * ASSUMED definition: the next column joins the current stripe while it is narrower than w_max and the Jaccard
 * similarity of its row pattern with the union of the stripe's patterns is >= rho.  colptr / rowval: 0-based int64
 * CSC structure, rows ascending.  Writes the 1-based spl (at most n+1 entries) and its length. */
int vbc_overlap_chunk(int64_t n, const int64_t *colptr, const int64_t *rowval, double rho, int w_max, int64_t *spl, int64_t *L_out);

/* ---- execution control ---------------------------------------------------------------------*/
int vbc_set_stream(vbc_mat *A, void *cuda_stream); /* a cudaStream_t; NULL = legacy default stream */
int vbc_csc_set_stream(vbc_csc *A, void *cuda_stream);
int vbc_sync(vbc_mat *A);
int vbc_set_option(vbc_mat *A, int option, int64_t value);
int vbc_get_option(const vbc_mat *A, int option, int64_t *value);
/* number of kernel launches this handle has made since creation (bench.py's gpu_launches) */
int vbc_launch_count(const vbc_mat *A, int64_t *count);

/* ---- multi-GPU: row-block partition with x replicated through peer memory ---------------------
 * north_star (e): stripes (the row blocks of A') are split across the GPUs of one box; every rank
 * needs the whole x for its gathers, so each iteration x_{t+1} <- alpha * A' x_t ends with an
 * all-gather of the y slices.  Here that all-gather is FUSED into the multiply: the adjoint kernel
 * stores each finished y segment straight into the next-x buffer of every rank (its own and, over
 * NVLink peer mappings, the others'), and a one-CTA flag kernel is the only cross-rank step.
 * One process per GPU; the handles are exchanged by the host (torch.distributed / MPI / files).
 * No reference counterpart: the reference is single-process. */
typedef struct vbc_peer vbc_peer;
#define VBC_IPC_HANDLE_BYTES 64
#define VBC_PEER_HANDLES 3 /* x buffer 0, x buffer 1, flag block */
#define VBC_MAX_PEERS 8

/* Allocates this rank's two x buffers (xlen elements of vt each, zero-filled) and flag block on
 * `device`, and writes VBC_PEER_HANDLES CUDA IPC handles to handles_out (3 * 64 bytes). */
int vbc_peer_create(vbc_peer **out, int vt, int64_t xlen, int rank, int nranks, int device, void *handles_out);
/* all_handles: nranks * 3 * 64 bytes, rank-major, as gathered from every rank.  Maps the peers. */
int vbc_peer_connect(vbc_peer *P, const void *all_handles);
/* Same-process variant (tests, one process driving several "ranks"): raw device pointers,
 * ptrs[r * 3 + k] = buffer k of rank r. */
int vbc_peer_connect_local(vbc_peer *P, void *const *ptrs);
/* raw device pointers of this rank's buffers (k = 0, 1: x buffers; 2: flags) and the current x index */
int vbc_peer_buffer(vbc_peer *P, int k, void **ptr);
int vbc_peer_current(const vbc_peer *P, int *cur);
/* One iteration: y = alpha * A' * x_cur on this rank's stripes, stored at element offset y_offset of
 * x_{1-cur} on EVERY rank; then signal + wait for all ranks (flags), then cur flips.  A->n columns
 * are written; A->m must equal xlen.  Enqueued on A's stream.  `barrier`: 3 = signal and wait
 * (normal), 1 = signal only, 2 = wait only, 0 = neither (same-process tests must not wait inside
 * one stream for a signal that a later launch of the same stream produces). */
int vbc_peer_spmv_step(vbc_peer *P, vbc_mat *A, double alpha, int64_t y_offset, int barrier);
/* Sparsity-aware replication (optional).  mask[c >> chunk_shift] (c = column inside this rank's slab,
 * nchunks bytes) has bit i set when the i-th destination reads that chunk of columns; destination 0 is
 * this rank itself, destination i is rank (rank + i) % nranks.  With a mask the fused epilogue sends a
 * y segment only to the ranks whose stripes gather from it (for a banded operator: the neighbours'
 * halos) -- x stays complete on every rank exactly where that rank reads it.  mask == NULL restores
 * full replication. */
int vbc_peer_set_mask(vbc_peer *P, const void *mask, int64_t nchunks, int chunk_shift);
/* Restrict the flag exchange to the ranks in `mask` (bit r = rank r).  With sparsity-aware replication a rank
 * only has to synchronise with the ranks it sends y segments to (they must have finished reading the buffer it
 * is about to overwrite) and the ranks it receives from (their segments must have landed) -- for a banded
 * operator its two neighbours instead of everyone.  The relation must be symmetric across ranks. */
int vbc_peer_set_neighbors(vbc_peer *P, unsigned mask);
/* Overlapping the flag exchange with the multiply (optional).  enable = 2: barrier = 3 steps become four
 * launches -- [stripes i0..i1) | wait for the peers' previous step | remaining stripes | signal -- so the
 * wait (and the drift between ranks) hides behind the first launch; the default (enable = 0) is
 * [all stripes | signal + wait].  enable = 1 (REMOVED, now the same as 0; it measured slower and quadrupled the
 * kernel's code size): vbc_peer_spmv_step(..., barrier = 3) launched ONE
 * kernel per iteration: it first runs the stripes [i0, i1) -- which must gather only from this rank's
 * own slice and (under the mask) feed only this rank -- then waits for the peers' flags of the
 * previous iteration, runs the remaining stripes, and the last CTA to finish publishes this rank's
 * flag.  The wait is hidden behind the work of [i0, i1), and ranks may drift by that much.  With full
 * replication pass i0 == i1 (nothing can run before the wait).
 * enable = 3 (needs a mask; experimental, not yet run on a GPU): the step is the PLAIN adjoint kernel writing the
 * rank's slice into its own next-x buffer, then ONE kernel in which every CTA copies its share of the column chunks
 * some other rank reads to those ranks and the last CTA to finish runs the signal + wait. */
int vbc_peer_set_fused_sync(vbc_peer *P, int enable, int64_t i0, int64_t i1);
/* the flag kernel alone (barrier = 1 | 2 | 3 as above); does not flip cur */
int vbc_peer_barrier(vbc_peer *P, void *cuda_stream, int barrier);
/* 0 if no flag wait has timed out since creation (checked after a stream sync by the caller) */
int vbc_peer_status(vbc_peer *P, int *timed_out);
void vbc_peer_destroy(vbc_peer *P);

#ifdef __cplusplus
}
#endif
#endif /* VBC_H */
