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

/* Element types.  The integer types (the reference's methods are generic over Tv, and test/runtests.jl:15-16 runs Bool and Int32
 * matrices) are packed, uploaded, downloaded and multiplied (vbc_spmv, vbc_csc_trspmv) with Julia's wrapping arithmetic, exact in
 * any summation order; alpha and beta must then be integers of the type (else VBC_EARG, the InexactError of `convert(eltype(y), α)`).
 * Bool is widened to Int32 by the host layers.  SpMM, the triangular solve, the mixed-type multiply and the multi-GPU entry points
 * are floating-point only (VBC_EARG). */
enum vbc_dtype { VBC_F32 = 0, VBC_F64 = 1, VBC_INT32 = 2, VBC_INT64 = 3 };
enum vbc_itype { VBC_I32 = 0, VBC_I64 = 1 };

enum vbc_status {
    VBC_OK = 0,
    VBC_EDIM = 1,   /* DimensionMismatch: multiply_1DVBC.jl:44-45,:139-140; multiply_VBC.jl:51-52,:152-153; TrSpMV.jl:3-4 */
    VBC_EARG = 2,   /* ArgumentError: SparseMatrixVBCs.jl:45-50, :72-79 (m,n >= 0; U,W > 0) + malformed inputs   */
    VBC_ELIMIT = 3, /* AssertionError `w <= W` constructors_1DVBC.jl:46 / constructors_VBC.jl:65, `u <= U` :58-60;
                       also: sizes beyond the device layout's 32-bit fields                                        */
    VBC_ECUDA = 4,  /* a CUDA runtime call failed (message has the CUDA error string)                            */
    VBC_ENCCL = 5,  /* an NCCL call failed, or libnccl could not be loaded (vbc_dist_* with the NCCL exchange)     */
    VBC_ENOMEM = 6  /* host or device allocation failed                                                          */
};

/* options for vbc_set_option */
enum vbc_option {
    VBC_OPT_ADJ_GROUP = 1,  /* lanes cooperating on one stripe in the adjoint kernel: 0 = auto, else 4|8|16|32 */
    VBC_OPT_FWD_GROUP = 2,  /* same for the forward (scatter) kernel                                         */
    VBC_OPT_GRID_MULT = 3,  /* CTAs per SM for the persistent grid-stride launch: 0 = auto                   */
    VBC_OPT_PARITY_MODE = 4, /* 1: multiply straight from the canonical Ti arrays (pos/idx/ofs/spl) with the
                                  generic kernel instead of the compact device layout                         */
    VBC_OPT_FWD_MODE = 5,    /* forward multiply y = A x (the reference's serial scatter, multiply_1DVBC.jl:9-83 / multiply_VBC.jl:3-87), made
                              * owner-computes at the first forward multiply:
                              *   0 = auto: uniform 2D blocks get a TRANSPOSED COPY of the values when device memory is plentiful (the
                              *       forward multiply is then the adjoint kernel on that copy), else the transposed unit index; 1D and
                              *       variable blocks use the atomic scatter kernel;
                              *   1 = always the atomic scatter kernel;  2 = the transposed unit index (16 B per unit, no second copy of
                              *       the values) whenever the layout allows it;  3 = the transposed copy, regardless of free memory    */
    VBC_OPT_SPMM_SIMT = 6,   /* Float64 adjoint SpMM: 0 = auto: FP64 tensor (DMMA m8n8k4) tiles, fed by TMA row gathers
                              * (cp.async.bulk.tensor tile::gather4) when every stripe is 8 wide, the matrix is in rows mode and
                              * the panels are 16-byte aligned with even k and leading dimensions, else by per-lane loads;
                              * 1 = the SIMT (DFMA) kernel; 2 = always per-lane loads; 3 = same as 0                             */
    VBC_OPT_E2E_PIPELINE = 7, /* host-vector adjoint multiplies: 1 (default) = x is uploaded in pieces on its own stream and each
                              * chunk of stripes starts as soon as the x rows it gathers from have arrived, while the y ranges of
                              * finished chunks are already on their way back (pays when the matrix is banded; any matrix stays
                              * correct: a chunk that reaches far simply waits for more of x); 0 = upload all of x first.  Either way
                              * only x[lo, hi), lo / hi the smallest / largest index any stripe gathers from, is uploaded.          */
    VBC_OPT_E2E_UPLOAD_ELEMS = 8, /* read-only: x elements the last host-vector multiply copied to the device */
    VBC_OPT_E2E_GRAPH = 9     /* host-vector adjoint multiplies: 1 (default) = when the same PINNED x and y (and alpha, beta) come back --
                              * an iterative caller's `mul!(y, A', x)` -- the second call captures the copies, chunk kernels and their
                              * dependencies into one CUDA graph on an internal stream and later calls replay it (one launch instead of
                              * ~50 runtime calls); pageable buffers and changing pointers take the eager path.  get: 2 once a graph exists. */
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
 * schedule (also done lazily by the first solve), extracts the diagonal blocks and reports the number of levels;
 * the solve is one cooperative persistent kernel that walks the row blocks in level order -- x itself carries the
 * dependencies (it starts as a sentinel, a consumer polls the entries it needs), so there are no flags or epochs
 * and a solve may be captured in a CUDA graph and replayed.  Stripes may be at most 32 columns wide (VBC_ELIMIT).
 * b and x have n entries; x != b (VBC_EARG).  A zero diagonal entry is reported by the analysis (VBC_EARG). */
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

/* ---- benchmark support: synthetic matrices generated on the device -----------------------------
 * The block-banded generator of the benchmark configs (modelled on the reference's own generators for its
 * time-model experiments, costs.jl:63-85 / :200-222): global stripe l holds a dense u x w block at every row
 * part l*K/L + offsets[t] inside [0, K); values = hash of the entry's global (row, col) and the seed, in [0, 1),
 * so any slab [l0, l1) can be produced by the rank that owns it and any entry re-derived on the host.
 * Writes a SparseMatrixCSC slab ((l1-l0)*w columns, K*u rows) into freshly allocated DEVICE arrays (1-based
 * Ti colptr / rowval, Tv nzval) for vbc_pack_csc_dev; release them with vbc_gen_free.  offsets ascend strictly. */
int vbc_gen_banded_csc(int vt, int it, int64_t K, int64_t L, int u, int w, const int64_t *offsets, int noffsets,
                       int64_t l0, int64_t l1, uint64_t seed, double diag_boost,
                       void **colptr, void **rowval, void **nzval, int64_t *nnz, int device);
int vbc_gen_free(void *colptr, void *rowval, void *nzval, int device);
/* need[c] = 1 when an adjoint multiply of A gathers from x[c << chunk_shift, (c+1) << chunk_shift) (host array of
 * nchunks >= ceil(m / 2^chunk_shift) bytes): the input of the sparsity-aware exchange plan (vbc_peer_set_mask). */
int vbc_read_chunks(vbc_mat *A, int chunk_shift, unsigned char *need, int64_t nchunks);

/* ---- execution control ---------------------------------------------------------------------*/
int vbc_set_stream(vbc_mat *A, void *cuda_stream); /* a cudaStream_t; NULL = legacy default stream */
int vbc_csc_set_stream(vbc_csc *A, void *cuda_stream);
int vbc_sync(vbc_mat *A);
int vbc_set_option(vbc_mat *A, int option, int64_t value);
int vbc_get_option(const vbc_mat *A, int option, int64_t *value);
/* number of kernel launches this handle has made since creation (bench.py's gpu_launches) */
int vbc_launch_count(const vbc_mat *A, int64_t *count);

/* ---- multi-GPU: row-block partition with the x exchange fused into the multiply ---------------
 * north_star (e): stripes (the row blocks of A') are split across the GPUs of one box; every rank
 * needs the x entries its stripes gather from, so each iteration x_{t+1} <- alpha * A' x_t ends with an
 * exchange of the y slices.  Here that exchange is FUSED into the multiply: one kernel per iteration
 * runs the rank's interior stripes like the single-GPU kernel, and its boundary stripes store their
 * results straight into the next-x buffers of the ranks that read them (own HBM + NVLink peer
 * mappings) and publish a per-step flag -- no collective call, no separate exchange launch.
 * Two front ends share it: vbc_peer_* (one process per GPU; the IPC handles are exchanged by the
 * host: torch.distributed / MPI / files) and vbc_dist_* (one process driving all GPUs, below).
 * No reference counterpart: the reference is single-process; the loop it scales is
 * the @threads stripe loop of multiply_1DVBC.jl:169-177 / multiply_VBC.jl:182-189. */
typedef struct vbc_peer vbc_peer;
#define VBC_IPC_HANDLE_BYTES 64
#define VBC_PEER_HANDLES 3 /* x buffer 0, x buffer 1, flag block */
#define VBC_MAX_PEERS 8

/* Allocates this rank's two x buffers (xlen elements of vt each, zero-filled) and flag block on
 * `device`, and writes VBC_PEER_HANDLES CUDA IPC handles to handles_out (3 * 64 bytes; may be NULL
 * when the ranks share a process). */
int vbc_peer_create(vbc_peer **out, int vt, int64_t xlen, int rank, int nranks, int device, void *handles_out);
/* all_handles: nranks * 3 * 64 bytes, rank-major, as gathered from every rank.  Maps the peers. */
int vbc_peer_connect(vbc_peer *P, const void *all_handles);
/* Same-process variant (vbc_dist_*, tests): raw device pointers, ptrs[r * 3 + k] = buffer k of rank r
 * (peer access between the devices must already be enabled). */
int vbc_peer_connect_local(vbc_peer *P, void *const *ptrs);
/* raw device pointers of this rank's buffers (k = 0, 1: x buffers; 2: flags) and the current x index */
int vbc_peer_buffer(vbc_peer *P, int k, void **ptr);
int vbc_peer_current(const vbc_peer *P, int *cur);
/* One iteration, one kernel launch on A's stream: y = alpha * A' * x_cur on this rank's stripes, stored at
 * element offset y_offset of x_{1-cur} on this rank and on every rank that reads it (all ranks without a
 * mask); then cur flips.  A->n columns are written; A->m must equal xlen.  `barrier` bit 2 (wait): the
 * boundary stripes start only when every neighbour has published its previous step (its halo has landed
 * here, and it no longer reads the buffer about to be overwritten); bit 1 (signal): the step is published
 * to the neighbours when the last boundary stripe has been stored.  3 = both (normal); 0 = neither
 * (same-process tests, or when vbc_peer_barrier is used between the steps instead).
 * After the last iteration, vbc_peer_barrier(P, stream, 2) waits until the neighbours' final halos are in. */
int vbc_peer_spmv_step(vbc_peer *P, vbc_mat *A, double alpha, int64_t y_offset, int barrier);
/* Sparsity-aware replication (optional).  mask[c >> chunk_shift] (c = column inside this rank's slab,
 * nchunks bytes) has bit i set when the i-th destination reads that chunk of columns; destination 0 is
 * this rank itself (always written), destination i is rank (rank + i) % nranks.  With a mask a y segment is
 * sent only to the ranks whose stripes gather from it (for a banded operator: the neighbours' halos) -- x
 * stays complete on every rank exactly where that rank reads it.  mask == NULL restores full replication. */
int vbc_peer_set_mask(vbc_peer *P, const void *mask, int64_t nchunks, int chunk_shift);
/* Restrict the flag exchange to the ranks in `mask` (bit r = rank r).  With sparsity-aware replication a rank
 * only has to synchronise with the ranks it sends y segments to (they must have finished reading the buffer it
 * is about to overwrite) and the ranks it receives from (their segments must have landed) -- for a banded
 * operator its two neighbours instead of everyone.  The relation must be symmetric across ranks. */
int vbc_peer_set_neighbors(vbc_peer *P, unsigned mask);
/* Interior stripes [i0, i1) (0-based, of the matrix passed to vbc_peer_spmv_step): stripes that gather only
 * from this rank's own slice of x and (under the mask) feed only this rank.  They run before and independently
 * of the flag exchange; all other stripes are boundary stripes.  Default: none (i0 == i1 == 0). */
int vbc_peer_set_interior(vbc_peer *P, int64_t i0, int64_t i1);
int vbc_peer_get_interior(const vbc_peer *P, int64_t *i0, int64_t *i1);
/* Derives the interior range on the device from the packed matrix and the mask set so far (the longest run of
 * stripes whose gathers stay inside x[y_offset, y_offset + n) and whose columns only this rank reads), installs
 * it and reports it (i0_out / i1_out may be NULL).  Without a mask, or with one rank, see vbc_peer_set_mask. */
int vbc_peer_auto_interior(vbc_peer *P, vbc_mat *A, int64_t y_offset, int64_t *i0_out, int64_t *i1_out);
/* the flag exchange alone, as one tiny kernel (barrier = 1 signal | 2 wait | 3 both); does not flip cur */
int vbc_peer_barrier(vbc_peer *P, void *cuda_stream, int barrier);
/* 0 if no flag wait has timed out since creation (checked after a stream sync by the caller) */
int vbc_peer_status(vbc_peer *P, int *timed_out);
/* stats[0] = steps published, [1] = ns spent spinning in in-kernel flag waits (summed over waiting lanes),
 * [2] = waits that had to spin at all, [3] = longest single wait in ns; reset != 0 clears [1..3].  Synchronous. */
int vbc_peer_wait_stats(vbc_peer *P, uint64_t stats[4], int reset);
void vbc_peer_destroy(vbc_peer *P);

/* ---- multi-GPU, one process driving all devices (the form a Julia host uses) -------------------
 * The iterated adjoint multiply x_{t+1} <- alpha * A' x_t of a SQUARE operator over `ngpus` devices of this box:
 * the stripes (row blocks of A') are split into contiguous ranges balanced by the reference's memory cost model
 * (costs.jl:10 / :140), every range is packed on its device from the host CSC + partitions (same arguments as
 * vbc_pack_csc, m == n), x lives in peer-mapped buffers and an iteration is one launch per device
 * (VBC_EXCH_FUSED: the exchange of the new x is fused into the multiply, see vbc_peer_*), or the plain multiply
 * followed by ncclAllGather (VBC_EXCH_NCCL: the unfused comparator; libnccl.so.2 is loaded at first use and its
 * failures are reported as VBC_ENCCL).  devices == NULL: devices 0..ngpus-1.  A device listed more than once is accepted
 * for testing on a small box: kernels of ranks that share a device must not wait for one another (nothing guarantees
 * they run at the same time), so such a handle launches every iteration without in-kernel flags and synchronises the
 * ranks on the host between iterations. */
typedef struct vbc_dist vbc_dist;
enum vbc_exchange { VBC_EXCH_FUSED = 0, VBC_EXCH_NCCL = 1 };
int vbc_dist_create(vbc_dist **out, int ngpus, const int *devices, int vt, int it, int64_t n, int U, int W,
                    const void *colptr, const void *rowval, const void *nzval,
                    const void *pi_spl, int64_t K, const void *phi_spl, int64_t L, int exchange);
/* the partition chosen: slice_len = padded slice length S; stripe_bounds[ngpus+1]; cost_per_gpu[ngpus] (bytes under
 * the memory model); interior[2*ngpus] = interior stripe range of every rank (any pointer may be NULL) */
int vbc_dist_info(const vbc_dist *D, int *ngpus, int64_t *slice_len, int64_t *stripe_bounds, int64_t *cost_per_gpu, int64_t *interior);
int vbc_dist_set_x(vbc_dist *D, const void *x);      /* host vector of n entries -> every device */
/* `iters` iterations, device-resident, synchronous on return; *ms_per_iter (may be NULL) = device time per iteration,
 * maximum over the devices (CUDA events).  Fused exchange: the iterations are captured into one CUDA graph per device. */
int vbc_dist_spmv_iter(vbc_dist *D, int iters, double alpha, double *ms_per_iter);
int vbc_dist_gather_x(vbc_dist *D, void *x);         /* current x (n entries) assembled from the owners' slices */
void vbc_dist_destroy(vbc_dist *D);

#ifdef __cplusplus
}
#endif
#endif /* VBC_H */
