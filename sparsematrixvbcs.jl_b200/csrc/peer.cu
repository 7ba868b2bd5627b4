// peer.cu -- row-partitioned multiply across the GPUs of one box: x replicated through peer
// memory (CUDA IPC over NVLink / NVSwitch), the all-gather fused into the adjoint kernel's
// epilogue (spmv.cu, PeerDst), and a flag kernel as the only cross-rank step.
#include <new>

#include "common.cuh"

struct vbc_peer {
    int vt = VBC_F64, rank = 0, nranks = 1, device = 0;
    int64_t xlen = 0;
    void *own[VBC_PEER_HANDLES] = {nullptr, nullptr, nullptr};               // x0, x1, flags (this rank)
    void *bufs[VBC_MAX_PEERS][VBC_PEER_HANDLES] = {};                        // every rank's buffers as mapped here
    bool ipc_opened[VBC_MAX_PEERS][VBC_PEER_HANDLES] = {};
    bool connected = false;
    int cur = 0;
    unsigned long long *d_epoch = nullptr; // device-side epoch counter: the flag kernel advances it itself, so a
                                           // captured CUDA graph of steps stays correct when replayed
    int *d_timeout = nullptr;
    unsigned char *d_mask = nullptr; // per column chunk of this rank's slice: which destinations read it
    int chunk_shift = 0;
    int64_t mask_len = 0;
    int fused_sync = 0;      // 0: multiply, then k_peer_flags; 1: flags inside the multiply kernel (removed); 2: split launches
                             // [stripes i0..i1] [wait] [the rest] [signal] so the wait hides behind the first launch;
                             // 3: plain multiply into the own buffer, then ONE kernel that pushes the chunks other ranks read
                             //    and does the flag exchange (sparsity-aware mode only; experimental, not yet run on a GPU)
    int i0 = 0, i1 = 0;      // stripes [i0, i1): no peer involved (run before the in-kernel wait)
    unsigned *d_done = nullptr;
    unsigned nbr_mask = 0xffffffffu; // ranks this rank exchanges flags with (bit r); default: everyone
    int *d_push = nullptr;           // fused_sync 3: the column chunks of this rank's slice that some OTHER rank reads
    int npush = 0;
    int64_t launches = 0;
};

namespace vbc {

struct FlagPtrs {
    unsigned long long *p[VBC_MAX_PEERS]; // flag block of every rank (p[me] is local)
};

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// One CTA of 32 threads; thread r talks to rank r.
//   signal: flags_r[me] = epoch  (release, system scope: this rank's earlier stores -- the y
//           segments written by the preceding multiply on this stream -- are visible first)
//   wait  : spin until flags_me[r] >= epoch for every r (acquire), with a wall-clock bound so a
//           dead peer cannot hang the GPU.
__global__ void k_peer_flags(const __grid_constant__ FlagPtrs f, const int me, const int nranks, unsigned long long *__restrict__ d_epoch,
                             const int do_signal, const int do_wait, int *__restrict__ timed_out, const unsigned nbr_mask)
{
    const int r = threadIdx.x;
    // epoch of this barrier: a signal opens a new epoch, a wait-only launch waits for the open one
    const unsigned long long epoch = *d_epoch + (do_signal ? 1ull : 0ull);
    __syncwarp();
    if (r == 0 && do_signal) *d_epoch = epoch;
    if (r >= nranks || !((nbr_mask >> r) & 1u)) return; // only the ranks this one sends to or receives from
    if (do_signal) {
        __threadfence_system();
        st_release_sys(f.p[r] + me, epoch);
    }
    if (do_wait) {
        unsigned long long t0;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        while (ld_acquire_sys(f.p[me] + r) < epoch) {
            unsigned long long t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > 4000000000ull) { // 4 s
                atomicExch(timed_out, 1);
                break;
            }
            __nanosleep(64);
        }
    }
}

// fused_sync 3.  The multiply has written this rank's whole y slice into its own next-x buffer with the plain
// kernel; only the column chunks some other rank gathers from (push[0..npush), usually a fraction of a percent for a
// banded operator) still have to travel.  Every CTA copies its share of those chunks to the ranks that read them;
// the last CTA to finish then runs the flag exchange of k_peer_flags (signal, wait).
struct PushDst {
    void *p[VBC_MAX_PEERS]; // destination i of the rotated order of vbc_peer_spmv_step (p[0] = own buffer, unused here)
    int n;
};
template <typename Tv>
__global__ void __launch_bounds__(256) k_peer_push_flags(const Tv *__restrict__ src, const __grid_constant__ PushDst dst, const unsigned char *__restrict__ mask,
                                                         const int chunk_shift, const int ncols, const int *__restrict__ push, const int npush,
                                                         const __grid_constant__ FlagPtrs f, const int me, const int nranks, unsigned long long *__restrict__ d_epoch,
                                                         unsigned *__restrict__ d_done, int *__restrict__ timed_out, const unsigned nbr_mask)
{
    for (int idx = blockIdx.x; idx < npush; idx += gridDim.x) {
        const int ch = push[idx];
        const unsigned mk = mask[ch];
        const int c0 = ch << chunk_shift, c1 = min(ncols, c0 + (1 << chunk_shift));
        for (int c = c0 + (int)threadIdx.x; c < c1; c += (int)blockDim.x) {
            const Tv v = src[c];
            for (int i = 1; i < dst.n; i++)
                if ((mk >> i) & 1u) reinterpret_cast<Tv *>(dst.p[i])[c] = v;
        }
    }
    __threadfence_system(); // this thread's peer stores are visible system-wide before the CTA is counted as done
    __shared__ bool last;
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned prev = atomicAdd(d_done, 1u);
        last = prev == gridDim.x - 1;
        if (last) *d_done = 0; // ready for the next launch (stream order)
    }
    __syncthreads();
    if (!last || threadIdx.x >= 32) return;
    __threadfence();
    const int r = threadIdx.x;
    const unsigned long long epoch = *d_epoch + 1ull;
    __syncwarp();
    if (r == 0) *d_epoch = epoch;
    if (r >= nranks || !((nbr_mask >> r) & 1u)) return;
    __threadfence_system();
    st_release_sys(f.p[r] + me, epoch);
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (ld_acquire_sys(f.p[me] + r) < epoch) {
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > 4000000000ull) { atomicExch(timed_out, 1); break; }
        __nanosleep(64);
    }
}

static int flags_launch(vbc_peer *P, cudaStream_t st, int barrier)
{
    if (!(barrier & 3)) return VBC_OK;
    FlagPtrs f;
    for (int r = 0; r < VBC_MAX_PEERS; r++) f.p[r] = r < P->nranks ? (unsigned long long *)P->bufs[r][2] : nullptr;
    k_peer_flags<<<1, 32, 0, st>>>(f, P->rank, P->nranks, P->d_epoch, barrier & 1, (barrier & 2) ? 1 : 0, P->d_timeout, P->nbr_mask | (1u << P->rank));
    P->launches++;
    VBC_CUDA(cudaGetLastError());
    return VBC_OK;
}

} // namespace vbc

using namespace vbc;

extern "C" {

int vbc_peer_create(vbc_peer **out, int vt, int64_t xlen, int rank, int nranks, int device, void *handles_out)
{
    if (!out) VBC_FAIL(VBC_EARG, "out is NULL");
    *out = nullptr;
    if (vt != VBC_F32 && vt != VBC_F64) VBC_FAIL(VBC_EARG, "vt must be VBC_F32 or VBC_F64");
    if (nranks < 1 || nranks > VBC_MAX_PEERS || rank < 0 || rank >= nranks) VBC_FAIL(VBC_EARG, "rank %d / nranks %d out of range (max %d)", rank, nranks, VBC_MAX_PEERS);
    if (xlen < 0) VBC_FAIL(VBC_EARG, "xlen must be >= 0");
    DeviceGuard guard(device);
    if (!guard.ok) VBC_FAIL(VBC_ECUDA, "cudaSetDevice(%d) failed", device);
    vbc_peer *P = new (std::nothrow) vbc_peer();
    if (!P) VBC_FAIL(VBC_ENOMEM, "host allocation failed");
    P->vt = vt; P->rank = rank; P->nranks = nranks; P->device = device; P->xlen = xlen;
    const size_t xb = vt_size(vt) * (size_t)(xlen > 0 ? xlen : 1);
    const size_t sizes[VBC_PEER_HANDLES] = {xb, xb, sizeof(unsigned long long) * VBC_MAX_PEERS};
    int rc = VBC_OK;
    for (int k = 0; k < VBC_PEER_HANDLES && rc == VBC_OK; k++) {
        if (cudaMalloc(&P->own[k], sizes[k]) != cudaSuccess || cudaMemset(P->own[k], 0, sizes[k]) != cudaSuccess) {
            set_error("vbc_peer_create: allocation of %zu bytes failed: %s", sizes[k], cudaGetErrorString(cudaGetLastError()));
            rc = VBC_ENOMEM;
        }
    }
    if (rc == VBC_OK && (cudaMalloc(&P->d_epoch, sizeof(unsigned long long)) != cudaSuccess || cudaMemset(P->d_epoch, 0, sizeof(unsigned long long)) != cudaSuccess)) {
        set_error("vbc_peer_create: epoch allocation failed");
        rc = VBC_ENOMEM;
    }
    if (rc == VBC_OK && (cudaMalloc(&P->d_timeout, sizeof(int)) != cudaSuccess || cudaMemset(P->d_timeout, 0, sizeof(int)) != cudaSuccess)) {
        set_error("vbc_peer_create: flag allocation failed");
        rc = VBC_ENOMEM;
    }
    if (rc == VBC_OK && handles_out) {
        for (int k = 0; k < VBC_PEER_HANDLES; k++) {
            cudaIpcMemHandle_t h;
            static_assert(sizeof(cudaIpcMemHandle_t) == VBC_IPC_HANDLE_BYTES, "IPC handle size");
            if (cudaIpcGetMemHandle(&h, P->own[k]) != cudaSuccess) {
                set_error("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(cudaGetLastError()));
                rc = VBC_ECUDA;
                break;
            }
            memcpy((char *)handles_out + (size_t)k * VBC_IPC_HANDLE_BYTES, &h, VBC_IPC_HANDLE_BYTES);
        }
    }
    if (rc != VBC_OK) { vbc_peer_destroy(P); return rc; }
    for (int k = 0; k < VBC_PEER_HANDLES; k++) P->bufs[rank][k] = P->own[k];
    if (nranks == 1) P->connected = true;
    *out = P;
    return VBC_OK;
}

int vbc_peer_connect(vbc_peer *P, const void *all_handles)
{
    if (!P || !all_handles) VBC_FAIL(VBC_EARG, "NULL argument");
    DeviceGuard guard(P->device);
    if (!guard.ok) VBC_FAIL(VBC_ECUDA, "cudaSetDevice(%d) failed", P->device);
    for (int r = 0; r < P->nranks; r++) {
        if (r == P->rank) continue;
        for (int k = 0; k < VBC_PEER_HANDLES; k++) {
            cudaIpcMemHandle_t h;
            memcpy(&h, (const char *)all_handles + ((size_t)r * VBC_PEER_HANDLES + k) * VBC_IPC_HANDLE_BYTES, VBC_IPC_HANDLE_BYTES);
            void *ptr = nullptr;
            cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) VBC_FAIL(VBC_ECUDA, "cudaIpcOpenMemHandle(rank %d, buffer %d) failed: %s", r, k, cudaGetErrorString(e));
            P->bufs[r][k] = ptr;
            P->ipc_opened[r][k] = true;
        }
    }
    P->connected = true;
    return VBC_OK;
}

int vbc_peer_connect_local(vbc_peer *P, void *const *ptrs)
{
    if (!P || !ptrs) VBC_FAIL(VBC_EARG, "NULL argument");
    for (int r = 0; r < P->nranks; r++)
        for (int k = 0; k < VBC_PEER_HANDLES; k++)
            if (r != P->rank) P->bufs[r][k] = ptrs[r * VBC_PEER_HANDLES + k];
    P->connected = true;
    return VBC_OK;
}

int vbc_peer_buffer(vbc_peer *P, int k, void **ptr)
{
    if (!P || !ptr || k < 0 || k >= VBC_PEER_HANDLES) VBC_FAIL(VBC_EARG, "bad argument");
    *ptr = P->own[k];
    return VBC_OK;
}

int vbc_peer_current(const vbc_peer *P, int *cur)
{
    if (!P || !cur) VBC_FAIL(VBC_EARG, "NULL argument");
    *cur = P->cur;
    return VBC_OK;
}

int vbc_peer_spmv_step(vbc_peer *P, vbc_mat *A, double alpha, int64_t y_offset, int barrier)
{
    if (!P || !A) VBC_FAIL(VBC_EARG, "NULL argument");
    if (!P->connected) VBC_FAIL(VBC_EARG, "vbc_peer_connect has not been called");
    if (A->vt != P->vt) VBC_FAIL(VBC_EARG, "matrix and exchange buffers have different element types");
    if (A->m != P->xlen) VBC_FAIL(VBC_EDIM, "DimensionMismatch: A' needs x of length %lld, exchange buffers hold %lld", (long long)A->m, (long long)P->xlen);
    if (y_offset < 0 || y_offset + A->n > P->xlen) VBC_FAIL(VBC_EDIM, "DimensionMismatch: y slice [%lld, %lld) outside x of length %lld", (long long)y_offset, (long long)(y_offset + A->n), (long long)P->xlen);
    DeviceGuard guard(P->device);
    if (!guard.ok) VBC_FAIL(VBC_ECUDA, "cudaSetDevice(%d) failed", P->device);
    const size_t tv = vt_size(P->vt);
    void *dst[VBC_MAX_PEERS];
    const int nxt = 1 - P->cur;
    // own buffer first: its stores are local; then the peers, starting after this rank so the
    // ranks do not all hit the same destination at the same moment
    int n = 0;
    for (int i = 0; i < P->nranks; i++) {
        const int r = (P->rank + i) % P->nranks;
        dst[n++] = (char *)P->bufs[r][nxt] + tv * (size_t)y_offset;
    }
    if (P->d_mask && A->n > 0 && ((A->n - 1) >> P->chunk_shift) >= P->mask_len)
        VBC_FAIL(VBC_EDIM, "peer mask covers %lld chunks, the y slice needs %lld", (long long)P->mask_len, (long long)(((A->n - 1) >> P->chunk_shift) + 1));
    // the mask is indexed by (column in the y slice) >> shift; the kernel indexes by slab column, so no offset is needed
    if (P->fused_sync == 3 && barrier == 3 && P->d_mask && P->d_done) {
        // plain kernel into the own next-x buffer, then push + flags in one launch
        VBC_TRY(launch_spmv(A, 1, alpha, P->own[P->cur], 0.0, dst[0]));
        PushDst pd;
        pd.n = n;
        for (int i = 0; i < VBC_MAX_PEERS; i++) pd.p[i] = i < n ? dst[i] : nullptr;
        FlagPtrs f;
        for (int r = 0; r < VBC_MAX_PEERS; r++) f.p[r] = r < P->nranks ? (unsigned long long *)P->bufs[r][2] : nullptr;
        int grid = P->npush < 1 ? 1 : (P->npush > 64 ? 64 : P->npush);
        const unsigned nbr = P->nbr_mask | (1u << P->rank);
        if (P->vt == VBC_F64)
            k_peer_push_flags<double><<<grid, 256, 0, A->stream>>>((const double *)dst[0], pd, P->d_mask, P->chunk_shift, (int)A->n, P->d_push, P->npush, f, P->rank,
                                                                   P->nranks, P->d_epoch, P->d_done, P->d_timeout, nbr);
        else
            k_peer_push_flags<float><<<grid, 256, 0, A->stream>>>((const float *)dst[0], pd, P->d_mask, P->chunk_shift, (int)A->n, P->d_push, P->npush, f, P->rank,
                                                                  P->nranks, P->d_epoch, P->d_done, P->d_timeout, nbr);
        P->launches++;
        VBC_CUDA(cudaGetLastError());
    } else if (P->fused_sync == 2 && barrier == 3 && P->i1 > P->i0) {
        const int L = (int)A->L;
        const int ra[4] = {P->i0, P->i1, 0, 0}, rc[4] = {0, P->i0, P->i1, L};
        VBC_TRY(launch_spmv_adj_peer(A, alpha, P->own[P->cur], n, dst, P->d_mask, P->chunk_shift, nullptr, ra));
        VBC_TRY(flags_launch(P, A->stream, 2)); // wait for the peers' previous step
        VBC_TRY(launch_spmv_adj_peer(A, alpha, P->own[P->cur], n, dst, P->d_mask, P->chunk_shift, nullptr, rc));
        VBC_TRY(flags_launch(P, A->stream, 1)); // publish this step
    } else {
        VBC_TRY(launch_spmv_adj_peer(A, alpha, P->own[P->cur], n, dst, P->d_mask, P->chunk_shift, nullptr, nullptr));
        VBC_TRY(flags_launch(P, A->stream, barrier));
    }
    P->cur = nxt;
    return VBC_OK;
}

int vbc_peer_set_mask(vbc_peer *P, const void *mask, int64_t nchunks, int chunk_shift)
{
    if (!P) VBC_FAIL(VBC_EARG, "NULL argument");
    DeviceGuard guard(P->device);
    cudaFree(P->d_mask);
    cudaFree(P->d_push);
    P->d_mask = nullptr; P->d_push = nullptr; P->npush = 0; P->mask_len = 0; P->chunk_shift = 0;
    if (!mask) return VBC_OK; // back to full replication
    if (nchunks < 1 || chunk_shift < 0 || chunk_shift > 30) VBC_FAIL(VBC_EARG, "bad mask geometry");
    VBC_CUDA(cudaMalloc(&P->d_mask, (size_t)nchunks));
    VBC_CUDA(cudaMemcpy(P->d_mask, mask, (size_t)nchunks, cudaMemcpyHostToDevice));
    P->mask_len = nchunks;
    P->chunk_shift = chunk_shift;
    { // chunks with a reader other than this rank (bit 0 is the own buffer in the rotated destination order)
        int *list = new (std::nothrow) int[(size_t)nchunks];
        if (!list) VBC_FAIL(VBC_ENOMEM, "host allocation failed");
        int np = 0;
        for (int64_t c = 0; c < nchunks; c++)
            if (((const unsigned char *)mask)[c] & 0xfeu) list[np++] = (int)c;
        cudaError_t e = cudaMalloc(&P->d_push, sizeof(int) * (size_t)(np > 0 ? np : 1));
        if (e == cudaSuccess && np > 0) e = cudaMemcpy(P->d_push, list, sizeof(int) * (size_t)np, cudaMemcpyHostToDevice);
        delete[] list;
        if (e != cudaSuccess) VBC_FAIL(VBC_ECUDA, "vbc_peer_set_mask: push list upload failed: %s", cudaGetErrorString(e));
        P->npush = np;
    }
    return VBC_OK;
}

int vbc_peer_set_fused_sync(vbc_peer *P, int enable, int64_t i0, int64_t i1)
{
    if (!P) VBC_FAIL(VBC_EARG, "NULL argument");
    if (i0 < 0 || i1 < i0 || enable < 0 || enable > 3) VBC_FAIL(VBC_EARG, "bad interior range / mode");
    if (enable == 1) enable = 0; // the in-kernel flag exchange lost to the separate flag kernel and quadrupled the kernel's code size; removed
    DeviceGuard guard(P->device);
    if (enable && !P->d_done) {
        VBC_CUDA(cudaMalloc(&P->d_done, sizeof(unsigned)));
        VBC_CUDA(cudaMemset(P->d_done, 0, sizeof(unsigned)));
    }
    P->fused_sync = enable;
    P->i0 = (int)i0;
    P->i1 = (int)i1;
    return VBC_OK;
}

int vbc_peer_set_neighbors(vbc_peer *P, unsigned mask)
{
    if (!P) VBC_FAIL(VBC_EARG, "NULL argument");
    P->nbr_mask = mask;
    return VBC_OK;
}

int vbc_peer_barrier(vbc_peer *P, void *cuda_stream, int barrier)
{
    if (!P) VBC_FAIL(VBC_EARG, "NULL argument");
    if (!P->connected) VBC_FAIL(VBC_EARG, "vbc_peer_connect has not been called");
    DeviceGuard guard(P->device);
    return flags_launch(P, (cudaStream_t)cuda_stream, barrier);
}

int vbc_peer_status(vbc_peer *P, int *timed_out)
{
    if (!P || !timed_out) VBC_FAIL(VBC_EARG, "NULL argument");
    DeviceGuard guard(P->device);
    VBC_CUDA(cudaMemcpy(timed_out, P->d_timeout, sizeof(int), cudaMemcpyDeviceToHost));
    return VBC_OK;
}

void vbc_peer_destroy(vbc_peer *P)
{
    if (!P) return;
    DeviceGuard guard(P->device);
    cudaDeviceSynchronize();
    for (int r = 0; r < VBC_MAX_PEERS; r++)
        for (int k = 0; k < VBC_PEER_HANDLES; k++)
            if (P->ipc_opened[r][k]) cudaIpcCloseMemHandle(P->bufs[r][k]);
    for (int k = 0; k < VBC_PEER_HANDLES; k++) cudaFree(P->own[k]);
    cudaFree(P->d_timeout);
    cudaFree(P->d_epoch);
    cudaFree(P->d_mask);
    cudaFree(P->d_done);
    cudaFree(P->d_push);
    delete P;
}

} // extern "C"
