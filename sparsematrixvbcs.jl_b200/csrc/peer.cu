// peer.cu -- row-partitioned multiply across the GPUs of one box: x lives in peer-mapped buffers
// (CUDA IPC between processes, or plain peer access inside one process), and the exchange of the new
// x is fused into the adjoint kernel (spmv.cu, k_spmv_adj_halo): boundary stripes store their results
// straight into the buffers of the ranks that read them over NVLink / NVSwitch and the last one
// publishes a flag; there is no collective call and no separate exchange launch in an iteration.
#include <new>

#include "common.cuh"

#define VBC_PEER_CTL_WORDS 8 // [1] finished boundary warps (running total), [2] epoch, [3] wait ns, [4] waits that spun, [5] longest wait

struct vbc_peer {
    int vt = VBC_F64, rank = 0, nranks = 1, device = 0;
    int64_t xlen = 0;
    void *own[VBC_PEER_HANDLES] = {nullptr, nullptr, nullptr};               // x0, x1, flags (this rank)
    void *bufs[VBC_MAX_PEERS][VBC_PEER_HANDLES] = {};                        // every rank's buffers as mapped here
    bool ipc_opened[VBC_MAX_PEERS][VBC_PEER_HANDLES] = {};
    bool connected = false;
    int cur = 0;
    unsigned long long *d_ctl = nullptr;   // device-side counters (VBC_PEER_CTL_WORDS): the kernels advance the epoch themselves,
                                           // so a captured CUDA graph of steps stays correct when replayed
    int *d_timeout = nullptr;
    unsigned char *d_mask = nullptr; // per column chunk of this rank's slice: which destinations read it
    int chunk_shift = 0;
    int64_t mask_len = 0;
    int i0 = 0, i1 = 0;              // interior stripes [i0, i1): gather only from this rank's slice, feed only this rank
    unsigned nbr_mask = 0xffffffffu; // ranks this rank exchanges flags with (bit r); default: everyone
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

// The flag exchange on its own (vbc_peer_barrier): one CTA of 32 threads; thread r talks to rank r.
//   signal: flags_r[me] = epoch  (release, system scope: this rank's earlier stores on this stream are visible first)
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
    if (r >= nranks || r == me || !((nbr_mask >> r) & 1u)) return; // only the ranks this one sends to or receives from
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

// interior[l] = 1 when stripe l gathers only from x[own_lo, own_hi) and every column chunk it writes is read by this rank alone
__global__ void __launch_bounds__(256) k_mark_interior(const StripeMeta *__restrict__ meta, const int *__restrict__ desc, const int L, const int reach,
                                                        const long long own_lo, const long long own_hi, const unsigned char *__restrict__ mask,
                                                        const int chunk_shift, const int full_replication, unsigned char *__restrict__ interior)
{
    const int l = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 3), lane = threadIdx.x & 7; // 8 lanes per stripe
    if (l >= L) return;
    const StripeMeta a = meta[l], b = meta[l + 1];
    bool ok = !full_replication;
    for (int q = a.pos + lane; ok && q < b.pos; q += 8) {
        const long long i = desc[q];
        ok = i >= own_lo && i + reach <= own_hi;
    }
    if (ok && mask != nullptr && b.col > a.col && lane == 0)
        for (int ch = a.col >> chunk_shift; ch <= ((b.col - 1) >> chunk_shift); ch++) ok = ok && ((mask[ch] | 1u) == 1u);
    ok = __all_sync(0xffu << ((threadIdx.x & 31) & ~7), ok);
    if (lane == 0) interior[l] = ok ? 1 : 0;
}

static int flags_launch(vbc_peer *P, cudaStream_t st, int barrier)
{
    if (!(barrier & 3)) return VBC_OK;
    FlagPtrs f;
    for (int r = 0; r < VBC_MAX_PEERS; r++) f.p[r] = r < P->nranks ? (unsigned long long *)P->bufs[r][2] : nullptr;
    k_peer_flags<<<1, 32, 0, st>>>(f, P->rank, P->nranks, P->d_ctl + 2, barrier & 1, (barrier & 2) ? 1 : 0, P->d_timeout, P->nbr_mask);
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
    if (rc == VBC_OK && (cudaMalloc(&P->d_ctl, sizeof(unsigned long long) * VBC_PEER_CTL_WORDS) != cudaSuccess ||
                         cudaMemset(P->d_ctl, 0, sizeof(unsigned long long) * VBC_PEER_CTL_WORDS) != cudaSuccess)) {
        set_error("vbc_peer_create: counter allocation failed");
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
    if (barrier < 0 || barrier > 3) VBC_FAIL(VBC_EARG, "barrier must be 0..3");
    if (A->vt != P->vt) VBC_FAIL(VBC_EARG, "matrix and exchange buffers have different element types");
    if (A->m != P->xlen) VBC_FAIL(VBC_EDIM, "DimensionMismatch: A' needs x of length %lld, exchange buffers hold %lld", (long long)A->m, (long long)P->xlen);
    if (y_offset < 0 || y_offset + A->n > P->xlen) VBC_FAIL(VBC_EDIM, "DimensionMismatch: y slice [%lld, %lld) outside x of length %lld", (long long)y_offset, (long long)(y_offset + A->n), (long long)P->xlen);
    if (P->i1 > A->L) VBC_FAIL(VBC_EDIM, "interior stripe range [%d, %d) outside the %lld stripes of the matrix", P->i0, P->i1, (long long)A->L);
    DeviceGuard guard(P->device);
    if (!guard.ok) VBC_FAIL(VBC_ECUDA, "cudaSetDevice(%d) failed", P->device);
    const size_t tv = vt_size(P->vt);
    const int nxt = 1 - P->cur;
    if (P->d_mask && A->n > 0 && ((A->n - 1) >> P->chunk_shift) >= P->mask_len)
        VBC_FAIL(VBC_EDIM, "peer mask covers %lld chunks, the y slice needs %lld", (long long)P->mask_len, (long long)(((A->n - 1) >> P->chunk_shift) + 1));
    HaloLaunch hl{};
    // own buffer first: its stores are local; then the peers, starting after this rank so the
    // ranks do not all hit the same destination at the same moment
    hl.n = P->nranks;
    for (int i = 0; i < P->nranks; i++) {
        const int r = (P->rank + i) % P->nranks;
        hl.dst[i] = (char *)P->bufs[r][nxt] + tv * (size_t)y_offset;
    }
    // the mask is indexed by (column in the y slice) >> shift; the kernel indexes by slab column, so no offset is needed
    hl.d_mask = P->d_mask; hl.chunk_shift = P->chunk_shift;
    hl.i0 = P->i0; hl.i1 = P->i1;
    hl.me = P->rank; hl.nranks = P->nranks;
    hl.do_signal = barrier & 1; hl.do_wait = (barrier & 2) ? 1 : 0;
    hl.nbr_mask = P->nbr_mask;
    for (int r = 0; r < VBC_MAX_PEERS; r++) hl.flags[r] = r < P->nranks ? (unsigned long long *)P->bufs[r][2] : nullptr;
    hl.ctl = P->d_ctl;
    hl.timed_out = P->d_timeout;
    VBC_TRY(launch_spmv_adj_halo(A, alpha, P->own[P->cur], &hl));
    P->launches++;
    P->cur = nxt;
    return VBC_OK;
}

int vbc_peer_set_mask(vbc_peer *P, const void *mask, int64_t nchunks, int chunk_shift)
{
    if (!P) VBC_FAIL(VBC_EARG, "NULL argument");
    DeviceGuard guard(P->device);
    cudaFree(P->d_mask);
    P->d_mask = nullptr; P->mask_len = 0; P->chunk_shift = 0;
    if (!mask) return VBC_OK; // back to full replication
    if (nchunks < 1 || chunk_shift < 0 || chunk_shift > 30) VBC_FAIL(VBC_EARG, "bad mask geometry");
    VBC_CUDA(cudaMalloc(&P->d_mask, (size_t)nchunks));
    VBC_CUDA(cudaMemcpy(P->d_mask, mask, (size_t)nchunks, cudaMemcpyHostToDevice));
    P->mask_len = nchunks;
    P->chunk_shift = chunk_shift;
    return VBC_OK;
}

int vbc_peer_set_interior(vbc_peer *P, int64_t i0, int64_t i1)
{
    if (!P) VBC_FAIL(VBC_EARG, "NULL argument");
    if (i0 < 0 || i1 < i0 || i1 > 0x7fffffff) VBC_FAIL(VBC_EARG, "bad interior range [%lld, %lld)", (long long)i0, (long long)i1);
    P->i0 = (int)i0;
    P->i1 = (int)i1;
    return VBC_OK;
}

int vbc_peer_get_interior(const vbc_peer *P, int64_t *i0, int64_t *i1)
{
    if (!P || !i0 || !i1) VBC_FAIL(VBC_EARG, "NULL argument");
    *i0 = P->i0;
    *i1 = P->i1;
    return VBC_OK;
}

int vbc_peer_auto_interior(vbc_peer *P, vbc_mat *A, int64_t y_offset, int64_t *i0_out, int64_t *i1_out)
{
    if (!P || !A) VBC_FAIL(VBC_EARG, "NULL argument");
    if (A->m != P->xlen || y_offset < 0 || y_offset + A->n > P->xlen) VBC_FAIL(VBC_EDIM, "DimensionMismatch: the matrix does not fit the exchange buffers");
    if (A->opt_parity) VBC_FAIL(VBC_EARG, "needs the compact layout (parity mode is on)");
    DeviceGuard guard(P->device);
    if (!guard.ok) VBC_FAIL(VBC_ECUDA, "cudaSetDevice(%d) failed", P->device);
    const int L = (int)A->L;
    int64_t best0 = 0, best1 = 0;
    if (L > 0 && P->nranks > 1) {
        unsigned char *d_f = nullptr;
        VBC_CUDA(cudaMalloc(&d_f, (size_t)L));
        unsigned char *h_f = new (std::nothrow) unsigned char[(size_t)L];
        if (!h_f) { cudaFree(d_f); VBC_FAIL(VBC_ENOMEM, "host allocation failed"); }
        const int reach = A->desc_mode == DESC_BLOCKS ? A->u0 : 1;
        const long long grid = ((long long)L * 8 + 255) / 256;
        // without a mask every stripe feeds every rank: nothing is interior
        k_mark_interior<<<(unsigned)grid, 256, 0, A->stream>>>(A->d_meta, A->d_desc, L, reach, (long long)y_offset, (long long)(y_offset + A->n), P->d_mask,
                                                                P->chunk_shift, P->d_mask == nullptr ? 1 : 0, d_f);
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpyAsync(h_f, d_f, (size_t)L, cudaMemcpyDeviceToHost, A->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(A->stream);
        cudaFree(d_f);
        if (e != cudaSuccess) { delete[] h_f; VBC_FAIL(VBC_ECUDA, "vbc_peer_auto_interior: %s", cudaGetErrorString(e)); }
        for (int64_t l = 0; l < L;) { // longest run of interior stripes
            if (!h_f[l]) { l++; continue; }
            int64_t e1 = l;
            while (e1 < L && h_f[e1]) e1++;
            if (e1 - l > best1 - best0) { best0 = l; best1 = e1; }
            l = e1;
        }
        delete[] h_f;
    } else if (P->nranks == 1) {
        best0 = 0; best1 = L; // a single rank has no boundary
    }
    P->i0 = (int)best0; P->i1 = (int)best1;
    if (i0_out) *i0_out = best0;
    if (i1_out) *i1_out = best1;
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

int vbc_peer_wait_stats(vbc_peer *P, uint64_t stats[4], int reset)
{
    if (!P || !stats) VBC_FAIL(VBC_EARG, "NULL argument");
    DeviceGuard guard(P->device);
    unsigned long long h[VBC_PEER_CTL_WORDS];
    VBC_CUDA(cudaMemcpy(h, P->d_ctl, sizeof(h), cudaMemcpyDeviceToHost));
    stats[0] = h[2]; stats[1] = h[3]; stats[2] = h[4]; stats[3] = h[5];
    if (reset) VBC_CUDA(cudaMemset(P->d_ctl + 3, 0, 3 * sizeof(unsigned long long)));
    return VBC_OK;
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
    cudaFree(P->d_ctl);
    cudaFree(P->d_mask);
    delete P;
}

} // extern "C"
