// dist.cu -- vbc_dist_*: the row-partitioned multiply driven by ONE process over several GPUs of a box
// (the form a Julia host uses: no torch, no MPI).  Host logic only; the device work is the pack kernels (pack.cu) and
// the fused multiply + exchange kernel (spmv.cu / peer.cu) through the public vbc_* / vbc_peer_* entry points.
//
//   * stripes are split into P contiguous ranges at the P-quantiles of the prefix sum of the reference's own memory cost
//     model (costs.jl:10 1D: 3|Ti| + rows (|Ti| + w |Tv|); costs.jl:140 2D: 3|Ti| + sum over blocks (|Ti| + u w |Tv|)),
//     computed from the CSC structure with the reference's own counting pass (last-seen array, constructors_1DVBC.jl:22-32 /
//     constructors_VBC.jl:31-47);
//   * x lives in padded per-rank-slice coordinates (rank r's slice at [r S, r S + len_r), S = max len), so the slabs' row
//     indices are remapped on the host before the pack kernel runs and slices of unequal length need no repacking;
//   * every device gets a vbc_mat (its slab) and a vbc_peer (its x buffers and flags), connected through plain peer access;
//     the exchange plan (who reads which chunk of whose slice) comes from vbc_read_chunks on each packed slab;
//   * an iteration is one launch per device; `iters` iterations are captured into one CUDA graph per device.
// Exchange VBC_EXCH_NCCL is the unfused comparator: plain multiply, then ncclAllGather of the padded slices (libnccl is
// loaded with dlopen at first use, so libvbc.so itself does not link it).
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <new>
#include <vector>

#include "common.cuh"

struct vbc_dist {
    int P = 1, vt = VBC_F64, it = VBC_I64, exchange = VBC_EXCH_FUSED;
    int64_t n = 0, S = 0, padded = 0;
    std::vector<int> dev;
    std::vector<int64_t> stripe_bounds, col_bounds, cost;
    std::vector<vbc_mat *> mat;
    std::vector<vbc_peer *> peer;
    std::vector<cudaStream_t> stream;
    std::vector<cudaEvent_t> ev0, ev1;
    std::vector<cudaGraphExec_t> graph;
    int graph_iters = 0, graph_cur = -1;
    bool lockstep = false;    // a device is listed more than once: its ranks' kernels may not wait for one another on the device
                              // (nothing guarantees they run at the same time), so every iteration is launched without in-kernel
                              // flags and the host synchronises all ranks between iterations
    double graph_alpha = 0.0;
    // NCCL comparator
    void *nccl_lib = nullptr;
    std::vector<void *> comm;
    std::vector<void *> ybuf; // per device: this rank's y slice (S elements)
    std::vector<int> cur;     // NCCL mode: current x buffer per device (the peer objects' buffers are reused)
    std::vector<unsigned char> interior_set;
};

namespace vbc {

// ---- NCCL through dlopen -------------------------------------------------------------------
typedef int (*nccl_CommInitAll_t)(void **comms, int ndev, const int *devlist);
typedef int (*nccl_CommDestroy_t)(void *comm);
typedef int (*nccl_AllGather_t)(const void *send, void *recv, size_t count, int dtype, void *comm, cudaStream_t s);
typedef int (*nccl_Group_t)(void);
typedef const char *(*nccl_GetErrorString_t)(int);
struct NcclApi {
    nccl_CommInitAll_t CommInitAll = nullptr;
    nccl_CommDestroy_t CommDestroy = nullptr;
    nccl_AllGather_t AllGather = nullptr;
    nccl_Group_t GroupStart = nullptr, GroupEnd = nullptr;
    nccl_GetErrorString_t GetErrorString = nullptr;
};
static NcclApi g_nccl;

static int load_nccl(vbc_dist *D)
{
    if (g_nccl.AllGather) return VBC_OK;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *nm : names)
        if ((h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL)) != nullptr) break;
    if (!h) VBC_FAIL(VBC_ENCCL, "libnccl.so.2 could not be loaded: %s", dlerror());
    D->nccl_lib = h;
    g_nccl.CommInitAll = (nccl_CommInitAll_t)dlsym(h, "ncclCommInitAll");
    g_nccl.CommDestroy = (nccl_CommDestroy_t)dlsym(h, "ncclCommDestroy");
    g_nccl.AllGather = (nccl_AllGather_t)dlsym(h, "ncclAllGather");
    g_nccl.GroupStart = (nccl_Group_t)dlsym(h, "ncclGroupStart");
    g_nccl.GroupEnd = (nccl_Group_t)dlsym(h, "ncclGroupEnd");
    g_nccl.GetErrorString = (nccl_GetErrorString_t)dlsym(h, "ncclGetErrorString");
    if (!g_nccl.CommInitAll || !g_nccl.CommDestroy || !g_nccl.AllGather || !g_nccl.GroupStart || !g_nccl.GroupEnd) {
        g_nccl = NcclApi{};
        VBC_FAIL(VBC_ENCCL, "libnccl is missing ncclCommInitAll / ncclAllGather / ncclGroupStart");
    }
    return VBC_OK;
}
#define VBC_NCCL(call)                                                                                                    \
    do {                                                                                                                  \
        int r__ = (call);                                                                                                 \
        if (r__ != 0) VBC_FAIL(VBC_ENCCL, "%s -> %s", #call, g_nccl.GetErrorString ? g_nccl.GetErrorString(r__) : "NCCL error"); \
    } while (0)

static inline int64_t rdi(const void *a, int it, int64_t i) { return it == VBC_I64 ? ((const int64_t *)a)[i] : (int64_t)((const int32_t *)a)[i]; }
static inline void wri(void *a, int it, int64_t i, int64_t v)
{
    if (it == VBC_I64) ((int64_t *)a)[i] = v;
    else ((int32_t *)a)[i] = (int32_t)v;
}

// per-stripe cost under the reference's memory model, from the CSC structure (the reference's counting pass)
static void stripe_costs(int it, int64_t m, int64_t n, const void *colptr, const void *rowval, const void *pi_spl, int64_t K, const void *phi_spl,
                         int64_t L, int64_t tv, std::vector<int64_t> &cost)
{
    const int64_t ti = (int64_t)it_size(it);
    cost.assign((size_t)L, 0);
    std::vector<int64_t> last((size_t)std::max<int64_t>(pi_spl ? K : m, 1), -1);
    std::vector<int64_t> part; // row -> part id (2D)
    if (pi_spl) {
        part.resize((size_t)std::max<int64_t>(m, 1));
        for (int64_t k = 0; k < K; k++)
            for (int64_t i = rdi(pi_spl, it, k) - 1; i < rdi(pi_spl, it, k + 1) - 1; i++) part[(size_t)i] = k;
    }
    (void)n;
    for (int64_t l = 0; l < L; l++) {
        const int64_t j0 = rdi(phi_spl, it, l) - 1, j1 = rdi(phi_spl, it, l + 1) - 1, w = j1 - j0;
        int64_t units = 0, rows = 0;
        for (int64_t j = j0; j < j1; j++)
            for (int64_t q = rdi(colptr, it, j) - 1; q < rdi(colptr, it, j + 1) - 1; q++) {
                const int64_t i = rdi(rowval, it, q) - 1;
                const int64_t unit = pi_spl ? part[(size_t)i] : i;
                if (last[(size_t)unit] != l) { // `hst[i] < l` of the reference
                    last[(size_t)unit] = l;
                    units++;
                    rows += pi_spl ? rdi(pi_spl, it, unit + 1) - rdi(pi_spl, it, unit) : 1;
                }
            }
        cost[(size_t)l] = 3 * ti + units * ti + rows * w * tv;
    }
}

} // namespace vbc

using namespace vbc;

extern "C" {

void vbc_dist_destroy(vbc_dist *D)
{
    if (!D) return;
    DeviceGuard guard(D->dev.empty() ? 0 : D->dev[0]);
    for (size_t r = 0; r < D->dev.size(); r++) {
        cudaSetDevice(D->dev[r]);
        cudaDeviceSynchronize();
        if (r < D->graph.size() && D->graph[r]) cudaGraphExecDestroy(D->graph[r]);
        if (r < D->comm.size() && D->comm[r] && g_nccl.CommDestroy) g_nccl.CommDestroy(D->comm[r]);
        if (r < D->ybuf.size()) cudaFree(D->ybuf[r]);
        if (r < D->peer.size()) vbc_peer_destroy(D->peer[r]);
        if (r < D->mat.size()) vbc_destroy(D->mat[r]);
        if (r < D->ev0.size() && D->ev0[r]) cudaEventDestroy(D->ev0[r]);
        if (r < D->ev1.size() && D->ev1[r]) cudaEventDestroy(D->ev1[r]);
        if (r < D->stream.size() && D->stream[r]) cudaStreamDestroy(D->stream[r]);
    }
    delete D;
}

int vbc_dist_create(vbc_dist **out, int ngpus, const int *devices, int vt, int it, int64_t n, int U, int W, const void *colptr, const void *rowval,
                    const void *nzval, const void *pi_spl, int64_t K, const void *phi_spl, int64_t L, int exchange)
{
    if (!out) VBC_FAIL(VBC_EARG, "out is NULL");
    *out = nullptr;
    if (ngpus < 1 || ngpus > VBC_MAX_PEERS) VBC_FAIL(VBC_EARG, "ngpus must be in 1..%d", VBC_MAX_PEERS);
    if ((vt != VBC_F32 && vt != VBC_F64) || (it != VBC_I32 && it != VBC_I64)) VBC_FAIL(VBC_EARG, "bad element / index type");
    if (exchange != VBC_EXCH_FUSED && exchange != VBC_EXCH_NCCL) VBC_FAIL(VBC_EARG, "exchange must be VBC_EXCH_FUSED or VBC_EXCH_NCCL");
    if (n < 0 || W <= 0 || L < 0 || (pi_spl && (U <= 0 || K < 0))) VBC_FAIL(VBC_EARG, "ArgumentError: bad shape");
    if (!colptr || !phi_spl || (!rowval && n > 0)) VBC_FAIL(VBC_EARG, "NULL array argument");
    if (rdi(phi_spl, it, 0) != 1 || rdi(phi_spl, it, L) != n + 1) VBC_FAIL(VBC_EARG, "Φ is not a SplitPartition of 1:%lld", (long long)n);
    if (pi_spl && (rdi(pi_spl, it, 0) != 1 || rdi(pi_spl, it, K) != n + 1)) VBC_FAIL(VBC_EARG, "Π is not a SplitPartition of 1:%lld (the iterated operator is square)", (long long)n);
    for (int64_t l = 0; l < L; l++)
        if (rdi(phi_spl, it, l + 1) < rdi(phi_spl, it, l)) VBC_FAIL(VBC_EARG, "Φ.spl is decreasing");
    if (rdi(colptr, it, 0) != 1) VBC_FAIL(VBC_EARG, "colptr[1] must be 1");
    const int64_t nnz = rdi(colptr, it, n) - 1;
    for (int64_t j = 0; j < n; j++)
        if (rdi(colptr, it, j + 1) < rdi(colptr, it, j)) VBC_FAIL(VBC_EARG, "colptr is decreasing at column %lld", (long long)(j + 1));
    for (int64_t q = 0; q < nnz; q++)
        if (rdi(rowval, it, q) < 1 || rdi(rowval, it, q) > n) VBC_FAIL(VBC_EARG, "rowval[%lld] out of 1:%lld", (long long)(q + 1), (long long)n);
    int ndev = 0;
    VBC_CUDA(cudaGetDeviceCount(&ndev));
    DeviceGuard guard(devices ? devices[0] : 0);
    vbc_dist *D = new (std::nothrow) vbc_dist();
    if (!D) VBC_FAIL(VBC_ENOMEM, "host allocation failed");
    D->P = ngpus; D->vt = vt; D->it = it; D->n = n; D->exchange = exchange;
    for (int r = 0; r < ngpus; r++) {
        const int d = devices ? devices[r] : r;
        if (d < 0 || d >= ndev) { delete D; VBC_FAIL(VBC_EARG, "device %d of rank %d does not exist (%d visible)", d, r, ndev); }
        D->dev.push_back(d);
    }
    const int P = ngpus;
    for (int r = 0; r < P; r++)
        for (int q = 0; q < r; q++)
            if (D->dev[(size_t)r] == D->dev[(size_t)q]) D->lockstep = true;
    if (D->lockstep && exchange == VBC_EXCH_NCCL) { delete D; VBC_FAIL(VBC_ENCCL, "the NCCL exchange needs one distinct device per rank (a device is listed twice)"); }
    const int64_t tv = (int64_t)vt_size(vt), ti = (int64_t)it_size(it);
    // ---- split the stripes by cost; a rank boundary must also be a row-part boundary (the slices of x are whole row parts)
    stripe_costs(it, n, n, colptr, rowval, pi_spl, K, phi_spl, L, tv, D->cost);
    std::vector<int64_t> pre((size_t)L + 1, 0);
    for (int64_t l = 0; l < L; l++) pre[(size_t)l + 1] = pre[(size_t)l] + D->cost[(size_t)l];
    std::vector<char> allowed((size_t)L + 1, 1);
    if (pi_spl) {
        std::vector<char> is_pi((size_t)n + 2, 0);
        for (int64_t k = 0; k <= K; k++) is_pi[(size_t)rdi(pi_spl, it, k)] = 1;
        for (int64_t l = 0; l <= L; l++) allowed[(size_t)l] = is_pi[(size_t)rdi(phi_spl, it, l)];
    }
    D->stripe_bounds.assign((size_t)P + 1, 0);
    D->stripe_bounds[(size_t)P] = L;
    for (int r = 1; r < P; r++) {
        const double target = (double)pre[(size_t)L] * r / P;
        int64_t k = std::lower_bound(pre.begin(), pre.end(), (int64_t)target) - pre.begin();
        if (k > 0 && std::abs((double)pre[(size_t)k - 1] - target) <= std::abs((double)pre[(size_t)std::min(k, L)] - target)) k--;
        k = std::min(std::max(k, D->stripe_bounds[(size_t)r - 1]), L);
        int64_t lo = k, hi = k; // nearest allowed boundary
        while (lo > D->stripe_bounds[(size_t)r - 1] && !allowed[(size_t)lo]) lo--;
        while (hi < L && !allowed[(size_t)hi]) hi++;
        if (allowed[(size_t)lo] && (!allowed[(size_t)hi] || k - lo <= hi - k)) k = lo;
        else if (allowed[(size_t)hi]) k = hi;
        else { delete D; VBC_FAIL(VBC_EARG, "no stripe boundary near the cost quantile of rank %d is also a row-part boundary of Π", r); }
        D->stripe_bounds[(size_t)r] = k;
    }
    D->col_bounds.resize((size_t)P + 1);
    for (int r = 0; r <= P; r++) D->col_bounds[(size_t)r] = rdi(phi_spl, it, D->stripe_bounds[(size_t)r]) - 1;
    int64_t S = 0;
    for (int r = 0; r < P; r++) S = std::max(S, D->col_bounds[(size_t)r + 1] - D->col_bounds[(size_t)r]);
    D->S = S;
    D->padded = S * P;
    auto to_padded = [&](int64_t i) { // global 0-based index -> padded
        const int64_t r = std::upper_bound(D->col_bounds.begin(), D->col_bounds.end(), i) - D->col_bounds.begin() - 1;
        return r * S + (i - D->col_bounds[(size_t)r]);
    };
    // Π in padded coordinates: every rank slice keeps its own parts; a padding gap becomes extra (empty) parts no taller than the tallest real one
    std::vector<int64_t> ppi; // 1-based spl
    if (pi_spl) {
        int64_t umax = 1;
        for (int64_t k = 0; k < K; k++) umax = std::max(umax, rdi(pi_spl, it, k + 1) - rdi(pi_spl, it, k));
        int64_t k = 0;
        for (int r = 0; r < P; r++) {
            const int64_t c0 = D->col_bounds[(size_t)r], c1 = D->col_bounds[(size_t)r + 1];
            while (k <= K && rdi(pi_spl, it, k) - 1 < c0) k++;
            for (; k <= K && rdi(pi_spl, it, k) - 1 <= c1; k++) {
                const int64_t v = r * S + (rdi(pi_spl, it, k) - 1 - c0) + 1;
                if (ppi.empty() || ppi.back() != v) ppi.push_back(v);
            }
            k--; // the boundary c1 is also the first boundary of the next slice
            const int64_t end = (r + 1) * S + 1;
            while (ppi.back() < end) ppi.push_back(std::min(end, ppi.back() + umax));
        }
        if (ppi.empty()) ppi.push_back(1);
    }
    // ---- per-rank slabs: pack on the owning device
    int rc = VBC_OK;
    D->mat.assign((size_t)P, nullptr);
    D->peer.assign((size_t)P, nullptr);
    D->stream.assign((size_t)P, nullptr);
    D->ev0.assign((size_t)P, nullptr);
    D->ev1.assign((size_t)P, nullptr);
    D->graph.assign((size_t)P, nullptr);
    for (int r = 0; r < P && rc == VBC_OK; r++) {
        const int64_t c0 = D->col_bounds[(size_t)r], c1 = D->col_bounds[(size_t)r + 1];
        const int64_t q0 = rdi(colptr, it, c0) - 1, q1 = rdi(colptr, it, c1) - 1;
        const int64_t l0 = D->stripe_bounds[(size_t)r], l1 = D->stripe_bounds[(size_t)r + 1];
        std::vector<char> cp((size_t)((c1 - c0 + 1) * ti)), rv((size_t)std::max<int64_t>((q1 - q0) * ti, 1)), ph((size_t)((l1 - l0 + 1) * ti));
        for (int64_t j = c0; j <= c1; j++) wri(cp.data(), it, j - c0, rdi(colptr, it, j) - q0);
        for (int64_t q = q0; q < q1; q++) wri(rv.data(), it, q - q0, to_padded(rdi(rowval, it, q) - 1) + 1);
        for (int64_t l = l0; l <= l1; l++) wri(ph.data(), it, l - l0, rdi(phi_spl, it, l) - c0);
        std::vector<char> pp;
        if (pi_spl) {
            pp.resize(ppi.size() * (size_t)ti);
            for (size_t k = 0; k < ppi.size(); k++) wri(pp.data(), it, (int64_t)k, ppi[k]);
        }
        rc = vbc_pack_csc(&D->mat[(size_t)r], vt, it, D->padded, c1 - c0, U, W, cp.data(), rv.data(), (const char *)nzval + (size_t)(q0 * tv),
                          pi_spl ? pp.data() : nullptr, pi_spl ? (int64_t)ppi.size() - 1 : 0, ph.data(), l1 - l0, D->dev[(size_t)r]);
        if (rc != VBC_OK) break;
        if (cudaSetDevice(D->dev[(size_t)r]) != cudaSuccess || cudaStreamCreateWithFlags(&D->stream[(size_t)r], cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreate(&D->ev0[(size_t)r]) != cudaSuccess || cudaEventCreate(&D->ev1[(size_t)r]) != cudaSuccess) { set_error("vbc_dist_create: stream / event creation failed"); rc = VBC_ECUDA; break; }
        rc = vbc_set_stream(D->mat[(size_t)r], D->stream[(size_t)r]);
        if (rc == VBC_OK) rc = vbc_peer_create(&D->peer[(size_t)r], vt, D->padded, r, P, D->dev[(size_t)r], nullptr);
    }
    // ---- peer access + buffer table
    if (rc == VBC_OK && P > 1) {
        for (int r = 0; r < P && rc == VBC_OK; r++)
            for (int q = 0; q < P && rc == VBC_OK; q++) {
                if (D->dev[(size_t)r] == D->dev[(size_t)q]) continue;
                int can = 0;
                cudaDeviceCanAccessPeer(&can, D->dev[(size_t)r], D->dev[(size_t)q]);
                if (!can) { set_error("device %d cannot access device %d as a peer", D->dev[(size_t)r], D->dev[(size_t)q]); rc = VBC_ECUDA; break; }
                cudaSetDevice(D->dev[(size_t)r]);
                const cudaError_t e = cudaDeviceEnablePeerAccess(D->dev[(size_t)q], 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { set_error("cudaDeviceEnablePeerAccess(%d -> %d): %s", D->dev[(size_t)r], D->dev[(size_t)q], cudaGetErrorString(e)); rc = VBC_ECUDA; }
                cudaGetLastError();
            }
        std::vector<void *> table((size_t)P * VBC_PEER_HANDLES, nullptr);
        for (int r = 0; r < P && rc == VBC_OK; r++)
            for (int k = 0; k < VBC_PEER_HANDLES && rc == VBC_OK; k++) rc = vbc_peer_buffer(D->peer[(size_t)r], k, &table[(size_t)r * VBC_PEER_HANDLES + k]);
        for (int r = 0; r < P && rc == VBC_OK; r++) rc = vbc_peer_connect_local(D->peer[(size_t)r], table.data());
    }
    // ---- exchange plan: who gathers from which 128-column chunk of whose slice
    if (rc == VBC_OK && P > 1 && exchange == VBC_EXCH_FUSED) {
        const int shift = 7;
        const int64_t C = 1LL << shift, ng = (D->padded + C - 1) / C;
        std::vector<std::vector<unsigned char>> need((size_t)P, std::vector<unsigned char>((size_t)std::max<int64_t>(ng, 1), 0));
        for (int r = 0; r < P && rc == VBC_OK; r++) rc = vbc_read_chunks(D->mat[(size_t)r], shift, need[(size_t)r].data(), ng);
        for (int r = 0; r < P && rc == VBC_OK; r++) {
            const int64_t nloc = D->col_bounds[(size_t)r + 1] - D->col_bounds[(size_t)r], nl = (nloc + C - 1) / C, yoff = r * S;
            std::vector<unsigned char> mask((size_t)std::max<int64_t>(nl, 1), 1);
            for (int64_t c = 0; c < nl; c++) {
                const int64_t lo = (yoff + c * C) >> shift, hi = std::min(yoff + (c + 1) * C - 1, D->padded - 1) >> shift;
                for (int i = 1; i < P; i++) {
                    const int q = (r + i) % P;
                    if (need[(size_t)q][(size_t)lo] | need[(size_t)q][(size_t)hi]) mask[(size_t)c] |= (unsigned char)(1u << i);
                }
            }
            unsigned nbr = 0;
            for (int q = 0; q < P; q++) {
                if (q == r) continue;
                const int64_t g0 = (q * S) >> shift, g1 = std::min((q + 1) * S - 1, D->padded - 1) >> shift;
                const int64_t m0 = (r * S) >> shift, m1 = std::min((r + 1) * S - 1, D->padded - 1) >> shift;
                bool link = false;
                for (int64_t c = g0; c <= g1 && !link; c++) link = need[(size_t)r][(size_t)c] != 0; // r gathers from q's slice
                for (int64_t c = m0; c <= m1 && !link; c++) link = need[(size_t)q][(size_t)c] != 0; // q gathers from r's slice
                if (link) nbr |= 1u << q;
            }
            if (nl > 0) rc = vbc_peer_set_mask(D->peer[(size_t)r], mask.data(), nl, shift);
            if (rc == VBC_OK) rc = vbc_peer_set_neighbors(D->peer[(size_t)r], nbr);
            if (rc == VBC_OK) rc = vbc_peer_auto_interior(D->peer[(size_t)r], D->mat[(size_t)r], yoff, nullptr, nullptr);
        }
    }
    if (rc == VBC_OK && P == 1) rc = vbc_peer_auto_interior(D->peer[0], D->mat[0], 0, nullptr, nullptr);
    // ---- NCCL comparator
    if (rc == VBC_OK && exchange == VBC_EXCH_NCCL) {
        rc = load_nccl(D);
        if (rc == VBC_OK) {
            D->comm.assign((size_t)P, nullptr);
            D->ybuf.assign((size_t)P, nullptr);
            D->cur.assign((size_t)P, 0);
            const int r0 = g_nccl.CommInitAll(D->comm.data(), P, D->dev.data());
            if (r0 != 0) { set_error("ncclCommInitAll(%d devices) -> %s", P, g_nccl.GetErrorString ? g_nccl.GetErrorString(r0) : "NCCL error"); rc = VBC_ENCCL; }
            for (int r = 0; r < P && rc == VBC_OK; r++) {
                cudaSetDevice(D->dev[(size_t)r]);
                if (cudaMalloc(&D->ybuf[(size_t)r], (size_t)(tv * std::max<int64_t>(S, 1))) != cudaSuccess || cudaMemset(D->ybuf[(size_t)r], 0, (size_t)(tv * std::max<int64_t>(S, 1))) != cudaSuccess) { set_error("vbc_dist_create: y slice allocation failed"); rc = VBC_ENOMEM; }
            }
        }
    }
    if (rc != VBC_OK) { vbc_dist_destroy(D); return rc; }
    *out = D;
    return VBC_OK;
}

int vbc_dist_info(const vbc_dist *D, int *ngpus, int64_t *slice_len, int64_t *stripe_bounds, int64_t *cost_per_gpu, int64_t *interior)
{
    if (!D) VBC_FAIL(VBC_EARG, "handle is NULL");
    if (ngpus) *ngpus = D->P;
    if (slice_len) *slice_len = D->S;
    for (int r = 0; r <= D->P && stripe_bounds; r++) stripe_bounds[r] = D->stripe_bounds[(size_t)r];
    for (int r = 0; r < D->P && cost_per_gpu; r++) {
        int64_t s = 0;
        for (int64_t l = D->stripe_bounds[(size_t)r]; l < D->stripe_bounds[(size_t)r + 1]; l++) s += D->cost[(size_t)l];
        cost_per_gpu[r] = s;
    }
    for (int r = 0; r < D->P && interior; r++) {
        int64_t i0 = 0, i1 = 0;
        VBC_TRY(vbc_peer_get_interior(D->peer[(size_t)r], &i0, &i1));
        interior[2 * r] = i0; interior[2 * r + 1] = i1;
    }
    return VBC_OK;
}

static int dist_current(vbc_dist *D, int r, void **buf)
{
    int cur = 0;
    if (D->exchange == VBC_EXCH_NCCL) cur = D->cur[(size_t)r];
    else VBC_TRY(vbc_peer_current(D->peer[(size_t)r], &cur));
    return vbc_peer_buffer(D->peer[(size_t)r], cur, buf);
}

int vbc_dist_set_x(vbc_dist *D, const void *x)
{
    if (!D || (!x && D->n > 0)) VBC_FAIL(VBC_EARG, "NULL argument");
    DeviceGuard guard(D->dev[0]);
    const size_t tv = vt_size(D->vt);
    for (int r = 0; r < D->P; r++) {
        void *buf = nullptr;
        VBC_TRY(dist_current(D, r, &buf));
        VBC_CUDA(cudaSetDevice(D->dev[(size_t)r]));
        VBC_CUDA(cudaStreamSynchronize(D->stream[(size_t)r]));
        VBC_CUDA(cudaMemset(buf, 0, tv * (size_t)std::max<int64_t>(D->padded, 1)));
        for (int q = 0; q < D->P; q++) { // every rank starts with the complete x (it reads any part of it in the first step)
            const int64_t c0 = D->col_bounds[(size_t)q], len = D->col_bounds[(size_t)q + 1] - c0;
            if (len > 0) VBC_CUDA(cudaMemcpy((char *)buf + tv * (size_t)(q * D->S), (const char *)x + tv * (size_t)c0, tv * (size_t)len, cudaMemcpyHostToDevice));
        }
    }
    return VBC_OK;
}

int vbc_dist_gather_x(vbc_dist *D, void *x)
{
    if (!D || (!x && D->n > 0)) VBC_FAIL(VBC_EARG, "NULL argument");
    DeviceGuard guard(D->dev[0]);
    const size_t tv = vt_size(D->vt);
    for (int r = 0; r < D->P; r++) { // every rank's OWN slice is always current
        void *buf = nullptr;
        VBC_TRY(dist_current(D, r, &buf));
        VBC_CUDA(cudaSetDevice(D->dev[(size_t)r]));
        VBC_CUDA(cudaStreamSynchronize(D->stream[(size_t)r]));
        const int64_t c0 = D->col_bounds[(size_t)r], len = D->col_bounds[(size_t)r + 1] - c0;
        if (len > 0) VBC_CUDA(cudaMemcpy((char *)x + tv * (size_t)c0, (const char *)buf + tv * (size_t)(r * D->S), tv * (size_t)len, cudaMemcpyDeviceToHost));
    }
    return VBC_OK;
}

// one iteration of the NCCL comparator on every device: plain multiply into the rank's y slice, then the all-gather
static int nccl_iteration(vbc_dist *D, double alpha)
{
    for (int r = 0; r < D->P; r++) {
        void *xb = nullptr;
        VBC_TRY(vbc_peer_buffer(D->peer[(size_t)r], D->cur[(size_t)r], &xb));
        const int64_t nloc = D->col_bounds[(size_t)r + 1] - D->col_bounds[(size_t)r];
        VBC_TRY(vbc_spmv(D->mat[(size_t)r], 1, alpha, xb, D->padded, 0.0, D->ybuf[(size_t)r], nloc, 1));
    }
    VBC_NCCL(g_nccl.GroupStart());
    for (int r = 0; r < D->P; r++) {
        void *xn = nullptr;
        VBC_TRY(vbc_peer_buffer(D->peer[(size_t)r], 1 - D->cur[(size_t)r], &xn));
        VBC_NCCL(g_nccl.AllGather(D->ybuf[(size_t)r], xn, (size_t)D->S, D->vt == VBC_F64 ? 8 : 7 /* ncclFloat64 : ncclFloat32 */, D->comm[(size_t)r], D->stream[(size_t)r]));
    }
    VBC_NCCL(g_nccl.GroupEnd());
    for (int r = 0; r < D->P; r++) D->cur[(size_t)r] ^= 1;
    return VBC_OK;
}

int vbc_dist_spmv_iter(vbc_dist *D, int iters, double alpha, double *ms_per_iter)
{
    if (!D) VBC_FAIL(VBC_EARG, "handle is NULL");
    if (iters < 0) VBC_FAIL(VBC_EARG, "iters must be >= 0");
    if (ms_per_iter) *ms_per_iter = 0.0;
    if (iters == 0 || D->n == 0) return VBC_OK;
    DeviceGuard guard(D->dev[0]);
    const int P = D->P;
    const bool fused = D->exchange == VBC_EXCH_FUSED;
    // fused: an even number of iterations is captured per device into one graph and replayed (the x buffers alternate, so an
    // even count leaves every pointer where the capture found it); an odd remainder runs as a plain launch
    if (fused && D->lockstep) { // ranks sharing a device: no kernel waits for another one; the host is the barrier
        for (int r = 0; r < P; r++) { VBC_CUDA(cudaSetDevice(D->dev[(size_t)r])); VBC_CUDA(cudaEventRecord(D->ev0[(size_t)r], D->stream[(size_t)r])); }
        for (int t = 0; t < iters; t++) {
            for (int r = 0; r < P; r++) VBC_TRY(vbc_peer_spmv_step(D->peer[(size_t)r], D->mat[(size_t)r], alpha, (int64_t)r * D->S, 0));
            for (int r = 0; r < P; r++) { VBC_CUDA(cudaSetDevice(D->dev[(size_t)r])); VBC_CUDA(cudaStreamSynchronize(D->stream[(size_t)r])); }
        }
        double worst = 0.0;
        for (int r = 0; r < P; r++) {
            VBC_CUDA(cudaSetDevice(D->dev[(size_t)r]));
            VBC_CUDA(cudaEventRecord(D->ev1[(size_t)r], D->stream[(size_t)r]));
            VBC_CUDA(cudaStreamSynchronize(D->stream[(size_t)r]));
            float ms = 0.f;
            VBC_CUDA(cudaEventElapsedTime(&ms, D->ev0[(size_t)r], D->ev1[(size_t)r]));
            worst = std::max(worst, (double)ms);
        }
        if (ms_per_iter) *ms_per_iter = worst / iters;
        return VBC_OK;
    }
    const int git = fused ? (iters & ~1) : 0;
    int cur0 = 0;
    if (fused) VBC_TRY(vbc_peer_current(D->peer[0], &cur0));
    if (git >= 2 && (D->graph_iters != git || D->graph_alpha != alpha || D->graph_cur != cur0)) { // a graph bakes in which x buffer it starts from
        for (int r = 0; r < P; r++)
            if (D->graph[(size_t)r]) { cudaSetDevice(D->dev[(size_t)r]); cudaGraphExecDestroy(D->graph[(size_t)r]); D->graph[(size_t)r] = nullptr; }
        for (int r = 0; r < P; r++) {
            VBC_CUDA(cudaSetDevice(D->dev[(size_t)r]));
            cudaGraph_t g = nullptr;
            VBC_CUDA(cudaStreamBeginCapture(D->stream[(size_t)r], cudaStreamCaptureModeThreadLocal));
            int rc = VBC_OK;
            for (int t = 0; t < git && rc == VBC_OK; t++) rc = vbc_peer_spmv_step(D->peer[(size_t)r], D->mat[(size_t)r], alpha, (int64_t)r * D->S, 3);
            const cudaError_t e = cudaStreamEndCapture(D->stream[(size_t)r], &g);
            if (rc != VBC_OK) { if (g) cudaGraphDestroy(g); return rc; }
            if (e != cudaSuccess) VBC_FAIL(VBC_ECUDA, "vbc_dist_spmv_iter: capture failed: %s", cudaGetErrorString(e));
            const cudaError_t e2 = cudaGraphInstantiate(&D->graph[(size_t)r], g, 0);
            cudaGraphDestroy(g);
            if (e2 != cudaSuccess) VBC_FAIL(VBC_ECUDA, "vbc_dist_spmv_iter: graph instantiation failed: %s", cudaGetErrorString(e2));
        }
        D->graph_iters = git;
        D->graph_alpha = alpha;
        D->graph_cur = cur0;
    }
    for (int r = 0; r < P; r++) {
        VBC_CUDA(cudaSetDevice(D->dev[(size_t)r]));
        VBC_CUDA(cudaEventRecord(D->ev0[(size_t)r], D->stream[(size_t)r]));
    }
    if (fused) {
        if (git >= 2)
            for (int r = 0; r < P; r++) { VBC_CUDA(cudaSetDevice(D->dev[(size_t)r])); VBC_CUDA(cudaGraphLaunch(D->graph[(size_t)r], D->stream[(size_t)r])); }
        for (int t = git; t < iters; t++)
            for (int r = 0; r < P; r++) VBC_TRY(vbc_peer_spmv_step(D->peer[(size_t)r], D->mat[(size_t)r], alpha, (int64_t)r * D->S, 3));
        for (int r = 0; r < P; r++) VBC_TRY(vbc_peer_barrier(D->peer[(size_t)r], D->stream[(size_t)r], 2)); // the final halos have landed
    } else {
        for (int t = 0; t < iters; t++) VBC_TRY(nccl_iteration(D, alpha));
    }
    double worst = 0.0;
    for (int r = 0; r < P; r++) {
        VBC_CUDA(cudaSetDevice(D->dev[(size_t)r]));
        VBC_CUDA(cudaEventRecord(D->ev1[(size_t)r], D->stream[(size_t)r]));
    }
    for (int r = 0; r < P; r++) {
        VBC_CUDA(cudaSetDevice(D->dev[(size_t)r]));
        VBC_CUDA(cudaStreamSynchronize(D->stream[(size_t)r]));
        float ms = 0.f;
        VBC_CUDA(cudaEventElapsedTime(&ms, D->ev0[(size_t)r], D->ev1[(size_t)r]));
        worst = std::max(worst, (double)ms);
        if (fused) {
            int to = 0;
            VBC_TRY(vbc_peer_status(D->peer[(size_t)r], &to));
            if (to) VBC_FAIL(VBC_ECUDA, "vbc_dist_spmv_iter: a flag wait timed out on rank %d", r);
        }
    }
    if (ms_per_iter) *ms_per_iter = worst / iters;
    return VBC_OK;
}

} // extern "C"
