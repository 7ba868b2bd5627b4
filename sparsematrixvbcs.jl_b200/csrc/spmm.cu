// spmm.cu -- VBC sparse matrix x dense matrix (k right-hand sides), north_star item (c).
//
//   Y <- alpha * op(A) * X + beta * Y,   X: cols(op(A)) x k,  Y: rows(op(A)) x k
//
// The reference declares `*(A, B::DenseMatrix)` (multiply_1DVBC.jl:184-185, multiply_VBC.jl:196-197)
// but none of its `mul!` methods accepts matrices, so the call is non-functional there:
// this is new functionality whose oracle is k independent `mul!`s, column by column.
//
// Kernel shape (row-major panels, one row of X / Y = k contiguous values): one warp per stripe, lane t
// owns right-hand sides t, t+32, ... (KT per lane).  Adjoint: a WB x KT register tile of the stripe's
// w x k output accumulates  val[r, dj] * X[row_r, c]  over the stored rows -- the X row is one coalesced
// load per warp, the val row is a warp-uniform (broadcast) load; the tile is stored once.  Forward:
// the stripe's X[j:j+w, c] tile sits in registers and each stored row ends in one coalesced
// red.global.add per right-hand side.  Every val byte is read once per 32*KT right-hand sides, so the
// flop/byte ratio grows with k (k = 32, Float64: 3.3 flop/B) -- DFMA throughput and the L2 gathers of X
// rows bound it, not HBM.  Column-major panels (Julia's layout) are transposed into row-major staging
// buffers on the device and back.
#include "common.cuh"
#include "walk.cuh"

namespace vbc {

// x index of stored row r of a stripe (any lane, no walk state)
template <int MODE>
__device__ __forceinline__ int row_xindex(const int *__restrict__ desc, const int pos0, const int r, const int u0, const int log2u)
{
    if (MODE == DESC_ROWS) return __ldg(desc + pos0 + r);
    if (log2u >= 0) return __ldg(desc + pos0 + (r >> log2u)) + (r & (u0 - 1));
    return __ldg(desc + pos0 + r / u0) + r % u0;
}

// Adjoint SpMM.  One warp per stripe; rows are taken RB at a time in a two-stage software pipeline: while
// batch t is multiplied, batch t+1's X rows (RB independent coalesced loads per lane), its slab of val
// (loaded coalesced, each value once per warp, parked in the other half of a shared-memory double buffer)
// and batch t+2's x indices are already in flight.  The FMAs read val with warp-uniform (broadcast) LDS,
// 128-bit when WB is even.
#ifndef VBC_SPMM_RB
#define VBC_SPMM_RB 4
#endif
template <typename Tv, int MODE, int WB, int KT>
__global__ void __launch_bounds__(256, (WB * KT <= 8 ? 3 : 2)) k_spmm_adj(const StripeMeta *__restrict__ meta, const int *__restrict__ desc,
                                                      const Tv *__restrict__ val, const Tv *__restrict__ X, const long long ldx,
                                                      Tv *__restrict__ Y, const long long ldy, const int L, const int k,
                                                      const int u0, const int log2u, const Tv alpha, const Tv beta)
{
    constexpr int RB = VBC_SPMM_RB;
    constexpr int VPL = (RB * WB + 31) / 32; // staged values per lane and batch
    __shared__ __align__(16) Tv vs_all[8][2][RB * WB];
    Tv(*vs)[RB * WB] = vs_all[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const int nwarps = (int)((gridDim.x * blockDim.x) >> 5);
    for (int l = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5); l < L; l += nwarps) {
        const StripeMeta a = ld_meta(meta + l), b = ld_meta(meta + l + 1);
        const int w = b.col - a.col;
        if (w <= 0) continue;
        const int R = (MODE == DESC_ROWS) ? (b.pos - a.pos) : (int)((b.ofs - a.ofs) / w);
        for (int kb = 0; kb < k; kb += 32 * KT) {
            for (int wb = 0; wb < w; wb += WB) {
                const int wc = min(WB, w - wb);
                Tv acc[WB][KT];
#pragma unroll
                for (int dj = 0; dj < WB; dj++)
#pragma unroll
                    for (int t = 0; t < KT; t++) acc[dj][t] = (Tv)0;

                // issue the loads of one batch: X rows into xn, val slab into vreg; the x indices were loaded a batch earlier
                Tv xn[RB][KT], vreg[VPL];
                int xi_next = lane < min(RB, R) ? row_xindex<MODE>(desc, a.pos, lane, u0, log2u) : 0;
                auto issue = [&](const int r) {
                    const int nr = min(RB, R - r);
                    const int myxi = xi_next;
                    xi_next = (lane < RB && r + RB + lane < R) ? row_xindex<MODE>(desc, a.pos, r + RB + lane, u0, log2u) : 0;
#pragma unroll
                    for (int j = 0; j < RB; j++) {
                        const int xi = __shfl_sync(0xffffffffu, myxi, j);
#pragma unroll
                        for (int t = 0; t < KT; t++) {
                            const int c = kb + t * 32 + lane;
                            xn[j][t] = (j < nr && c < k) ? __ldg(X + (long long)xi * ldx + c) : (Tv)0;
                        }
                    }
#pragma unroll
                    for (int q = 0; q < VPL; q++) {
                        const int i = lane + 32 * q, j = i / WB, dj = i % WB;
                        vreg[q] = (i < RB * WB && j < nr && dj < wc) ? __ldcs(val + a.ofs + (long long)(r + j) * w + wb + dj) : (Tv)0;
                    }
                };
                int buf = 0;
                if (R > 0) issue(0);
                for (int r = 0; r < R; r += RB) {
                    // park batch r's values; take over its X rows
                    Tv xc[RB][KT];
#pragma unroll
                    for (int j = 0; j < RB; j++)
#pragma unroll
                        for (int t = 0; t < KT; t++) xc[j][t] = xn[j][t];
#pragma unroll
                    for (int q = 0; q < VPL; q++)
                        if (lane + 32 * q < RB * WB) vs[buf][lane + 32 * q] = vreg[q];
                    if (r + RB < R) issue(r + RB); // next batch's loads fly while this one is multiplied
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < RB; j++) {
                        if constexpr (WB % 2 == 0 && sizeof(Tv) == 8) {
#pragma unroll
                            for (int dj = 0; dj < WB; dj += 2) {
                                const double2 v2 = *reinterpret_cast<const double2 *>(&vs[buf][j * WB + dj]); // warp-uniform LDS.128
#pragma unroll
                                for (int t = 0; t < KT; t++) {
                                    acc[dj][t] = fma((Tv)v2.x, xc[j][t], acc[dj][t]);
                                    acc[dj + 1][t] = fma((Tv)v2.y, xc[j][t], acc[dj + 1][t]);
                                }
                            }
                        } else {
#pragma unroll
                            for (int dj = 0; dj < WB; dj++) {
                                const Tv v = vs[buf][j * WB + dj];
#pragma unroll
                                for (int t = 0; t < KT; t++) acc[dj][t] = fma(v, xc[j][t], acc[dj][t]);
                            }
                        }
                    }
                    buf ^= 1; // the other half is free: it was read two batches ago, and every lane has passed a __syncwarp since
                }
                __syncwarp();
#pragma unroll
                for (int dj = 0; dj < WB; dj++) {
                    if (dj < wc) {
                        Tv *yp = Y + (long long)(a.col + wb + dj) * ldy;
#pragma unroll
                        for (int t = 0; t < KT; t++) {
                            const int c = kb + t * 32 + lane;
                            if (c < k) yp[c] = (beta == (Tv)0) ? alpha * acc[dj][t] : alpha * acc[dj][t] + beta * yp[c];
                        }
                    }
                }
            }
        }
    }
}

// Adjoint SpMM on the FP64 tensor path (DMMA).  For one stripe, Y[j:j+w, :] = V' * Xg with V the stripe's R x w
// slab and Xg the R gathered rows of X: a real contraction over the stripe's rows, so it maps onto
// mma.sync.m8n8k4.f64 tiles -- A = 8 stripe columns x 4 stripe rows (read straight from val, one value per
// lane), B = 4 gathered X rows x 8 right-hand sides, C = 8 x 8 accumulators (2 per lane and N-tile).  One
// warp per stripe, 4 N-tiles (32 right-hand sides) per pass: per 4 stripe rows a lane issues 1 val load,
// 1 index load, 4 X loads and 4 DMMAs -- about 5 instructions per stripe row instead of ~45 in the SIMT
// kernel, which is what bounds that one (issue slots, not HBM or DFMA rate).  No shared memory.
// (On Blackwell FP64 tensor math still goes through the mma.sync-class DMMA path; tcgen05 has no FP64 kind.)
#ifndef VBC_DMMA_MINB
#define VBC_DMMA_MINB 4
#endif
#ifndef VBC_DMMA_KS
#define VBC_DMMA_KS 2
#endif
#ifndef VBC_DMMA_VEC_DEFAULT
#define VBC_DMMA_VEC_DEFAULT 0 // 1: the 256-bit X-row loads are the default whenever the panels are aligned
#endif
__device__ __forceinline__ void dmma_m8n8k4(double (&c)[2], const double a, const double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// VEC: the right-hand sides are permuted over the N-tiles -- fragment column n of tile nt is right-hand side
// kb + 4 n + nt -- so the four B values of a lane are 32 contiguous bytes of one X row (one 256-bit load, the
// eight lanes of a k-row cover the whole 256-byte row) and its C values are two runs of 4 contiguous Y entries.
// Needs k % 4 == 0 and 32-byte aligned rows (checked at launch).
__device__ __forceinline__ void ld_row4(const double *p, double (&v)[4])
{
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
}

template <int MODE, bool VEC>
__global__ void __launch_bounds__(256, VBC_DMMA_MINB) k_spmm_adj_dmma(const StripeMeta *__restrict__ meta, const int *__restrict__ desc,
                                                        const double *__restrict__ val, const double *__restrict__ X, const long long ldx,
                                                        double *__restrict__ Y, const long long ldy, const int L, const int k,
                                                        const int u0, const int log2u, const double alpha, const double beta)
{
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int nwarps = (int)((gridDim.x * blockDim.x) >> 5);
    for (int l = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5); l < L; l += nwarps) {
        const StripeMeta a = ld_meta(meta + l), b = ld_meta(meta + l + 1);
        const int w = b.col - a.col;
        if (w <= 0) continue;
        const int R = (MODE == DESC_ROWS) ? (b.pos - a.pos) : (int)((b.ofs - a.ofs) / w);
        for (int kb = 0; kb < k; kb += 32) {
            for (int wb = 0; wb < w; wb += 8) {
                double c[4][2];
#pragma unroll
                for (int nt = 0; nt < 4; nt++) { c[nt][0] = 0.0; c[nt][1] = 0.0; }
                const bool arow = wb + g < w; // this lane's A row (stripe column) exists
                const double *vp = val + a.ofs + (long long)t * w + wb + g;
                int xi_next = t < R ? row_xindex<MODE>(desc, a.pos, t, u0, log2u) : -1;
                constexpr int KS = VBC_DMMA_KS; // k-steps (of 4 stripe rows) per iteration: all their loads are issued first
                for (int r = 0; r < R; r += 4 * KS) {
                    int xi[KS];
                    xi[0] = xi_next;
#pragma unroll
                    for (int q = 1; q < KS; q++) xi[q] = (r + 4 * q + t < R) ? row_xindex<MODE>(desc, a.pos, r + 4 * q + t, u0, log2u) : -1;
                    xi_next = (r + 4 * KS + t < R) ? row_xindex<MODE>(desc, a.pos, r + 4 * KS + t, u0, log2u) : -1;
                    double av[KS], bv[KS][4];
#pragma unroll
                    for (int q = 0; q < KS; q++) av[q] = (arow && xi[q] >= 0) ? __ldcs(vp + 4 * q * (long long)w) : 0.0;
                    vp += 4 * KS * (long long)w;
#pragma unroll
                    for (int q = 0; q < KS; q++) {
                        if constexpr (VEC) {
                            const int col = kb + 4 * g;
                            if (xi[q] >= 0 && col < k) ld_row4(X + (long long)xi[q] * ldx + col, bv[q]);
                            else { bv[q][0] = 0.0; bv[q][1] = 0.0; bv[q][2] = 0.0; bv[q][3] = 0.0; }
                        } else {
#pragma unroll
                            for (int nt = 0; nt < 4; nt++) {
                                const int col = kb + nt * 8 + g;
                                bv[q][nt] = (xi[q] >= 0 && col < k) ? __ldg(X + (long long)xi[q] * ldx + col) : 0.0;
                            }
                        }
                    }
#pragma unroll
                    for (int q = 0; q < KS; q++)
#pragma unroll
                        for (int nt = 0; nt < 4; nt++) dmma_m8n8k4(c[nt], av[q], bv[q][nt]);
                }
                if (arow) {
                    double *yp = Y + (long long)(a.col + wb + g) * ldy;
                    if constexpr (VEC) {
#pragma unroll
                        for (int e = 0; e < 2; e++) {
                            const int col = kb + 4 * (2 * t + e);
                            if (col < k) {
                                double2 *y2 = reinterpret_cast<double2 *>(yp + col);
                                double2 lo = make_double2(alpha * c[0][e], alpha * c[1][e]), hi = make_double2(alpha * c[2][e], alpha * c[3][e]);
                                if (beta != 0.0) {
                                    const double2 o0 = y2[0], o1 = y2[1];
                                    lo.x += beta * o0.x; lo.y += beta * o0.y; hi.x += beta * o1.x; hi.y += beta * o1.y;
                                }
                                y2[0] = lo; y2[1] = hi;
                            }
                        }
                    } else {
#pragma unroll
                        for (int nt = 0; nt < 4; nt++) {
#pragma unroll
                            for (int e = 0; e < 2; e++) {
                                const int col = kb + nt * 8 + 2 * t + e;
                                if (col < k) yp[col] = (beta == 0.0) ? alpha * c[nt][e] : alpha * c[nt][e] + beta * yp[col];
                            }
                        }
                    }
                }
            }
        }
    }
}

// Adjoint SpMM, DMMA tiles fed through shared memory by the bulk-copy engine (TMA).
//
// ncu on the two kernels above (profiles/r01_spmm_ncu.md): both sit at ~90 % of the L1 data pipe
// (l1tex__data_pipe_lsu_wavefronts) with DRAM and L2 far below their peaks.  The m8n8k4 B fragment puts FOUR
// DIFFERENT gathered X rows in adjacent lanes, so every 32-byte sector of a register load is its own
// data-pipe wavefront (one sector per clock and SM), hit or miss, scalar or 256-bit.  Here the X rows never
// pass through the load/store data pipe: each stored row is one `cp.async.bulk` (global -> shared, 256 bytes,
// completion counted on an mbarrier), the stripe's slab of val is one more bulk copy per chunk, and the
// fragments are read back with conflict-free LDS.64 (rows padded to 288 bytes).  One warp = one private
// ring of ST stages x CH rows; the ring keeps running across stripe boundaries, row indices and stripe
// meta are fetched one step ahead, so nothing in the loop waits on a global load.
#ifndef VBC_TMA_CH
#define VBC_TMA_CH 16     // rows per stage (multiple of 4, at most 32)
#endif
#ifndef VBC_TMA_ST
#define VBC_TMA_ST 2      // stages per warp
#endif
#ifndef VBC_TMA_MINB
#define VBC_TMA_MINB 2    // CTAs per SM the register allocation is bounded for (the ring: 8 warps x ST x (CH x 352) bytes per CTA)
#endif
#define VBC_TMA_ROWB 288  // bytes between staged rows: 256 + 32, i.e. 8 banks of shift per row -> the 4 rows x 8 doubles of a half-warp hit 32 distinct banks
#define VBC_TMA_STAGE (VBC_TMA_CH * VBC_TMA_ROWB + VBC_TMA_CH * 8 * 8)
#define VBC_TMA_SMEM (8 * VBC_TMA_ST * VBC_TMA_STAGE + 8 * VBC_TMA_ST * 8)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(const uint32_t bar, const uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(const uint32_t bar, const uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(const uint32_t dst, const void *src, const uint32_t bytes, const uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(const uint32_t bar, const uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}

template <int MODE>
__global__ void __launch_bounds__(256, VBC_TMA_MINB) k_spmm_adj_tma(const StripeMeta *__restrict__ meta, const int *__restrict__ desc,
                                                       const double *__restrict__ val, const double *__restrict__ X, const long long ldx,
                                                       double *__restrict__ Y, const long long ldy, const int L, const int k,
                                                       const int u0, const int log2u, const double alpha, const double beta)
{
    constexpr int CH = VBC_TMA_CH, ST = VBC_TMA_ST, ROWB = VBC_TMA_ROWB, STAGE = VBC_TMA_STAGE;
    extern __shared__ __align__(128) unsigned char tma_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    unsigned char *ring = tma_smem + (size_t)warp * (ST * STAGE);
    const uint32_t ring_u32 = smem_u32(ring);
    const uint32_t bar_u32 = smem_u32(tma_smem + 8 * ST * STAGE) + (uint32_t)warp * ST * 8;
    if (lane == 0)
        for (int s = 0; s < ST; s++) mbar_init(bar_u32 + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    const int kb = (int)blockIdx.y * 32;
    const int kc = min(32, k - kb);
    const uint32_t rowbytes = (uint32_t)kc * 8u;
    const double *Xk = X + kb;
    const int nwarps = (int)((gridDim.x * blockDim.x) >> 5);
    const int l0 = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);

    // ---- producer: a cursor (stripe pl, row pr) over this warp's stripes, one chunk of up to CH rows per step
    int pl = l0, pr = 0, pR = 0, pw = 0, xi_pref = 0, issued = 0;
    StripeMeta pa{}, na{}, nb{};
    auto rows_of = [&](const StripeMeta &a, const StripeMeta &b, int &w) {
        w = b.col - a.col;
        return (w > 0) ? ((MODE == DESC_ROWS) ? (b.pos - a.pos) : (int)((b.ofs - a.ofs) / w)) : 0;
    };
    auto fetch_meta = [&](const int l) { // the meta of stripe l rides ahead of its use
        if (l < L) { na = ld_meta(meta + l); nb = ld_meta(meta + l + 1); }
    };
    auto settle = [&]() { // pl points at a stripe whose meta is in (na, nb): skip stripes without rows, then prefetch indices
        while (pl < L) {
            pa = na;
            pR = rows_of(na, nb, pw);
            fetch_meta(pl + nwarps);
            if (pR > 0) break;
            pl += nwarps;
        }
        pr = 0;
        xi_pref = (pl < L && lane < min(CH, pR)) ? row_xindex<MODE>(desc, pa.pos, lane, u0, log2u) : 0;
    };
    fetch_meta(pl);
    settle();
    auto issue = [&]() {
        if (pl >= L) return;
        const int s = issued % ST;
        const int nr = min(CH, pR - pr);
        const uint32_t bar = bar_u32 + 8u * s, dst = ring_u32 + (uint32_t)(s * STAGE);
        const bool val_bulk = ((pa.ofs | (long long)pw) & 1) == 0; // 16-byte aligned slab chunks
        const int xi = xi_pref;
        if (lane == 0) mbar_arrive_expect_tx(bar, (uint32_t)nr * rowbytes + (val_bulk ? (uint32_t)(nr * pw) * 8u : 0u));
        __syncwarp();
        if (lane < nr) bulk_g2s(dst + (uint32_t)(lane * ROWB), Xk + (long long)xi * ldx, rowbytes, bar);
        if (lane == 0 && val_bulk) bulk_g2s(dst + (uint32_t)(CH * ROWB), val + pa.ofs + (long long)pr * pw, (uint32_t)(nr * pw) * 8u, bar);
        issued++;
        pr += CH;
        if (pr >= pR) { pl += nwarps; settle(); }
        else xi_pref = (lane < min(CH, pR - pr)) ? row_xindex<MODE>(desc, pa.pos, pr + lane, u0, log2u) : 0;
    };
    for (int i = 0; i < ST; i++) issue();

    // ---- consumer: the same stripes in the same order
    int consumed = 0;
    for (int l = l0; l < L; l += nwarps) {
        const StripeMeta a = ld_meta(meta + l), b = ld_meta(meta + l + 1);
        int w;
        const int R = rows_of(a, b, w);
        if (w <= 0) continue;
        const bool val_bulk = ((a.ofs | (long long)w) & 1) == 0;
        const bool arow = g < w; // this lane's A row (stripe column) exists
        double c[4][2];
#pragma unroll
        for (int nt = 0; nt < 4; nt++) { c[nt][0] = 0.0; c[nt][1] = 0.0; }
        for (int r = 0; r < R; r += CH) {
            const int s = consumed % ST;
            const uint32_t parity = (uint32_t)((consumed / ST) & 1);
            const unsigned char *sp = ring + s * STAGE;
            double av[CH / 4];
            if (!val_bulk) { // odd slab start or odd width: the A values come straight from global
#pragma unroll
                for (int q = 0; q < CH / 4; q++)
                    av[q] = (arow && r + 4 * q + t < R) ? __ldcs(val + a.ofs + (long long)(r + 4 * q + t) * w + g) : 0.0;
            }
            {
                unsigned spins = 0;
                while (!mbar_try_wait(bar_u32 + 8u * s, parity))
                    if (++spins > (1u << 22)) __trap(); // a lost copy must not hang the GPU
            }
            if (val_bulk) {
                const double *vs = reinterpret_cast<const double *>(sp + CH * ROWB);
#pragma unroll
                for (int q = 0; q < CH / 4; q++) av[q] = (arow && r + 4 * q + t < R) ? vs[(4 * q + t) * w + g] : 0.0;
            }
#pragma unroll
            for (int q = 0; q < CH / 4; q++) {
                const bool rok = r + 4 * q + t < R;
                const unsigned char *rp = sp + (4 * q + t) * ROWB + g * 8;
#pragma unroll
                for (int nt = 0; nt < 4; nt++) {
                    const double bv = (rok && nt * 8 + g < kc) ? *reinterpret_cast<const double *>(rp + nt * 64) : 0.0;
                    dmma_m8n8k4(c[nt], av[q], bv);
                }
            }
            __syncwarp();
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // the stage is handed back to the copy engine
            consumed++;
            issue();
        }
        if (arow) {
            double *yp = Y + (long long)(a.col + g) * ldy;
#pragma unroll
            for (int nt = 0; nt < 4; nt++) {
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const int col = kb + nt * 8 + 2 * t + e;
                    if (col < k) yp[col] = (beta == 0.0) ? alpha * c[nt][e] : alpha * c[nt][e] + beta * yp[col];
                }
            }
        }
    }
}

// Same ring, filled with LDGSTS (cp.async, 16 bytes per lane) instead of bulk copies.  NOT YET RUN ON A GPU
// (written after the round's GPU budget was spent; option 5, never selected automatically).  Why it should beat
// the bulk-copy kernel: ncu shows that kernel spending a third of its instructions and most of its "wait" stalls
// in the compiler's per-lane ELECT / R2UR / UBLKCP loop (one 256-byte copy per ~10 instructions, operands in
// uniform registers).  Here 16 adjacent lanes copy one X row, 32 lanes two rows per instruction, with ordinary
// per-lane addresses: 8 instructions per 16-row chunk, and the natural lane order keeps the data pipe at one
// wavefront per 128 bytes.  Completion is counted with cp.async groups (one group per chunk, always committed).
__device__ __forceinline__ void cp_async16(const uint32_t dst, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(256, VBC_TMA_MINB) k_spmm_adj_cpasync(const StripeMeta *__restrict__ meta, const int *__restrict__ desc,
                                                           const double *__restrict__ val, const double *__restrict__ X, const long long ldx,
                                                           double *__restrict__ Y, const long long ldy, const int L, const int k,
                                                           const int u0, const int log2u, const double alpha, const double beta)
{
    constexpr int CH = VBC_TMA_CH, ST = VBC_TMA_ST, ROWB = VBC_TMA_ROWB, STAGE = VBC_TMA_STAGE;
    extern __shared__ __align__(128) unsigned char tma_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    unsigned char *ring = tma_smem + (size_t)warp * (ST * STAGE);
    const uint32_t ring_u32 = smem_u32(ring);
    const int kb = (int)blockIdx.y * 32;
    const int kc = min(32, k - kb);
    const int ppr = kc >> 1; // 16-byte pieces per staged row
    const double *Xk = X + kb;
    const int nwarps = (int)((gridDim.x * blockDim.x) >> 5);
    const int l0 = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);

    int pl = l0, pr = 0, pR = 0, pw = 0, xi_pref = 0, issued = 0;
    StripeMeta pa{}, na{}, nb{};
    auto rows_of = [&](const StripeMeta &a, const StripeMeta &b, int &w) {
        w = b.col - a.col;
        return (w > 0) ? ((MODE == DESC_ROWS) ? (b.pos - a.pos) : (int)((b.ofs - a.ofs) / w)) : 0;
    };
    auto fetch_meta = [&](const int l) {
        if (l < L) { na = ld_meta(meta + l); nb = ld_meta(meta + l + 1); }
    };
    auto settle = [&]() {
        while (pl < L) {
            pa = na;
            pR = rows_of(na, nb, pw);
            fetch_meta(pl + nwarps);
            if (pR > 0) break;
            pl += nwarps;
        }
        pr = 0;
        xi_pref = (pl < L && lane < min(CH, pR)) ? row_xindex<MODE>(desc, pa.pos, lane, u0, log2u) : 0;
    };
    fetch_meta(pl);
    settle();
    auto issue = [&]() {
        if (pl < L) {
            const int s = issued % ST;
            const int nr = min(CH, pR - pr);
            const uint32_t dst = ring_u32 + (uint32_t)(s * STAGE);
            const int npieces = nr * ppr;
            for (int i0 = 0; i0 < npieces; i0 += 32) { // uniform trip count: the shuffle below needs the whole warp
                const int i = i0 + lane;
                const int row = (ppr == 16) ? (i >> 4) : (i / ppr);
                const int piece = i - row * ppr;
                const int xi = __shfl_sync(0xffffffffu, xi_pref, min(row, nr - 1));
                if (i < npieces) cp_async16(dst + (uint32_t)(row * ROWB + piece * 16), Xk + (long long)xi * ldx + 2 * piece);
            }
            if (((pa.ofs | (long long)pw) & 1) == 0) { // the chunk of the stripe's slab, when 16-byte aligned
                const double *src = val + pa.ofs + (long long)pr * pw;
                for (int i = lane; i < (nr * pw) >> 1; i += 32) cp_async16(dst + (uint32_t)(CH * ROWB + i * 16), src + 2 * i);
            }
            issued++;
            pr += CH;
            if (pr >= pR) { pl += nwarps; settle(); }
            else xi_pref = (lane < min(CH, pR - pr)) ? row_xindex<MODE>(desc, pa.pos, pr + lane, u0, log2u) : 0;
        }
        asm volatile("cp.async.commit_group;" ::: "memory"); // one group per call, empty at the tail, so the wait below can count
    };
    for (int i = 0; i < ST; i++) issue();

    int consumed = 0;
    for (int l = l0; l < L; l += nwarps) {
        const StripeMeta a = ld_meta(meta + l), b = ld_meta(meta + l + 1);
        int w;
        const int R = rows_of(a, b, w);
        if (w <= 0) continue;
        const bool val_staged = ((a.ofs | (long long)w) & 1) == 0;
        const bool arow = g < w;
        double c[4][2];
#pragma unroll
        for (int nt = 0; nt < 4; nt++) { c[nt][0] = 0.0; c[nt][1] = 0.0; }
        for (int r = 0; r < R; r += CH) {
            const int s = consumed % ST;
            const unsigned char *sp = ring + s * STAGE;
            double av[CH / 4];
            if (!val_staged) {
#pragma unroll
                for (int q = 0; q < CH / 4; q++)
                    av[q] = (arow && r + 4 * q + t < R) ? __ldcs(val + a.ofs + (long long)(r + 4 * q + t) * w + g) : 0.0;
            }
            asm volatile("cp.async.wait_group %0;" ::"n"(ST - 1) : "memory"); // this lane's copies of the oldest chunk have landed
            __syncwarp();                                                     // ... and so have every other lane's
            if (val_staged) {
                const double *vs = reinterpret_cast<const double *>(sp + CH * ROWB);
#pragma unroll
                for (int q = 0; q < CH / 4; q++) av[q] = (arow && r + 4 * q + t < R) ? vs[(4 * q + t) * w + g] : 0.0;
            }
#pragma unroll
            for (int q = 0; q < CH / 4; q++) {
                const bool rok = r + 4 * q + t < R;
                const unsigned char *rp = sp + (4 * q + t) * ROWB + g * 8;
#pragma unroll
                for (int nt = 0; nt < 4; nt++) {
                    const double bv = (rok && nt * 8 + g < kc) ? *reinterpret_cast<const double *>(rp + nt * 64) : 0.0;
                    dmma_m8n8k4(c[nt], av[q], bv);
                }
            }
            __syncwarp(); // every lane is done reading the stage before it is refilled
            consumed++;
            issue();
        }
        if (arow) {
            double *yp = Y + (long long)(a.col + g) * ldy;
#pragma unroll
            for (int nt = 0; nt < 4; nt++) {
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const int col = kb + nt * 8 + 2 * t + e;
                    if (col < k) yp[col] = (beta == 0.0) ? alpha * c[nt][e] : alpha * c[nt][e] + beta * yp[col];
                }
            }
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

template <typename Tv, int MODE, int WB, int KT>
__global__ void __launch_bounds__(256) k_spmm_fwd(const StripeMeta *__restrict__ meta, const int *__restrict__ desc,
                                                   const Tv *__restrict__ val, const Tv *__restrict__ X, const long long ldx,
                                                   Tv *__restrict__ Y, const long long ldy, const int L, const int k,
                                                   const int u0, const int log2u, const Tv alpha)
{
    const int lane = threadIdx.x & 31;
    const int nwarps = (int)((gridDim.x * blockDim.x) >> 5);
    for (int l = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5); l < L; l += nwarps) {
        const StripeMeta a = ld_meta(meta + l), b = ld_meta(meta + l + 1);
        const int w = b.col - a.col;
        if (w <= 0 || b.ofs == a.ofs) continue;
        const int R = (MODE == DESC_ROWS) ? (b.pos - a.pos) : (int)((b.ofs - a.ofs) / w);
        for (int kb = 0; kb < k; kb += 32 * KT) {
            for (int wb = 0; wb < w; wb += WB) {
                Tv xs[WB][KT]; // X[j + wb + dj, c]
#pragma unroll
                for (int dj = 0; dj < WB; dj++)
#pragma unroll
                    for (int t = 0; t < KT; t++) {
                        const int c = kb + t * 32 + lane;
                        xs[dj][t] = (wb + dj < w && c < k) ? __ldg(X + (long long)(a.col + wb + dj) * ldx + c) : (Tv)0;
                    }
                RowWalk<MODE> walk;
                walk.init(desc, a.pos, 0, 1, u0, log2u);
                const Tv *vp = val + a.ofs + wb;
                for (int r = 0; r < R; r++) {
                    const int xi = walk.next();
                    Tv p[KT];
#pragma unroll
                    for (int t = 0; t < KT; t++) p[t] = (Tv)0;
#pragma unroll
                    for (int dj = 0; dj < WB; dj++) {
                        const Tv v = (wb + dj < w) ? __ldg(vp + dj) : (Tv)0;
#pragma unroll
                        for (int t = 0; t < KT; t++) p[t] = fma(v, xs[dj][t], p[t]);
                    }
                    vp += w;
#pragma unroll
                    for (int t = 0; t < KT; t++) {
                        const int c = kb + t * 32 + lane;
                        if (c < k) atomicAdd(Y + (long long)xi * ldy + c, alpha * p[t]);
                    }
                }
            }
        }
    }
}

// Y[r, 0:k) <- beta * Y[r, 0:k)
template <typename Tv>
__global__ void __launch_bounds__(256) k_scale_panel(Tv *__restrict__ Y, const long long ldy, const long long rows, const int k, const Tv beta)
{
    const long long total = rows * k;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        Tv *p = Y + (i / k) * ldy + (i % k);
        *p = (beta == (Tv)0) ? (Tv)0 : beta * *p;
    }
}

// out[c * ldo + r] = in[r * ldi + c]   (rows x cols input, tiled through shared memory)
template <typename Tv>
__global__ void __launch_bounds__(256) k_transpose(const Tv *__restrict__ in, const long long ldi, const long long rows, const long long cols,
                                                    Tv *__restrict__ out, const long long ldo)
{
    __shared__ Tv tile[32][33];
    const long long r0 = (long long)blockIdx.x * 32, c0 = (long long)blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5; // 32 x 8
    for (int i = ty; i < 32; i += 8)
        if (r0 + i < rows && c0 + tx < cols) tile[i][tx] = in[(r0 + i) * ldi + c0 + tx];
    __syncthreads();
    for (int i = ty; i < 32; i += 8)
        if (c0 + i < cols && r0 + tx < rows) out[(c0 + i) * ldo + r0 + tx] = tile[tx][i];
}

template <typename Tv, int MODE>
static int launch_spmm_mode(vbc_mat *A, int trans, int k, Tv alpha, const Tv *X, long long ldx, Tv beta, Tv *Y, long long ldy)
{
    const int L = (int)A->L;
    int log2u = -1;
    if (A->u0 > 0 && !(A->u0 & (A->u0 - 1))) { log2u = 0; while ((1 << log2u) < A->u0) log2u++; }
    int64_t grid = (int64_t)A->sm_count * 6;
    const int64_t need = ((int64_t)L * 32 + 255) / 256;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    const bool wide = A->W > 8;
    const bool k2 = k > 32;
    const bool narrow = A->W <= 4;
#define SPMM_LAUNCH(KERNEL, WBv, KTv, ...) KERNEL<Tv, MODE, WBv, KTv><<<(unsigned)grid, 256, 0, A->stream>>>(__VA_ARGS__)
    if (trans) {
        if (L == 0) return VBC_OK;
        if constexpr (sizeof(Tv) == 8) {
            if (A->opt_spmm_simt != 1) { // Float64: tensor (DMMA) tiles
                int64_t g2 = (int64_t)A->sm_count * 8;
                if (g2 > need) g2 = need;
                if (g2 < 1) g2 = 1;
                // tiles fed through shared memory (4: bulk copies, 5: cp.async -- experimental): stripes at most 8 wide, 16-byte aligned X rows
                if ((A->opt_spmm_simt == 4 || A->opt_spmm_simt == 5) && A->W <= 8 && (k % 2) == 0 && (ldx % 2) == 0 && ((uintptr_t)X % 16) == 0) {
                    const bool bulk = A->opt_spmm_simt == 4;
                    static bool attr_set[2][2][64] = {}; // per kernel instance and device; first call, outside any stream capture
                    if (!attr_set[bulk][MODE][A->device & 63]) {
                        if (bulk) VBC_CUDA(cudaFuncSetAttribute(k_spmm_adj_tma<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, VBC_TMA_SMEM));
                        else VBC_CUDA(cudaFuncSetAttribute(k_spmm_adj_cpasync<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, VBC_TMA_SMEM));
                        attr_set[bulk][MODE][A->device & 63] = true;
                    }
                    int64_t g3 = (int64_t)A->sm_count * VBC_TMA_MINB;
                    if (g3 > need) g3 = need;
                    if (g3 < 1) g3 = 1;
                    dim3 grid3((unsigned)g3, (unsigned)((k + 31) / 32));
                    if (bulk)
                        k_spmm_adj_tma<MODE><<<grid3, 256, VBC_TMA_SMEM, A->stream>>>(A->d_meta, A->d_desc, (const double *)A->d_val, (const double *)X, ldx,
                                                                                     (double *)Y, ldy, L, k, A->u0, log2u, (double)alpha, (double)beta);
                    else
                        k_spmm_adj_cpasync<MODE><<<grid3, 256, VBC_TMA_SMEM, A->stream>>>(A->d_meta, A->d_desc, (const double *)A->d_val, (const double *)X, ldx,
                                                                                         (double *)Y, ldy, L, k, A->u0, log2u, (double)alpha, (double)beta);
                    A->launches++;
                    VBC_CUDA(cudaGetLastError());
                    return VBC_OK;
                }
                // 256-bit X-row loads when every row of both panels is 32-byte aligned (opt_spmm_simt: 2 forces the scalar loads, 3 the vector loads)
                const bool aligned = (k % 4) == 0 && (ldx % 4) == 0 && (ldy % 4) == 0 && ((uintptr_t)X % 32) == 0 && ((uintptr_t)Y % 32) == 0;
                const bool vec = aligned && (A->opt_spmm_simt == 3 || (A->opt_spmm_simt == 0 && VBC_DMMA_VEC_DEFAULT));
                if (vec)
                    k_spmm_adj_dmma<MODE, true><<<(unsigned)g2, 256, 0, A->stream>>>(A->d_meta, A->d_desc, (const double *)A->d_val, (const double *)X, ldx,
                                                                                 (double *)Y, ldy, L, k, A->u0, log2u, (double)alpha, (double)beta);
                else
                    k_spmm_adj_dmma<MODE, false><<<(unsigned)g2, 256, 0, A->stream>>>(A->d_meta, A->d_desc, (const double *)A->d_val, (const double *)X, ldx,
                                                                                  (double *)Y, ldy, L, k, A->u0, log2u, (double)alpha, (double)beta);
                A->launches++;
                VBC_CUDA(cudaGetLastError());
                return VBC_OK;
            }
        }
        if (wide) { if (k2) SPMM_LAUNCH(k_spmm_adj, 16, 2, A->d_meta, A->d_desc, (const Tv *)A->d_val, X, ldx, Y, ldy, L, k, A->u0, log2u, alpha, beta);
                    else    SPMM_LAUNCH(k_spmm_adj, 16, 1, A->d_meta, A->d_desc, (const Tv *)A->d_val, X, ldx, Y, ldy, L, k, A->u0, log2u, alpha, beta); }
        else if (narrow) { if (k2) SPMM_LAUNCH(k_spmm_adj, 4, 2, A->d_meta, A->d_desc, (const Tv *)A->d_val, X, ldx, Y, ldy, L, k, A->u0, log2u, alpha, beta);
                    else    SPMM_LAUNCH(k_spmm_adj, 4, 1, A->d_meta, A->d_desc, (const Tv *)A->d_val, X, ldx, Y, ldy, L, k, A->u0, log2u, alpha, beta); }
        else      { if (k2) SPMM_LAUNCH(k_spmm_adj, 8, 2, A->d_meta, A->d_desc, (const Tv *)A->d_val, X, ldx, Y, ldy, L, k, A->u0, log2u, alpha, beta);
                    else    SPMM_LAUNCH(k_spmm_adj, 8, 1, A->d_meta, A->d_desc, (const Tv *)A->d_val, X, ldx, Y, ldy, L, k, A->u0, log2u, alpha, beta); }
        A->launches++;
    } else {
        if (A->m > 0 && beta != (Tv)1) {
            int64_t g = (A->m * (int64_t)k + 255) / 256;
            if (g > (int64_t)A->sm_count * 8) g = (int64_t)A->sm_count * 8;
            k_scale_panel<Tv><<<(unsigned)g, 256, 0, A->stream>>>(Y, ldy, A->m, k, beta);
            A->launches++;
        }
        if (L == 0 || A->nval == 0) return VBC_OK;
        if (wide) { if (k2) SPMM_LAUNCH(k_spmm_fwd, 16, 2, A->d_meta, A->d_desc, (const Tv *)A->d_val, X, ldx, Y, ldy, L, k, A->u0, log2u, alpha);
                    else    SPMM_LAUNCH(k_spmm_fwd, 16, 1, A->d_meta, A->d_desc, (const Tv *)A->d_val, X, ldx, Y, ldy, L, k, A->u0, log2u, alpha); }
        else      { if (k2) SPMM_LAUNCH(k_spmm_fwd, 8, 2, A->d_meta, A->d_desc, (const Tv *)A->d_val, X, ldx, Y, ldy, L, k, A->u0, log2u, alpha);
                    else    SPMM_LAUNCH(k_spmm_fwd, 8, 1, A->d_meta, A->d_desc, (const Tv *)A->d_val, X, ldx, Y, ldy, L, k, A->u0, log2u, alpha); }
        A->launches++;
    }
#undef SPMM_LAUNCH
    VBC_CUDA(cudaGetLastError());
    return VBC_OK;
}

template <typename Tv>
static int transpose_launch(vbc_mat *A, const Tv *in, long long ldi, long long rows, long long cols, Tv *out, long long ldo)
{
    if (rows == 0 || cols == 0) return VBC_OK;
    dim3 grid((unsigned)((rows + 31) / 32), (unsigned)((cols + 31) / 32));
    k_transpose<Tv><<<grid, 256, 0, A->stream>>>(in, ldi, rows, cols, out, ldo);
    A->launches++;
    VBC_CUDA(cudaGetLastError());
    return VBC_OK;
}

// device pointers; layout 0 = row-major panels (X[i*ldx + c]), 1 = column-major (X[c*ldx + i])
template <typename Tv>
static int spmm_t(vbc_mat *A, int trans, int64_t k, double alpha_d, const Tv *X, int64_t ldx, double beta_d, Tv *Y, int64_t ldy, int layout)
{
    const Tv alpha = (Tv)alpha_d, beta = (Tv)beta_d;
    const int64_t xr = trans ? A->m : A->n, yr = trans ? A->n : A->m;
    if (layout == 0) {
        return A->desc_mode == DESC_ROWS ? launch_spmm_mode<Tv, DESC_ROWS>(A, trans, (int)k, alpha, X, ldx, beta, Y, ldy)
                                         : launch_spmm_mode<Tv, DESC_BLOCKS>(A, trans, (int)k, alpha, X, ldx, beta, Y, ldy);
    }
    // column-major: stage through row-major buffers kept in the handle (grow-only), so repeated calls pay
    // no allocation and stay stream-ordered
    const int64_t nx = xr * k, ny = yr * k;
    if (A->px_cap < nx * (int64_t)sizeof(Tv)) {
        VBC_CUDA(cudaStreamSynchronize(A->stream));
        cudaFree(A->d_px); A->d_px = nullptr; A->px_cap = 0;
        VBC_CUDA(cudaMalloc(&A->d_px, sizeof(Tv) * (size_t)(nx > 0 ? nx : 1)));
        A->px_cap = nx * (int64_t)sizeof(Tv);
    }
    if (A->py_cap < ny * (int64_t)sizeof(Tv)) {
        VBC_CUDA(cudaStreamSynchronize(A->stream));
        cudaFree(A->d_py); A->d_py = nullptr; A->py_cap = 0;
        VBC_CUDA(cudaMalloc(&A->d_py, sizeof(Tv) * (size_t)(ny > 0 ? ny : 1)));
        A->py_cap = ny * (int64_t)sizeof(Tv);
    }
    Tv *Xr = (Tv *)A->d_px, *Yr = (Tv *)A->d_py;
    VBC_TRY(transpose_launch<Tv>(A, X, ldx, k, xr, Xr, k)); // input viewed as k rows (columns of X) of length xr
    if (beta != (Tv)0) VBC_TRY(transpose_launch<Tv>(A, Y, ldy, k, yr, Yr, k));
    int rc;
    if (A->desc_mode == DESC_ROWS) rc = launch_spmm_mode<Tv, DESC_ROWS>(A, trans, (int)k, alpha, Xr, k, beta, Yr, k);
    else rc = launch_spmm_mode<Tv, DESC_BLOCKS>(A, trans, (int)k, alpha, Xr, k, beta, Yr, k);
    VBC_TRY(rc);
    return transpose_launch<Tv>(A, Yr, k, yr, k, Y, ldy);
}

int launch_spmm(vbc_mat *A, int trans, int64_t k, double alpha, const void *X, int64_t ldx, double beta, void *Y, int64_t ldy, int layout)
{
    if (A->opt_parity) VBC_FAIL(VBC_EARG, "spmm needs the compact layout (parity mode is on)");
    return A->vt == VBC_F64 ? spmm_t<double>(A, trans, k, alpha, (const double *)X, ldx, beta, (double *)Y, ldy, layout)
                            : spmm_t<float>(A, trans, k, alpha, (const float *)X, ldx, beta, (float *)Y, ldy, layout);
}

} // namespace vbc

using namespace vbc;

extern "C" int vbc_spmm(vbc_mat *A, int trans, int64_t k, double alpha, const void *X, int64_t ldx, double beta, void *Y, int64_t ldy,
                        int layout, int on_device)
{
    if (!A) VBC_FAIL(VBC_EARG, "matrix handle is NULL");
    if (k < 0 || k > (1 << 20)) VBC_FAIL(VBC_EARG, "k out of range");
    if (layout != 0 && layout != 1) VBC_FAIL(VBC_EARG, "layout must be 0 (row-major panels) or 1 (column-major)");
    const int64_t xr = trans ? A->m : A->n, yr = trans ? A->n : A->m;
    const int64_t min_ldx = layout == 0 ? k : xr, min_ldy = layout == 0 ? k : yr;
    if (ldx < min_ldx || ldy < min_ldy) VBC_FAIL(VBC_EDIM, "DimensionMismatch: leading dimensions (%lld, %lld) smaller than (%lld, %lld)", (long long)ldx, (long long)ldy, (long long)min_ldx, (long long)min_ldy);
    if (k == 0) return VBC_OK;
    if ((!X && xr > 0) || (!Y && yr > 0)) VBC_FAIL(VBC_EARG, "NULL panel");
    DeviceGuard guard(A->device);
    if (!guard.ok) VBC_FAIL(VBC_ECUDA, "cudaSetDevice(%d) failed", A->device);
    if (on_device) return launch_spmm(A, trans, k, alpha, X, ldx, beta, Y, ldy, layout);
    // host panels: copy the ld-strided storage as is (outer x ld elements)
    const size_t tv = vt_size(A->vt);
    const int64_t xo = layout == 0 ? xr : k, yo = layout == 0 ? yr : k; // outer extents
    void *dX = nullptr, *dY = nullptr;
    VBC_CUDA(cudaMalloc(&dX, tv * (size_t)(xo * ldx > 0 ? xo * ldx : 1)));
    if (cudaMalloc(&dY, tv * (size_t)(yo * ldy > 0 ? yo * ldy : 1)) != cudaSuccess) { cudaFree(dX); VBC_FAIL(VBC_ENOMEM, "spmm panel allocation failed"); }
    int rc = VBC_OK;
    cudaError_t e = cudaSuccess;
    if (xo * ldx > 0) e = cudaMemcpyAsync(dX, X, tv * (size_t)(xo * ldx), cudaMemcpyHostToDevice, A->stream);
    if (e == cudaSuccess && yo * ldy > 0) e = cudaMemcpyAsync(dY, Y, tv * (size_t)(yo * ldy), cudaMemcpyHostToDevice, A->stream); // keeps ld padding intact
    if (e == cudaSuccess) rc = launch_spmm(A, trans, k, alpha, dX, ldx, beta, dY, ldy, layout);
    if (e == cudaSuccess && rc == VBC_OK && yo * ldy > 0) e = cudaMemcpyAsync(Y, dY, tv * (size_t)(yo * ldy), cudaMemcpyDeviceToHost, A->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(A->stream);
    cudaFree(dX); cudaFree(dY);
    if (e != cudaSuccess) VBC_FAIL(VBC_ECUDA, "vbc_spmm: %s", cudaGetErrorString(e));
    return rc;
}
