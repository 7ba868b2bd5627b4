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
#include <cuda.h>
#include <limits.h>
#include <stdlib.h>

#include "common.cuh"
#include "walk.cuh"

namespace vbc {

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
__device__ __forceinline__ void dmma_m8n8k4(double (&c)[2], const double a, const double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

template <int MODE>
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
#pragma unroll
                        for (int nt = 0; nt < 4; nt++) {
                            const int col = kb + nt * 8 + g;
                            bv[q][nt] = (xi[q] >= 0 && col < k) ? __ldg(X + (long long)xi[q] * ldx + col) : 0.0;
                        }
                    }
#pragma unroll
                    for (int q = 0; q < KS; q++)
#pragma unroll
                        for (int nt = 0; nt < 4; nt++) dmma_m8n8k4(c[nt], av[q], bv[q][nt]);
                }
                if (arow) {
                    double *yp = Y + (long long)(a.col + wb + g) * ldy;
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

// The same tiles with the X rows loaded the way the memory system wants them and handed to the fragment layout by shuffles.
// In k_spmm_adj_dmma the four lanes of a quad (the contraction slots) read four DIFFERENT gathered rows, so the coalescer
// serves every 32-byte sector on its own: 8 L1 wavefronts per fragment load, 10 per stored row, and the L1 data pipe (90 %
// busy) bounds the kernel.  Here lane 8t' + g' loads 16 bytes of row t' -- eight adjacent lanes cover one 128-byte line, a
// warp-wide LDG.128 is four full lines -- and the fragment lane 4g + t fetches its four values from loader lane 8t + g with
// eight SHFL (the val tile stays on per-lane loads: routing it the same way cost 12 us).  The columns of the four N-tiles are permuted to make that a pure lane permutation: tile 0 / 1 = even / odd
// columns of 0..15, tile 2 / 3 = even / odd columns of 16..31; a lane's accumulators are then columns 4t..4t+3 and
// 16+4t..19+4t of output row g, stored as four 16-byte vectors.  Needs 16-byte aligned panels with even k and leading
// dimensions (else the kernel above runs).
template <int MODE>
__global__ void __launch_bounds__(256, VBC_DMMA_MINB) k_spmm_adj_dmma_v(const StripeMeta *__restrict__ meta, const int *__restrict__ desc,
                                                          const double *__restrict__ val, const double *__restrict__ X, const long long ldx,
                                                          double *__restrict__ Y, const long long ldy, const int L, const int k,
                                                          const int u0, const int log2u, const double alpha, const double beta)
{
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int lt = lane >> 3, lg = lane & 7; // loader role: row lt of a k-step, columns 2*lg, 2*lg + 1 of both 16-column halves
    const int src = 8 * t + g;               // the loader lane that holds this lane's fragment values
    const int nwarps = (int)((gridDim.x * blockDim.x) >> 5);
    for (int l = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5); l < L; l += nwarps) {
        const StripeMeta a = ld_meta(meta + l), b = ld_meta(meta + l + 1);
        const int w = b.col - a.col;
        if (w <= 0) continue;
        const int R = (MODE == DESC_ROWS) ? (b.pos - a.pos) : (int)((b.ofs - a.ofs) / w);
        for (int kb = 0; kb < k; kb += 32) {
            const int c0 = kb + 2 * lg, c1 = c0 + 16;
            const bool ok0 = c0 < k, ok1 = c1 < k; // k is even: a 16-byte vector never straddles the panel's edge
            for (int wb = 0; wb < w; wb += 8) {
                double c[4][2];
#pragma unroll
                for (int nt = 0; nt < 4; nt++) { c[nt][0] = 0.0; c[nt][1] = 0.0; }
                const bool arow = wb + g < w;
                const double *vp = val + a.ofs + (long long)t * w + wb + g;
                int xi_next = lt < R ? row_xindex<MODE>(desc, a.pos, lt, u0, log2u) : -1;
                constexpr int KS = VBC_DMMA_KS;
                // (a second register set that is loaded while this one is multiplied was tried: 317 us -- spills at the register
                // budget of 3 CTAs per SM, and the wait for one set also waits for the loads of the other: they share scoreboards)
                for (int r = 0; r < R; r += 4 * KS) {
                    int xi[KS];
                    xi[0] = xi_next;
#pragma unroll
                    for (int q = 1; q < KS; q++) xi[q] = (r + 4 * q + lt < R) ? row_xindex<MODE>(desc, a.pos, r + 4 * q + lt, u0, log2u) : -1;
                    xi_next = (r + 4 * KS + lt < R) ? row_xindex<MODE>(desc, a.pos, r + 4 * KS + lt, u0, log2u) : -1;
                    double av[KS];
                    double2 x0[KS], x1[KS];
#pragma unroll
                    for (int q = 0; q < KS; q++) av[q] = (arow && r + 4 * q + t < R) ? __ldcs(vp + 4 * q * (long long)w) : 0.0;
                    vp += 4 * KS * (long long)w;
#pragma unroll
                    for (int q = 0; q < KS; q++) {
                        const double *xp = X + (long long)(xi[q] >= 0 ? xi[q] : 0) * ldx;
                        x0[q] = (xi[q] >= 0 && ok0) ? __ldg(reinterpret_cast<const double2 *>(xp + c0)) : make_double2(0.0, 0.0);
                        x1[q] = (xi[q] >= 0 && ok1) ? __ldg(reinterpret_cast<const double2 *>(xp + c1)) : make_double2(0.0, 0.0);
                    }
#pragma unroll
                    for (int q = 0; q < KS; q++) {
                        const double b0 = __shfl_sync(0xffffffffu, x0[q].x, src), b1 = __shfl_sync(0xffffffffu, x0[q].y, src);
                        const double b2 = __shfl_sync(0xffffffffu, x1[q].x, src), b3 = __shfl_sync(0xffffffffu, x1[q].y, src);
                        dmma_m8n8k4(c[0], av[q], b0);
                        dmma_m8n8k4(c[1], av[q], b1);
                        dmma_m8n8k4(c[2], av[q], b2);
                        dmma_m8n8k4(c[3], av[q], b3);
                    }
                }
                if (arow) { // accumulator (nt, e) is column 4t + 2e + (nt & 1) of half nt >> 1
                    double *yp = Y + (long long)(a.col + wb + g) * ldy + kb + 4 * t;
#pragma unroll
                    for (int hf = 0; hf < 2; hf++) {
#pragma unroll
                        for (int e = 0; e < 2; e++) {
                            const int col = kb + 16 * hf + 4 * t + 2 * e;
                            if (col < k) {
                                double2 *p2 = reinterpret_cast<double2 *>(yp + 16 * hf + 2 * e);
                                double2 o = make_double2(alpha * c[2 * hf][e], alpha * c[2 * hf + 1][e]);
                                if (beta != 0.0) { const double2 old = *p2; o.x += beta * old.x; o.y += beta * old.y; }
                                *p2 = o;
                            }
                        }
                    }
                }
            }
        }
    }
}

// ---- Adjoint SpMM fed by TMA row gathers (Float64, rows mode, uniform stripe width 8, panels of <= 32 right-hand sides) -------------
// Same row-stream decomposition as k_spmm_adj_stream, same FP64 tensor tiles as k_spmm_adj_dmma, but the gathered X rows never
// pass through the load/store unit as per-lane requests (the L1 data pipe is what bounds both of those kernels):
//   * one `cp.async.bulk.tensor.2d ... tile::gather4` fetches four arbitrary rows of X (row indices straight from desc) as a
//     [4][16] Float64 tile; the tensor map is encoded with a {16, 1} box and SWIZZLE_128B, so that an 8-row x 128-byte atom
//     holds rows r = 0..7 with their 16-byte chunks XORed by r;
//   * a k-step of the m8n8k4 tile takes rows {0, 3, 4, 7} or {1, 2, 5, 6} of the atom: the four lanes of a quad (contraction
//     slots t = 0..3) then read chunks whose XOR patterns differ in bits 1-2 -- every B-fragment LDS.64 is conflict-free
//     (2 wavefronts for 256 bytes; the same fragment costs 8 when loaded from global memory);
//   * the 8 x 8 values of the chunk's val rows arrive by one 1-D bulk copy on the same mbarrier (A fragments: 2-way conflicts).
// Every warp runs its own ring of S stages (8 rows each) with one mbarrier per stage: no CTA-wide synchronisation at all.
__device__ __forceinline__ void mbar_init(const unsigned bar, const unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(const unsigned bar, const unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_try_wait(const unsigned bar, const unsigned parity)
{
    unsigned ok;
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void tma_gather4(const unsigned dst, const CUtensorMap *tm, const int c0, const int r0, const int r1, const int r2, const int r3, const unsigned bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                 ::"r"(dst), "l"(tm), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar) : "memory");
}
__device__ __forceinline__ bool elect_one()
{
    unsigned pred;
    asm volatile("{ .reg .pred p; elect.sync _|p, 0xffffffff; selp.u32 %0, 1, 0, p; }" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void bulk_copy_g2s(const unsigned dst, const void *src, const unsigned bytes, const unsigned bar)
{
    asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

#ifndef VBC_TMA_S
#define VBC_TMA_S 4
#endif
#ifndef VBC_TMA_MINB
#define VBC_TMA_MINB 2
#endif
constexpr int TMA_S = VBC_TMA_S;                           // stages per warp
constexpr int TMA_CR = 8;                                  // rows per stage
constexpr int TMA_STAGE_X = 2048, TMA_STAGE_V = 512;       // bytes per stage: two swizzle atoms of X, 8 x 8 values
constexpr int TMA_WARP_BYTES = TMA_S * (TMA_STAGE_X + TMA_STAGE_V);
constexpr int TMA_SMEM_BYTES = 8 * TMA_WARP_BYTES + 8 * TMA_S * 8 + 1024; // + mbarriers + alignment slack

template <bool FULL>
__global__ void __launch_bounds__(256, VBC_TMA_MINB) k_spmm_adj_tma(const __grid_constant__ CUtensorMap tmX, const StripeMeta *__restrict__ meta,
                                                         const int *__restrict__ desc, const double *__restrict__ val, double *__restrict__ Y,
                                                         const long long ldy, const int L, const int nunits, const double ratio, const int k,
                                                         const int kb, const double alpha, const double beta)
{
    constexpr int S = TMA_S, CR = TMA_CR, W = 8;
    extern __shared__ unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int wq = (int)__reduce_max_sync(0xffffffffu, threadIdx.x >> 5); // the warp's index as a value known to be warp-uniform
    const int nwarps = (int)gridDim.x * 8, wid = (int)blockIdx.x * 8 + wq;
    const unsigned smem0 = ((unsigned)__cvta_generic_to_shared(smem_raw) + 1023u) & ~1023u;
    const unsigned xs = smem0 + (unsigned)wq * (S * TMA_STAGE_X);                        // [S][2 column halves][8 rows][128 B], 128-byte swizzle
    const unsigned vsm = smem0 + 8u * S * TMA_STAGE_X + (unsigned)wq * (S * TMA_STAGE_V); // [S][8 rows][8 values]
    const unsigned bars = smem0 + 8u * S * (TMA_STAGE_X + TMA_STAGE_V) + (unsigned)wq * (S * 8);
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < S; s++) mbar_init(bars + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
    if (wid >= nunits) return;

    // Units of <= 32 adjacent stripes are dealt round-robin (all warps work in one sliding window of X: L2 reuse of the gathered
    // rows); a unit is a flat run of stored rows [P, E), taken in chunks of 8 rows = one stage.  The requests of chunk c + S - 1 go
    // out while chunk c is multiplied; the pipeline drains at the end of a unit (a few per warp) and the next unit's bounds and
    // stripe ends are fetched while this one runs.
    auto unit_lo = [&](const int u) { return u >= nunits ? L : min(L, (int)((double)u * ratio)); };
    int u = wid;
    int l0n = unit_lo(u), l1n = unit_lo(u + 1);
    int Pn = __ldg(&meta[l0n].pos), En = __ldg(&meta[l1n].pos);
    int segsn = lane < l1n - l0n ? __ldg(&meta[l0n + 1 + lane].pos) : 0;

    double c[4][2];
#pragma unroll
    for (int nt = 0; nt < 4; nt++) c[nt][0] = c[nt][1] = 0.0;
    // this lane's rows in the two k-steps of a chunk and the swizzled byte offsets of its B-fragment elements (n-tiles 0 / 1; 2 and 3: + 1024)
    const int r_ks0 = 2 * t + (t & 1), r_ks1 = 2 * t + 1 - (t & 1); // {0, 3, 4, 7} and {1, 2, 5, 6}
    const unsigned xo0 = (unsigned)(r_ks0 * 128 + ((((g >> 1) ^ r_ks0) & 7) << 4) + (g & 1) * 8);
    const unsigned xo1 = (unsigned)(r_ks1 * 128 + ((((g >> 1) ^ r_ks1) & 7) << 4) + (g & 1) * 8);
    const unsigned vo0 = (unsigned)(r_ks0 * 64 + g * 8), vo1 = (unsigned)(r_ks1 * 64 + g * 8);
    int stage = 0, istage = 0; // stage multiplied next / stage filled next (both advance once per chunk, in the same order)
    unsigned parity = 0;

    for (; u < nunits; u += nwarps) {
        const int l0 = l0n, l1 = l1n, P = Pn, E = En, segs = segsn;
        if (u + nwarps < nunits) {
            l0n = unit_lo(u + nwarps); l1n = unit_lo(u + nwarps + 1);
            Pn = __ldg(&meta[l0n].pos); En = __ldg(&meta[l1n].pos);
            segsn = lane < l1n - l0n ? __ldg(&meta[l0n + 1 + lane].pos) : 0;
        }
        const int nch = (E - P + CR - 1) / CR;
        auto ldidx = [&](const int ci) { const int row = P + CR * ci + lane; return (lane < CR && row < E) ? __ldcs(desc + row) : 0; };
        // requests of chunk ci, all from ONE elected lane with operands the compiler knows to be warp-uniform: every row index is
        // extracted with a warp reduction (REDUX writes a uniform register), so UTMALDG / UBLKCP take their operands without the
        // ELECT / R2UR / branch waterfall that per-lane values need
        auto issue = [&](const int ci, const int idx) {
            if (ci >= nch) return;
            int r[CR];
#pragma unroll
            for (int j = 0; j < CR; j++) r[j] = __reduce_add_sync(0xffffffffu, lane == j ? idx : 0);
            const int n = min(CR, E - P - CR * ci);
            const unsigned bar = bars + 8u * (unsigned)istage, dst = xs + (unsigned)istage * TMA_STAGE_X;
            if (elect_one()) {
                mbar_expect_tx(bar, (unsigned)(TMA_STAGE_X + n * W * 8));
                bulk_copy_g2s(vsm + (unsigned)istage * TMA_STAGE_V, val + (long long)(P + CR * ci) * W, (unsigned)(n * W * 8), bar);
                tma_gather4(dst, &tmX, kb, r[0], r[1], r[2], r[3], bar);
                tma_gather4(dst + 512u, &tmX, kb, r[4], r[5], r[6], r[7], bar);
                tma_gather4(dst + 1024u, &tmX, kb + 16, r[0], r[1], r[2], r[3], bar);
                tma_gather4(dst + 1536u, &tmX, kb + 16, r[4], r[5], r[6], r[7], bar);
            }
            istage = istage == S - 1 ? 0 : istage + 1;
        };
        auto flush = [&](const int l) { // stripe l is complete: Y[l*8 + g, panel] <- alpha * c + beta * Y
#pragma unroll
            for (int nt = 0; nt < 4; nt++) {
                const int col = kb + nt * 8 + 2 * t;
                if (FULL || col < k) {
                    double2 *yp = reinterpret_cast<double2 *>(Y + ((long long)l * W + g) * ldy + col);
                    double2 o = make_double2(alpha * c[nt][0], alpha * c[nt][1]);
                    if (beta != 0.0) { const double2 old = *yp; o.x += beta * old.x; o.y += beta * old.y; }
                    *yp = o;
                }
                c[nt][0] = c[nt][1] = 0.0;
            }
        };
#pragma unroll
        for (int ci = 0; ci < S - 1; ci++) issue(ci, ldidx(ci));
        int idxA = ldidx(S - 1), idxB = ldidx(S); // row indices are fetched two chunks before their requests go out
        int l = l0, seg_end = __shfl_sync(0xffffffffu, segs, 0);
        for (int ci = 0; ci < nch; ci++) {
            issue(ci + S - 1, idxA); // into the stage multiplied in the previous iteration
            idxA = idxB; idxB = ldidx(ci + S + 1);
            const unsigned bar = bars + 8u * (unsigned)stage;
            for (unsigned spins = 0; !mbar_try_wait(bar, parity); spins++)
                if (spins > (1u << 24)) __trap(); // a request that never completes must not hang the GPU
            const unsigned xb = xs + (unsigned)stage * TMA_STAGE_X, vb = vsm + (unsigned)stage * TMA_STAGE_V;
            double a0, a1, b0[4], b1[4];
            asm volatile("ld.shared.f64 %0, [%1];" : "=d"(a0) : "r"(vb + vo0) : "memory");
            asm volatile("ld.shared.f64 %0, [%1];" : "=d"(a1) : "r"(vb + vo1) : "memory");
#pragma unroll
            for (int nt = 0; nt < 4; nt++) {
                const unsigned ofs = (nt >> 1) * 1024u;
                asm volatile("ld.shared.f64 %0, [%1];" : "=d"(b0[nt]) : "r"(xb + ofs + ((nt & 1) ? (xo0 ^ 64u) : xo0)) : "memory");
                asm volatile("ld.shared.f64 %0, [%1];" : "=d"(b1[nt]) : "r"(xb + ofs + ((nt & 1) ? (xo1 ^ 64u) : xo1)) : "memory");
            }
            const int Pc = P + CR * ci;
            if (seg_end > Pc + CR) { // all eight rows belong to the open stripe
#pragma unroll
                for (int nt = 0; nt < 4; nt++) dmma_m8n8k4(c[nt], a0, b0[nt]);
#pragma unroll
                for (int nt = 0; nt < 4; nt++) dmma_m8n8k4(c[nt], a1, b1[nt]);
            } else {
                int lo = Pc;
                for (;;) {
                    const int hi = min(seg_end, Pc + CR);
                    if (hi > lo) { // rows [lo, hi) of the chunk belong to stripe l: everything else is masked on both operands
                        const bool m0 = Pc + r_ks0 >= lo && Pc + r_ks0 < hi, m1 = Pc + r_ks1 >= lo && Pc + r_ks1 < hi;
                        const double am0 = m0 ? a0 : 0.0, am1 = m1 ? a1 : 0.0;
#pragma unroll
                        for (int nt = 0; nt < 4; nt++) dmma_m8n8k4(c[nt], am0, m0 ? b0[nt] : 0.0);
#pragma unroll
                        for (int nt = 0; nt < 4; nt++) dmma_m8n8k4(c[nt], am1, m1 ? b1[nt] : 0.0);
                    }
                    if (seg_end > Pc + CR) break; // the stripe continues in the next chunk
                    flush(l); l++;
                    if (l == l1) break;
                    seg_end = __shfl_sync(0xffffffffu, segs, l - l0);
                    lo = hi;
                }
            }
            __syncwarp(); // every lane has read the stage before the next iteration overwrites it
            stage = stage == S - 1 ? 0 : stage + 1;
            if (stage == 0) parity ^= 1u;
        }
        while (l < l1) { flush(l); l++; } // stripes without stored rows at the end of the unit (or a unit without rows)
    }
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
                constexpr int RU = (WB * KT <= 16) ? 2 : 1; // stored rows whose values are loaded before the first product
                for (int r = 0; r < R; r += RU) {
                    int xi[RU];
                    Tv v[RU][WB];
#pragma unroll
                    for (int q = 0; q < RU; q++) {
                        xi[q] = walk.next_if(r + q < R);
#pragma unroll
                        for (int dj = 0; dj < WB; dj++) v[q][dj] = (r + q < R && wb + dj < w) ? __ldg(vp + dj) : (Tv)0;
                        vp += w;
                    }
#pragma unroll
                    for (int q = 0; q < RU; q++) {
                        Tv p[KT];
#pragma unroll
                        for (int t = 0; t < KT; t++) p[t] = (Tv)0;
#pragma unroll
                        for (int dj = 0; dj < WB; dj++)
#pragma unroll
                            for (int t = 0; t < KT; t++) p[t] = fma(v[q][dj], xs[dj][t], p[t]);
                        if (r + q < R) {
#pragma unroll
                            for (int t = 0; t < KT; t++) {
                                const int c = kb + t * 32 + lane;
                                if (c < k) atomicAdd(Y + (long long)xi[q] * ldy + c, alpha * p[t]);
                            }
                        }
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
            // one stripe width 8 over the whole matrix, rows mode, 16-byte aligned even-k panels: the TMA-fed kernel applies
            const int Wu = A->w_uniform;
            const bool streamable = MODE == DESC_ROWS && Wu == 8 && A->nval == A->ndesc * Wu && A->n == (int64_t)L * Wu &&
                                    (k % 2) == 0 && (ldx % 2) == 0 && (ldy % 2) == 0 && ldx < (1ll << 28) && ((uintptr_t)X % 16) == 0 && ((uintptr_t)Y % 16) == 0;
            if ((A->opt_spmm_simt == 0 || A->opt_spmm_simt == 3) && streamable && Wu == 8) { // TMA-fed tensor tiles
                typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                             const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
                static EncodeFn encode = nullptr;
                static int tma_state_dev[64] = {}; // per device (the shared-memory attribute belongs to its context): 0 untried, 1 ready, -1 unavailable
                int &tma_state = tma_state_dev[A->device & 63];
                if (tma_state == 0) {
                    cudaDriverEntryPointQueryResult qres;
                    void *fn = nullptr;
                    tma_state = 1;
                    if (!encode) {
                        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) tma_state = -1;
                        else encode = (EncodeFn)fn;
                    }
                    if (tma_state == 1 && (cudaFuncSetAttribute(k_spmm_adj_tma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TMA_SMEM_BYTES) != cudaSuccess ||
                                           cudaFuncSetAttribute(k_spmm_adj_tma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TMA_SMEM_BYTES) != cudaSuccess)) tma_state = -1;
                    if (tma_state < 0) cudaGetLastError();
                }
                CUtensorMap tm;
                bool tma_ok = tma_state == 1;
                if (tma_ok) {
                    const cuuint64_t gdim[2] = {(cuuint64_t)k, (cuuint64_t)A->m}, gstr[1] = {(cuuint64_t)ldx * 8};
                    const cuuint32_t box[2] = {16, 1}, estr[2] = {1, 1};
                    tma_ok = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<Tv *>(X), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
                }
                if (tma_ok) { // (not available: the per-lane-load tiles below run instead)
                int64_t g2 = (int64_t)A->sm_count * VBC_TMA_MINB;
                const int64_t nw = g2 * 8;
                const double avg_rows = (double)A->ndesc / (double)L;
                static int unit_rows = -1;
                if (unit_rows < 0) { const char *e = getenv("VBC_TMA_UNIT_ROWS"); unit_rows = e ? atoi(e) : 288; if (unit_rows < 8) unit_rows = 8; }
                int64_t U0 = (int64_t)((double)unit_rows / (avg_rows > 1.0 ? avg_rows : 1.0) + 0.5); // stripes per unit: the pipeline drains once per unit
                if (U0 < 1) U0 = 1;
                if (U0 > 24) U0 = 24;
                int64_t q = (L + nw * U0 / 2) / (nw * U0);
                if (q < 1) q = 1;
                while ((double)L / (double)(nw * q) > 30.0) q++;
                int64_t nunits = nw * q;
                if (nunits > L) { nunits = L; g2 = (nunits + 7) / 8; }
                const double ratio = (double)L / (double)nunits;
                for (int kb = 0; kb < k; kb += 32) {
                    if (k - kb >= 32) k_spmm_adj_tma<true><<<(unsigned)g2, 256, TMA_SMEM_BYTES, A->stream>>>(tm, A->d_meta, A->d_desc, (const double *)A->d_val, (double *)Y, ldy, L, (int)nunits, ratio, k, kb, (double)alpha, (double)beta);
                    else k_spmm_adj_tma<false><<<(unsigned)g2, 256, TMA_SMEM_BYTES, A->stream>>>(tm, A->d_meta, A->d_desc, (const double *)A->d_val, (double *)Y, ldy, L, (int)nunits, ratio, k, kb, (double)alpha, (double)beta);
                    A->launches++;
                }
                VBC_CUDA(cudaGetLastError());
                return VBC_OK;
                }
            }
            if (A->opt_spmm_simt != 1) { // Float64: tensor (DMMA) tiles
                int64_t g2 = (int64_t)A->sm_count * 8;
                if (g2 > need) g2 = need;
                if (g2 < 1) g2 = 1;
                const bool vec = (k % 2) == 0 && (ldx % 2) == 0 && (ldy % 2) == 0 && ((uintptr_t)X % 16) == 0 && ((uintptr_t)Y % 16) == 0 && !getenv("VBC_SPMM_NOVEC");
                if (vec) k_spmm_adj_dmma_v<MODE><<<(unsigned)g2, 256, 0, A->stream>>>(A->d_meta, A->d_desc, (const double *)A->d_val, (const double *)X, ldx,
                                                                          (double *)Y, ldy, L, k, A->u0, log2u, (double)alpha, (double)beta);
                else k_spmm_adj_dmma<MODE><<<(unsigned)g2, 256, 0, A->stream>>>(A->d_meta, A->d_desc, (const double *)A->d_val, (const double *)X, ldx,
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
    if (!vt_is_float(A->vt)) VBC_FAIL(VBC_EARG, "ArgumentError: spmm is offered for Float32 / Float64 matrices");
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
    // host panels: an ld-strided buffer holds (outer - 1) * ld + inner elements (BLAS convention), so the copies are 2-D:
    // `inner` elements per row of the panel, pitch ld on both sides; the padding between rows is neither read nor written
    const size_t tv = vt_size(A->vt);
    const int64_t xo = layout == 0 ? xr : k, yo = layout == 0 ? yr : k; // outer extents
    const int64_t xi = layout == 0 ? k : xr, yi = layout == 0 ? k : yr; // inner extents
    void *dX = nullptr, *dY = nullptr;
    VBC_CUDA(cudaMalloc(&dX, tv * (size_t)(xo * ldx > 0 ? xo * ldx : 1)));
    if (cudaMalloc(&dY, tv * (size_t)(yo * ldy > 0 ? yo * ldy : 1)) != cudaSuccess) { cudaFree(dX); VBC_FAIL(VBC_ENOMEM, "spmm panel allocation failed"); }
    int rc = VBC_OK;
    cudaError_t e = cudaSuccess;
    if (xo > 0 && xi > 0) e = cudaMemcpy2DAsync(dX, tv * (size_t)ldx, X, tv * (size_t)ldx, tv * (size_t)xi, (size_t)xo, cudaMemcpyHostToDevice, A->stream);
    if (e == cudaSuccess && yo > 0 && yi > 0 && beta != 0.0) e = cudaMemcpy2DAsync(dY, tv * (size_t)ldy, Y, tv * (size_t)ldy, tv * (size_t)yi, (size_t)yo, cudaMemcpyHostToDevice, A->stream);
    if (e == cudaSuccess) rc = launch_spmm(A, trans, k, alpha, dX, ldx, beta, dY, ldy, layout);
    if (e == cudaSuccess && rc == VBC_OK && yo > 0 && yi > 0) e = cudaMemcpy2DAsync(Y, tv * (size_t)ldy, dY, tv * (size_t)ldy, tv * (size_t)yi, (size_t)yo, cudaMemcpyDeviceToHost, A->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(A->stream);
    cudaFree(dX); cudaFree(dY);
    if (e != cudaSuccess) VBC_FAIL(VBC_ECUDA, "vbc_spmm: %s", cudaGetErrorString(e));
    return rc;
}
