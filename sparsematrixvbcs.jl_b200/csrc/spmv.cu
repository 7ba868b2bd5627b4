// spmv.cu -- VBC sparse matrix-vector multiply kernels for sm_100a.
//
//   adjoint  y <- alpha * A' x + beta * y   replaces  mul!(y, B', x, α, β)
//            /root/reference/src/multiply_1DVBC.jl:85-180 (_1DVBR_mul! :90-134)
//            /root/reference/src/multiply_VBC.jl:89-192   (_VBR_mul!   :93-147)
//   forward  y <- alpha * A x + beta * y    replaces  mul!(y, B, x, α, β)
//            /root/reference/src/multiply_1DVBC.jl:9-83, multiply_VBC.jl:3-87
//
// HBM-bound streaming kernels.  A stripe l is a dense R_l x w_l row-major slab of `val`
// (1D: R_l stored rows; 2D: the rows of its blocks, block after block), and all slabs are
// contiguous.  G lanes (a "group", G in {8,32}) cooperate on one stripe: consecutive lanes read
// consecutive 16-byte vectors of the slab (fully coalesced), so lane v always holds the same
// column-vector c = v mod cpr and rows r0, r0+rps, ...  Adjoint: per-lane FMAs into EPV
// accumulators, then a shuffle reduction over the lanes that share c -- the owner-computes
// "row block" kernel.  Forward: per-row shuffle reduction over the cpr lanes of a row, then
// one fp atomic add per stored row (the scatter orientation, y[idx[Q]] += ...).
//
// Where the reference unrolls one body per width class (le_nest over ws, util.jl:28-38), the
// kernels here instantiate one body per (elements-per-load EPV, vectors-per-row CPR) class
// and select it per stripe.
#include "common.cuh"
#include "walk.cuh"

namespace vbc {

// ---- compile-time tuning knobs (tools/build_variants.sh builds one library per setting) -------
#ifndef VBC_ADJ_UNR
#define VBC_ADJ_UNR 4      // independent row-steps in flight per lane in the adjoint main loop
#endif
#ifndef VBC_ADJ_MINB
#define VBC_ADJ_MINB 4     // __launch_bounds__ min CTAs/SM for the adjoint kernel (caps registers at 64); 0 = compiler default
                           // (tools/tune.py sweep, profiles/r01_tuning.md: 4 CTAs x 64 regs beats 5 x 48 and 3 x 80)
#endif
#if VBC_ADJ_MINB > 0
#define VBC_ADJ_BOUNDS __launch_bounds__(256, VBC_ADJ_MINB)
#else
#define VBC_ADJ_BOUNDS __launch_bounds__(256)
#endif
#ifndef VBC_LD_MODE
#define VBC_LD_MODE 0      // val loads: 0 = ld.global.cs (streaming), 1 = ld.global.nc, 2 = ld.global.nc.L1::no_allocate
#endif
#ifndef VBC_WIDE_LD
#define VBC_WIDE_LD 0      // 1: 256-bit loads (sm_100+ LDG.256) for Float64 stripes whose width is a multiple of 4
#endif

// ---- vector loads ---------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ T ld_val(const T *p)
{
#if VBC_LD_MODE == 0
    return __ldcs(p);
#else
    return __ldg(p);
#endif
}
#if VBC_LD_MODE == 2
template <> __device__ __forceinline__ double2 ld_val<double2>(const double2 *p)
{
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}
template <> __device__ __forceinline__ float4 ld_val<float4>(const float4 *p)
{
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
#endif
template <typename Tv, int EPV> struct Ld;
template <> struct Ld<double, 1> { static __device__ __forceinline__ void s(const double *p, double (&v)[1]) { v[0] = ld_val(p); } };
template <> struct Ld<double, 2> { static __device__ __forceinline__ void s(const double *p, double (&v)[2]) { const double2 t = ld_val(reinterpret_cast<const double2 *>(p)); v[0] = t.x; v[1] = t.y; } };
template <> struct Ld<double, 4> { static __device__ __forceinline__ void s(const double *p, double (&v)[4]) {
    asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p)); } };
template <> struct Ld<float, 1> { static __device__ __forceinline__ void s(const float *p, float (&v)[1]) { v[0] = ld_val(p); } };
template <> struct Ld<float, 2> { static __device__ __forceinline__ void s(const float *p, float (&v)[2]) { const float2 t = ld_val(reinterpret_cast<const float2 *>(p)); v[0] = t.x; v[1] = t.y; } };
template <> struct Ld<float, 4> { static __device__ __forceinline__ void s(const float *p, float (&v)[4]) { const float4 t = ld_val(reinterpret_cast<const float4 *>(p)); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; } };

// Destinations of the adjoint result.  n == 0: the caller's y (alpha/beta applied).  n > 0: the
// fused all-gather of the row-partitioned multiply -- every finished y segment is stored into the
// next-x buffer of each of the n ranks (own HBM and NVLink peer mappings); beta is not applied.
struct PeerDst {
    int n;
    void *p[VBC_MAX_PEERS];
    const unsigned char *mask; // optional: mask[(col >> chunk_shift)] bit i set <=> destination i reads that column chunk
    int chunk_shift;
    // in-kernel cross-rank synchronisation (fused flag exchange); sync_n == 0: none (the caller launches k_peer_flags)
    int sync_n, me;                               // ranks, this rank
    unsigned long long *flags[VBC_MAX_PEERS];     // flag block of every rank (flags[me] is local)
    unsigned long long *d_epoch;                  // epoch signalled at the end of this rank's previous step
    unsigned *d_done;                             // CTAs finished in this launch
    int *timed_out;
    int i0, i1;                                   // stripes [i0, i1) need nothing from other ranks and feed only this rank
    int ra0, ra1, rb0, rb1;                       // sync_n == 0: the stripe ranges this launch covers ([ra0,ra1) then [rb0,rb1))
};

__device__ __forceinline__ void st_release_sys_u64(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

template <typename Tv, int EPV, bool PEER>
__device__ __forceinline__ void store_y(Tv *__restrict__ y, const PeerDst &dst, const int col, const Tv (&acc)[EPV], const Tv alpha, const Tv beta)
{
    if constexpr (PEER) {
        Tv v[EPV];
#pragma unroll
        for (int e = 0; e < EPV; e++) v[e] = alpha * acc[e];
        for (int i = 0; i < dst.n; i++) {
            Tv *yp = reinterpret_cast<Tv *>(dst.p[i]) + col;
#pragma unroll
            for (int e = 0; e < EPV; e++) yp[e] = v[e];
        }
    } else {
        Tv *yp = y + col;
#pragma unroll
        for (int e = 0; e < EPV; e++) yp[e] = (beta == (Tv)0) ? alpha * acc[e] : alpha * acc[e] + beta * yp[e];
    }
}

// ---- adjoint ----------------------------------------------------------------------------------
// CPR > 0: compile-time vectors per row (power of two, <= G).  CPR == 0: runtime cpr <= G.
template <typename Tv, int G, int MODE, int EPV, int CPR, bool PEER>
__device__ __forceinline__ void adj_stripe(const StripeMeta a, const StripeMeta b, const int w, const int lane, const unsigned gmask,
                                           const int *__restrict__ desc, const Tv *__restrict__ val, const Tv *__restrict__ x,
                                           Tv *__restrict__ y, const PeerDst &dst, const int u0, const int log2u, const Tv alpha, const Tv beta)
{
    int cpr, rps, c, r0;
    if constexpr (CPR > 0) { cpr = CPR; rps = G / CPR; c = lane % CPR; r0 = lane / CPR; }
    else { cpr = w / EPV; rps = small_div(G, cpr); r0 = small_div(lane, cpr); c = lane - r0 * cpr; }
    int R;
    if (MODE == DESC_ROWS) R = b.pos - a.pos;
    else { const int nv = (int)(b.ofs - a.ofs); if constexpr (CPR > 0) R = nv / (CPR * EPV); else R = nv / w; }
    const bool active = (CPR > 0) || (lane < rps * cpr);
    Tv acc[EPV];
#pragma unroll
    for (int e = 0; e < EPV; e++) acc[e] = (Tv)0;
    int r = active ? r0 : R;
    const Tv *vp = val + a.ofs + (long long)r0 * w + c * EPV;
    const int vstride = rps * w;
    RowWalk<MODE> walk;
    walk.init(desc, a.pos, r0, rps, u0, log2u);
    // Batches of UNR independent row-steps: all loads of a batch are issued before the first FMA;
    // the last batch is predicated instead of falling into a serial remainder loop.
    constexpr int UNR = VBC_ADJ_UNR;
    for (; r < R; r += UNR * rps) {
        Tv v[UNR][EPV];
        int xi[UNR];
        Tv xv[UNR];
        bool ok[UNR];
#pragma unroll
        for (int k = 0; k < UNR; k++) {
            ok[k] = r + k * rps < R;
            if (ok[k]) Ld<Tv, EPV>::s(vp, v[k]);
            else {
#pragma unroll
                for (int e = 0; e < EPV; e++) v[k][e] = (Tv)0;
            }
            vp += vstride;
        }
#pragma unroll
        for (int k = 0; k < UNR; k++) xi[k] = walk.next_if(ok[k]);
#pragma unroll
        for (int k = 0; k < UNR; k++) xv[k] = ok[k] ? __ldg(x + xi[k]) : (Tv)0;
#pragma unroll
        for (int k = 0; k < UNR; k++)
#pragma unroll
            for (int e = 0; e < EPV; e++) acc[e] = fma(v[k][e], xv[k], acc[e]);
    }
    // sum the lanes that hold the same column-vector c
    if constexpr (CPR > 0) {
#pragma unroll
        for (int d = CPR; d < G; d <<= 1)
#pragma unroll
            for (int e = 0; e < EPV; e++) acc[e] += __shfl_xor_sync(gmask, acc[e], d, G);
    } else {
        for (int d = cpr; d < G; d <<= 1)
#pragma unroll
            for (int e = 0; e < EPV; e++) {
                const Tv t = __shfl_down_sync(gmask, acc[e], d, G);
                if (lane + d < G) acc[e] += t;
            }
    }
    // y[j + Δj] = tmp[Δj]  (multiply_1DVBC.jl:114-116), with BLAS alpha/beta
    if (lane < cpr) store_y<Tv, EPV, PEER>(y, dst, a.col + lane * EPV, acc, alpha, beta);
}

// wide stripes (more vectors per row than lanes): one lane per column, serial over rows
template <typename Tv, int G, int MODE, bool PEER>
__device__ __noinline__ void adj_stripe_wide(const StripeMeta a, const StripeMeta b, const int w, const int lane,
                                             const int *__restrict__ desc, const Tv *__restrict__ val, const Tv *__restrict__ x,
                                             Tv *__restrict__ y, const PeerDst &dst, const int u0, const Tv alpha, const Tv beta)
{
    const int R = (MODE == DESC_ROWS) ? (b.pos - a.pos) : (int)((b.ofs - a.ofs) / w);
    for (int col = lane; col < w; col += G) {
        Tv acc = (Tv)0;
        RowWalk<MODE> walk;
        walk.init(desc, a.pos, 0, 1, u0, -1);
        const Tv *vp = val + a.ofs + col;
        for (int r = 0; r < R; r++) { acc = fma(__ldcs(vp), __ldg(x + walk.next()), acc); vp += w; }
        const Tv accv[1] = {acc};
        store_y<Tv, 1, PEER>(y, dst, a.col + col, accv, alpha, beta);
    }
}

template <typename Tv, int G, int MODE, int EPV, bool PEER>
__device__ __forceinline__ void adj_dispatch_cpr(const StripeMeta a, const StripeMeta b, const int w, const int lane, const unsigned gmask,
                                                 const int *__restrict__ desc, const Tv *__restrict__ val, const Tv *__restrict__ x,
                                                 Tv *__restrict__ y, const PeerDst &dst, const int u0, const int log2u, const Tv alpha, const Tv beta)
{
    const int cpr = w / EPV;
    if (cpr == 1) adj_stripe<Tv, G, MODE, EPV, 1, PEER>(a, b, w, lane, gmask, desc, val, x, y, dst, u0, log2u, alpha, beta);
    else if (cpr == 2) adj_stripe<Tv, G, MODE, EPV, 2, PEER>(a, b, w, lane, gmask, desc, val, x, y, dst, u0, log2u, alpha, beta);
    else if (cpr == 4) adj_stripe<Tv, G, MODE, EPV, 4, PEER>(a, b, w, lane, gmask, desc, val, x, y, dst, u0, log2u, alpha, beta);
    else if (cpr <= G) adj_stripe<Tv, G, MODE, EPV, 0, PEER>(a, b, w, lane, gmask, desc, val, x, y, dst, u0, log2u, alpha, beta);
    else adj_stripe_wide<Tv, G, MODE, PEER>(a, b, w, lane, desc, val, x, y, dst, u0, alpha, beta);
}

template <typename Tv, int G, int MODE, bool PEER>
__global__ void VBC_ADJ_BOUNDS k_spmv_adj(const StripeMeta *__restrict__ meta, const int *__restrict__ desc,
                                                   const Tv *__restrict__ val, const Tv *__restrict__ x, Tv *__restrict__ y,
                                                   const __grid_constant__ PeerDst dst, const int *__restrict__ order,
                                                   const int L, const int u0, const int log2u, const Tv alpha, const Tv beta)
{
    constexpr int VE = 16 / (int)sizeof(Tv);
    const int lane = threadIdx.x % G;
    const unsigned gmask = group_mask<G>();
    const int ngroups = (int)((gridDim.x * blockDim.x) / G);
    if constexpr (PEER) {
        // Fused all-gather: the warp's groups own adjacent stripes, i.e. one contiguous run of columns.
        // Their results are staged in shared memory and flushed to every destination (own next-x buffer
        // and the peers' over NVLink) with full-warp coalesced stores -- 128-256 B per store instruction
        // instead of one 16-32 B store per group, which is what NVLink write packets want.
        constexpr int GPW = 32 / G;
        __shared__ Tv stage_all[8][GPW * 32];
        Tv *stage = stage_all[threadIdx.x >> 5];
        const int lane32 = threadIdx.x & 31, gid = lane32 / G;
        const int nwarps = ngroups / GPW;
        const PeerDst none{};
        const int warp0 = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
        auto run_range = [&](const int lo, const int hi) {
        for (int lbase = lo + warp0 * GPW; lbase < hi; lbase += nwarps * GPW) {
            const int l = lbase + gid;
            const int lend = min(lbase + GPW, hi);
            const int colbase = ld_meta(meta + lbase).col, colend = ld_meta(meta + lend).col;
            // sparsity-aware replication: when every column chunk of this run is read by this rank only, the
            // results go straight to the own next-x buffer -- no staging, no flush (the common case for a banded operator)
            bool self_only = false;
            if (dst.mask != nullptr && colend > colbase) {
                self_only = true;
                for (int ch = colbase >> dst.chunk_shift; ch <= ((colend - 1) >> dst.chunk_shift); ch++) self_only = self_only && (__ldg(dst.mask + ch) == 1);
            }
            if (l < hi) {
                const StripeMeta a = ld_meta(meta + l), b = ld_meta(meta + l + 1);
                const int w = b.col - a.col;
                Tv *ys = self_only ? reinterpret_cast<Tv *>(dst.p[0]) : stage - colbase; // the stripe bodies store y[a.col + ...]
                if (w > 0) {
                    if ((w % VE) == 0 && (a.ofs % VE) == 0)
                        adj_dispatch_cpr<Tv, G, MODE, VE, false>(a, b, w, lane, gmask, desc, val, x, ys, none, u0, log2u, alpha, (Tv)0);
                    else if (VE == 4 && (w % 2) == 0 && (a.ofs % 2) == 0)
                        adj_dispatch_cpr<Tv, G, MODE, (VE == 4 ? 2 : 1), false>(a, b, w, lane, gmask, desc, val, x, ys, none, u0, log2u, alpha, (Tv)0);
                    else
                        adj_dispatch_cpr<Tv, G, MODE, 1, false>(a, b, w, lane, gmask, desc, val, x, ys, none, u0, log2u, alpha, (Tv)0);
                }
            }
            if (self_only) continue; // warp-uniform
            __syncwarp();
            const int ncols = colend - colbase;
            if (dst.mask == nullptr) { // full replication
                for (int i = 0; i < dst.n; i++) {
                    Tv *d = reinterpret_cast<Tv *>(dst.p[i]) + colbase;
                    for (int c = lane32; c < ncols; c += 32) d[c] = stage[c];
                }
            } else { // sparsity-aware replication: a destination gets only the column chunks it gathers from
                for (int c = lane32; c < ncols; c += 32) {
                    const unsigned mk = __ldg(dst.mask + ((colbase + c) >> dst.chunk_shift));
                    const Tv v = stage[c];
                    for (int i = 0; i < dst.n; i++)
                        if ((mk >> i) & 1u) reinterpret_cast<Tv *>(dst.p[i])[colbase + c] = v;
                }
            }
            __syncwarp();
        }
        };
        // one copy of the stripe bodies: the (at most two) ranges of this launch are walked by the same loop
        for (int part = 0; part < 2; part++) run_range(part ? dst.rb0 : dst.ra0, min(part ? dst.rb1 : dst.ra1, L));
        return;
    }
    int l = (int)((blockIdx.x * blockDim.x + threadIdx.x) / G);
    if (l >= L) return;
    // order == null: stripes in index order, next stripe's meta prefetched.  order != null (mixed widths):
    // position l holds stripe order[l], stripes of one body class are adjacent, so a warp's groups agree.
    StripeMeta na, nb;
    if (order == nullptr) { na = ld_meta(meta + l); nb = ld_meta(meta + l + 1); }
    for (; l < L; l += ngroups) {
        StripeMeta a, b;
        if (order == nullptr) {
            a = na; b = nb;
            if (l + ngroups < L) { na = ld_meta(meta + l + ngroups); nb = ld_meta(meta + l + ngroups + 1); } // next stripe's meta rides along
        } else {
            const int ls = __ldg(order + l);
            a = ld_meta(meta + ls); b = ld_meta(meta + ls + 1);
        }
        const int w = b.col - a.col;
        if (w <= 0) continue;
#if VBC_WIDE_LD
        if (sizeof(Tv) == 8 && (w % 4) == 0 && (a.ofs % 4) == 0)
            adj_dispatch_cpr<Tv, G, MODE, (sizeof(Tv) == 8 ? 4 : VE), false>(a, b, w, lane, gmask, desc, val, x, y, dst, u0, log2u, alpha, beta);
        else
#endif
        if ((w % VE) == 0 && (a.ofs % VE) == 0)
            adj_dispatch_cpr<Tv, G, MODE, VE, false>(a, b, w, lane, gmask, desc, val, x, y, dst, u0, log2u, alpha, beta);
        else if (VE == 4 && (w % 2) == 0 && (a.ofs % 2) == 0)
            adj_dispatch_cpr<Tv, G, MODE, (VE == 4 ? 2 : 1), false>(a, b, w, lane, gmask, desc, val, x, y, dst, u0, log2u, alpha, beta);
        else
            adj_dispatch_cpr<Tv, G, MODE, 1, false>(a, b, w, lane, gmask, desc, val, x, y, dst, u0, log2u, alpha, beta);
    }
}

// ---- forward (scatter) ------------------------------------------------------------------------
template <typename Tv, int G, int MODE, int EPV, int CPR>
__device__ __forceinline__ void fwd_stripe(const StripeMeta a, const StripeMeta b, const int w, const int lane, const unsigned gmask,
                                           const int *__restrict__ desc, const Tv *__restrict__ val, const Tv *__restrict__ x,
                                           Tv *__restrict__ y, const int u0, const int log2u, const Tv alpha)
{
    int cpr, rps, c, r0;
    if constexpr (CPR > 0) { cpr = CPR; rps = G / CPR; c = lane % CPR; r0 = lane / CPR; }
    else { cpr = w / EPV; rps = small_div(G, cpr); r0 = small_div(lane, cpr); c = lane - r0 * cpr; }
    int R;
    if (MODE == DESC_ROWS) R = b.pos - a.pos;
    else { const int nv = (int)(b.ofs - a.ofs); if constexpr (CPR > 0) R = nv / (CPR * EPV); else R = nv / w; }
    const bool active = (CPR > 0) || (lane < rps * cpr);
    // tmp = x[j : j+w)  (multiply_1DVBC.jl:27)
    Tv xs[EPV];
#pragma unroll
    for (int e = 0; e < EPV; e++) xs[e] = active ? __ldg(x + a.col + c * EPV + e) : (Tv)0;
    const Tv *vp = val + a.ofs + (long long)r0 * w + c * EPV;
    const int vstride = rps * w;
    RowWalk<MODE> walk;
    walk.init(desc, a.pos, r0, rps, u0, log2u);
    for (int rb = 0; rb < R; rb += rps) { // group-uniform trip count: every lane joins the shuffles
        const bool ok = active && (rb + r0 < R);
        Tv p = (Tv)0;
        int xi = 0;
        if (ok) {
            Tv v[EPV];
            Ld<Tv, EPV>::s(vp, v);
            xi = walk.next();
#pragma unroll
            for (int e = 0; e < EPV; e++) p = fma(v[e], xs[e], p);
        }
        vp += vstride;
        if constexpr (CPR > 0) {
#pragma unroll
            for (int d = 1; d < CPR; d <<= 1) p += __shfl_xor_sync(gmask, p, d, G);
        } else {
            for (int d = 1; d < cpr; d <<= 1) {
                const Tv t = __shfl_down_sync(gmask, p, d, G);
                if (c + d < cpr) p += t;
            }
        }
        if (ok && c == 0) atomicAdd(y + xi, alpha * p); // y[idx[Q]] += ...  (multiply_1DVBC.jl:34)
    }
}

template <typename Tv, int G, int MODE>
__device__ __noinline__ void fwd_stripe_wide(const StripeMeta a, const StripeMeta b, const int w, const int lane, const unsigned gmask,
                                             const int *__restrict__ desc, const Tv *__restrict__ val, const Tv *__restrict__ x,
                                             Tv *__restrict__ y, const int u0, const Tv alpha)
{
    const int R = (MODE == DESC_ROWS) ? (b.pos - a.pos) : (int)((b.ofs - a.ofs) / w);
    RowWalk<MODE> walk;
    walk.init(desc, a.pos, 0, 1, u0, -1);
    for (int r = 0; r < R; r++) {
        const int xi = walk.next();
        Tv p = (Tv)0;
        for (int col = lane; col < w; col += G) p = fma(__ldcs(val + a.ofs + (long long)r * w + col), __ldg(x + a.col + col), p);
#pragma unroll
        for (int d = 1; d < G; d <<= 1) p += __shfl_xor_sync(gmask, p, d, G);
        if (lane == 0) atomicAdd(y + xi, alpha * p);
    }
}

template <typename Tv, int G, int MODE, int EPV>
__device__ __forceinline__ void fwd_dispatch_cpr(const StripeMeta a, const StripeMeta b, const int w, const int lane, const unsigned gmask,
                                                 const int *__restrict__ desc, const Tv *__restrict__ val, const Tv *__restrict__ x,
                                                 Tv *__restrict__ y, const int u0, const int log2u, const Tv alpha)
{
    const int cpr = w / EPV;
    if (cpr == 1) fwd_stripe<Tv, G, MODE, EPV, 1>(a, b, w, lane, gmask, desc, val, x, y, u0, log2u, alpha);
    else if (cpr == 2) fwd_stripe<Tv, G, MODE, EPV, 2>(a, b, w, lane, gmask, desc, val, x, y, u0, log2u, alpha);
    else if (cpr == 4) fwd_stripe<Tv, G, MODE, EPV, 4>(a, b, w, lane, gmask, desc, val, x, y, u0, log2u, alpha);
    else if (cpr <= G) fwd_stripe<Tv, G, MODE, EPV, 0>(a, b, w, lane, gmask, desc, val, x, y, u0, log2u, alpha);
    else fwd_stripe_wide<Tv, G, MODE>(a, b, w, lane, gmask, desc, val, x, y, u0, alpha);
}

template <typename Tv, int G, int MODE>
__global__ void __launch_bounds__(256) k_spmv_fwd(const StripeMeta *__restrict__ meta, const int *__restrict__ desc,
                                                   const Tv *__restrict__ val, const Tv *__restrict__ x, Tv *__restrict__ y,
                                                   const int *__restrict__ order, const int L, const int u0, const int log2u, const Tv alpha)
{
    constexpr int VE = 16 / (int)sizeof(Tv);
    const int lane = threadIdx.x % G;
    const unsigned gmask = group_mask<G>();
    const int ngroups = (int)((gridDim.x * blockDim.x) / G);
    for (int l = (int)((blockIdx.x * blockDim.x + threadIdx.x) / G); l < L; l += ngroups) {
        const int ls = order ? __ldg(order + l) : l;
        const StripeMeta a = ld_meta(meta + ls), b = ld_meta(meta + ls + 1);
        const int w = b.col - a.col;
        if (w <= 0 || b.ofs == a.ofs) continue;
        if ((w % VE) == 0 && (a.ofs % VE) == 0)
            fwd_dispatch_cpr<Tv, G, MODE, VE>(a, b, w, lane, gmask, desc, val, x, y, u0, log2u, alpha);
        else if (VE == 4 && (w % 2) == 0 && (a.ofs % 2) == 0)
            fwd_dispatch_cpr<Tv, G, MODE, (VE == 4 ? 2 : 1)>(a, b, w, lane, gmask, desc, val, x, y, u0, log2u, alpha);
        else
            fwd_dispatch_cpr<Tv, G, MODE, 1>(a, b, w, lane, gmask, desc, val, x, y, u0, log2u, alpha);
    }
}

// y <- beta * y (beta == 0 stores zeros without reading y)   multiply_1DVBC.jl:50-52
template <typename Tv>
__global__ void __launch_bounds__(256) k_scale(Tv *__restrict__ y, const int64_t len, const Tv beta)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (int64_t)gridDim.x * blockDim.x)
        y[i] = (beta == (Tv)0) ? (Tv)0 : beta * y[i];
}

// ---- parity mode: multiply straight from the canonical Ti arrays, the reference's loops verbatim
// (one thread per (stripe, column) for the adjoint; one thread per stripe for the forward).
template <typename Tv, typename Ti>
__global__ void __launch_bounds__(128) k_parity_adj(const Ti *__restrict__ phi, const Ti *__restrict__ pi, const Ti *__restrict__ pos,
                                                     const Ti *__restrict__ idx, const Ti *__restrict__ ofs, const Tv *__restrict__ val,
                                                     const Tv *__restrict__ x, Tv *__restrict__ y, const int64_t n, const int64_t L,
                                                     const int *__restrict__ c2s_unused, const int ndim, const Tv alpha, const Tv beta)
{
    (void)c2s_unused;
    const int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= L) return;
    const int64_t j = (int64_t)phi[l] - 1, w = (int64_t)phi[l + 1] - 1 - j;
    for (int64_t dj = 0; dj < w; dj++) {
        Tv tmp = (Tv)0;
        int64_t q = (int64_t)ofs[l] - 1;
        for (int64_t Q = (int64_t)pos[l] - 1; Q < (int64_t)pos[l + 1] - 1; Q++) {
            if (ndim == 1) { tmp = fma(val[q + dj], x[(int64_t)idx[Q] - 1], tmp); q += w; }
            else {
                const int64_t k = (int64_t)idx[Q] - 1, i = (int64_t)pi[k] - 1, u = (int64_t)pi[k + 1] - 1 - i;
                for (int64_t di = 0; di < u; di++) tmp = fma(val[q + w * di + dj], x[i + di], tmp);
                q += u * w;
            }
        }
        y[j + dj] = (beta == (Tv)0) ? alpha * tmp : alpha * tmp + beta * y[j + dj];
    }
    (void)n;
}

template <typename Tv, typename Ti>
__global__ void __launch_bounds__(128) k_parity_fwd(const Ti *__restrict__ phi, const Ti *__restrict__ pi, const Ti *__restrict__ pos,
                                                     const Ti *__restrict__ idx, const Ti *__restrict__ ofs, const Tv *__restrict__ val,
                                                     const Tv *__restrict__ x, Tv *__restrict__ y, const int64_t L, const int ndim, const Tv alpha)
{
    const int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= L) return;
    const int64_t j = (int64_t)phi[l] - 1, w = (int64_t)phi[l + 1] - 1 - j;
    int64_t q = (int64_t)ofs[l] - 1;
    for (int64_t Q = (int64_t)pos[l] - 1; Q < (int64_t)pos[l + 1] - 1; Q++) {
        int64_t i, u;
        if (ndim == 1) { i = (int64_t)idx[Q] - 1; u = 1; }
        else { const int64_t k = (int64_t)idx[Q] - 1; i = (int64_t)pi[k] - 1; u = (int64_t)pi[k + 1] - 1 - i; }
        for (int64_t di = 0; di < u; di++) {
            Tv s = (Tv)0;
            for (int64_t dj = 0; dj < w; dj++) s = fma(val[q + w * di + dj], x[j + dj], s);
            atomicAdd(y + i + di, alpha * s);
        }
        q += u * w;
    }
}

// ---- launchers ------------------------------------------------------------------------------
static int ilog2_exact(int v)
{
    if (v <= 0 || (v & (v - 1))) return -1;
    int l = 0;
    while ((1 << l) < v) l++;
    return l;
}

static int auto_group(const vbc_mat *A)
{
    // ~ (bytes of an average stripe) / 16 B vectors; aim for >= 4 vectors per lane
    if (A->L == 0) return 8;
    // tools/perf_table.py: 4 lanes win below ~80 16-byte vectors per stripe (52 -> 36.7 vs 38.5 us), 8 lanes from
    // ~100 to a few hundred (104 -> 68.2 vs 69.7; 200 -> 71.5 vs 75.0 at 32 lanes)
    const double vec_per_stripe = (double)A->nval * (double)vt_size(A->vt) / 16.0 / (double)A->L;
    int G = vec_per_stripe < 80.0 ? 4 : (vec_per_stripe < 512.0 ? 8 : 32);
    // few stripes (C1: L = 1250): widen the groups until the grid has ~1 CTA of 256 threads per SM x 4, as long as
    // a stripe still has >= 2 vectors per lane (4.0 vs 7.1 us on C1 at 32 vs 8 lanes)
    while (G < 32 && (double)A->L * G < 148.0 * 4 * 256 && vec_per_stripe >= 4.0 * G) G *= 2;
    return G;
}

template <typename Tv, int G, int MODE, bool PEER>
static int launch_adj_t(vbc_mat *A, Tv alpha, const Tv *x, Tv beta, Tv *y, const PeerDst &dst)
{
    int occ = 0;
    VBC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_spmv_adj<Tv, G, MODE, PEER>, 256, 0));
    if (occ < 1) occ = 1;
    if (A->opt_grid_mult > 0) occ = A->opt_grid_mult;
    int64_t grid = (int64_t)A->sm_count * occ;
    // a stripe range [l0, l1) is launched by offsetting meta: its entries are absolute (value offset, descriptor, column)
    const bool ranged = !PEER && A->range_l0 >= 0 && A->d_order == nullptr;
    const int64_t l0 = ranged ? A->range_l0 : 0, l1 = ranged ? A->range_l1 : A->L;
    const int64_t need = ((l1 - l0) * G + 255) / 256;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    k_spmv_adj<Tv, G, MODE, PEER><<<(unsigned)grid, 256, 0, A->stream>>>(A->d_meta + l0, A->d_desc, (const Tv *)A->d_val, x, y, dst, PEER ? nullptr : A->d_order, (int)(l1 - l0), A->u0, ilog2_exact(A->u0), alpha, beta);
    A->launches++;
    VBC_CUDA(cudaGetLastError());
    return VBC_OK;
}

template <typename Tv, int G, int MODE>
static int launch_fwd_t(vbc_mat *A, Tv alpha, const Tv *x, Tv *y)
{
    int occ = 0;
    VBC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_spmv_fwd<Tv, G, MODE>, 256, 0));
    if (occ < 1) occ = 1;
    if (A->opt_grid_mult > 0) occ = A->opt_grid_mult;
    int64_t grid = (int64_t)A->sm_count * occ;
    const int64_t need = (A->L * G + 255) / 256;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    k_spmv_fwd<Tv, G, MODE><<<(unsigned)grid, 256, 0, A->stream>>>(A->d_meta, A->d_desc, (const Tv *)A->d_val, x, y, A->d_order, (int)A->L, A->u0, ilog2_exact(A->u0), alpha);
    A->launches++;
    VBC_CUDA(cudaGetLastError());
    return VBC_OK;
}

template <typename Tv, bool PEER>
static int launch_adj_any(vbc_mat *A, Tv alpha, const Tv *x, Tv beta, Tv *y, const PeerDst &dst)
{
    const bool rows = A->desc_mode == DESC_ROWS;
    const int G = A->opt_adj_group ? A->opt_adj_group : auto_group(A);
    if (G >= 32) return rows ? launch_adj_t<Tv, 32, DESC_ROWS, PEER>(A, alpha, x, beta, y, dst) : launch_adj_t<Tv, 32, DESC_BLOCKS, PEER>(A, alpha, x, beta, y, dst);
    if (G >= 16) return rows ? launch_adj_t<Tv, 16, DESC_ROWS, PEER>(A, alpha, x, beta, y, dst) : launch_adj_t<Tv, 16, DESC_BLOCKS, PEER>(A, alpha, x, beta, y, dst);
    if (G >= 8) return rows ? launch_adj_t<Tv, 8, DESC_ROWS, PEER>(A, alpha, x, beta, y, dst) : launch_adj_t<Tv, 8, DESC_BLOCKS, PEER>(A, alpha, x, beta, y, dst);
    return rows ? launch_adj_t<Tv, 4, DESC_ROWS, PEER>(A, alpha, x, beta, y, dst) : launch_adj_t<Tv, 4, DESC_BLOCKS, PEER>(A, alpha, x, beta, y, dst);
}

template <typename Tv>
static int scale_y(vbc_mat *A, Tv *y, int64_t len, Tv beta)
{
    if (len == 0 || beta == (Tv)1) return VBC_OK;
    if (beta == (Tv)0) { VBC_CUDA(cudaMemsetAsync(y, 0, sizeof(Tv) * (size_t)len, A->stream)); return VBC_OK; }
    int64_t g = (len + 255) / 256;
    if (g > (int64_t)A->sm_count * 8) g = (int64_t)A->sm_count * 8;
    k_scale<Tv><<<(unsigned)g, 256, 0, A->stream>>>(y, len, beta);
    A->launches++;
    VBC_CUDA(cudaGetLastError());
    return VBC_OK;
}

template <typename Tv, typename Ti>
static int launch_parity(vbc_mat *A, int trans, Tv alpha, const Tv *x, Tv beta, Tv *y)
{
    const unsigned g = (unsigned)((A->L + 127) / 128 > 0 ? (A->L + 127) / 128 : 1);
    if (trans) {
        if (A->L > 0) {
            k_parity_adj<Tv, Ti><<<g, 128, 0, A->stream>>>((const Ti *)A->d_phi_spl, (const Ti *)A->d_pi_spl, (const Ti *)A->d_pos, (const Ti *)A->d_idx,
                                                           (const Ti *)A->d_ofs, (const Tv *)A->d_val, x, y, A->n, A->L, nullptr, A->ndim, alpha, beta);
            A->launches++;
        }
    } else {
        VBC_TRY(scale_y<Tv>(A, y, A->m, beta));
        if (A->L > 0) {
            k_parity_fwd<Tv, Ti><<<g, 128, 0, A->stream>>>((const Ti *)A->d_phi_spl, (const Ti *)A->d_pi_spl, (const Ti *)A->d_pos, (const Ti *)A->d_idx,
                                                           (const Ti *)A->d_ofs, (const Tv *)A->d_val, x, y, A->L, A->ndim, alpha);
            A->launches++;
        }
    }
    VBC_CUDA(cudaGetLastError());
    return VBC_OK;
}

template <typename Tv>
static int launch_spmv_t(vbc_mat *A, int trans, double alpha_d, const void *xv, double beta_d, void *yv)
{
    const Tv alpha = (Tv)alpha_d, beta = (Tv)beta_d;
    const Tv *x = (const Tv *)xv;
    Tv *y = (Tv *)yv;
    if (A->opt_parity) {
        if (A->it == VBC_I64) return launch_parity<Tv, int64_t>(A, trans, alpha, x, beta, y);
        return launch_parity<Tv, int32_t>(A, trans, alpha, x, beta, y);
    }
    const bool rows = A->desc_mode == DESC_ROWS;
    if (trans) {
        if (A->L == 0) return VBC_OK; // n == 0: nothing to write
        return launch_adj_any<Tv, false>(A, alpha, x, beta, y, PeerDst{});
    }
    if (A->opt_fwd_atomic != 1) { // owner-computes forward through the transposed unit index (fwdt.cu)
        VBC_TRY(ensure_tindex(A));
        if (A->tindex && (A->opt_fwd_atomic == 2 || A->desc_mode == DESC_BLOCKS)) return launch_fwdt(A, alpha_d, xv, beta_d, yv);
    }
    VBC_TRY(scale_y<Tv>(A, y, A->m, beta));
    if (A->L == 0 || A->nval == 0) return VBC_OK;
    int G = A->opt_fwd_group ? A->opt_fwd_group : 32; // the scatter kernel likes whole warps per stripe (110 vs 131 us on C3)
    if (G < 8) G = 8;
    if (G >= 32) return rows ? launch_fwd_t<Tv, 32, DESC_ROWS>(A, alpha, x, y) : launch_fwd_t<Tv, 32, DESC_BLOCKS>(A, alpha, x, y);
    return rows ? launch_fwd_t<Tv, 8, DESC_ROWS>(A, alpha, x, y) : launch_fwd_t<Tv, 8, DESC_BLOCKS>(A, alpha, x, y);
}

// adjoint multiply whose result goes to `n` destination buffers (each already offset to this
// rank's first column): the compute half of vbc_peer_spmv_step.
int launch_spmv_adj_peer(vbc_mat *A, double alpha, const void *d_x, int n, void *const *dst_ptrs, const unsigned char *d_mask, int chunk_shift,
                         const PeerSyncArgs *sync, const int *ranges)
{
    if (A->opt_parity) VBC_FAIL(VBC_EARG, "peer multiply needs the compact layout (parity mode is on)");
    if (n < 1 || n > VBC_MAX_PEERS) VBC_FAIL(VBC_EARG, "peer count %d out of 1..%d", n, VBC_MAX_PEERS);
    if (A->L == 0) return VBC_OK;
    PeerDst dst{};
    dst.n = n;
    for (int i = 0; i < VBC_MAX_PEERS; i++) dst.p[i] = i < n ? dst_ptrs[i] : nullptr;
    dst.mask = d_mask;
    dst.chunk_shift = chunk_shift;
    dst.sync_n = 0;
    dst.ra0 = 0; dst.ra1 = (int)A->L; dst.rb0 = dst.rb1 = 0;
    if (ranges) { dst.ra0 = ranges[0]; dst.ra1 = ranges[1]; dst.rb0 = ranges[2]; dst.rb1 = ranges[3]; }
    if (sync) {
        dst.sync_n = sync->nranks; dst.me = sync->me;
        for (int r = 0; r < VBC_MAX_PEERS; r++) dst.flags[r] = r < sync->nranks ? sync->flags[r] : nullptr;
        dst.d_epoch = sync->d_epoch; dst.d_done = sync->d_done; dst.timed_out = sync->timed_out;
        dst.i0 = sync->i0 < 0 ? 0 : (sync->i0 > (int)A->L ? (int)A->L : sync->i0);
        dst.i1 = sync->i1 < dst.i0 ? dst.i0 : (sync->i1 > (int)A->L ? (int)A->L : sync->i1);
    }
    if (A->vt == VBC_F64) return launch_adj_any<double, true>(A, alpha, (const double *)d_x, 0.0, nullptr, dst);
    return launch_adj_any<float, true>(A, (float)alpha, (const float *)d_x, 0.0f, nullptr, dst);
}

int launch_spmv(vbc_mat *A, int trans, double alpha, const void *d_x, double beta, void *d_y)
{
    return A->vt == VBC_F64 ? launch_spmv_t<double>(A, trans, alpha, d_x, beta, d_y) : launch_spmv_t<float>(A, trans, alpha, d_x, beta, d_y);
}

} // namespace vbc
