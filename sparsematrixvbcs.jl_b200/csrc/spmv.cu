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
#include <stdlib.h>

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
#ifndef VBC_HALO_RUN_INLINE
#define VBC_HALO_RUN_INLINE __forceinline__ // the boundary-run body of k_spmv_adj_halo: inlined (a call in that kernel slows its interior loop)
#endif
#ifndef VBC_FLAT_ALL
#define VBC_FLAT_ALL 0    // 1: in the FLAT kernels every stripe (aligned ones too) takes the flat-slab body
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

// Row-partitioned multiply with the x exchange fused into the kernel (peer.cu, vbc_peer_spmv_step).
// Destination 0 is this rank's own next-x buffer, destination i the buffer of rank (me + i) % n, mapped
// over NVLink; all pointers are already offset to this rank's first column.
struct HaloArgs {
    int n;                                   // destinations
    void *p[VBC_MAX_PEERS];
    const unsigned char *mask;               // optional: mask[col >> chunk_shift] bit i set <=> destination i reads that column chunk
    int chunk_shift;
    int i0, i1;                              // interior stripes [i0, i1): gather only from this rank's slice, feed only this rank
    int nrunsA, nruns;                       // boundary runs of 32/G adjacent stripes: [0, nrunsA) cover [0, i0), the rest cover [i1, L)
    int first_warp, nbwarps;                 // boundary run r belongs to warp (first_warp + r) % warps; nbwarps = min(nruns, warps) warps own any
    int ifence, imid;                        // interior is walked as [i0, ifence) (then boundary warps count themselves done), [ifence, imid) by
                                             // every warp and [imid, i1) by the warps without boundary runs
    int stage_elems;                         // staging row per warp, in elements: 32/G stripes x W columns
    int me, nranks, do_wait, do_signal;
    unsigned nbr_mask;                       // ranks this rank exchanges flags with
    unsigned long long *flags[VBC_MAX_PEERS]; // flag block of every rank (flags[me] is local)
    unsigned long long *ctl;                 // [1] boundary warps finished (running total), [2] epoch (steps published so far), [3..5] wait statistics
    int *timed_out;
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
__device__ __forceinline__ unsigned long long ld_relaxed_gpu_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// x gathers.  XC = false: read-only path (x does not change while the kernel runs).  XC = true: the halo part of x is
// written by other GPUs while this kernel runs (and only read after an acquire of their flag): L2 is the point of
// coherence for those peer writes, so the load must not be served from a line L1 might still hold.
template <typename Tv, bool XC> __device__ __forceinline__ Tv ld_x(const Tv *p)
{
    if constexpr (XC) return __ldcg(p);
    else return __ldg(p);
}

template <typename Tv, int EPV>
__device__ __forceinline__ void store_y(Tv *__restrict__ y, const int col, const Tv (&acc)[EPV], const Tv alpha, const Tv beta)
{
    Tv *yp = y + col;
#pragma unroll
    for (int e = 0; e < EPV; e++) yp[e] = (beta == (Tv)0) ? alpha * acc[e] : alpha * acc[e] + beta * yp[e];
}

// ---- adjoint ----------------------------------------------------------------------------------
// CPR > 0: compile-time vectors per row (power of two, <= G).  CPR == 0: runtime cpr <= G.
template <typename Tv, int G, int MODE, int EPV, int CPR, bool XC>
__device__ __forceinline__ void adj_stripe(const StripeMeta a, const StripeMeta b, const int w, const int lane, const unsigned gmask,
                                           const int *__restrict__ desc, const Tv *__restrict__ val, const Tv *__restrict__ x,
                                           Tv *__restrict__ y, const int u0, const int log2u, const Tv alpha, const Tv beta)
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
        for (int k = 0; k < UNR; k++) xv[k] = ok[k] ? ld_x<Tv, XC>(x + xi[k]) : (Tv)0;
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
    if (lane < cpr) store_y<Tv, EPV>(y, a.col + lane * EPV, acc, alpha, beta);
}

// Unaligned stripes (Float64: odd width, or a slab that starts on an odd element): the slab is read as what it is in
// memory -- one contiguous array -- with ALIGNED 16-byte loads, whatever the row length.  S = per * floor(G / per) lanes
// (per = w for odd w, w/2 for even w) read consecutive vectors; 2S elements are a whole number of rows, so the two elements
// of a lane always fall on the same two columns (c0, c1) and on rows that advance by a fixed step.  The x values of the
// stripe's rows are gathered ONCE per row into the group's shared-memory row (one descriptor and one x load per stored row
// instead of one per element) and read back per element.  The element before the slab (odd start) and the one after it are
// masked.  Lanes that share (c0, c1) are summed with the strided shuffle tree of the generic body; for odd w column j is
// the sum of one lane's first and another lane's second accumulator (one more shuffle).
// One 8-byte load per element with its own descriptor and x loads was 0.58 of the HBM peak on the C2v matrix.
template <int G, int MODE, bool XC>
__device__ __forceinline__ void adj_stripe_flat(const StripeMeta a, const StripeMeta b, const int w, const int R, const int lane, const unsigned gmask,
                                                const int *__restrict__ desc, const double *__restrict__ val, const double *__restrict__ x,
                                                double *__restrict__ y, const int u0, const int log2u, const double alpha, const double beta,
                                                double *__restrict__ xs)
{
    const int n = (int)(b.ofs - a.ofs);
    const int shift = (int)(a.ofs & 1);
    const int odd = w & 1;
    const int per = odd ? w : (w >> 1);
    const int reps = small_div(G, per);
    const int S = reps * per;             // active lanes
    const int rstep = odd ? 2 * reps : reps; // rows per step (2S / w)
    // this lane's two elements in step 0, relative to the slab: q0 (-1: the element before an odd-aligned slab), q0 + 1
    const int q0 = 2 * lane - shift;
    int ra, c0;
    if (q0 < 0) { ra = -1; c0 = w - 1; } else { ra = small_div(q0, w); c0 = q0 - ra * w; }
    int rb = small_div(q0 + 1, w);
    const int c1 = q0 + 1 - rb * w;
    const double2 *vp = reinterpret_cast<const double2 *>(val + (a.ofs - shift)) + lane;
    double acc0 = 0.0, acc1 = 0.0;
    constexpr int UNR = VBC_ADJ_UNR;
    int p = lane < S ? q0 : n;
    double2 v[UNR];
    auto load_batch = [&]() {
#pragma unroll
        for (int k = 0; k < UNR; k++) {
            v[k] = (p + k * 2 * S < n) ? ld_val(vp) : make_double2(0.0, 0.0);
            vp += S;
        }
    };
    load_batch(); // the first values are on their way while the descriptors and x values of the rows are fetched
    for (int i = lane; i < R; i += G) xs[i] = ld_x<double, XC>(x + row_xindex<MODE>(desc, a.pos, i, u0, log2u));
    __syncwarp(gmask);
    while (p < n) {
#pragma unroll
        for (int k = 0; k < UNR; k++) {
            const int pk = p + k * 2 * S;
            const bool m0 = pk >= 0 && pk < n, m1 = pk + 1 < n;
            const double x0 = m0 ? xs[ra + k * rstep] : 0.0, x1 = m1 ? xs[rb + k * rstep] : 0.0;
            if (m0) acc0 = fma(v[k].x, x0, acc0);
            if (m1) acc1 = fma(v[k].y, x1, acc1);
        }
        ra += UNR * rstep; rb += UNR * rstep;
        p += UNR * 2 * S;
        if (p < n) load_batch();
    }
    for (int d = per; d < G; d <<= 1) {
        const double t0 = __shfl_down_sync(gmask, acc0, d, G), t1 = __shfl_down_sync(gmask, acc1, d, G);
        if (lane + d < G) { acc0 += t0; acc1 += t1; }
    }
    double *yp = y + a.col;
    if (odd) { // column c0 = this lane's first accumulator + the second accumulator of lane (lane + (w-1)/2) mod w
        int src = lane + ((w - 1) >> 1);
        if (src >= w) src -= w;
        const double t = __shfl_sync(gmask, acc1, src, G);
        if (lane < w) { const double s = acc0 + t; yp[c0] = (beta == 0.0) ? alpha * s : alpha * s + beta * yp[c0]; }
    } else if (lane < per) {
        yp[c0] = (beta == 0.0) ? alpha * acc0 : alpha * acc0 + beta * yp[c0];
        yp[c1] = (beta == 0.0) ? alpha * acc1 : alpha * acc1 + beta * yp[c1];
    }
    __syncwarp(gmask); // the group's x row is rewritten by the next stripe
}

// The same for Float32 (four elements per 16-byte vector): a stripe is unaligned when its width or its start is not a multiple
// of four.  per = w / gcd(w, 4) lanes make a period (4 * per elements = 4 / gcd whole rows), every lane keeps four accumulators
// with fixed columns; after the strided shuffle tree the period's 4 * per partial sums go through the group's shared-memory row
// (free once the loop is done) and lane j adds the 4 / gcd of them that belong to column j.
template <int G, int MODE, bool XC>
__device__ __forceinline__ void adj_stripe_flat4(const StripeMeta a, const StripeMeta b, const int w, const int R, const int lane, const unsigned gmask,
                                                 const int *__restrict__ desc, const float *__restrict__ val, const float *__restrict__ x,
                                                 float *__restrict__ y, const int u0, const int log2u, const float alpha, const float beta,
                                                 float *__restrict__ xs)
{
    const int n = (int)(b.ofs - a.ofs);
    const int shift = (int)(a.ofs & 3);
    const int lg = (w & 3) == 0 ? 2 : ((w & 1) == 0 ? 1 : 0); // log2 gcd(w, 4)
    const int per = w >> lg;
    const int reps = small_div(G, per);
    const int S = reps * per;            // active lanes
    const int m = 4 >> lg;               // rows per period = partial sums per column
    const int rstep = reps * m;          // rows per step (4S / w)
    const int q0 = 4 * lane - shift;     // this lane's first element in step 0, relative to the slab (negative: before the slab)
    int r[4];
#pragma unroll
    for (int e = 0; e < 4; e++) r[e] = q0 + e < 0 ? -small_div(w - 1 - (q0 + e), w) : small_div(q0 + e, w); // floor((q0 + e) / w): up to three elements precede the slab
    const float4 *vp = reinterpret_cast<const float4 *>(val + (a.ofs - shift)) + lane;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    constexpr int UNR = VBC_ADJ_UNR;
    int p = lane < S ? q0 : n;
    float4 v[UNR];
    auto load_batch = [&]() {
#pragma unroll
        for (int k = 0; k < UNR; k++) {
            v[k] = (p + k * 4 * S < n) ? ld_val(vp) : make_float4(0.f, 0.f, 0.f, 0.f);
            vp += S;
        }
    };
    load_batch();
    for (int i = lane; i < R; i += G) xs[i] = ld_x<float, XC>(x + row_xindex<MODE>(desc, a.pos, i, u0, log2u));
    __syncwarp(gmask);
    while (p < n) {
#pragma unroll
        for (int k = 0; k < UNR; k++) {
            const int pk = p + k * 4 * S;
            const float ve[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const bool ok = pk + e >= 0 && pk + e < n;
                const float xv = ok ? xs[r[e] + k * rstep] : 0.f;
                if (ok) acc[e] = fmaf(ve[e], xv, acc[e]);
            }
        }
#pragma unroll
        for (int e = 0; e < 4; e++) r[e] += UNR * rstep;
        p += UNR * 4 * S;
        if (p < n) load_batch();
    }
    __syncwarp(gmask);
    for (int d = per; d < G; d <<= 1) {
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const float t = __shfl_down_sync(gmask, acc[e], d, G);
            if (lane + d < G) acc[e] += t;
        }
    }
    if (lane < per) {
#pragma unroll
        for (int e = 0; e < 4; e++) xs[4 * lane + e] = acc[e]; // partial sum of period element u = 4 * lane + e, column (u - shift) mod w
    }
    __syncwarp(gmask);
    float *yp = y + a.col;
    for (int j = lane; j < w; j += G) {
        int u = j + shift;
        u -= small_div(u, w) * w;
        float sum = 0.f;
        for (int t = 0; t < m; t++) sum += xs[u + t * w];
        yp[j] = (beta == 0.f) ? alpha * sum : alpha * sum + beta * yp[j];
    }
    __syncwarp(gmask); // the group's row is rewritten by the next stripe
}

// wide stripes (more vectors per row than lanes): one lane per column, serial over rows
template <typename Tv, int G, int MODE, bool XC>
__device__ __noinline__ void adj_stripe_wide(const StripeMeta a, const StripeMeta b, const int w, const int lane,
                                             const int *__restrict__ desc, const Tv *__restrict__ val, const Tv *__restrict__ x,
                                             Tv *__restrict__ y, const int u0, const Tv alpha, const Tv beta)
{
    const int R = (MODE == DESC_ROWS) ? (b.pos - a.pos) : (int)((b.ofs - a.ofs) / w);
    for (int col = lane; col < w; col += G) {
        Tv acc = (Tv)0;
        RowWalk<MODE> walk;
        walk.init(desc, a.pos, 0, 1, u0, -1);
        const Tv *vp = val + a.ofs + col;
        for (int r = 0; r < R; r++) { acc = fma(__ldcs(vp), ld_x<Tv, XC>(x + walk.next()), acc); vp += w; }
        const Tv accv[1] = {acc};
        store_y<Tv, 1>(y, a.col + col, accv, alpha, beta);
    }
}

template <typename Tv, int G, int MODE, int EPV, bool XC>
__device__ __forceinline__ void adj_dispatch_cpr(const StripeMeta a, const StripeMeta b, const int w, const int lane, const unsigned gmask,
                                                 const int *__restrict__ desc, const Tv *__restrict__ val, const Tv *__restrict__ x,
                                                 Tv *__restrict__ y, const int u0, const int log2u, const Tv alpha, const Tv beta)
{
    const int cpr = w / EPV;
    if (cpr == 1) adj_stripe<Tv, G, MODE, EPV, 1, XC>(a, b, w, lane, gmask, desc, val, x, y, u0, log2u, alpha, beta);
    else if (cpr == 2) adj_stripe<Tv, G, MODE, EPV, 2, XC>(a, b, w, lane, gmask, desc, val, x, y, u0, log2u, alpha, beta);
    else if (cpr == 4) adj_stripe<Tv, G, MODE, EPV, 4, XC>(a, b, w, lane, gmask, desc, val, x, y, u0, log2u, alpha, beta);
    else if (cpr <= G) adj_stripe<Tv, G, MODE, EPV, 0, XC>(a, b, w, lane, gmask, desc, val, x, y, u0, log2u, alpha, beta);
    else adj_stripe_wide<Tv, G, MODE, XC>(a, b, w, lane, desc, val, x, y, u0, alpha, beta);
}

// one stripe, body class chosen from its width and the alignment of its slab (the analogue of le_nest, util.jl:28-38)
constexpr int FLAT_XCAP_PER_LANE = 16; // shared-memory x row of a group in the FLAT kernels: 16 * G values (32 KB per CTA of 256 threads)

template <typename Tv, int G, int MODE, bool XC, bool FLAT = false>
__device__ __forceinline__ void adj_one_stripe(const StripeMeta a, const StripeMeta b, const int lane, const unsigned gmask,
                                               const int *__restrict__ desc, const Tv *__restrict__ val, const Tv *__restrict__ x,
                                               Tv *__restrict__ y, const int u0, const int log2u, const Tv alpha, const Tv beta, Tv *__restrict__ xs = nullptr)
{
    constexpr int VE = 16 / (int)sizeof(Tv);
    const int w = b.col - a.col;
    if (w <= 0) return;
    if constexpr (FLAT && sizeof(Tv) == 4) {
        if ((w & 3) || (a.ofs & 3)) {
            const int R = (MODE == DESC_ROWS) ? (b.pos - a.pos) : (int)((b.ofs - a.ofs) / w);
            const int per = (w & 3) == 0 ? (w >> 2) : ((w & 1) == 0 ? (w >> 1) : w);
            if (per <= G && R <= FLAT_XCAP_PER_LANE * G) {
                adj_stripe_flat4<G, MODE, XC>(a, b, w, R, lane, gmask, desc, (const float *)val, (const float *)x, (float *)y, u0, log2u, (float)alpha, (float)beta, (float *)xs);
                return;
            }
        }
    }
    if constexpr (FLAT && sizeof(Tv) == 8) {
        if (VBC_FLAT_ALL || (w & 1) || (a.ofs & 1)) {
            const int R = (MODE == DESC_ROWS) ? (b.pos - a.pos) : (int)((b.ofs - a.ofs) / w);
            if (((w & 1) ? w : (w >> 1)) <= G && R <= FLAT_XCAP_PER_LANE * G) {
                adj_stripe_flat<G, MODE, XC>(a, b, w, R, lane, gmask, desc, (const double *)val, (const double *)x, (double *)y, u0, log2u, (double)alpha, (double)beta, (double *)xs);
                return;
            }
        }
    }
#if VBC_WIDE_LD
    if (sizeof(Tv) == 8 && (w % 4) == 0 && (a.ofs % 4) == 0)
        adj_dispatch_cpr<Tv, G, MODE, (sizeof(Tv) == 8 ? 4 : VE), XC>(a, b, w, lane, gmask, desc, val, x, y, u0, log2u, alpha, beta);
    else
#endif
    if ((w % VE) == 0 && (a.ofs % VE) == 0)
        adj_dispatch_cpr<Tv, G, MODE, VE, XC>(a, b, w, lane, gmask, desc, val, x, y, u0, log2u, alpha, beta);
    else if (VE == 4 && (w % 2) == 0 && (a.ofs % 2) == 0)
        adj_dispatch_cpr<Tv, G, MODE, (VE == 4 ? 2 : 1), XC>(a, b, w, lane, gmask, desc, val, x, y, u0, log2u, alpha, beta);
    else
        adj_dispatch_cpr<Tv, G, MODE, 1, XC>(a, b, w, lane, gmask, desc, val, x, y, u0, log2u, alpha, beta);
}

// stripes [lo, hi) dealt round-robin to the groups of the grid.  order == null: stripes in index order, next stripe's
// meta prefetched.  order != null (mixed widths): position l holds stripe order[l], stripes of one body class are
// adjacent, so a warp's groups agree.  (One loop for both, so the stripe bodies exist once in the kernel.)
template <typename Tv, int G, int MODE, bool FLAT = false>
__device__ __forceinline__ void adj_range(const StripeMeta *__restrict__ meta, const int *__restrict__ order, const int lo, const int hi,
                                          const int group, const int ngroups,
                                          const int *__restrict__ desc, const Tv *__restrict__ val, const Tv *__restrict__ x,
                                          Tv *__restrict__ y, const int u0, const int log2u, const Tv alpha, const Tv beta, Tv *__restrict__ xs = nullptr)
{
    const int lane = threadIdx.x % G;
    const unsigned gmask = group_mask<G>();
    int l = lo + group;
    if (l >= hi) return;
    if constexpr (FLAT) { // the kernels for matrices with unaligned stripes: meta prefetched in both modes (order mode: through the index fetched one stripe earlier)
        StripeMeta na, nb;
        int lsn = 0;
        if (order == nullptr) { na = ld_meta(meta + l); nb = ld_meta(meta + l + 1); }
        else {
            const int ls = __ldg(order + l);
            na = ld_meta(meta + ls); nb = ld_meta(meta + ls + 1);
            if (l + ngroups < hi) lsn = __ldg(order + l + ngroups);
        }
        for (; l < hi; l += ngroups) {
            const StripeMeta a = na, b = nb;
            if (l + ngroups < hi) {
                if (order == nullptr) { na = ld_meta(meta + l + ngroups); nb = ld_meta(meta + l + ngroups + 1); }
                else {
                    na = ld_meta(meta + lsn); nb = ld_meta(meta + lsn + 1);
                    if (l + 2 * ngroups < hi) lsn = __ldg(order + l + 2 * ngroups);
                }
            }
            adj_one_stripe<Tv, G, MODE, false, true>(a, b, lane, gmask, desc, val, x, y, u0, log2u, alpha, beta, xs);
        }
        return;
    }
    StripeMeta na, nb;
    if (order == nullptr) { na = ld_meta(meta + l); nb = ld_meta(meta + l + 1); }
    for (; l < hi; l += ngroups) {
        StripeMeta a, b;
        if (order == nullptr) {
            a = na; b = nb;
            if (l + ngroups < hi) { na = ld_meta(meta + l + ngroups); nb = ld_meta(meta + l + ngroups + 1); } // next stripe's meta rides along
        } else {
            const int ls = __ldg(order + l);
            a = ld_meta(meta + ls); b = ld_meta(meta + ls + 1);
        }
        adj_one_stripe<Tv, G, MODE, false>(a, b, lane, gmask, desc, val, x, y, u0, log2u, alpha, beta);
    }
}

template <typename Tv, int G, int MODE>
__global__ void VBC_ADJ_BOUNDS k_spmv_adj(const StripeMeta *__restrict__ meta, const int *__restrict__ desc,
                                          const Tv *__restrict__ val, const Tv *__restrict__ x, Tv *__restrict__ y,
                                          const int *__restrict__ order, const int L, const int u0, const int log2u, const Tv alpha, const Tv beta)
{
    adj_range<Tv, G, MODE>(meta, order, 0, L, (int)((blockIdx.x * blockDim.x + threadIdx.x) / G), (int)((gridDim.x * blockDim.x) / G),
                           desc, val, x, y, u0, log2u, alpha, beta);
}

// the same loop for matrices that hold unaligned stripes: those take adj_stripe_flat / adj_stripe_flat4, with one shared-memory x row per group
template <typename Tv, int G, int MODE>
__global__ void VBC_ADJ_BOUNDS k_spmv_adj_flat(const StripeMeta *__restrict__ meta, const int *__restrict__ desc,
                                               const Tv *__restrict__ val, const Tv *__restrict__ x, Tv *__restrict__ y,
                                               const int *__restrict__ order, const int L, const int u0, const int log2u, const Tv alpha, const Tv beta)
{
    __shared__ Tv xs_all[256 / G][FLAT_XCAP_PER_LANE * G];
    adj_range<Tv, G, MODE, true>(meta, order, 0, L, (int)((blockIdx.x * blockDim.x + threadIdx.x) / G), (int)((gridDim.x * blockDim.x) / G),
                                 desc, val, x, y, u0, log2u, alpha, beta, xs_all[threadIdx.x / G]);
}

// LONG stripes (pack.cu, build_class_order: more than 2 K values and more than 16 times the average stripe -- a dense column
// group, or a dense row of A in the transposed copy): one CTA per stripe instead of one group of lanes.  T = 256 - 256 % w
// threads read consecutive values of the slab (coalesced); T is a multiple of w, so a thread keeps one column and walks rows
// r0, r0 + T/w, ...; the per-thread sums of a column are added up in shared memory in a fixed order (deterministic).
template <typename Tv, int MODE>
__global__ void __launch_bounds__(256) k_spmv_adj_long(const StripeMeta *__restrict__ meta, const int *__restrict__ stripes, const int *__restrict__ desc,
                                                        const Tv *__restrict__ val, const Tv *__restrict__ x, Tv *__restrict__ y,
                                                        const int u0, const int log2u, const Tv alpha, const Tv beta)
{
    __shared__ Tv part[256];
    const int l = __ldg(stripes + blockIdx.x);
    const StripeMeta a = ld_meta(meta + l), b = ld_meta(meta + l + 1);
    const int w = b.col - a.col;
    if (w <= 0 || w > 256) return; // (wider than a CTA: cannot be a packed stripe; uploaded ones stay with the generic kernel's result)
    const int R = (MODE == DESC_ROWS) ? (b.pos - a.pos) : (int)((b.ofs - a.ofs) / w);
    const int T = 256 - 256 % w, rstep = T / w;
    const int tid = (int)threadIdx.x, c = tid % w;
    Tv acc = (Tv)0;
    if (tid < T) {
        const Tv *vp = val + a.ofs + tid;
        constexpr int UNR = 8; // independent row-steps in flight per thread
        for (int r = tid / w; r < R; r += UNR * rstep) {
            Tv v[UNR], xv[UNR];
            int xi[UNR];
#pragma unroll
            for (int k = 0; k < UNR; k++) { const bool ok = r + k * rstep < R; v[k] = ok ? __ldcs(vp + (long long)k * T) : (Tv)0; xi[k] = ok ? row_xindex<MODE>(desc, a.pos, r + k * rstep, u0, log2u) : -1; }
#pragma unroll
            for (int k = 0; k < UNR; k++) xv[k] = xi[k] >= 0 ? __ldg(x + xi[k]) : (Tv)0;
#pragma unroll
            for (int k = 0; k < UNR; k++) acc = fma(v[k], xv[k], acc);
            vp += (long long)UNR * T;
        }
    }
    part[tid] = acc;
    __syncthreads();
    if (tid < w) {
        Tv s = (Tv)0;
        for (int t = tid; t < T; t += w) s += part[t];
        Tv *yp = y + a.col + c;
        *yp = (beta == (Tv)0) ? alpha * s : alpha * s + beta * *yp;
    }
}

// ---- adjoint with the x exchange of the row-partitioned iteration fused in (north_star (e)) -------------------------
// One launch per iteration x_{t+1} <- alpha * A' x_t on this rank's stripes.
//   boundary stripes (for a banded operator the few thousand at both ends of the slab) read x entries that the neighbours
//     produced in their previous step and produce entries the neighbours read in their next one.  They are cut into runs
//     of 32/G adjacent stripes and dealt to the warps as the continuation of the interior round-robin, but every warp
//     runs its boundary runs FIRST: the halo leaves for the neighbours in the first microseconds of the launch, spread
//     over as many warps as there are runs, while all other warps already stream interior stripes.  A boundary warp
//     first waits (acquire, system scope) until every neighbour has published the previous step -- that happened a whole
//     kernel ago, the wait only spins when the ranks have drifted by more than one step -- computes its stripes with
//     L2-coherent x loads, stages each run's results in shared memory and stores them with full-warp coalesced stores into
//     the own next-x buffer and into the buffers of exactly the ranks that gather from those columns (mask), over NVLink.
//     The warp that finishes last publishes the step to the neighbours (release, system scope).
//   interior stripes [i0, i1) gather only from this rank's own slice of x and feed only this rank: exactly the loop of
//     the plain kernel, storing into the own next-x buffer.  Nothing here depends on another GPU.
// The epoch (ctl[2]) is read by a boundary warp before it counts itself done and written only after all have: no reset,
// no host-side counter, so a captured CUDA graph of steps replays correctly.
template <typename Tv, int G, int MODE>
__device__ VBC_HALO_RUN_INLINE void adj_boundary_run(const StripeMeta *__restrict__ meta, const int lbase, const int lend, const int *__restrict__ desc,
                                              const Tv *__restrict__ val, const Tv *__restrict__ x, Tv *__restrict__ stage, const HaloArgs &h,
                                              const int u0, const int log2u, const Tv alpha)
{
    const int lane = threadIdx.x % G, lane32 = threadIdx.x & 31, gid = lane32 / G;
    const unsigned gmask = group_mask<G>();
    const int colbase = ld_meta(meta + lbase).col, colend = ld_meta(meta + lend).col;
    const int l = lbase + gid;
    if (l < lend) // the stripe bodies store y[a.col + ...]: point them at the staging row of this run
        adj_one_stripe<Tv, G, MODE, true>(ld_meta(meta + l), ld_meta(meta + l + 1), lane, gmask, desc, val, x, stage - colbase, u0, log2u, alpha, (Tv)0);
    __syncwarp();
    const int ncols = colend - colbase;
    if (h.mask == nullptr) { // full replication: every destination gets the run
        for (int i = 0; i < h.n; i++) {
            Tv *d = reinterpret_cast<Tv *>(h.p[i]) + colbase;
            for (int c = lane32; c < ncols; c += 32) d[c] = stage[c];
        }
    } else { // sparsity-aware replication: a destination gets only the column chunks it gathers from
        for (int c = lane32; c < ncols; c += 32) {
            const unsigned mk = __ldg(h.mask + ((colbase + c) >> h.chunk_shift)) | 1u;
            const Tv v = stage[c];
            for (int i = 0; i < h.n; i++)
                if ((mk >> i) & 1u) reinterpret_cast<Tv *>(h.p[i])[colbase + c] = v;
        }
    }
    __syncwarp();
}

template <typename Tv, int G, int MODE>
__global__ void VBC_ADJ_BOUNDS k_spmv_adj_halo(const StripeMeta *__restrict__ meta, const int *__restrict__ desc,
                                               const Tv *__restrict__ val, const Tv *__restrict__ x, const __grid_constant__ HaloArgs h,
                                               const int L, const int u0, const int log2u, const Tv alpha)
{
    constexpr int GPW = 32 / G;
    extern __shared__ __align__(16) unsigned char halo_smem[]; // 8 warps x (32/G stripes x W columns) staged results
    const int lane32 = threadIdx.x & 31, gid = lane32 / G;
    const int warp = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int nwarps = (int)((gridDim.x * blockDim.x) >> 5);
    int w2 = warp - h.first_warp; // position in the deal of the boundary runs
    if (w2 < 0) w2 += nwarps;
    const bool bwarp = w2 < h.nruns; // this warp owns boundary runs w2, w2 + nwarps, ...
    unsigned long long epoch = 0;
    if (bwarp) {
        Tv *stage = reinterpret_cast<Tv *>(halo_smem) + (threadIdx.x >> 5) * h.stage_elems;
        // the neighbours' flags and this rank's epoch are fetched together; the comparison waits for both
        unsigned long long flag = ~0ull;
        const bool waits = h.do_wait && lane32 < h.nranks && lane32 != h.me && ((h.nbr_mask >> lane32) & 1u);
        if (waits) flag = ld_acquire_sys_u64(h.flags[h.me] + lane32);
        epoch = ld_relaxed_gpu_u64(h.ctl + 2); // steps this rank has published: the neighbours must have published as many
        if (waits && flag < epoch) {
            unsigned long long t0, t1;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
            t1 = t0;
            while (ld_acquire_sys_u64(h.flags[h.me] + lane32) < epoch) {
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                if (t1 - t0 > 4000000000ull) { atomicExch(h.timed_out, 1); break; } // 4 s: a dead peer must not hang the GPU
                __nanosleep(32);
            }
            // wait statistics (vbc_peer_wait_stats): total ns, waits that spun, longest
            atomicAdd(h.ctl + 3, t1 - t0);
            atomicAdd(h.ctl + 4, 1ull);
            atomicMax(h.ctl + 5, t1 - t0);
        }
        __syncwarp();
        for (int r = w2; r < h.nruns; r += nwarps) {
            int lbase, lend;
            if (r < h.nrunsA) { lbase = r * GPW; lend = min(lbase + GPW, h.i0); }
            else { lbase = h.i1 + (r - h.nrunsA) * GPW; lend = min(lbase + GPW, L); }
            adj_boundary_run<Tv, G, MODE>(meta, lbase, lend, desc, val, x, stage, h, u0, log2u, alpha);
        }
    }
    // Interior stripes, in three parts walked by ONE copy of the loop:
    //   [i0, ifence)   every warp; after it a boundary warp counts itself done (fence + atomic) -- by then its peer stores
    //                  have long been acknowledged, so the fence has nothing to wait for, and the step is still published
    //                  tens of microseconds before any neighbour needs it;
    //   [ifence, imid) every warp;
    //   [imid, i1)     only the warps WITHOUT boundary runs: a boundary run costs about two interior runs (L2-coherent
    //                  loads, staging, peer stores, fence), so the boundary warps get that much less of the interior deal
    //                  and all warps finish together.
#pragma unroll 1
    for (int part = 0; part < 3; part++) {
        int lo, hi, group, ngroups;
        if (part < 2) {
            lo = part == 0 ? h.i0 : h.ifence; hi = part == 0 ? h.ifence : h.imid;
            group = warp * GPW + gid; ngroups = nwarps * GPW;
        } else {
            if (bwarp) break;
            lo = h.imid; hi = h.i1;
            group = (w2 - h.nbwarps) * GPW + gid; ngroups = (nwarps - h.nbwarps) * GPW; // dealt over the warps without boundary runs
        }
        adj_range<Tv, G, MODE>(meta, nullptr, lo, hi, group, ngroups, desc, val, x, reinterpret_cast<Tv *>(h.p[0]), u0, log2u, alpha, (Tv)0);
        if (part == 0 && bwarp) {
            __syncwarp();
            __threadfence_system(); // this lane's stores (own HBM and peers) are visible system-wide before the warp counts as done
            __syncwarp();
            unsigned long long fin = 0;
            if (lane32 == 0) fin = atomicAdd(h.ctl + 1, 1ull);
            fin = __shfl_sync(0xffffffffu, fin, 0);
            if ((fin % (unsigned long long)h.nbwarps) == (unsigned long long)(h.nbwarps - 1)) { // last boundary warp of this launch: publish the step
                // (do_signal == 0: a separate flag kernel publishes and advances the epoch -- vbc_peer_barrier)
                __threadfence_system();
                if (h.do_signal && lane32 == 0) *reinterpret_cast<volatile unsigned long long *>(h.ctl + 2) = epoch + 1ull;
                if (h.do_signal && lane32 < h.nranks && lane32 != h.me && ((h.nbr_mask >> lane32) & 1u))
                    st_release_sys_u64(h.flags[lane32] + h.me, epoch + 1ull);
            }
        }
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
    // batches of UNR row-steps (group-uniform trip count: every lane joins the shuffles): all loads of a batch are issued before
    // the first product, then the per-row sums are reduced and scattered -- one load in flight per lane left this kernel
    // waiting for memory between any two atomics
    constexpr int UNR = 4;
    for (int rb = 0; rb < R; rb += UNR * rps) {
        Tv v[UNR][EPV], p[UNR];
        int xi[UNR];
        bool ok[UNR];
#pragma unroll
        for (int k = 0; k < UNR; k++) {
            ok[k] = active && (rb + k * rps + r0 < R);
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
        for (int k = 0; k < UNR; k++) {
            p[k] = (Tv)0;
#pragma unroll
            for (int e = 0; e < EPV; e++) p[k] = fma(v[k][e], xs[e], p[k]);
        }
        __syncwarp(gmask);
#pragma unroll
        for (int k = 0; k < UNR; k++) {
            if constexpr (CPR > 0) {
#pragma unroll
                for (int d = 1; d < CPR; d <<= 1) p[k] += __shfl_xor_sync(gmask, p[k], d, G);
            } else {
                for (int d = 1; d < cpr; d <<= 1) {
                    const Tv t = __shfl_down_sync(gmask, p[k], d, G);
                    if (c + d < cpr) p[k] += t;
                }
            }
        }
#pragma unroll
        for (int k = 0; k < UNR; k++)
            if (ok[k] && c == 0) atomicAdd(y + xi[k], alpha * p[k]); // y[idx[Q]] += ...  (multiply_1DVBC.jl:34)
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
    // mixed widths (stripes regrouped by body class): the unaligned classes use one lane per column ELEMENT, and with 8 lanes
    // a 5..7-wide stripe leaves lanes idle; 16 lanes hold two or three rows per step (C2v: 238 -> 214 us)
    if (A->d_order != nullptr && A->nclasses > 1 && G == 8) G = 16;
    // unaligned stripes: the flat-slab bodies need at least 8 lanes (natural 1..6-wide blocks, 1 KB stripes: 32 -> 23.5 us)
    if (A->has_unaligned && !A->opt_no_flat && G == 4) G = 8;
    return G;
}

template <typename Tv, int G, int MODE>
static int launch_adj_t(vbc_mat *A, Tv alpha, const Tv *x, Tv beta, Tv *y)
{
    int occ = 0;
    VBC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_spmv_adj<Tv, G, MODE>, 256, 0));
    if (occ < 1) occ = 1;
    if (A->opt_grid_mult > 0) occ = A->opt_grid_mult;
    int64_t grid = (int64_t)A->sm_count * occ;
    // a stripe range [l0, l1) is launched by offsetting meta: its entries are absolute (value offset, descriptor, column)
    const bool ranged = A->range_l0 >= 0 && A->d_order == nullptr;
    const int64_t n_long = A->d_order != nullptr ? A->n_long : 0; // the last n_long entries of the order: one CTA each, below
    const int64_t l0 = ranged ? A->range_l0 : 0, l1 = (ranged ? A->range_l1 : A->L) - n_long;
    auto launch_long = [&]() -> int {
        if (n_long == 0) return VBC_OK;
        k_spmv_adj_long<Tv, MODE><<<(unsigned)n_long, 256, 0, A->stream>>>(A->d_meta, A->d_order + (A->L - n_long), A->d_desc, (const Tv *)A->d_val, x, y, A->u0, ilog2_exact(A->u0), alpha, beta);
        A->launches++;
        VBC_CUDA(cudaGetLastError());
        return VBC_OK;
    };
    const int64_t need = ((l1 - l0) * G + 255) / 256;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    if constexpr (G >= 8) {
        if (A->has_unaligned && !A->opt_no_flat) { // stripes that are not 16-byte aligned: the flat-slab bodies (one shared-memory x row per group)
            int occf = 0;
            VBC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occf, k_spmv_adj_flat<Tv, G, MODE>, 256, 0));
            if (occf < 1) occf = 1;
            if (A->opt_grid_mult > 0) occf = A->opt_grid_mult;
            int64_t gridf = (int64_t)A->sm_count * occf;
            if (gridf > need) gridf = need;
            if (gridf < 1) gridf = 1;
            k_spmv_adj_flat<Tv, G, MODE><<<(unsigned)gridf, 256, 0, A->stream>>>(A->d_meta + l0, A->d_desc, (const Tv *)A->d_val, x, y, A->d_order, (int)(l1 - l0), A->u0, ilog2_exact(A->u0), alpha, beta);
            A->launches++;
            VBC_CUDA(cudaGetLastError());
            return launch_long();
        }
    }
    if (l1 > l0) {
        k_spmv_adj<Tv, G, MODE><<<(unsigned)grid, 256, 0, A->stream>>>(A->d_meta + l0, A->d_desc, (const Tv *)A->d_val, x, y, A->d_order, (int)(l1 - l0), A->u0, ilog2_exact(A->u0), alpha, beta);
        A->launches++;
    }
    VBC_CUDA(cudaGetLastError());
    return launch_long();
}

// the fused multiply + exchange: same persistent grid as the plain kernel
template <typename Tv, int G, int MODE>
static int launch_halo_t(vbc_mat *A, Tv alpha, const Tv *x, HaloArgs &h)
{
    int occ = 0;
    VBC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_spmv_adj_halo<Tv, G, MODE>, 256, 8 * (32 / G) * 32 * sizeof(Tv)));
    if (occ < 1) occ = 1;
    int64_t grid = (int64_t)A->sm_count * occ;
    const int64_t need = (A->L * G + 255) / 256;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    constexpr int GPW = 32 / G;
    const int L = (int)A->L;
    h.nrunsA = (h.i0 + GPW - 1) / GPW;
    h.nruns = h.nrunsA + (L - h.i1 + GPW - 1) / GPW;
    const int64_t warps = grid * 8;
    // interior runs are dealt round-robin from warp 0 (adj_range); the boundary runs continue that deal, so the warps that
    // got one interior run fewer take the boundary runs
    const int64_t int_runs = (h.i1 - h.i0 + GPW - 1) / GPW;
    h.first_warp = (int)(int_runs % warps);
    h.nbwarps = (int)(h.nruns < warps ? h.nruns : warps);
    const int64_t ngroups = warps * GPW;
    // a boundary run is weighted as `bw` interior runs: the boundary warps take part in `common` rounds of the interior deal,
    // the rest of the interior goes to the other warps only
    static int bw_env = -1;
    if (bw_env < 0) { const char *e = getenv("VBC_HALO_BOUNDARY_WEIGHT"); bw_env = e ? atoi(e) : 3; if (bw_env < 0 || bw_env > 16) bw_env = 3; }
    const int64_t bw = bw_env;
    int64_t common = int_runs / warps; // rounds every warp can take part in
    if (h.nbwarps > 0 && h.nbwarps < warps) {
        const int64_t rb = (h.nruns + h.nbwarps - 1) / h.nbwarps; // boundary runs per boundary warp
        int64_t c = (int_runs + bw * h.nruns) / warps - bw * rb;
        if (c < 0) c = 0;
        if (c < common) common = c;
    } else if (h.nbwarps >= warps) {
        common = (int_runs + warps - 1) / warps; // every warp is a boundary warp: the whole interior is common
    }
    const int64_t mid = (int64_t)h.i0 + common * ngroups;
    h.imid = (int)(mid < h.i1 ? mid : h.i1);
    const int64_t fence_at = (int64_t)h.i0 + 2 * ngroups; // two rounds of the interior deal
    h.ifence = (int)(fence_at < h.imid ? fence_at : h.imid);
    const int wmax = A->W < 1 ? 1 : (A->W > 32 ? 32 : A->W);
    h.stage_elems = GPW * wmax;
    const size_t smem = (size_t)8 * h.stage_elems * sizeof(Tv);
    k_spmv_adj_halo<Tv, G, MODE><<<(unsigned)grid, 256, smem, A->stream>>>(A->d_meta, A->d_desc, (const Tv *)A->d_val, x, h, L, A->u0, ilog2_exact(A->u0), alpha);
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

template <typename Tv>
static int launch_adj_any(vbc_mat *A, Tv alpha, const Tv *x, Tv beta, Tv *y)
{
    const bool rows = A->desc_mode == DESC_ROWS;
    const int G = A->opt_adj_group ? A->opt_adj_group : auto_group(A);
    if (G >= 32) return rows ? launch_adj_t<Tv, 32, DESC_ROWS>(A, alpha, x, beta, y) : launch_adj_t<Tv, 32, DESC_BLOCKS>(A, alpha, x, beta, y);
    if (G >= 16) return rows ? launch_adj_t<Tv, 16, DESC_ROWS>(A, alpha, x, beta, y) : launch_adj_t<Tv, 16, DESC_BLOCKS>(A, alpha, x, beta, y);
    if (G >= 8) return rows ? launch_adj_t<Tv, 8, DESC_ROWS>(A, alpha, x, beta, y) : launch_adj_t<Tv, 8, DESC_BLOCKS>(A, alpha, x, beta, y);
    return rows ? launch_adj_t<Tv, 4, DESC_ROWS>(A, alpha, x, beta, y) : launch_adj_t<Tv, 4, DESC_BLOCKS>(A, alpha, x, beta, y);
}

template <typename Tv>
static int launch_halo_any(vbc_mat *A, Tv alpha, const Tv *x, HaloArgs &h)
{
    const bool rows = A->desc_mode == DESC_ROWS;
    const int G = A->opt_adj_group ? A->opt_adj_group : auto_group(A);
    if (G >= 32) return rows ? launch_halo_t<Tv, 32, DESC_ROWS>(A, alpha, x, h) : launch_halo_t<Tv, 32, DESC_BLOCKS>(A, alpha, x, h);
    if (G >= 16) return rows ? launch_halo_t<Tv, 16, DESC_ROWS>(A, alpha, x, h) : launch_halo_t<Tv, 16, DESC_BLOCKS>(A, alpha, x, h);
    if (G >= 8) return rows ? launch_halo_t<Tv, 8, DESC_ROWS>(A, alpha, x, h) : launch_halo_t<Tv, 8, DESC_BLOCKS>(A, alpha, x, h);
    return rows ? launch_halo_t<Tv, 4, DESC_ROWS>(A, alpha, x, h) : launch_halo_t<Tv, 4, DESC_BLOCKS>(A, alpha, x, h);
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
        return launch_adj_any<Tv>(A, alpha, x, beta, y);
    }
    if (A->opt_fwd_atomic != 1) { // owner-computes forward through the transposed unit index (fwdt.cu)
        VBC_TRY(ensure_tindex(A));
        if (A->tindex && (tindex_kind(A) == 2 || A->opt_fwd_atomic == 2 || A->desc_mode == DESC_BLOCKS)) return launch_fwdt(A, alpha_d, xv, beta_d, yv);
    }
    VBC_TRY(scale_y<Tv>(A, y, A->m, beta));
    if (A->L == 0 || A->nval == 0) return VBC_OK;
    int G = A->opt_fwd_group ? A->opt_fwd_group : 32; // the scatter kernel likes whole warps per stripe (110 vs 131 us on C3)
    if (G < 8) G = 8;
    if (G >= 32) return rows ? launch_fwd_t<Tv, 32, DESC_ROWS>(A, alpha, x, y) : launch_fwd_t<Tv, 32, DESC_BLOCKS>(A, alpha, x, y);
    return rows ? launch_fwd_t<Tv, 8, DESC_ROWS>(A, alpha, x, y) : launch_fwd_t<Tv, 8, DESC_BLOCKS>(A, alpha, x, y);
}

// One step of the row-partitioned iteration: the adjoint multiply of this rank's slab with the exchange of the new x
// fused in (k_spmv_adj_halo).  dst_ptrs[i]: next-x buffer of destination i, already offset to this rank's first column.
int launch_spmv_adj_halo(vbc_mat *A, double alpha, const void *d_x, const HaloLaunch *hl)
{
    if (A->opt_parity) VBC_FAIL(VBC_EARG, "peer multiply needs the compact layout (parity mode is on)");
    if (hl->n < 1 || hl->n > VBC_MAX_PEERS) VBC_FAIL(VBC_EARG, "peer count %d out of 1..%d", hl->n, VBC_MAX_PEERS);
    if (A->W > 32) VBC_FAIL(VBC_ELIMIT, "the fused multiply + exchange stages a run of stripes in shared memory: stripes may be at most 32 columns wide (W = %d)", A->W);
    if (A->d_order != nullptr) VBC_FAIL(VBC_ELIMIT, "the fused multiply + exchange walks the stripes in index order: matrices whose stripes were regrouped by width class are not supported");
    if (A->L == 0) return VBC_OK;
    HaloArgs h{};
    h.n = hl->n;
    for (int i = 0; i < VBC_MAX_PEERS; i++) h.p[i] = i < hl->n ? hl->dst[i] : nullptr;
    h.mask = hl->d_mask;
    h.chunk_shift = hl->chunk_shift;
    const int L = (int)A->L;
    h.i0 = hl->i0 < 0 ? 0 : (hl->i0 > L ? L : hl->i0);
    h.i1 = hl->i1 < h.i0 ? h.i0 : (hl->i1 > L ? L : hl->i1);
    h.me = hl->me; h.nranks = hl->nranks; h.do_wait = hl->do_wait; h.do_signal = hl->do_signal; h.nbr_mask = hl->nbr_mask;
    for (int r = 0; r < VBC_MAX_PEERS; r++) h.flags[r] = r < hl->nranks ? hl->flags[r] : nullptr;
    h.ctl = hl->ctl;
    h.timed_out = hl->timed_out;
    if (!vt_is_float(A->vt)) VBC_FAIL(VBC_EARG, "ArgumentError: the fused multiply + exchange is offered for Float32 / Float64 matrices");
    if (A->vt == VBC_F64) return launch_halo_any<double>(A, alpha, (const double *)d_x, h);
    return launch_halo_any<float>(A, (float)alpha, (const float *)d_x, h);
}

int launch_spmv(vbc_mat *A, int trans, double alpha, const void *d_x, double beta, void *d_y)
{
    if (!vt_is_float(A->vt)) return launch_spmv_int(A, trans, alpha, d_x, beta, d_y);
    return A->vt == VBC_F64 ? launch_spmv_t<double>(A, trans, alpha, d_x, beta, d_y) : launch_spmv_t<float>(A, trans, alpha, d_x, beta, d_y);
}

} // namespace vbc
