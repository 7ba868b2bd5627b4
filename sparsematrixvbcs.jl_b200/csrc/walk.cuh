// walk.cuh -- device helpers shared by the multiply kernels (spmv.cu, spmm.cu, trsv.cu): stripe meta
// loads, group masks, and the per-lane walk over a stripe's x indices.
#pragma once
#include "common.cuh"

namespace vbc {

__device__ __forceinline__ StripeMeta ld_meta(const StripeMeta *p)
{
    const int4 t = __ldg(reinterpret_cast<const int4 *>(p));
    StripeMeta s;
    s.ofs = (long long)(((unsigned long long)(unsigned)t.y << 32) | (unsigned)t.x);
    s.pos = t.z;
    s.col = t.w;
    return s;
}

template <int G> __device__ __forceinline__ unsigned group_mask()
{
    if constexpr (G == 32) return 0xffffffffu;
    else return ((1u << G) - 1u) << (((threadIdx.x & 31) / G) * G);
}

// x-index stream of a stripe for one lane: row r0, then r0+rps, ...
template <int MODE> struct RowWalk;
template <> struct RowWalk<DESC_ROWS> {
    const int *dp; int step;
    __device__ __forceinline__ void init(const int *desc, int pos0, int r0, int rps, int, int) { dp = desc + pos0 + r0; step = rps; }
    __device__ __forceinline__ int next() { const int xi = __ldcs(dp); dp += step; return xi; }
    __device__ __forceinline__ int next_if(bool ok) { const int xi = ok ? __ldcs(dp) : 0; dp += step; return xi; }
};
template <> struct RowWalk<DESC_BLOCKS> {
    const int *bp; int di, qb, rb, u0;
    __device__ __forceinline__ void init(const int *desc, int pos0, int r0, int rps, int u0_, int log2u)
    {
        u0 = u0_;
        if (log2u >= 0) { bp = desc + pos0 + (r0 >> log2u); di = r0 & (u0 - 1); qb = rps >> log2u; rb = rps & (u0 - 1); }
        else { bp = desc + pos0 + r0 / u0; di = r0 % u0; qb = rps / u0; rb = rps % u0; }
    }
    __device__ __forceinline__ int next()
    {
        const int xi = __ldg(bp) + di;
        bp += qb; di += rb;
        if (di >= u0) { di -= u0; ++bp; }
        return xi;
    }
    __device__ __forceinline__ int next_if(bool ok)
    {
        const int xi = ok ? __ldg(bp) + di : 0;
        bp += qb; di += rb;
        if (di >= u0) { di -= u0; ++bp; }
        return xi;
    }
};

} // namespace vbc
