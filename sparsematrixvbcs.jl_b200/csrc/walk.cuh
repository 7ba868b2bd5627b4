// walk.cuh -- device helpers shared by the multiply kernels (spmv.cu, spmm.cu, trsv.cu): stripe meta
// loads, group masks, and the per-lane walk over a stripe's x indices.
#pragma once
#include "common.cuh"

namespace vbc {

// floor(a / b) for 0 <= a < 2048, 1 <= b <= 32 without an integer division: (a * ceil(2^16 / b)) >> 16
// (exact on that range; checked exhaustively).  The generic stripe bodies divide lane ids and group
// sizes by the runtime vectors-per-row count once per stripe; a real division costs ~20 instructions.
__constant__ unsigned short c_inv16[33] = {0, 0 /* b = 1 handled below */, 32768, 21846, 16384, 13108, 10923, 9363, 8192, 7282, 6554, 5958, 5462,
                                           5042, 4682, 4370, 4096, 3856, 3641, 3450, 3277, 3121, 2979, 2850, 2731, 2622, 2521, 2428,
                                           2341, 2260, 2185, 2115, 2048};
__device__ __forceinline__ int small_div(const int a, const int b)
{
    return b == 1 ? a : (int)(((unsigned)a * (unsigned)c_inv16[b]) >> 16);
}

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

// x index of stored row r of a stripe (any lane, no walk state)
template <int MODE>
__device__ __forceinline__ int row_xindex(const int *__restrict__ desc, const int pos0, const int r, const int u0, const int log2u)
{
    if (MODE == DESC_ROWS) return __ldg(desc + pos0 + r);
    if (log2u >= 0) return __ldg(desc + pos0 + (r >> log2u)) + (r & (u0 - 1));
    return __ldg(desc + pos0 + r / u0) + r % u0;
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
