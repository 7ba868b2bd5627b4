// hostdp.cu -- host-side dynamic program of the contiguous partitioners (no device code).
//
// ChainPartitioners' DynamicTotalChunker (used by the reference at constructors_1DVBC.jl:1-2,
// constructors_VBC.jl:5-7, test/runtests.jl:22-23, bin/test_table.jl:66-69) minimises the total of a
// per-stripe cost over all contiguous partitions with stripes at most W wide.  ChainPartitioners is not
// vendored, so this is the textbook O(n W) recurrence on a caller-supplied cost table, ties broken
// towards the NARROWER last stripe (parity with ChainPartitioners' tie-breaking is unpinned).
#include "common.cuh"

extern "C" int vbc_dp_chunk(int64_t n, int W, const double *cost, int64_t *spl, int64_t *L_out)
{
    // cost[(b * W) + (w - 1)] = cost of the stripe of columns [b - w, b)  (0-based b in 1..n), w in 1..W
    if (n < 0 || W < 1 || !spl || !L_out || (n > 0 && !cost)) VBC_FAIL(VBC_EARG, "bad argument");
    double *best = (double *)malloc(sizeof(double) * (size_t)(n + 1));
    int *arg = (int *)malloc(sizeof(int) * (size_t)(n + 1));
    if (!best || !arg) { free(best); free(arg); VBC_FAIL(VBC_ENOMEM, "host allocation failed"); }
    best[0] = 0.0;
    arg[0] = 0;
    for (int64_t b = 1; b <= n; b++) {
        double bb = 0.0;
        int ba = 0;
        const int wmax = b < W ? (int)b : W;
        for (int w = 1; w <= wmax; w++) {
            const double c = best[b - w] + cost[(size_t)b * W + (w - 1)];
            if (ba == 0 || c < bb) { bb = c; ba = w; }
        }
        best[b] = bb;
        arg[b] = ba;
    }
    int64_t L = 0;
    for (int64_t b = n; b > 0; b -= arg[b]) L++;
    spl[L] = n + 1;
    int64_t l = L;
    for (int64_t b = n; b > 0; b -= arg[b]) spl[--l] = b - arg[b] + 1; // 1-based first column of the stripe
    if (n == 0) spl[0] = 1;
    *L_out = L;
    free(best);
    free(arg);
    return VBC_OK;
}
