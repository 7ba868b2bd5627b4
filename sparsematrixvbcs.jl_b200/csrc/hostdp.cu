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

// Greedy overlap chunker (stand-in for ChainPartitioners' OverlapChunker(rho, w_max), test/runtests.jl:21,
// bin/test_table.jl:66).  ASSUMED definition (the package is not vendored): walk the columns left to right; the
// next column joins the current stripe while the stripe is narrower than w_max and the Jaccard similarity of the
// column's row pattern with the union of the stripe's patterns is >= rho.  colptr/rowval: 0-based int64 CSC
// structure with ascending rows.  Writes the 1-based spl and its length.
extern "C" int vbc_overlap_chunk(int64_t n, const int64_t *colptr, const int64_t *rowval, double rho, int w_max, int64_t *spl, int64_t *L_out)
{
    if (n < 0 || w_max < 1 || !spl || !L_out || (n > 0 && (!colptr || !rowval))) VBC_FAIL(VBC_EARG, "bad argument");
    int64_t L = 0;
    int64_t cap = 16, ulen = 0;
    int64_t *uni = (int64_t *)malloc(sizeof(int64_t) * (size_t)cap), *tmp = (int64_t *)malloc(sizeof(int64_t) * (size_t)cap);
    if (!uni || !tmp) { free(uni); free(tmp); VBC_FAIL(VBC_ENOMEM, "host allocation failed"); }
    int width = 0;
    for (int64_t j = 0; j < n; j++) {
        const int64_t b = colptr[j], e = colptr[j + 1], len = e - b;
        bool join = false;
        int64_t inter = 0;
        if (width > 0 && width < w_max) {
            int64_t p = 0, q = b; // |pattern ∩ union| by a merge of two ascending lists
            while (p < ulen && q < e) {
                if (uni[p] < rowval[q]) p++;
                else if (uni[p] > rowval[q]) q++;
                else { inter++; p++; q++; }
            }
            const int64_t un = ulen + len - inter;
            join = un == 0 ? true : ((double)inter >= rho * (double)un);
        }
        if (!join) { // open a new stripe at column j
            spl[L++] = j + 1;
            width = 0;
            ulen = 0;
        }
        // union <- union ∪ pattern(j)
        const int64_t need = ulen + len;
        if (need > cap) {
            while (cap < need) cap *= 2;
            int64_t *nu = (int64_t *)realloc(uni, sizeof(int64_t) * (size_t)cap), *nt = (int64_t *)realloc(tmp, sizeof(int64_t) * (size_t)cap);
            if (!nu || !nt) { free(nu ? nu : uni); free(nt ? nt : tmp); VBC_FAIL(VBC_ENOMEM, "host allocation failed"); }
            uni = nu; tmp = nt;
        }
        int64_t p = 0, q = b, o = 0;
        while (p < ulen || q < e) {
            if (q >= e || (p < ulen && uni[p] < rowval[q])) tmp[o++] = uni[p++];
            else if (p >= ulen || uni[p] > rowval[q]) tmp[o++] = rowval[q++];
            else { tmp[o++] = uni[p]; p++; q++; }
        }
        int64_t *sw = uni; uni = tmp; tmp = sw;
        ulen = o;
        width++;
    }
    spl[L] = n + 1;
    *L_out = L;
    free(uni);
    free(tmp);
    return VBC_OK;
}
