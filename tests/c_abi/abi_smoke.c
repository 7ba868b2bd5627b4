/* Plain-C client of libvbc.so: what a foreign-language binding sees.  Built and run by tests/test_abi_symbols.py
 * on a CPU-only box: it checks that include/vbc.h is valid C99, that the library links, and that argument
 * validation answers with the documented status codes before any CUDA call.  With a GPU (argv[1] == "gpu") it also
 * packs the worked 4 x 5 example of SURVEY.md Appendix A and multiplies. */
#include <stdio.h>
#include <string.h>

#include "vbc.h"

#define CHECK(cond)                                                                  \
    do {                                                                             \
        if (!(cond)) { printf("FAILED %s:%d: %s  [%s]\n", __FILE__, __LINE__, #cond, vbc_last_error()); return 1; } \
    } while (0)

int main(int argc, char **argv)
{
    vbc_mat *A = NULL;
    int64_t one = 1;
    CHECK(vbc_version() >= 100);
    CHECK(vbc_last_error() != NULL);
    /* ArgumentError paths (SparseMatrixVBCs.jl:45-50) */
    CHECK(vbc_pack_csc(&A, VBC_F64, VBC_I64, -1, 0, 0, 4, &one, &one, &one, NULL, 0, &one, 0, 0) == VBC_EARG);
    CHECK(A == NULL);
    CHECK(vbc_pack_csc(&A, VBC_F64, VBC_I64, 0, 0, 0, 0, &one, &one, &one, NULL, 0, &one, 0, 0) == VBC_EARG);
    CHECK(strstr(vbc_last_error(), "W must be > 0") != NULL);
    CHECK(vbc_spmv(NULL, 0, 1.0, NULL, 0, 0.0, NULL, 0, 0) == VBC_EARG);
    CHECK(vbc_spmv_mixed(NULL, 0, 1.0, NULL, 0, 0.0, NULL, 0, VBC_F64, 0) == VBC_EARG);
    CHECK(vbc_spmm(NULL, 0, 1, 1.0, NULL, 1, 0.0, NULL, 1, 0, 0) == VBC_EARG);
    vbc_destroy(NULL); /* a no-op */
    if (argc > 1 && strcmp(argv[1], "gpu") == 0) {
        /* A = [1 0 0 2 0; 0 3 0 0 0; 4 0 5 0 6; 0 0 0 7 0], Phi = {1:2, 3:4, 5:5}, W = 2 */
        int64_t colptr[6] = {1, 3, 4, 5, 7, 8}, rowval[7] = {1, 3, 2, 3, 1, 4, 3}, spl[4] = {1, 3, 5, 6};
        double nzval[7] = {1, 4, 3, 5, 2, 7, 6}, x[4] = {1, 1, 1, 1}, y[5] = {0, 0, 0, 0, 0}, xf[5] = {1, 1, 1, 1, 1}, yf[4];
        int64_t nidx = 0, nval = 0;
        int ndev = 0;
        CHECK(vbc_device_count(&ndev) == VBC_OK && ndev > 0);
        CHECK(vbc_pack_csc(&A, VBC_F64, VBC_I64, 4, 5, 0, 2, colptr, rowval, nzval, NULL, 0, spl, 3, 0) == VBC_OK);
        CHECK(vbc_sizes(A, &nidx, &nval) == VBC_OK);
        CHECK(nidx == 7 && nval == 13);
        CHECK(vbc_spmv(A, 1, 1.0, x, 4, 0.0, y, 5, 0) == VBC_OK);       /* y = A' 1 = column sums */
        CHECK(y[0] == 5 && y[1] == 3 && y[2] == 5 && y[3] == 9 && y[4] == 6);
        CHECK(vbc_spmv(A, 0, 1.0, xf, 5, 0.0, yf, 4, 0) == VBC_OK);     /* y = A 1 = row sums */
        CHECK(yf[0] == 3 && yf[1] == 3 && yf[2] == 15 && yf[3] == 7);
        CHECK(vbc_spmv(A, 0, 1.0, x, 4, 0.0, yf, 4, 0) == VBC_EDIM);    /* DimensionMismatch */
        {   /* cost model, format bytes, options: what the reference-side binding queries */
            int64_t cost[3], rowterm = -1, bytes[3], opt = -1;
            CHECK(vbc_memory_cost(A, cost, &rowterm) == VBC_OK && rowterm == 0);
            CHECK(cost[0] == 3 * 8 + 3 * 8 + 6 * 8);                    /* costs.jl:10: 3|Ti| + rows (|Ti| + w |Tv|), stripe 1 has 3 rows x 2 */
            CHECK(vbc_format_bytes(A, bytes) == VBC_OK && bytes[0] == 8 * (3 * 4 + 7 + 13));
            CHECK(vbc_get_option(A, VBC_OPT_E2E_PIPELINE, &opt) == VBC_OK && opt == 1);
            CHECK(vbc_set_option(A, VBC_OPT_ADJ_GROUP, 5) == VBC_EARG);
        }
        vbc_destroy(A);
        {   /* vbc_dist_*: the row-partitioned iteration driven by one process (one rank here; two ranks sharing the GPU run in lockstep) */
            /* S = tridiagonal 8 x 8 (2 on the diagonal, 1 beside it), Phi = 4 stripes of 2 columns */
            int64_t cp[9], rv[22], ph[5] = {1, 3, 5, 7, 9}, bounds[3], interior[4];
            double nz[22], x0[8], x1[8], ms = -1.0;
            int q = 0, P = 0, ranks, dev2[2] = {0, 0};
            for (int j = 0; j < 8; j++) {
                cp[j] = q + 1;
                for (int i = j - 1; i <= j + 1; i++)
                    if (i >= 0 && i < 8) { rv[q] = i + 1; nz[q] = (i == j) ? 2.0 : 1.0; q++; }
            }
            cp[8] = q + 1;
            for (ranks = 1; ranks <= 2; ranks++) {
                vbc_dist *D = NULL;
                for (int j = 0; j < 8; j++) x0[j] = 1.0;
                CHECK(vbc_dist_create(&D, ranks, dev2, VBC_F64, VBC_I64, 8, 0, 2, cp, rv, nz, NULL, 0, ph, 4, VBC_EXCH_FUSED) == VBC_OK);
                CHECK(vbc_dist_info(D, &P, NULL, bounds, NULL, interior) == VBC_OK && P == ranks && bounds[0] == 0 && bounds[ranks] == 4);
                CHECK(vbc_dist_set_x(D, x0) == VBC_OK);
                CHECK(vbc_dist_spmv_iter(D, 2, 0.25, &ms) == VBC_OK && ms > 0.0);
                CHECK(vbc_dist_gather_x(D, x1) == VBC_OK);
                /* (S'/4)^2 * ones: interior entries 1, the ends 0.625 and 0.9375 */
                CHECK(x1[3] == 1.0 && x1[4] == 1.0 && x1[0] == 0.625 && x1[1] == 0.9375 && x1[6] == 0.9375 && x1[7] == 0.625);
                vbc_dist_destroy(D);
            }
            {
                vbc_dist *D = NULL;
                CHECK(vbc_dist_create(&D, 2, dev2, VBC_F64, VBC_I64, 8, 0, 2, cp, rv, nz, NULL, 0, ph, 4, VBC_EXCH_NCCL) == VBC_ENCCL); /* one device twice */
                CHECK(D == NULL);
            }
        }
    }
    printf("abi_smoke ok\n");
    return 0;
}
