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
        vbc_destroy(A);
    }
    printf("abi_smoke ok\n");
    return 0;
}
