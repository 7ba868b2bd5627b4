/* oracle/vbc_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C + OpenMP) of the blocked sparse multiply path of
 * ulysses4ever/SparseMatrixVBCs.jl:
 *     src/constructors_1DVBC.jl   1D pack (general :9-92, strict :94-143)
 *     src/constructors_VBC.jl     2D pack (:15-133)
 *     src/multiply_1DVBC.jl       forward (:9-83) / adjoint (:85-180) multiply
 *     src/multiply_VBC.jl         forward (:3-87) / adjoint (:89-192) multiply
 *     src/TrSpMV.jl               CSC transposed SpMV (:1-20)
 *     src/costs.jl                memory cost models (:10, :140)
 * Each function in vbc_oracle_body.inc cites the lines it follows.
 *
 * Who may use this: tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs -- as the CHECKER or the timed CPU baseline, never as the shipped
 * compute path.  The product (libvbc.so) neither links nor calls anything in this directory.
 *
 * PARITY STATUS: the reference is Julia and there is no Julia toolchain in this image, so the
 * reference itself cannot be executed here.  The oracle is pinned against what the reference's
 * own tests hold for this path: the six literal matrices of test/matrices.jl:4-9 and the
 * exact one-hot invariants of test/runtests.jl:29-53 and :63-87 (B*e_j == A*e_j and
 * B'*e_i == A'*e_i for every unit vector, both orientations, 1D and 2D), plus the random-x
 * `isapprox` check of bin/test_table.jl:42/:84/:126.  The reference stores no golden
 * pos/idx/ofs/val arrays or y vectors, so the packed arrays themselves are "parity unpinned"
 * by reference OUTPUTS; they are pinned structurally (SURVEY.md Appendix A/B known answers).
 *
 * Build:  make -C oracle        (gcc -O3 -march=native -ffast-math -fopenmp, mirroring the
 *                                reference's @fastmath/@inbounds/SIMD.jl code generation)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* Loop schedule of the adjoint multiply.  0 (default) = schedule(dynamic, 1): one stripe per grab from a shared
 * counter, the reference's discipline (`atomic_add!(l′, 1)` inside `@threads`, multiply_1DVBC.jl:169-177).
 * 1 = schedule(static): NOT what the reference does; offered so a benchmark can also state what the same loops
 * reach without the shared counter. */
static int g_static_schedule = 0;
void vbc_oracle_set_static_schedule(int on) { g_static_schedule = on ? 1 : 0; }
static void vbc_oracle_apply_schedule(void)
{
#ifdef _OPENMP
    if (g_static_schedule) omp_set_schedule(omp_sched_static, 0);
    else omp_set_schedule(omp_sched_dynamic, 1);
#endif
}

#define CAT3_(a, b, c) a##_##b##_##c
#define CAT3(a, b, c) CAT3_(a, b, c)

#define TV double
#define TI int64_t
#define FN(name) CAT3(name, f64, i64)
#include "vbc_oracle_body.inc"
#undef TV
#undef TI
#undef FN

#define TV double
#define TI int32_t
#define FN(name) CAT3(name, f64, i32)
#include "vbc_oracle_body.inc"
#undef TV
#undef TI
#undef FN

#define TV float
#define TI int64_t
#define FN(name) CAT3(name, f32, i64)
#include "vbc_oracle_body.inc"
#undef TV
#undef TI
#undef FN

#define TV float
#define TI int32_t
#define FN(name) CAT3(name, f32, i32)
#include "vbc_oracle_body.inc"
#undef TV
#undef TI
#undef FN

/* Integer element types (test/runtests.jl:15-16 packs and multiplies Bool and Int32 matrices).  Julia's Int32 / Int64
 * arithmetic wraps; unsigned C arithmetic is the same bits without undefined behaviour. */
#define TV uint32_t
#define TI int64_t
#define FN(name) CAT3(name, s32, i64)
#include "vbc_oracle_body.inc"
#undef TV
#undef TI
#undef FN

#define TV uint32_t
#define TI int32_t
#define FN(name) CAT3(name, s32, i32)
#include "vbc_oracle_body.inc"
#undef TV
#undef TI
#undef FN

#define TV uint64_t
#define TI int64_t
#define FN(name) CAT3(name, s64, i64)
#include "vbc_oracle_body.inc"
#undef TV
#undef TI
#undef FN

#define TV uint64_t
#define TI int32_t
#define FN(name) CAT3(name, s64, i32)
#include "vbc_oracle_body.inc"
#undef TV
#undef TI
#undef FN

int vbc_oracle_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
