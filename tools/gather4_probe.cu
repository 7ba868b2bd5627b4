// gather4_probe.cu -- what does cp.async.bulk.tensor.2d ... tile::gather4 need from the tensor map?  (no PTX manual offline)
// Builds a [rows x 32] Float64 tensor with a row pitch of ld elements, encodes it with box {32, boxrows} and issues one gather4 of
// four arbitrary rows; prints whether the 4 x 256-byte tile in shared memory equals those rows.  Not part of the library.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/gather4_probe tools/gather4_probe.cu && /tmp/gather4_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                             const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void k_probe(const __grid_constant__ CUtensorMap tmap, int r0, int r1, int r2, int r3, int c0, double *out, int *status)
{
    __shared__ __align__(128) double tile[4 * 32];
    __shared__ __align__(8) unsigned long long bar;
    const unsigned bar_a = (unsigned)__cvta_generic_to_shared(&bar), tile_a = (unsigned)__cvta_generic_to_shared(tile);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 128; i += 32) tile[i] = -1.0;
    __syncwarp();
    if (threadIdx.x == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(1024) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                     ::"r"(tile_a), "l"(&tmap), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar_a) : "memory");
    }
    __syncwarp();
    unsigned ok = 0;
    long long spins = 0;
    while (!ok && spins < 20000000) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(bar_a), "r"(0) : "memory");
        spins++;
    }
    if (threadIdx.x == 0) *status = ok ? 1 : -1;
    for (int i = threadIdx.x; i < 128; i += 32) out[i] = tile[i];
}

int main()
{
    EncodeFn encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&encode, cudaEnableDefault, &qres) != cudaSuccess || !encode) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
    const int rows = 1000, k = 32, ld = 40;
    double *h = (double *)malloc(sizeof(double) * rows * ld), *d, *dout;
    for (int i = 0; i < rows * ld; i++) h[i] = (double)(i / ld) * 1000.0 + (i % ld);
    cudaMalloc(&d, sizeof(double) * rows * ld);
    cudaMemcpy(d, h, sizeof(double) * rows * ld, cudaMemcpyHostToDevice);
    cudaMalloc(&dout, sizeof(double) * 128);
    int *dstat; cudaMalloc(&dstat, 4);
    const int R[4] = {5, 100, 7, 999};
    for (int boxrows = 1; boxrows <= 4; boxrows += 3) {
        CUtensorMap tm;
        cuuint64_t gdim[2] = {(cuuint64_t)k, (cuuint64_t)rows}, gstr[1] = {(cuuint64_t)ld * 8};
        cuuint32_t box[2] = {32, (cuuint32_t)boxrows}, estr[2] = {1, 1};
        CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, d, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("box {32, %d}: encode -> %d\n", boxrows, (int)r);
        if (r != CUDA_SUCCESS) continue;
        cudaMemset(dstat, 0, 4);
        k_probe<<<1, 32>>>(tm, R[0], R[1], R[2], R[3], 0, dout, dstat);
        cudaError_t e = cudaDeviceSynchronize();
        printf("  launch: %s\n", cudaGetErrorString(e));
        if (e != cudaSuccess) return 2;
        double out[128]; int st;
        cudaMemcpy(out, dout, sizeof(out), cudaMemcpyDeviceToHost);
        cudaMemcpy(&st, dstat, 4, cudaMemcpyDeviceToHost);
        int good = 0;
        for (int t = 0; t < 4; t++) for (int c = 0; c < 32; c++) good += out[t * 32 + c] == R[t] * 1000.0 + c;
        printf("  barrier %s, %d / 128 values as expected for a dense [4][32] tile; first of each row: %.0f %.0f %.0f %.0f\n", st == 1 ? "completed" : "TIMED OUT", good, out[0], out[32], out[64], out[96]);
    }
    return 0;
}
