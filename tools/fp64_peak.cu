// fp64_peak.cu -- measures the FP64 ceilings SpMM is judged against: DFMA (SIMT) and DMMA m8n8k4 (tensor path)
// throughput on the whole chip.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/fp64_peak tools/fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dfma(double *out, int iters)
{
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double b = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; i++) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

__global__ void k_dmma(double *out, int iters)
{
    double c[8][2];
    for (int i = 0; i < 8; i++) { c[i][0] = 0; c[i][1] = 0; }
    const double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[j][0]), "+d"(c[j][1]) : "d"(a), "d"(b));
    }
    double s = 0;
    for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main()
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double *out;
    cudaMalloc(&out, sizeof(double) * sms * 8 * 256);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int rep = 0; rep < 2; rep++) {
        cudaEventRecord(e0); k_dfma<<<sms * 8, 256>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double flops = 2.0 * 8 * iters * (double)sms * 8 * 256;
        printf("DFMA  : %.1f TFLOP/s (%.2f ms)\n", flops / ms / 1e9, ms);
        cudaEventRecord(e0); k_dmma<<<sms * 8, 256>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        flops = 2.0 * 256 * 8 * iters * (double)sms * 8 * 8; // 256 FMA per warp-level mma, 8 per iter, 8 warps per CTA
        printf("DMMA  : %.1f TFLOP/s (%.2f ms)\n", flops / ms / 1e9, ms);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
