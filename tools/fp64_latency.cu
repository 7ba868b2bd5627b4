// Dependent-chain latency of the FP64 / FP32 pipes and of a 64-bit warp shuffle on this GPU (one warp, clock64 around a
// 4096-long chain).  nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/fp64_latency.cu -o /tmp/fp64_latency
#include <cstdio>
#include <cuda_runtime.h>
template <typename T> __global__ void chain_fma(T *out, T a, T b, long long *cyc)
{
    T x = a;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < 4096; i++) x = fma(x, b, a);
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
template <typename T> __global__ void chain_add(T *out, T a, long long *cyc)
{
    T x = a;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < 4096; i++) x = x + a;
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
template <typename T> __global__ void chain_shfl(T *out, T a, long long *cyc)
{
    T x = a + threadIdx.x;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < 4096; i++) x = __shfl_down_sync(0xffffffffu, x, 4);
    long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main()
{
    double *od; float *of; long long *c, h;
    cudaMalloc(&od, 256); cudaMalloc(&of, 128); cudaMalloc(&c, 8);
    for (int rep = 0; rep < 2; rep++) {
        chain_fma<double><<<1, 32>>>(od, 1.0000001, 0.9999999, c); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); if (rep) printf("DFMA dependent latency  %.1f cycles\n", h / 4096.0);
        chain_add<double><<<1, 32>>>(od, 1.0000001, c); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); if (rep) printf("DADD dependent latency  %.1f cycles\n", h / 4096.0);
        chain_fma<float><<<1, 32>>>(of, 1.0000001f, 0.9999999f, c); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); if (rep) printf("FFMA dependent latency  %.1f cycles\n", h / 4096.0);
        chain_shfl<double><<<1, 32>>>(od, 1.0, c); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); if (rep) printf("SHFL.64 dependent latency %.1f cycles\n", h / 4096.0);
        chain_shfl<float><<<1, 32>>>(of, 1.0f, c); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); if (rep) printf("SHFL.32 dependent latency %.1f cycles\n", h / 4096.0);
    }
    return 0;
}
