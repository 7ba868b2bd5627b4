// Does the FP64 datapath of sm_100 go to sleep?  One warp alternates an idle gap (integer spin on clock64) with a burst of 12
// dependent DFMAs (the dependency chain of one level of the triangular solve) and reports the cycles per burst against the gap.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/fp64_wake.cu -o build/fp64_wake
#include <cstdio>
#include <cuda_runtime.h>
template <typename T> __global__ void burst(T *out, T a, T b, long long gap, long long *cyc)
{
    T x = a + (T)threadIdx.x;
    long long total = 0;
    for (int it = 0; it < 200; it++) {
        const long long s = clock64();
        while (clock64() - s < gap) { }
        long long t0, t1;
        asm volatile("mov.u64 %0, %%clock64;" : "=l"(t0) :: "memory");
#pragma unroll
        for (int i = 0; i < 12; i++) x = fma(x, b, a);
        if (x == (T)12345.678) out[32] = x; // consume x before the clock is read
        asm volatile("mov.u64 %0, %%clock64;" : "=l"(t1) :: "memory");
        total += t1 - t0;
    }
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) *cyc = total / 200;
}
int main()
{
    double *od; float *of; long long *c, h;
    cudaMalloc(&od, 512); cudaMalloc(&of, 256); cudaMalloc(&c, 8);
    const long long gaps[] = {0, 50, 200, 1000, 5000, 20000, 100000};
    for (long long g : gaps) {
        burst<double><<<1, 32>>>(od, 1.0000001, 0.9999999, g, c); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost); const long long d = h;
        burst<float><<<1, 32>>>(of, 1.0000001f, 0.9999999f, g, c); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
        printf("idle gap %7lld cycles: 12 dependent DFMA = %5lld cycles, 12 dependent FFMA = %5lld cycles\n", g, d, h);
    }
    return 0;
}
