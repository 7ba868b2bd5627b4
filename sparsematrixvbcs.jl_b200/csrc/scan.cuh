// scan.cuh -- device-wide exclusive prefix sum over int64 (three small kernels; used by the
// pack path to turn per-stripe counts into pos/ofs, constructors_1DVBC.jl:23,:31).
#pragma once
#include "common.cuh"

namespace vbc {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ long long warp_incl_scan(long long v, int lane)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        long long t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

// exclusive scan of `v` across the block; returns the exclusive prefix, writes the block total.
__device__ __forceinline__ long long block_excl_scan(long long v, long long *total, long long *smem /* 32 */)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    long long incl = warp_incl_scan(v, lane);
    if (lane == 31) smem[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        long long s = (lane < (blockDim.x >> 5)) ? smem[lane] : 0;
        long long si = warp_incl_scan(s, lane);
        smem[lane] = si - s; // exclusive per-warp offset
        if (lane == 31) smem[32] = si;
    }
    __syncthreads();
    long long r = incl - v + smem[warp];
    *total = smem[32];
    __syncthreads();
    return r;
}

static __global__ void __launch_bounds__(SCAN_THREADS) scan_tile_sums(const long long *__restrict__ in, long long *__restrict__ tile_sums, int64_t N)
{
    __shared__ long long smem[33];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    long long s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++)
        if (base + i < N) s += in[base + i];
    long long total;
    block_excl_scan(s, &total, smem);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// one block: exclusive scan of tile_sums[ntiles] in place; grand total to *total_out.
static __global__ void __launch_bounds__(SCAN_THREADS) scan_tile_offsets(long long *__restrict__ tile_sums, int64_t ntiles, long long *__restrict__ total_out)
{
    __shared__ long long smem[33];
    long long carry = 0;
    for (int64_t base = 0; base < ntiles; base += SCAN_THREADS) {
        const int64_t i = base + threadIdx.x;
        long long v = (i < ntiles) ? tile_sums[i] : 0;
        long long total;
        long long ex = block_excl_scan(v, &total, smem);
        if (i < ntiles) tile_sums[i] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) *total_out = carry;
}

// out[i] = add + sum(in[0..i))  for i in [0, N]  (N+1 outputs: the last one is the grand total + add)
template <typename To>
__global__ void __launch_bounds__(SCAN_THREADS) scan_apply(const long long *__restrict__ in, const long long *__restrict__ tile_offs, To *__restrict__ out, int64_t N, long long add)
{
    __shared__ long long smem[33];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    long long v[SCAN_ITEMS];
    long long s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        v[i] = (base + i < N) ? in[base + i] : 0;
        s += v[i];
    }
    long long total;
    long long ex = block_excl_scan(s, &total, smem) + tile_offs[blockIdx.x] + add;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        if (base + i < N) out[base + i] = (To)ex;
        ex += v[i];
        if (base + i == N - 1) out[N] = (To)ex;
    }
}

// Host driver.  d_in[N] -> d_out[N+1] (exclusive, offset by `add`); *h_total = sum(d_in).
// d_tmp must hold ceil(N / SCAN_TILE) + 1 int64.
template <typename To>
static int exclusive_scan(const long long *d_in, To *d_out, int64_t N, long long add, long long *d_tmp, long long *h_total, cudaStream_t st, int64_t *launches)
{
    if (N == 0) {
        To one = (To)add;
        VBC_CUDA(cudaMemcpyAsync(d_out, &one, sizeof(To), cudaMemcpyHostToDevice, st));
        VBC_CUDA(cudaStreamSynchronize(st));
        *h_total = 0;
        return VBC_OK;
    }
    const int64_t ntiles = (N + SCAN_TILE - 1) / SCAN_TILE;
    scan_tile_sums<<<(unsigned)ntiles, SCAN_THREADS, 0, st>>>(d_in, d_tmp, N);
    scan_tile_offsets<<<1, SCAN_THREADS, 0, st>>>(d_tmp, ntiles, d_tmp + ntiles);
    scan_apply<To><<<(unsigned)ntiles, SCAN_THREADS, 0, st>>>(d_in, d_tmp, d_out, N, add);
    if (launches) *launches += 3;
    VBC_CUDA(cudaGetLastError());
    VBC_CUDA(cudaMemcpyAsync(h_total, d_tmp + ntiles, sizeof(long long), cudaMemcpyDeviceToHost, st));
    VBC_CUDA(cudaStreamSynchronize(st));
    return VBC_OK;
}

static inline int64_t scan_tmp_elems(int64_t N) { return (N + SCAN_TILE - 1) / SCAN_TILE + 2; }

} // namespace vbc
