// gen.cu -- benchmark support: synthetic block-banded matrices generated ON THE DEVICE, slab by slab.
//
// The generator is the one of the host module synth.py (`banded_blocks`), itself modelled on the reference's
// generators for its time-model experiments (/root/reference/src/costs.jl:63-85 1D, :200-222 2D: dense u x w blocks at
// distinct (row part, stripe) positions, values rand(Tv) in [0, 1), EquiChunker partitions): stripe l of the global
// matrix holds a dense u x w block at every row part  l*K/L + offsets[t]  that falls inside [0, K).  Values come from a
// counter-based hash -- splitmix64 of (row * n_global + col) xor seed -- so any slab can be generated independently by
// the rank that owns it (BASELINE.json configs[4] has 2 * 10^9 nonzeros: a host CSC would be 24-32 GB and minutes of
// CPU time) and any entry can be re-derived on the host for parity.  Output: a SparseMatrixCSC slab in device memory
// (1-based colptr / rowval of type Ti, nzval of type Tv), ready for vbc_pack_csc_dev.
#include "common.cuh"
#include "scan.cuh"

namespace vbc {

__host__ __device__ __forceinline__ unsigned long long splitmix64(unsigned long long z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

template <typename Tv> __device__ __forceinline__ Tv hash_value(unsigned long long h);
template <> __device__ __forceinline__ float hash_value<float>(unsigned long long h) { return (float)(h >> 40) * 5.9604644775390625e-8f; }   // 2^-24
template <> __device__ __forceinline__ double hash_value<double>(unsigned long long h) { return (double)(h >> 11) * 1.1102230246251565e-16; } // 2^-53

struct GenOffsets { long long off[64]; int n; };

// blocks of stripe l (global index): valid offsets are those with 0 <= centre + off < K
__device__ __forceinline__ int stripe_blocks(const GenOffsets &g, const long long centre, const long long K)
{
    int c = 0;
    for (int t = 0; t < g.n; t++) { const long long k = centre + g.off[t]; c += (k >= 0 && k < K) ? 1 : 0; }
    return c;
}

__global__ void __launch_bounds__(256) k_gen_counts(const __grid_constant__ GenOffsets g, const long long K, const long long L, const long long l0,
                                                     const long long ncols, const int u, const int w, long long *__restrict__ cnt)
{
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ncols) return;
    const long long l = l0 + j / w;
    cnt[j] = (long long)stripe_blocks(g, (l * K) / L, K) * u;
}

// one thread per (column, block): writes the u entries of that block's column
template <typename Ti, typename Tv>
__global__ void __launch_bounds__(256) k_gen_fill(const __grid_constant__ GenOffsets g, const long long K, const long long L, const long long l0,
                                                   const long long ncols, const int u, const int w, const unsigned long long n_global,
                                                   const unsigned long long seed, const double diag_boost, const Ti *__restrict__ colptr,
                                                   Ti *__restrict__ rowval, Tv *__restrict__ nzval)
{
    const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long j = id / g.n;
    const int t = (int)(id - j * g.n);
    if (j >= ncols) return;
    const long long l = l0 + j / w, centre = (l * K) / L;
    const long long k = centre + g.off[t];
    if (k < 0 || k >= K) return;
    int before = 0; // valid blocks of this stripe with a smaller offset (offsets ascend)
    for (int s = 0; s < t; s++) { const long long ks = centre + g.off[s]; before += (ks >= 0 && ks < K) ? 1 : 0; }
    const long long col = l0 * w + j; // global column
    long long q = (long long)colptr[j] - 1 + (long long)before * u;
    for (int di = 0; di < u; di++, q++) {
        const long long row = k * u + di;
        rowval[q] = (Ti)(row + 1);
        Tv v = hash_value<Tv>(splitmix64(((unsigned long long)row * n_global + (unsigned long long)col) ^ seed));
        if (diag_boost != 0.0 && row == col) v += (Tv)diag_boost;
        nzval[q] = v;
    }
}

// mark the chunks of x an adjoint multiply of this matrix gathers from
__global__ void __launch_bounds__(256) k_read_chunks(const int *__restrict__ desc, const long long ndesc, const int reach, const long long m,
                                                      const int chunk_shift, unsigned char *__restrict__ need)
{
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < ndesc; q += (long long)gridDim.x * blockDim.x) {
        const long long i0 = desc[q];
        long long i1 = i0 + reach - 1;
        if (i1 >= m) i1 = m - 1;
        for (long long c = i0 >> chunk_shift; c <= (i1 >> chunk_shift); c++) need[c] = 1;
    }
}

template <typename Ti, typename Tv>
static int gen_t(int64_t K, int64_t L, int u, int w, const GenOffsets &g, int64_t l0, int64_t l1, uint64_t seed, double diag_boost,
                 void **colptr_out, void **rowval_out, void **nzval_out, int64_t *nnz_out)
{
    const int64_t ncols = (l1 - l0) * w;
    Ti *colptr = nullptr, *rowval = nullptr;
    Tv *nzval = nullptr;
    long long *cnt = nullptr, *tmp = nullptr;
    int rc = VBC_OK;
    long long total = 0;
    int64_t launches = 0;
    do {
        if (cudaMalloc(&colptr, sizeof(Ti) * (size_t)(ncols + 1)) != cudaSuccess || cudaMalloc(&cnt, sizeof(long long) * (size_t)(ncols > 0 ? ncols : 1)) != cudaSuccess ||
            cudaMalloc(&tmp, sizeof(long long) * (size_t)scan_tmp_elems(ncols)) != cudaSuccess) { set_error("generator: allocation failed"); rc = VBC_ENOMEM; break; }
        if (ncols > 0) k_gen_counts<<<(unsigned)((ncols + 255) / 256), 256>>>(g, K, L, l0, ncols, u, w, cnt);
        if ((rc = exclusive_scan<Ti>(cnt, colptr, ncols, 1, tmp, &total, 0, &launches)) != VBC_OK) break;
        if (sizeof(Ti) == 4 && total + 1 > 0x7fffffffLL) { set_error("generator: %lld nonzeros overflow Ti=Int32", total); rc = VBC_ELIMIT; break; }
        if (cudaMalloc(&rowval, sizeof(Ti) * (size_t)(total > 0 ? total : 1)) != cudaSuccess || cudaMalloc(&nzval, sizeof(Tv) * (size_t)(total > 0 ? total : 1)) != cudaSuccess) {
            set_error("generator: allocation of %lld nonzeros failed", total); rc = VBC_ENOMEM; break;
        }
        const long long work = (long long)ncols * g.n;
        if (work > 0) k_gen_fill<Ti, Tv><<<(unsigned)((work + 255) / 256), 256>>>(g, K, L, l0, ncols, u, w, (unsigned long long)(L * w), seed, diag_boost, colptr, rowval, nzval);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { set_error("generator kernels failed: %s", cudaGetErrorString(e)); rc = VBC_ECUDA; }
    } while (0);
    cudaFree(cnt); cudaFree(tmp);
    if (rc != VBC_OK) { cudaGetLastError(); cudaFree(colptr); cudaFree(rowval); cudaFree(nzval); return rc; }
    *colptr_out = colptr; *rowval_out = rowval; *nzval_out = nzval; *nnz_out = total;
    return VBC_OK;
}

} // namespace vbc

using namespace vbc;

extern "C" {

int vbc_gen_banded_csc(int vt, int it, int64_t K, int64_t L, int u, int w, const int64_t *offsets, int noffsets, int64_t l0, int64_t l1,
                       uint64_t seed, double diag_boost, void **colptr, void **rowval, void **nzval, int64_t *nnz, int device)
{
    if (!offsets || !colptr || !rowval || !nzval || !nnz) VBC_FAIL(VBC_EARG, "NULL argument");
    if ((vt != VBC_F32 && vt != VBC_F64) || (it != VBC_I32 && it != VBC_I64)) VBC_FAIL(VBC_EARG, "bad element / index type");
    if (K < 1 || L < 1 || u < 1 || w < 1 || noffsets < 1 || noffsets > 64 || l0 < 0 || l1 < l0 || l1 > L) VBC_FAIL(VBC_EARG, "bad generator geometry");
    GenOffsets g;
    g.n = noffsets;
    for (int t = 0; t < noffsets; t++) {
        g.off[t] = offsets[t];
        if (t > 0 && offsets[t] <= offsets[t - 1]) VBC_FAIL(VBC_EARG, "offsets must ascend strictly");
    }
    DeviceGuard guard(device);
    if (!guard.ok) VBC_FAIL(VBC_ECUDA, "cudaSetDevice(%d) failed", device);
    if (it == VBC_I64) return vt == VBC_F64 ? gen_t<int64_t, double>(K, L, u, w, g, l0, l1, seed, diag_boost, colptr, rowval, nzval, nnz)
                                             : gen_t<int64_t, float>(K, L, u, w, g, l0, l1, seed, diag_boost, colptr, rowval, nzval, nnz);
    return vt == VBC_F64 ? gen_t<int32_t, double>(K, L, u, w, g, l0, l1, seed, diag_boost, colptr, rowval, nzval, nnz)
                         : gen_t<int32_t, float>(K, L, u, w, g, l0, l1, seed, diag_boost, colptr, rowval, nzval, nnz);
}

int vbc_gen_free(void *colptr, void *rowval, void *nzval, int device)
{
    DeviceGuard guard(device);
    if (!guard.ok) VBC_FAIL(VBC_ECUDA, "cudaSetDevice(%d) failed", device);
    cudaFree(colptr); cudaFree(rowval); cudaFree(nzval);
    return VBC_OK;
}

int vbc_read_chunks(vbc_mat *A, int chunk_shift, unsigned char *need, int64_t nchunks)
{
    if (!A || !need) VBC_FAIL(VBC_EARG, "NULL argument");
    if (chunk_shift < 0 || chunk_shift > 30 || nchunks < ((A->m + (1LL << chunk_shift) - 1) >> chunk_shift)) VBC_FAIL(VBC_EARG, "bad chunk geometry");
    if (A->opt_parity) VBC_FAIL(VBC_EARG, "needs the compact layout (parity mode is on)");
    DeviceGuard guard(A->device);
    if (!guard.ok) VBC_FAIL(VBC_ECUDA, "cudaSetDevice(%d) failed", A->device);
    memset(need, 0, (size_t)nchunks);
    if (A->ndesc == 0 || A->m == 0) return VBC_OK;
    unsigned char *d = nullptr;
    VBC_CUDA(cudaMalloc(&d, (size_t)nchunks));
    cudaError_t e = cudaMemsetAsync(d, 0, (size_t)nchunks, A->stream);
    if (e == cudaSuccess) {
        long long g = (A->ndesc + 255) / 256;
        if (g > (long long)A->sm_count * 16) g = (long long)A->sm_count * 16;
        k_read_chunks<<<(unsigned)g, 256, 0, A->stream>>>(A->d_desc, A->ndesc, A->desc_mode == DESC_BLOCKS ? A->u0 : 1, A->m, chunk_shift, d);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(need, d, (size_t)nchunks, cudaMemcpyDeviceToHost, A->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(A->stream);
    cudaFree(d);
    if (e != cudaSuccess) VBC_FAIL(VBC_ECUDA, "vbc_read_chunks: %s", cudaGetErrorString(e));
    return VBC_OK;
}

} // extern "C"
