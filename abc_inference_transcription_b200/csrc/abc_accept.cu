// abc_accept.cu -- A1 on the device: the per-gene lists v[sortperm(err[v])] of scripts/accepted_particles.jl:20-24.
//
// The scoring kernels append accepted (gene, particle, err) tuples in arrival order (warp-aggregated atomics).  The reference
// orders every gene's accepted particles by ascending error with a stable sort over ascending indices, i.e. by the key
// (gene, err, particle).  Three stable LSD radix passes (CUB) over a 32-bit permutation produce exactly that order:
// by particle, then by the error's order-preserving bit pattern, then by gene.  sm_100a.
#include "abc_common.cuh"
#include "abc_internal.h"
#include <cub/cub.cuh>

__global__ void accept_keys_particle_kernel(const long long* __restrict__ particle, size_t n, unsigned long long* __restrict__ key,
                                            uint32_t* __restrict__ perm) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    key[i] = (unsigned long long)particle[i];
    perm[i] = (uint32_t)i;
}

// IEEE-754 total order on the bit patterns: negative values flip all bits, the others flip the sign bit
__global__ void accept_keys_err_kernel(const double* __restrict__ err, const uint32_t* __restrict__ perm, size_t n,
                                       unsigned long long* __restrict__ key) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long b = (unsigned long long)__double_as_longlong(err[perm[i]]);
    key[i] = (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

__global__ void accept_keys_gene_kernel(const int32_t* __restrict__ gene, const uint32_t* __restrict__ perm, size_t n,
                                        uint32_t* __restrict__ key) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    key[i] = (uint32_t)gene[perm[i]];
}

__global__ void accept_gather_kernel(const long long* __restrict__ particle, const double* __restrict__ err,
                                     const uint32_t* __restrict__ perm, size_t n, long long* __restrict__ out_idx,
                                     double* __restrict__ out_err) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t j = perm[i];
    out_idx[i] = particle[j];
    out_err[i] = err[j];
}

size_t abc_accept_sort_temp_bytes(size_t total) {
    size_t b64 = 0, b32 = 0;
    cub::DoubleBuffer<unsigned long long> k64(nullptr, nullptr);
    cub::DoubleBuffer<uint32_t> k32(nullptr, nullptr), v(nullptr, nullptr);
    cub::DeviceRadixSort::SortPairs(nullptr, b64, k64, v, (int)total);
    cub::DeviceRadixSort::SortPairs(nullptr, b32, k32, v, (int)total);
    return b64 > b32 ? b64 : b32;
}

// key64 / key32 / perm: two buffers of `total` elements each; returns the sorted lists in out_idx / out_err
int abc_launch_accept_sort(const int32_t* d_gene, const long long* d_particle, const double* d_err, size_t total, int G,
                           unsigned long long* d_key64[2], uint32_t* d_key32[2], uint32_t* d_perm[2], void* d_temp,
                           size_t temp_bytes, long long* d_out_idx, double* d_out_err, int* n_launches, cudaStream_t st) {
    if (total == 0) return ABC_OK;
    if (total >= 0x7FFFFFFFull) { abc_set_error("too many accepted tuples for one sort (%zu)", total); return ABC_ERR_ARG; }
    const int n = (int)total, threads = 256;
    const unsigned blocks = (unsigned)((total + threads - 1) / threads);
    cub::DoubleBuffer<unsigned long long> k64(d_key64[0], d_key64[1]);
    cub::DoubleBuffer<uint32_t> k32(d_key32[0], d_key32[1]), perm(d_perm[0], d_perm[1]);
    accept_keys_particle_kernel<<<blocks, threads, 0, st>>>(d_particle, total, k64.Current(), perm.Current());
    ABC_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(d_temp, temp_bytes, k64, perm, n, 0, 64, st));
    accept_keys_err_kernel<<<blocks, threads, 0, st>>>(d_err, perm.Current(), total, k64.Current());
    ABC_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(d_temp, temp_bytes, k64, perm, n, 0, 64, st));
    int gene_bits = 1;
    while ((1 << gene_bits) < G && gene_bits < 31) ++gene_bits;
    accept_keys_gene_kernel<<<blocks, threads, 0, st>>>(d_gene, perm.Current(), total, k32.Current());
    ABC_CUDA_CHECK(cub::DeviceRadixSort::SortPairs(d_temp, temp_bytes, k32, perm, n, 0, gene_bits, st));
    accept_gather_kernel<<<blocks, threads, 0, st>>>(d_particle, d_err, perm.Current(), total, d_out_idx, d_out_err);
    ABC_CUDA_CHECK(cudaGetLastError());
    if (n_launches) *n_launches = 4 + 3;      // own kernels + the three radix sorts (several CUB kernels each)
    return ABC_OK;
}
