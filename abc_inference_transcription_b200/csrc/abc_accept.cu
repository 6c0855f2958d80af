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

// ------------------------------------------------------------------------------------------------
// posterior summaries over the ordered accepted lists (scripts/posterior_kinetics.jl:10-33).
//   MAP  = parameters of the first accepted index = smallest error           (posterior_kinetics.jl:14)
//   mean = mean over the accepted parameter rows, summed in list order        (posterior_kinetics.jl:18-22)
//   lo / hi = quantile(1 - q), quantile(q) with Julia's default definition (julia_quantile below; posterior_kinetics.jl:26-33)
#include <cub/device/device_segmented_sort.cuh>

__global__ void posterior_map_mean_kernel(const long long* __restrict__ idx, const long long* __restrict__ offsets,
                                          const double* __restrict__ theta, long long n, int P, long long particle_offset,
                                          int G, double* __restrict__ map, double* __restrict__ mean, int* __restrict__ bad) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)G * P) return;
    const int g = (int)(t / P), p = (int)(t % P);
    const long long b = offsets[g], e = offsets[g + 1];
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    if (e <= b) {
        if (map) map[t] = nan;
        if (mean) mean[t] = nan;
        return;
    }
    double s = 0.0, first = nan;
    for (long long k = b; k < e; ++k) {
        const long long row = idx[k] - 1 - particle_offset;
        if (row < 0 || row >= n) { atomicExch(bad, 1); return; }
        const double v = theta[row * P + p];
        if (k == b) first = v;
        s = __dadd_rn(s, v);
    }
    if (map) map[t] = first;
    if (mean) mean[t] = __ddiv_rn(s, (double)(e - b));
}

__global__ void posterior_gather_kernel(const long long* __restrict__ idx, size_t total, const double* __restrict__ theta,
                                        long long n, int P, int p, long long particle_offset, double* __restrict__ vals) {
    const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= total) return;
    const long long row = idx[k] - 1 - particle_offset;
    vals[k] = (row >= 0 && row < n) ? theta[row * P + p] : __longlong_as_double(0x7ff8000000000000ll);
}

// Statistics.jl _quantile with alpha = beta = 1 (the default of quantile(v, p)):
//   aleph = n p + (1 - p);  j = clamp(trunc(aleph), 1, n - 1);  gamma = clamp(aleph - j, 0, 1);  v[j] + gamma (v[j+1] - v[j])
__device__ __forceinline__ double julia_quantile(const double* v, long long cnt, double pr) {
    if (cnt == 1) return v[0];
    const double aleph = __dadd_rn(__dmul_rn((double)cnt, pr), __dadd_rn(1.0, -pr));
    long long j = (long long)trunc(aleph);
    j = j < 1 ? 1 : (j > cnt - 1 ? cnt - 1 : j);
    double gam = __dadd_rn(aleph, -(double)j);
    gam = gam < 0.0 ? 0.0 : (gam > 1.0 ? 1.0 : gam);
    const double a = v[j - 1], b = v[j];
    return __dadd_rn(a, __dmul_rn(gam, __dadd_rn(b, -a)));
}

__global__ void posterior_quantile_kernel(const double* __restrict__ sorted, const long long* __restrict__ offsets, int G, int P,
                                          int p, double q, double* __restrict__ lo, double* __restrict__ hi) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    const long long b = offsets[g], cnt = offsets[g + 1] - b;
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    double l = nan, u = nan;
    if (cnt > 0) {
        l = julia_quantile(sorted + b, cnt, __dadd_rn(1.0, -q));
        u = julia_quantile(sorted + b, cnt, q);
    }
    if (lo) lo[(size_t)g * P + p] = l;
    if (hi) hi[(size_t)g * P + p] = u;
}

size_t abc_posterior_temp_bytes(size_t total, int G) {
    size_t bytes = 0;
    cub::DoubleBuffer<double> keys(nullptr, nullptr);
    cub::DeviceSegmentedSort::SortKeys(nullptr, bytes, keys, (int)total, G, (const long long*)nullptr, (const long long*)nullptr);
    return bytes;
}

// d_idx: ordered accepted lists (abc_launch_accept_sort), d_offsets: [G+1] on the device, d_theta: [n][P] on the device.
// d_vals: two buffers of `total` doubles.  Outputs [G][P] on the device, any of them may be NULL.
int abc_launch_posterior(const long long* d_idx, const long long* d_offsets, size_t total, int G, const double* d_theta,
                         long long n, int P, long long particle_offset, double q, double* d_vals[2], void* d_temp,
                         size_t temp_bytes, int* d_bad, double* d_map, double* d_mean, double* d_lo, double* d_hi,
                         int* n_launches, cudaStream_t st) {
    if (total >= 0x7FFFFFFFull) { abc_set_error("too many accepted tuples for one summary (%zu)", total); return ABC_ERR_ARG; }
    int launches = 0;
    ABC_CUDA_CHECK(cudaMemsetAsync(d_bad, 0, sizeof(int), st));
    {
        const long long items = (long long)G * P;
        posterior_map_mean_kernel<<<(unsigned)((items + 127) / 128), 128, 0, st>>>(d_idx, d_offsets, d_theta, n, P, particle_offset,
                                                                                G, d_map, d_mean, d_bad);
        launches++;
    }
    if ((d_lo || d_hi) && total > 0) {
        const unsigned blocks = (unsigned)((total + 255) / 256);
        for (int p = 0; p < P; ++p) {
            cub::DoubleBuffer<double> keys(d_vals[0], d_vals[1]);
            posterior_gather_kernel<<<blocks, 256, 0, st>>>(d_idx, total, d_theta, n, P, p, particle_offset, keys.Current());
            ABC_CUDA_CHECK(cub::DeviceSegmentedSort::SortKeys(d_temp, temp_bytes, keys, (int)total, G, d_offsets, d_offsets + 1, st));
            posterior_quantile_kernel<<<(G + 127) / 128, 128, 0, st>>>(keys.Current(), d_offsets, G, P, p, q, d_lo, d_hi);
            launches += 3;
        }
    } else if (d_lo || d_hi) {
        for (int p = 0; p < P; ++p) {
            posterior_quantile_kernel<<<(G + 127) / 128, 128, 0, st>>>(nullptr, d_offsets, G, P, p, q, d_lo, d_hi);
            launches++;
        }
    }
    ABC_CUDA_CHECK(cudaGetLastError());
    if (n_launches) *n_launches = launches;
    return ABC_OK;
}

// ------------------------------------------------------------------------------------------------
// get_model_probs (scripts/model_probs.jl:1-28, constant_model_probs.jl:1-28, non_constant_model_probs.jl:1-30)
// and the per-gene case split around it (model_probs.jl:42-54): acceptance-count ratios of K hypotheses with bootstrap
// percentile bounds.  The reference resamples the vector of sum(l) model labels with replacement n_bootstraps times; here
// one warp draws the sum(l) labels of one (gene, bootstrap) from Philox (64 random bits per label: index = floor(u * sum(l)),
// class by the cumulative counts) -- the same resampling, integer arithmetic only, reproducible for a given seed.
#define ABC_DOM_BOOT 2u
#define BOOT_MAXK 8
#define BOOT_MAXB 256

__global__ void boot_resample_kernel(const long long* __restrict__ counts, int K, int G, int B, uint32_t k0, uint32_t k1,
                                     double* __restrict__ stats) {
    const long long w = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= (long long)G * B) return;
    const int g = (int)(w / B), b = (int)(w % B);
    unsigned long long cum[BOOT_MAXK];
    unsigned long long tot = 0;
    int nz = 0;
    for (int k = 0; k < K; ++k) {
        const unsigned long long c = (unsigned long long)counts[(long long)k * G + g];
        tot += c; cum[k] = tot;
        nz += (c > 0);
    }
    if (nz < 2) return;            // decided without a bootstrap (model_probs.jl:42-50)
    unsigned long long cnt[BOOT_MAXK];
    for (int k = 0; k < K; ++k) cnt[k] = 0;
    const unsigned long long nblk = (tot + 1) >> 1;       // one Philox block = two labels
    for (unsigned long long j = lane; j < nblk; j += 32) {
        const uint4 r = philox4x32_10((uint32_t)j, (uint32_t)g, (uint32_t)b, ABC_DOM_BOOT << 29, k0, k1);
        const unsigned long long u0 = ((unsigned long long)r.x << 32) | r.y, u1 = ((unsigned long long)r.z << 32) | r.w;
        const unsigned long long i0 = __umul64hi(u0, tot), i1 = __umul64hi(u1, tot);
        const bool second = 2 * j + 1 < tot;
        for (int k = 0; k < K; ++k) {
            const bool below0 = i0 < cum[k] && (k == 0 || i0 >= cum[k - 1]);
            const bool below1 = second && i1 < cum[k] && (k == 0 || i1 >= cum[k - 1]);
            cnt[k] += (below0 ? 1ull : 0ull) + (below1 ? 1ull : 0ull);
        }
    }
    for (int k = 0; k < K; ++k) {
        unsigned long long c = cnt[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        if (lane == 0) stats[((long long)g * K + k) * B + b] = __ddiv_rn((double)c, (double)tot);     // sample_l ./ sum(sample_l)
    }
}

__global__ void boot_bounds_kernel(const long long* __restrict__ counts, int K, int G, int B, double alpha,
                                   const double* __restrict__ stats, double* __restrict__ prob, double* __restrict__ lb,
                                   double* __restrict__ ub) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)G * K) return;
    const int g = (int)(t / K), k = (int)(t % K);
    unsigned long long tot = 0;
    int nz = 0;
    for (int q = 0; q < K; ++q) { const unsigned long long c = (unsigned long long)counts[(long long)q * G + g]; tot += c; nz += (c > 0); }
    const unsigned long long mine = (unsigned long long)counts[(long long)k * G + g];
    double p = 0.0, l = 0.0, u = 0.0;
    if (nz == 1) {
        if (mine > 0) { p = 1.0; l = 1.0; u = 1.0; }
    } else if (nz >= 2) {
        p = __ddiv_rn((double)mine, (double)tot);
        double v[BOOT_MAXB];
        const double* s = stats + ((long long)g * K + k) * B;
        for (int b = 0; b < B; ++b) {            // insertion sort = sort(stats, dims = 1)
            const double x = s[b];
            int j = b;
            while (j > 0 && v[j - 1] > x) { v[j] = v[j - 1]; --j; }
            v[j] = x;
        }
        l = julia_quantile(v, B, 1.0 - alpha);
        u = julia_quantile(v, B, alpha);
    }
    prob[t] = p; lb[t] = l; ub[t] = u;
}

int abc_launch_model_probs(const long long* d_counts, int K, int G, int B, double alpha, uint64_t seed, double* d_stats,
                           double* d_prob, double* d_lb, double* d_ub, int* n_launches, cudaStream_t st) {
    if (K < 1 || K > BOOT_MAXK || B < 1 || B > BOOT_MAXB || G < 1) {
        abc_set_error("abc_model_probs: K in 1..%d, n_bootstraps in 1..%d", BOOT_MAXK, BOOT_MAXB);
        return ABC_ERR_ARG;
    }
    const int threads = 256;
    const long long warps = (long long)G * B;
    boot_resample_kernel<<<(unsigned)((warps * 32 + threads - 1) / threads), threads, 0, st>>>(d_counts, K, G, B, (uint32_t)seed,
                                                                                            (uint32_t)(seed >> 32), d_stats);
    boot_bounds_kernel<<<(unsigned)(((long long)G * K + 127) / 128), 128, 0, st>>>(d_counts, K, G, B, alpha, d_stats, d_prob, d_lb, d_ub);
    ABC_CUDA_CHECK(cudaGetLastError());
    if (n_launches) *n_launches = 2;
    return ABC_OK;
}
