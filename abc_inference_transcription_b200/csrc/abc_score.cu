// abc_score.cu -- error scoring of the particles x genes matrix with fused eps-acceptance.  sm_100a.
//
// Replaces nlsqerror_part / compute_trunc_errors (scripts/compute_errors.jl:30-70) and the
// findall(x -> x <= eps) of scripts/accepted_particles.jl:20.  Bit-exact FP64: every operation is an
// explicitly rounded IEEE op in the reference's order (no FMA contraction, correctly rounded divide):
//     err = 0.0
//     for group l in (pulse_mean, pulse_ff, chase_mean, chase_ff, ratio, mean_corr, corr_mean):
//         e = 0.0;  for i: e += (d[i]-s[i])^2 / (se[i]^2 + 0.1^2*d[i]^2 + eps_i);   err += e/53
//     if err > 10.0: err = 10.0          (NaN passes through, compute_errors.jl:62-64)
// The per-gene denominators den[i] = (se^2 + sigma^2 d^2) + eps_i depend on the data only and are
// computed once by abc_prepare_data_kernel with the same operation order.
//
// Early exit: every term is >= 0 or NaN and rounding is monotone, so once the running
// total of completed groups exceeds 10.0 the final value is exactly 10.0 -- unless a later term is
// NaN, which the reference lets through unclipped.  A NaN term can only come from a NaN statistic of
// the particle (then the particle's whole row is NaN: every gene uses all 53 statistics) or from
// non-finite data of the gene.  So: particles with a NaN statistic are answered NaN directly, genes
// with non-finite data are evaluated in full, and otherwise a warp leaves a gene as soon as ALL its
// lanes are past 10.0.
//
// Mapping: one thread per particle (its 53 statistics live in registers), genes streamed through
// shared memory in tiles (broadcast reads), results staged per tile and written as full 128-byte
// lines in either layout.
#include "abc_common.cuh"
#include "abc_internal.h"

#define SC_THREADS 128
#define SC_GT 16          // genes per shared-memory tile

__global__ void abc_prepare_data_kernel(const double* __restrict__ d, const double* __restrict__ se, long long n,
                                        double* __restrict__ den) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double sigma = 0.1;
    const double sig2 = __dmul_rn(sigma, sigma);               // 0.010000000000000002
    const double di = d[i], si = se[i];
    const double eps = (__dadd_rn(si, di) != 0.0) ? 0.0 : 0.0001;
    den[i] = __dadd_rn(__dadd_rn(__dmul_rn(si, si), __dmul_rn(sig2, __dmul_rn(di, di))), eps);
}

int abc_launch_prepare_data(const double* d_d, const double* d_se, int G, double* d_den, double* d_rden,
                            cudaStream_t st) {
    (void)d_rden;
    long long n = (long long)G * ABC_NSTATS;
    if (n <= 0) return ABC_OK;
    abc_prepare_data_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_d, d_se, n, d_den);
    ABC_CUDA_CHECK(cudaGetLastError());
    return ABC_OK;
}

template <int LO, int HI>
__device__ __forceinline__ double group_err(const double (&s)[ABC_NSTATS], const double* __restrict__ gd,
                                            const double* __restrict__ gden) {
    double e = 0.0;
#pragma unroll
    for (int t = LO; t < HI; ++t) {
        const double diff = __dadd_rn(gd[t], -s[t]);
        const double num = __dmul_rn(diff, diff);
        e = __dadd_rn(e, __ddiv_rn(num, gden[t]));
    }
    return __ddiv_rn(e, 53.0);
}

__global__ void __launch_bounds__(SC_THREADS)
abc_score_kernel(const AbcScoreArgs a) {
    __shared__ double sh_d[SC_GT][ABC_NSTATS];
    __shared__ double sh_den[SC_GT][ABC_NSTATS];
    __shared__ double sh_err[SC_GT][SC_THREADS + 1];
    __shared__ int sh_bad[SC_GT];
    const int tid = threadIdx.x, lane = tid & 31;
    const long long n_tiles = (a.n + SC_THREADS - 1) / SC_THREADS;
    // blockIdx.y owns a contiguous range of gene tiles
    const int n_gtiles = (a.G + SC_GT - 1) / SC_GT;
    const int gt_per = (n_gtiles + gridDim.y - 1) / gridDim.y;
    const int g_lo = min(a.G, (int)blockIdx.y * gt_per * SC_GT);
    const int g_hi = min(a.G, g_lo + gt_per * SC_GT);

    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long i0 = tile * SC_THREADS;
        const long long i = i0 + tid;
        const bool live = i < a.n;
        double s[ABC_NSTATS];
        {
            const double* sp = a.stats + (live ? i : (a.n - 1)) * ABC_NSTATS;
#pragma unroll
            for (int t = 0; t < ABC_NSTATS; ++t) s[t] = sp[t];
        }
        bool rnan = false;
#pragma unroll
        for (int t = 0; t < ABC_NSTATS; ++t) rnan = rnan || (s[t] != s[t]);
        for (int g0 = g_lo; g0 < g_hi; g0 += SC_GT) {
            const int gt = min(SC_GT, g_hi - g0);
            __syncthreads();
            if (tid < SC_GT) sh_bad[tid] = 0;
            __syncthreads();
            for (int k = tid; k < gt * ABC_NSTATS; k += SC_THREADS) {
                const double dv = a.d[(long long)g0 * ABC_NSTATS + k];
                const double nv = a.den[(long long)g0 * ABC_NSTATS + k];
                (&sh_d[0][0])[k] = dv;
                (&sh_den[0][0])[k] = nv;
                // non-finite data (x - x != 0 for Inf and NaN): no early exit for this gene
                if (__dadd_rn(dv, -dv) != 0.0 || __dadd_rn(nv, -nv) != 0.0) sh_bad[k / ABC_NSTATS] = 1;
            }
            __syncthreads();
            for (int gg = 0; gg < gt; ++gg) {
                const double* gd = sh_d[gg];
                const double* gden = sh_den[gg];
                const bool gbad = sh_bad[gg] != 0;       // warp uniform
                double err = 0.0;
                // group order and sizes: compute_errors.jl:55-61
                err = __dadd_rn(err, group_err<0, 5>(s, gd, gden));
                if (gbad || !__all_sync(0xffffffffu, rnan || err > 10.0)) {
                    err = __dadd_rn(err, group_err<5, 10>(s, gd, gden));
                    if (gbad || !__all_sync(0xffffffffu, rnan || err > 10.0)) {
                        err = __dadd_rn(err, group_err<10, 15>(s, gd, gden));
                        if (gbad || !__all_sync(0xffffffffu, rnan || err > 10.0)) {
                            err = __dadd_rn(err, group_err<15, 20>(s, gd, gden));
                            if (gbad || !__all_sync(0xffffffffu, rnan || err > 10.0)) {
                                err = __dadd_rn(err, group_err<20, 31>(s, gd, gden));
                                if (gbad || !__all_sync(0xffffffffu, rnan || err > 10.0)) {
                                    err = __dadd_rn(err, group_err<31, 42>(s, gd, gden));
                                    if (gbad || !__all_sync(0xffffffffu, rnan || err > 10.0)) {
                                        err = __dadd_rn(err, group_err<42, 53>(s, gd, gden));
                                    }
                                }
                            }
                        }
                    }
                }
                if (err > 10.0) err = 10.0;
                if (rnan) err = __longlong_as_double(0x7ff8000000000000ll);
                sh_err[gg][tid] = err;
                // fused threshold acceptance (accepted_particles.jl:20): err <= eps, NaN never accepted
                const bool acc = live && (err <= a.eps);
                const unsigned mask = __ballot_sync(0xffffffffu, acc);
                if (mask != 0u) {
                    const int leader = __ffs(mask) - 1;
                    unsigned long long base = 0;
                    if (lane == leader) {
                        const unsigned long long c = (unsigned long long)__popc(mask);
                        base = atomicAdd(a.acc_count, c);
                        atomicAdd(a.counts + (g0 + gg), c);
                    }
                    base = __shfl_sync(0xffffffffu, base, leader);
                    if (acc) {
                        const unsigned long long slot = base + (unsigned long long)__popc(mask & ((1u << lane) - 1u));
                        if ((long long)slot < a.acc_capacity) {
                            a.acc_gene[slot] = g0 + gg;
                            a.acc_particle[slot] = a.particle_offset + i + 1;   // 1-based like Julia
                            a.acc_err[slot] = err;
                        }
                    }
                }
            }
            __syncthreads();
            if (a.err != nullptr) {
                if (a.err_layout == ABC_ERR_GENE_MAJOR) {
                    for (int gg = 0; gg < gt; ++gg)
                        if (live) a.err[(long long)(g0 + gg) * a.n + i] = sh_err[gg][tid];
                } else {
                    // particle-major: each particle row gets gt contiguous doubles (one 128 B line for gt = 16)
                    for (int k = tid; k < SC_THREADS * SC_GT; k += SC_THREADS) {
                        const int pp = k / SC_GT, gg = k % SC_GT;
                        if (gg < gt && i0 + pp < a.n)
                            a.err[(i0 + pp) * (long long)a.G + g0 + gg] = sh_err[gg][pp];
                    }
                }
            }
        }
    }
}

int abc_launch_score(const AbcScoreArgs& a, int sm_count, cudaStream_t st) {
    if (a.n <= 0 || a.G <= 0) return ABC_OK;
    long long n_tiles = (a.n + SC_THREADS - 1) / SC_THREADS;
    int per_sm = 0;
    ABC_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, abc_score_kernel, SC_THREADS, 0));
    if (per_sm < 1) per_sm = 1;
    long long gx = (long long)sm_count * per_sm;
    if (gx > n_tiles) gx = n_tiles;
    // small batches: split the genes over blockIdx.y so that the grid still fills the chip
    const int n_gtiles = (a.G + SC_GT - 1) / SC_GT;
    long long gy = (4ll * sm_count * per_sm + gx - 1) / gx;
    if (gy > n_gtiles) gy = n_gtiles;
    if (gy > 65535) gy = 65535;
    if (gy < 1) gy = 1;
    abc_score_kernel<<<dim3((unsigned)gx, (unsigned)gy), SC_THREADS, 0, st>>>(a);
    ABC_CUDA_CHECK(cudaGetLastError());
    return ABC_OK;
}
