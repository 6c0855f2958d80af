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

// FP32 pre-filter constants (see abc_score2_kernel): with w = 1/(53 den),
//   sum_t w (d-s)^2 = sum_t w d^2 - sum_t (2 w d) s + sum_t w s^2
// fbw[g][t] = (2 w d, w); fa[g] = (sum_{t<15} w d^2, sum_{t>=15} w d^2).  A gene whose data is not finite
// or badly scaled gets NaN constants, which sends all its pairs to the exact path.
#define SC2_T1 15   // terms of stage 1: groups pulse_mean, pulse_ff, chase_mean
__global__ void abc_prepare_filter_kernel(const double* __restrict__ d, const double* __restrict__ den, int G,
                                          float2* __restrict__ fbw, float2* __restrict__ fa) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    double a1 = 0.0, a2 = 0.0;
    bool bad = false;
    for (int t = 0; t < ABC_NSTATS; ++t) {
        const double dv = d[(long long)g * ABC_NSTATS + t], nv = den[(long long)g * ABC_NSTATS + t];
        if (!(nv >= 1e-30 && nv <= 1e30) || !(fabs(dv) <= 1e15)) bad = true;
        const double w = 1.0 / (53.0 * nv);
        fbw[(long long)g * ABC_NSTATS + t] = make_float2((float)(2.0 * w * dv), (float)w);
        if (t < SC2_T1) a1 += w * dv * dv; else a2 += w * dv * dv;
    }
    const float nanf_ = __int_as_float(0x7fc00000);
    fa[g] = bad ? make_float2(nanf_, nanf_) : make_float2((float)a1, (float)a2);
}

int abc_launch_prepare_data(const double* d_d, const double* d_se, int G, double* d_den, float2* d_fbw, float2* d_fa,
                            cudaStream_t st) {
    long long n = (long long)G * ABC_NSTATS;
    if (n <= 0) return ABC_OK;
    abc_prepare_data_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_d, d_se, n, d_den);
    ABC_CUDA_CHECK(cudaGetLastError());
    abc_prepare_filter_kernel<<<(G + 127) / 128, 128, 0, st>>>(d_d, d_den, G, d_fbw, d_fa);
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


// ------------------------------------------------------------------------------------------------
// abc_score2_kernel: three-stage scoring for eps < 10 (the reference's eps is 4.8).
//   stage 1 (all pairs, FP32): rigorous lower bound of the first 15 terms; "surely > 10" pairs are answered
//            10.0 at once (91 % of prior-particle x gene pairs on the real data), the rest are queued;
//   stage 2 (queued pairs, FP32, compacted): lower bound over all 53 terms; survivors (1.7 %) are queued;
//   stage 3 (queued pairs, FP64, compacted): the exact reference arithmetic (same code as abc_score_kernel).
// Soundness of the bound: every term is >= 0, so any partial sum bounds the total from below; the FP32
// evaluation of the expanded form differs from the exact partial sum by < 1e-3 + 1e-5 * p (w d^2 <= 1/(53 sigma^2)
// = 1.89 bounds the cancelling magnitudes; |s| >> |d| makes the sum huge and the error relative), so
// p32 > 10.01 implies the exact total exceeds 10.0, which the reference clips to exactly 10.0.  NaN or Inf in
// p32 never satisfies the test and falls through to the exact path; particles with a NaN statistic are
// answered NaN directly (every gene uses all 53 statistics, compute_errors.jl:58-64).
#define SC2_THREADS 128
#define SC2_WARPS (SC2_THREADS / 32)
#define SC2_PPT 2
#define SC2_TILE (SC2_THREADS * SC2_PPT)
#define SC2_GT 32
#define SC2_SUB 4
#define SC2_Q1W 512                      // stage-1 queue entries per warp (warp private: no atomics)
#define SC2_Q2 2048
#define SC2_SURE 10.01f

struct Score2Smem {
    float2 bw[SC2_GT][SC2_T1];
    float a1[SC2_GT];
    unsigned int q1_id[SC2_WARPS][SC2_Q1W];
    float q1_p[SC2_WARPS][SC2_Q1W];
    unsigned int q2_id[SC2_Q2];
    unsigned int mask[SC2_GT][SC2_TILE / 32 + 1];
    int q1cnt[SC2_WARPS];
    int q2n;
    int any_nan;
};

// particle prologue: FP32 copy of the statistics and a per-particle "has a NaN statistic" flag
__global__ void abc_score_prep_kernel(const double* __restrict__ stats, long long n, float* __restrict__ fstats,
                                      unsigned char* __restrict__ rnan) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool nn = false;
    for (int t = 0; t < ABC_NSTATS; ++t) {
        const double v = stats[i * ABC_NSTATS + t];
        nn = nn || (v != v);
        fstats[i * ABC_NSTATS + t] = (float)v;
    }
    rnan[i] = nn ? 1 : 0;
}

__device__ __forceinline__ double exact_pair_error(const double* __restrict__ sp, const double* __restrict__ gd,
                                                   const double* __restrict__ gden) {
    // compute_errors.jl:30-43, 55-64: the reference's operation order, explicit rounding
    const int off[8] = {0, 5, 10, 15, 20, 31, 42, 53};
    double err = 0.0;
#pragma unroll 1
    for (int l = 0; l < 7; ++l) {
        double e = 0.0;
#pragma unroll 1
        for (int t = off[l]; t < off[l + 1]; ++t) {
            const double diff = __dadd_rn(gd[t], -sp[t]);
            e = __dadd_rn(e, __ddiv_rn(__dmul_rn(diff, diff), gden[t]));
        }
        err = __dadd_rn(err, __ddiv_rn(e, 53.0));
    }
    return (err > 10.0) ? 10.0 : err;
}

template <int LAYOUT>
__device__ __forceinline__ void score2_store(const AbcScoreArgs& a, long long i, int g, double v) {
    if (LAYOUT == ABC_ERR_GENE_MAJOR) a.err[(long long)g * a.n + i] = v;
    else if (LAYOUT == ABC_ERR_PARTICLE_MAJOR) a.err[i * (long long)a.G + g] = v;
}

// stage 3 on one q2 entry: exact FP64; fused eps-acceptance
template <int LAYOUT>
__device__ __forceinline__ void score2_exact_round(const AbcScoreArgs& a, Score2Smem& sm, long long i0, int e) {
    const unsigned int id = sm.q2_id[e];
    const int g = (int)(id >> 16);
    const long long i = i0 + (long long)(id & 0xffffu);
    const double err = exact_pair_error(a.stats + i * ABC_NSTATS, a.d + (long long)g * ABC_NSTATS,
                                        a.den + (long long)g * ABC_NSTATS);
    score2_store<LAYOUT>(a, i, g, err);
    if (err <= a.eps) {
        const unsigned long long slot = atomicAdd(a.acc_count, 1ull);
        atomicAdd(a.counts + g, 1ull);
        if ((long long)slot < a.acc_capacity) {
            a.acc_gene[slot] = g;
            a.acc_particle[slot] = a.particle_offset + i + 1;
            a.acc_err[slot] = err;
        }
    }
}

// drain q2: full rounds of SC2_THREADS entries; the remainder stays queued unless `final`
template <int LAYOUT>
__device__ __noinline__ void score2_drain_q2(const AbcScoreArgs& a, Score2Smem& sm, long long i0, bool final) {
    __syncthreads();
    const int n2 = sm.q2n;
    const int full = final ? n2 : (n2 / SC2_THREADS) * SC2_THREADS;
    for (int e = threadIdx.x; e < full; e += SC2_THREADS) score2_exact_round<LAYOUT>(a, sm, i0, e);
    __syncthreads();
    const int rem = n2 - full;
    unsigned int keep = 0;
    if ((int)threadIdx.x < rem) keep = sm.q2_id[full + threadIdx.x];
    __syncthreads();
    if ((int)threadIdx.x < rem) sm.q2_id[threadIdx.x] = keep;
    if (threadIdx.x == 0) sm.q2n = rem;
    __syncthreads();
}

// drain the warp-private stage-1 queues: stage 2 = FP32 lower bound over the remaining 38 terms, one queued
// pair per thread.  Callers have published their fill counts in sm.q1cnt and reset their registers.
template <int LAYOUT>
__device__ __noinline__ void score2_drain_q1(const AbcScoreArgs& a, Score2Smem& sm, const float2* __restrict__ fbw,
                                const float2* __restrict__ fa, long long i0, bool final) {
    __syncthreads();
    int c0 = sm.q1cnt[0], c1 = c0 + sm.q1cnt[1], c2 = c1 + sm.q1cnt[2], total = c2 + sm.q1cnt[3];
    __syncthreads();                       // every thread has read the counts before they are reset below
    for (int base = 0; base < total; base += SC2_THREADS) {
        if (sm.q2n > SC2_Q2 - SC2_THREADS) score2_drain_q2<LAYOUT>(a, sm, i0, false);      // block uniform
        const int e = base + threadIdx.x;
        if (e < total) {
            const int w = (e >= c2) ? 3 : (e >= c1) ? 2 : (e >= c0) ? 1 : 0;
            const int k = e - ((w == 3) ? c2 : (w == 2) ? c1 : (w == 1) ? c0 : 0);
            const unsigned int id = sm.q1_id[w][k];
            const int g = (int)(id >> 16);
            const long long i = i0 + (long long)(id & 0xffffu);
            const float* sp = a.fstats + i * ABC_NSTATS + SC2_T1;
            const float2* bw = fbw + (long long)g * ABC_NSTATS + SC2_T1;
            float p = sm.q1_p[w][k] + fa[g].y;
#pragma unroll
            for (int t = 0; t < ABC_NSTATS - SC2_T1; ++t) {
                const float sv = sp[t];
                const float2 c = bw[t];
                p = __fmaf_rn(-c.x, sv, p);
                p = __fmaf_rn(c.y, __fmul_rn(sv, sv), p);
            }
            if (p > SC2_SURE) score2_store<LAYOUT>(a, i, g, 10.0);
            else sm.q2_id[atomicAdd(&sm.q2n, 1)] = id;
        }
        __syncthreads();
    }
    if (threadIdx.x < SC2_WARPS) sm.q1cnt[threadIdx.x] = 0;
    __syncthreads();
    if (final) score2_drain_q2<LAYOUT>(a, sm, i0, true);
}

template <int LAYOUT>
__global__ void __launch_bounds__(SC2_THREADS, 4)
abc_score2_kernel(const AbcScoreArgs a, const float2* __restrict__ fbw, const float2* __restrict__ fa) {
    __shared__ Score2Smem sm;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned int lt_mask = (1u << lane) - 1u;
    const long long n_tiles = (a.n + SC2_TILE - 1) / SC2_TILE;
    const int n_gtiles = (a.G + SC2_GT - 1) / SC2_GT;
    const int gt_per = (n_gtiles + gridDim.y - 1) / gridDim.y;
    const int g_lo = min(a.G, (int)blockIdx.y * gt_per * SC2_GT);
    const int g_hi = min(a.G, g_lo + gt_per * SC2_GT);
    const double qnan = __longlong_as_double(0x7ff8000000000000ll);
    constexpr bool pmajor = (LAYOUT == ABC_ERR_PARTICLE_MAJOR);
    constexpr bool gmajor = (LAYOUT == ABC_ERR_GENE_MAJOR);
    if (tid < SC2_WARPS) sm.q1cnt[tid] = 0;
    if (tid == 0) { sm.q2n = 0; sm.any_nan = 0; }
    int q1w = 0;                                      // fill of this warp's queue (warp uniform)

    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long i0 = tile * SC2_TILE;
        const int rows = (int)min((long long)SC2_TILE, a.n - i0);
        float s[SC2_PPT][SC2_T1], s2[SC2_PPT][SC2_T1];
        bool rnan[SC2_PPT];
        unsigned int live_bal[SC2_PPT], nan_bal[SC2_PPT];
        __syncthreads();
        if (tid == 0) sm.any_nan = 0;
        __syncthreads();
#pragma unroll
        for (int pp = 0; pp < SC2_PPT; ++pp) {
            const long long i = i0 + pp * SC2_THREADS + tid;
            const bool live = i < a.n;
            const long long ii = live ? i : (a.n - 1);
            const float* sp = a.fstats + ii * ABC_NSTATS;
#pragma unroll
            for (int t = 0; t < SC2_T1; ++t) { s[pp][t] = sp[t]; s2[pp][t] = __fmul_rn(s[pp][t], s[pp][t]); }
            rnan[pp] = a.rnan[ii] != 0;
            live_bal[pp] = __ballot_sync(0xffffffffu, live);
            nan_bal[pp] = __ballot_sync(0xffffffffu, rnan[pp] && live);
            if (nan_bal[pp] != 0u && lane == 0) sm.any_nan = 1;
        }
        for (int g0 = g_lo; g0 < g_hi; g0 += SC2_GT) {
            const int gt = min(SC2_GT, g_hi - g0);
            __syncthreads();
            for (int k = tid; k < gt * SC2_T1; k += SC2_THREADS)
                sm.bw[k / SC2_T1][k % SC2_T1] = fbw[(long long)(g0 + k / SC2_T1) * ABC_NSTATS + (k % SC2_T1)];
            if (tid < gt) sm.a1[tid] = fa[g0 + tid].x;
            __syncthreads();
            for (int gs = 0; gs < gt; gs += SC2_SUB) {
                const int ge = min(gt, gs + SC2_SUB);
                for (int gg = gs; gg < ge; ++gg) {
                    float p[SC2_PPT];
#pragma unroll
                    for (int pp = 0; pp < SC2_PPT; ++pp) p[pp] = sm.a1[gg];
#pragma unroll
                    for (int t = 0; t < SC2_T1; ++t) {
                        const float2 c = sm.bw[gg][t];
#pragma unroll
                        for (int pp = 0; pp < SC2_PPT; ++pp) {
                            p[pp] = __fmaf_rn(-c.x, s[pp][t], p[pp]);
                            p[pp] = __fmaf_rn(c.y, s2[pp][t], p[pp]);
                        }
                    }
#pragma unroll
                    for (int pp = 0; pp < SC2_PPT; ++pp) {
                        const bool sure = rnan[pp] || (p[pp] > SC2_SURE);
                        const unsigned int bal = __ballot_sync(0xffffffffu, sure);
                        if (gmajor) {
                            if (sure && ((live_bal[pp] >> lane) & 1u))
                                a.err[(long long)(g0 + gg) * a.n + i0 + pp * SC2_THREADS + tid] = rnan[pp] ? qnan : 10.0;
                        } else if (pmajor && lane == 0) {
                            sm.mask[gg][pp * SC2_WARPS + warp] = bal;
                        }
                        const unsigned int need = ~bal & live_bal[pp];
                        if (!sure && ((need >> lane) & 1u)) {
                            const int slot = q1w + __popc(need & lt_mask);
                            sm.q1_id[warp][slot] = ((unsigned int)(g0 + gg) << 16) | (unsigned int)(pp * SC2_THREADS + tid);
                            sm.q1_p[warp][slot] = p[pp];
                        }
                        q1w += __popc(need);
                    }
                }
                // a sub-batch adds at most SC2_SUB * SC2_PPT * 32 = 256 entries per warp
                const bool want_drain = __syncthreads_or(q1w > SC2_Q1W - SC2_SUB * SC2_PPT * 32);
                if (want_drain) {
                    if (lane == 0) sm.q1cnt[warp] = q1w;
                    q1w = 0;
                    score2_drain_q1<LAYOUT>(a, sm, fbw, fa, i0, false);
                }
            }
            // particle-major: the "surely 10.0" values of this gene tile as contiguous row segments
            if (pmajor) {
                __syncthreads();
                unsigned int mw[SC2_TILE / 32];
#pragma unroll
                for (int w = 0; w < SC2_TILE / 32; ++w) mw[w] = (lane < gt) ? sm.mask[lane][w] : 0u;
                double* rowp = a.err + (i0 + warp) * (long long)a.G + g0 + lane;
#pragma unroll
                for (int w = 0; w < SC2_TILE / 32; ++w) {
#pragma unroll
                    for (int j = 0; j < 32 / SC2_WARPS; ++j) {
                        const int b = warp + SC2_WARPS * j, r = 32 * w + b;
                        if (r < rows && ((mw[w] >> b) & 1u)) rowp[(long long)(32 * w + SC2_WARPS * j) * a.G] = 10.0;
                    }
                }
                if (sm.any_nan) {        // rare: rows of particles with a NaN statistic are NaN, not 10.0
                    for (int r = warp; r < rows; r += SC2_WARPS)
                        if (a.rnan[i0 + r] && lane < gt) a.err[(i0 + r) * (long long)a.G + g0 + lane] = qnan;
                }
            }
        }
        if (lane == 0) sm.q1cnt[warp] = q1w;
        q1w = 0;
        score2_drain_q1<LAYOUT>(a, sm, fbw, fa, i0, true);      // end of the particle tile: empty both queues
    }
}

int abc_launch_score_prep(const double* d_stats, int64_t n, float* d_fstats, unsigned char* d_rnan, cudaStream_t st) {
    if (n <= 0) return ABC_OK;
    abc_score_prep_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(d_stats, (long long)n, d_fstats, d_rnan);
    ABC_CUDA_CHECK(cudaGetLastError());
    return ABC_OK;
}

int abc_launch_score(const AbcScoreArgs& a, int sm_count, cudaStream_t st) {
    if (a.n <= 0 || a.G <= 0) return ABC_OK;
    if (a.fbw != nullptr && a.fa != nullptr && a.fstats != nullptr && a.eps < 10.0 && a.G < 65536 && !a.force_reference_kernel) {
        const long long nt = (a.n + SC2_TILE - 1) / SC2_TILE;
        int per = 0;
        const int lay = (a.err == nullptr) ? ABC_ERR_NONE : a.err_layout;
        ABC_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, abc_score2_kernel<ABC_ERR_PARTICLE_MAJOR>, SC2_THREADS, 0));
        if (per < 1) per = 1;
        long long gx2 = (long long)sm_count * per;
        if (gx2 > nt) gx2 = nt;
        const int ngt = (a.G + SC2_GT - 1) / SC2_GT;
        long long gy2 = (4ll * sm_count * per + gx2 - 1) / gx2;
        if (gy2 > ngt) gy2 = ngt;
        if (gy2 > 65535) gy2 = 65535;
        if (gy2 < 1) gy2 = 1;
        const dim3 grid2((unsigned)gx2, (unsigned)gy2);
        if (lay == ABC_ERR_NONE) abc_score2_kernel<ABC_ERR_NONE><<<grid2, SC2_THREADS, 0, st>>>(a, a.fbw, a.fa);
        else if (lay == ABC_ERR_GENE_MAJOR) abc_score2_kernel<ABC_ERR_GENE_MAJOR><<<grid2, SC2_THREADS, 0, st>>>(a, a.fbw, a.fa);
        else abc_score2_kernel<ABC_ERR_PARTICLE_MAJOR><<<grid2, SC2_THREADS, 0, st>>>(a, a.fbw, a.fa);
        ABC_CUDA_CHECK(cudaGetLastError());
        return ABC_OK;
    }
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
