// abc_score3.cu -- tile-pruned error scoring (compute_errors.jl:30-70 + accepted_particles.jl:20).  sm_100a.
//
// Same results, bit for bit, as abc_score_kernel (abc_score.cu); the work is organised so that the error
// matrix is written at memory speed and arithmetic is spent only where a pair can be below 10.0:
//
//   * once per data set (host, abc_score3_build): genes are clustered into tiles of 32 by recursive bisection on
//     log sqrt(w_t), w_t = 1/(53 den_t); per tile and statistic t < 31 the box [amin, amax] x [bmin, bmax] of
//     a = sqrt(w) d, b = sqrt(w) is stored, so that for every gene of the tile and every s >= 0
//         sqrt(w_t) |d_t - s_t|  >=  max(0, amin - bmax s, bmin s - amax);
//   * abc_score3_classify_kernel: one thread per particle evaluates that bound against every tile (FP32) and
//     emits one "live" bit per (tile, particle): bound <= 10.01.  A cleared bit proves err > 10 for the 32 pairs,
//     which compute_errors.jl:62-64 clips to exactly 10.0;
//   * abc_score3_tile_kernel: one CTA per (tile, block of 2048 particles) first writes its share of the matrix
//     with 10.0 (NaN rows for particles with a NaN statistic) using 16-byte stores, then visits only the live
//     particles of its tile: lane = gene, four particles per pass, packed FP32 FMAs (FFMA2):
//         stage 1: lower bound over the first 15 terms; items without a lane <= 10.01 are finished;
//         stage 2: the remaining 38 terms; pairs still <= 10.01 are queued;
//         stage 3: queued pairs, 32 per round: the reference's FP64 arithmetic in the reference's order;
//     values below 10.0 overwrite the fill, and eps-acceptance is fused into stage 3.
//
// Soundness of the FP32 bounds: every term is >= 0, so partial sums and box bounds bound the total from below.
// With a = fl32(sqrt(w) d), b = fl32(sqrt(w)), s32 = fl32(s) and t = fma(-b, s32, a):
// |t - sqrt(w)(d - s)| <= 2^-24 (3 sqrt(w)|d| + 3 |T|) with sqrt(w)|d| <= sqrt(100/53), hence for an exact total
// P <= 10 the computed sum is below 10.0002 < 10.01.  NaN never compares "> 10.01" and falls through to stage 3.
//
// Stage 3 divides by the per-gene constant den with its precomputed correctly rounded reciprocal y = RN(1/den):
// q0 = RN(x y), two Markstein corrections q <- fma(fma(-den, q, x), y, q); the second one starts from a faithful
// quotient and therefore returns RN(x/den) (Markstein 1990).  Operands outside [1e-250, 1e200], NaN, and genes with
// non-finite or extreme data use div.rn.f64.
//
// Particle-major output: the 32 genes of a tile are scattered over a row, so the fill is done on contiguous
// slices of the block's rows instead; stage 3 runs as the next kernel on the stream, which orders its stores after the fill.
//
// Second half of the file ("tensor-core filter"): the same decision taken by a TF32 GEMM on the tensor cores
// (tcgen05.mma, accumulators in tensor memory) in front of the same FP64 stage; option score_mma_filter, DESIGN.md 6.2.
#include "abc_common.cuh"
#include "abc_internal.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#ifndef S3_FILL_EVICT_FIRST
#define S3_FILL_EVICT_FIRST 1
#endif
#define S3_TG 32
#define S3_NN 31          // statistics used by the tile bound: means, Fano factors, ratios (non-negative)
#define S3_WARPS 5                       // compute warps
#define S3_THREADS (S3_WARPS * 32 + 32)  // + one service warp: table loads and the fill, both through the TMA engine
#define S3_FILL_DOUBLES 512              // 4 KB of 10.0 in shared memory: the source of the bulk stores
#define S3_PB 2048        // particles per CTA
#define S3_NP 4           // particles per warp pass
#define S3_NPAIR 27       // term pairs (53 terms + one zero term)
#define S3_DEPTH 3        // staging ring: the statistics of a pass are requested two passes ahead
#define S3_SROW 56        // floats per staged particle (53 statistics, padded to whole float4)
#define S3_SURE 10.01f
#define S3C_THREADS 128
#define S3C_TCH 24

// ------------------------------------------------------------------------------------------------ host tables
static float f32_down(double x) {
    float f = (float)x;
    if ((double)f > x) f = nextafterf(f, -INFINITY);
    return f;
}
static float f32_up(double x) {
    float f = (float)x;
    if ((double)f < x) f = nextafterf(f, INFINITY);
    return f;
}

void abc_score3_build(const double* d, const double* den, int G, AbcScore3Host& out) {
    const int ntiles = (G + S3_TG - 1) / S3_TG;
    std::vector<unsigned char> ok((size_t)G, 1);
    std::vector<double> feat((size_t)G * S3_NN, 0.0);
    for (int g = 0; g < G; ++g) {
        for (int t = 0; t < ABC_NSTATS; ++t) {
            const double dv = d[(size_t)g * ABC_NSTATS + t], nv = den[(size_t)g * ABC_NSTATS + t];
            if (!(nv >= 1e-30 && nv <= 1e30) || !(std::fabs(dv) <= 1e15)) ok[g] = 0;
        }
        if (ok[g])
            for (int t = 0; t < S3_NN; ++t) feat[(size_t)g * S3_NN + t] = -0.5 * std::log(53.0 * den[(size_t)g * ABC_NSTATS + t]);
    }
    // recursive bisection: split the widest feature at a multiple of the tile size nearest the median
    std::vector<int> idx((size_t)G);
    for (int g = 0; g < G; ++g) idx[g] = g;
    std::vector<std::pair<int, int>> stack, leaves;
    stack.push_back({0, G});
    while (!stack.empty()) {
        const auto r = stack.back();
        stack.pop_back();
        const int len = r.second - r.first;
        if (len <= S3_TG) { leaves.push_back(r); continue; }
        int best = 0;
        double spread = -1.0;
        for (int t = 0; t < S3_NN; ++t) {
            double lo = INFINITY, hi = -INFINITY;
            for (int k = r.first; k < r.second; ++k) {
                const double v = feat[(size_t)idx[k] * S3_NN + t];
                lo = std::min(lo, v); hi = std::max(hi, v);
            }
            if (hi - lo > spread) { spread = hi - lo; best = t; }
        }
        std::stable_sort(idx.begin() + r.first, idx.begin() + r.second, [&](int x, int y) {
            return feat[(size_t)x * S3_NN + best] < feat[(size_t)y * S3_NN + best];
        });
        int h = std::max(S3_TG, ((len / 2 + S3_TG - 1) / S3_TG) * S3_TG);
        if (h >= len) h = len / 2;
        stack.push_back({r.first + h, r.second});
        stack.push_back({r.first, r.first + h});
    }
    std::sort(leaves.begin(), leaves.end());
    out.ntiles = ntiles;
    out.tb.assign((size_t)ntiles * S3_NN * 4, 0.f);
    out.ab.assign((size_t)ntiles * S3_NPAIR * S3_TG * 4, 0.f);
    out.wt.assign((size_t)ntiles * 6 * ABC_NSTATS * S3_TG, 0u);
    out.gidx.assign((size_t)ntiles * S3_TG, -1);
    out.okmask.assign((size_t)ntiles, 0u);
    const float qn = std::nanf("");
    for (int T = 0; T < ntiles && T < (int)leaves.size(); ++T) {
        const int lo = leaves[T].first, cnt = leaves[T].second - leaves[T].first;
        bool tile_ok = true;
        for (int l = 0; l < cnt; ++l) {
            const int g = idx[lo + l];
            out.gidx[(size_t)T * S3_TG + l] = g;
            if (ok[g]) out.okmask[T] |= 1u << l; else tile_ok = false;
            for (int t = 0; t < ABC_NSTATS; ++t) {
                const double dv = d[(size_t)g * ABC_NSTATS + t], nv = den[(size_t)g * ABC_NSTATS + t];
                // stage-3 constants d, den, RN(1/den) as separate high and low words: a 32-bit table row is one
                // conflict-free shared-memory wavefront for any set of gene slots
                const double w3[3] = {dv, nv, 1.0 / nv};
                for (int q = 0; q < 3; ++q) {
                    uint64_t bits;
                    memcpy(&bits, &w3[q], 8);
                    out.wt[(((size_t)T * 6 + 2 * q) * ABC_NSTATS + t) * S3_TG + l] = (uint32_t)(bits >> 32);
                    out.wt[(((size_t)T * 6 + 2 * q + 1) * ABC_NSTATS + t) * S3_TG + l] = (uint32_t)bits;
                }
                float a = qn, b = qn;
                if (ok[g]) { const double sw = std::sqrt(1.0 / (53.0 * nv)); a = (float)(sw * dv); b = (float)sw; }
                // terms (2j, 2j+1) packed for FFMA2: (a_2j, a_2j+1, -b_2j, -b_2j+1); term 52 is paired with zeros
                float* pk = &out.ab[(((size_t)T * S3_NPAIR + t / 2) * S3_TG + l) * 4];
                pk[t & 1] = a; pk[2 + (t & 1)] = -b;
            }
        }
        for (int t = 0; t < S3_NN; ++t) {
            double amin = INFINITY, amax = -INFINITY, bmin = INFINITY, bmax = -INFINITY;
            bool use = tile_ok && cnt > 0;
            for (int l = 0; l < cnt && use; ++l) {
                const int g = idx[lo + l];
                const double dv = d[(size_t)g * ABC_NSTATS + t], nv = den[(size_t)g * ABC_NSTATS + t];
                if (!(dv >= 0.0)) { use = false; break; }
                const double sw = std::sqrt(1.0 / (53.0 * nv)), av = sw * dv;
                amin = std::min(amin, av); amax = std::max(amax, av);
                bmin = std::min(bmin, sw); bmax = std::max(bmax, sw);
            }
            float* o = &out.tb[((size_t)T * S3_NN + t) * 4];
            if (use) {
                // (amin, -bmax, bmin, -amax), each rounded so that the bound can only shrink
                o[0] = f32_down(amin * (1.0 - 1e-12)); o[1] = -f32_up(bmax * (1.0 + 1e-12));
                o[2] = f32_down(bmin * (1.0 - 1e-12)); o[3] = -f32_up(amax * (1.0 + 1e-12));
            } else {
                o[0] = -INFINITY; o[1] = 0.f; o[2] = 0.f; o[3] = -INFINITY;      // contributes 0
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ classification
// One thread = two particles (tile constants are read once for both).  Also writes the FP32 copy of the
// statistics used by stages 1-2 and the per-particle NaN bits.
__global__ void __launch_bounds__(S3C_THREADS)
abc_score3_classify_kernel(const double* __restrict__ stats, long long n, int ntiles, const float4* __restrict__ tb,
                           float* __restrict__ fstats, unsigned int* __restrict__ live, unsigned int* __restrict__ nanw,
                           long long W, float sure) {
    __shared__ float4 sh[S3C_TCH][S3_NN];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long base = (long long)blockIdx.x * (2 * S3C_THREADS);
    float sx[2][S3_NN];
    bool alive[2];
    long long word[2];
#pragma unroll
    for (int pp = 0; pp < 2; ++pp) {
        const long long i = base + pp * S3C_THREADS + tid;
        const bool in = i < n;
        const double* sp = stats + (in ? i : (n - 1)) * ABC_NSTATS;
        bool nn = false;
#pragma unroll
        for (int t = 0; t < ABC_NSTATS; ++t) {
            const double v = sp[t];
            nn = nn || (v != v);
            const float f = (float)v;
            if (in) fstats[i * ABC_NSTATS + t] = f;
            if (t < S3_NN) sx[pp][t] = fmaxf(f, 0.f);
        }
        word[pp] = (base + pp * S3C_THREADS) / 32 + warp;
        const unsigned int nb = __ballot_sync(0xffffffffu, in && nn);
        if (lane == 0 && word[pp] < W) nanw[word[pp]] = nb;
        alive[pp] = in && !nn;
    }
    for (int t0 = 0; t0 < ntiles; t0 += S3C_TCH) {
        const int nt = min(S3C_TCH, ntiles - t0);
        __syncthreads();
        for (int k = tid; k < nt * S3_NN; k += S3C_THREADS) (&sh[0][0])[k] = tb[(long long)t0 * S3_NN + k];
        __syncthreads();
        for (int tt = 0; tt < nt; ++tt) {
            float lb0 = 0.f, lb1 = 0.f;
#pragma unroll
            for (int t = 0; t < S3_NN; ++t) {
                const float4 c = sh[tt][t];
                const float u0 = fmaxf(fmaxf(__fmaf_rn(c.y, sx[0][t], c.x), __fmaf_rn(c.z, sx[0][t], c.w)), 0.f);
                const float u1 = fmaxf(fmaxf(__fmaf_rn(c.y, sx[1][t], c.x), __fmaf_rn(c.z, sx[1][t], c.w)), 0.f);
                lb0 = __fmaf_rn(u0, u0, lb0);
                lb1 = __fmaf_rn(u1, u1, lb1);
            }
            const unsigned int b0 = __ballot_sync(0xffffffffu, alive[0] && !(lb0 > sure));
            const unsigned int b1 = __ballot_sync(0xffffffffu, alive[1] && !(lb1 > sure));
            if (lane == 0) {
                if (word[0] < W) live[(long long)(t0 + tt) * W + word[0]] = b0;
                if (word[1] < W) live[(long long)(t0 + tt) * W + word[1]] = b1;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ tile kernel
struct S3Smem {
    float4 ab[S3_NPAIR][S3_TG];                // (a_2j, a_2j+1, -b_2j, -b_2j+1) per (term pair, gene slot): held in registers
    float st[S3_WARPS][S3_DEPTH][S3_NP][S3_SROW];   // statistics of the four particles of a pass: cp.async ring per warp
    unsigned short list[S3_PB + 2 * S3_WARPS * S3_NP];   // live particles of this (tile, block), padded
    alignas(16) double tens[S3_FILL_DOUBLES];  // 10.0: source of the fill's bulk stores
    unsigned long long mbar;                   // completion of the table load
    int gidx[S3_TG];
    unsigned int livew[S3_PB / 32];
    int cnt[S3_PB / 32];
    int nlist;
    unsigned int qcount;                       // pairs queued for stage 3 by this CTA
    int next;                                  // next unassigned entry of the list (passes are handed out dynamically)
    int warps_done;
};

// shared memory of the stage-3 kernel
struct S3ExactSmem {
    unsigned int wt[6][ABC_NSTATS][S3_TG];     // d, den, 1/den as (high, low) words
    unsigned long long mbar;
    int gidx[S3_TG];
};

// ---- TMA / mbarrier helpers (PTX ISA: cp.async.bulk, mbarrier)
__device__ __forceinline__ unsigned int s3_saddr(const void* p) { return (unsigned int)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void s3_mbar_init(unsigned long long* mbar, unsigned int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s3_saddr(mbar)), "r"(count) : "memory");
}
__device__ __forceinline__ void s3_fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void s3_mbar_expect_tx(unsigned long long* mbar, unsigned int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s3_saddr(mbar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void s3_bulk_g2s(void* dst, const void* src, unsigned int bytes, unsigned long long* mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(s3_saddr(dst)), "l"(src), "r"(bytes), "r"(s3_saddr(mbar)) : "memory");
}
__device__ __forceinline__ void s3_mbar_wait(unsigned long long* mbar, unsigned int phase) {
    unsigned int ok;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(s3_saddr(mbar)), "r"(phase) : "memory");
    } while (!ok);
}
// The background is written once and not read again on the device: evict-first keeps 3.6 GB of streaming stores from
// displacing the statistics, tables and operands that the kernels running beside them re-read from L2.
__device__ __forceinline__ void s3_bulk_s2g(void* dst, const void* src, unsigned int bytes) {
#if S3_FILL_EVICT_FIRST
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
                 ::"l"(dst), "r"(s3_saddr(src)), "r"(bytes), "l"(pol) : "memory");
#else
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(s3_saddr(src)), "r"(bytes) : "memory");
#endif
}
__device__ __forceinline__ void s3_bulk_commit_wait() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
// read-once data (tables, sign-bit words): do not displace the statistics rows that stage 3 re-reads from L1
__device__ __forceinline__ uint4 s3_ldg_stream(const uint4* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ unsigned int s3_ld_acquire(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned int s3_ld_relaxed(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// fill p[0 .. len) with 10.0 by one warp: bulk stores (shared -> global) of up to 8 KB from sm.tens, at most one
// plain store at either end for 16-byte alignment.  Returns after the lane has ISSUED its stores.
template <int CH = S3_FILL_DOUBLES>
__device__ __forceinline__ void s3_fill_tma(double* __restrict__ p, long long len, const double* tens, int lane) {
    if (len <= 0) return;
    const long long head = (long long)((reinterpret_cast<unsigned long long>(p) >> 3) & 1ull);
    if (head && lane == 0) p[0] = 10.0;
    const long long mid = (len - head) & ~1ll;
    for (long long c = (long long)lane * CH; c < mid; c += 32ll * CH) {
        const long long m = min((long long)CH, mid - c);
        s3_bulk_s2g(p + head + c, tens, (unsigned int)(m * 8));
    }
    if (head + mid < len && lane == 0) p[len - 1] = 10.0;
}

__device__ __forceinline__ bool s3_fast_range(double x) {
    // hi words of ~1.4e-250 and ~1.5e200: also false for NaN/Inf (x is never negative)
    return ((unsigned int)__double2hiint(x) - 0x0C100000u) <= (0x69800000u - 0x0C100000u);
}

// x / den through the stored reciprocal (two Markstein corrections, see the header); operands must be in range
__device__ __forceinline__ double s3_div_fast(double x, double den, double rcp) {
    const double q0 = __dmul_rn(x, rcp);
    const double q1 = __fma_rn(__fma_rn(-den, q0, x), rcp, q0);
    return __fma_rn(__fma_rn(-den, q1, x), rcp, q1);
}

__device__ __noinline__ double s3_div_slow(double x, double den) { return __ddiv_rn(x, den); }

__device__ __forceinline__ double s3_word(const S3ExactSmem& sm, int q, int t, int gl) {
    return __hiloint2double((int)sm.wt[2 * q][t][gl], (int)sm.wt[2 * q + 1][t][gl]);
}

// compute_errors.jl:30-43: N consecutive terms of one group, e <- e + (d-s)^2/den in the reference's order.
// The N quotients are independent (loads and Markstein chains overlap); only the final additions are ordered.
template <int N>
__device__ __forceinline__ double s3_chunk(double e, const double* __restrict__ sp, const S3ExactSmem& sm, int gl, int t0, bool ok) {
    double xx[N], q[N];
    bool fast = ok;
#pragma unroll
    for (int j = 0; j < N; ++j) {
        const double diff = __dadd_rn(s3_word(sm, 0, t0 + j, gl), -sp[t0 + j]);
        xx[j] = __dmul_rn(diff, diff);
        fast = fast && s3_fast_range(xx[j]);
        q[j] = s3_div_fast(xx[j], s3_word(sm, 1, t0 + j, gl), s3_word(sm, 2, t0 + j, gl));
    }
    if (!fast) {
#pragma unroll
        for (int j = 0; j < N; ++j)
            if (!(ok && s3_fast_range(xx[j]))) q[j] = s3_div_slow(xx[j], s3_word(sm, 1, t0 + j, gl));
    }
#pragma unroll
    for (int j = 0; j < N; ++j) e = __dadd_rn(e, q[j]);
    return e;
}

__device__ __forceinline__ double s3_div53(double e) {
    if (s3_fast_range(e)) return s3_div_fast(e, 53.0, 1.0 / 53.0);
    return s3_div_slow(e, 53.0);
}

// compute_errors.jl:55-64 for one (particle, gene slot); loops stay rolled (instruction cache)
__device__ __forceinline__ double s3_exact(const double* __restrict__ sp, const S3ExactSmem& sm, int gl, bool ok) {
    double err = 0.0;
#pragma unroll 1
    for (int l = 0; l < 4; ++l)                       // pulse_mean, pulse_ff, chase_mean, chase_ff
        err = __dadd_rn(err, s3_div53(s3_chunk<5>(0.0, sp, sm, gl, 5 * l, ok)));
#pragma unroll 1
    for (int l = 0; l < 3; ++l) {                     // ratio, mean_corr, corr_mean
        double e = 0.0;
#pragma unroll 1
        for (int c = 0; c < 2; ++c) e = s3_chunk<4>(e, sp, sm, gl, 20 + 11 * l + 4 * c, ok);
        e = s3_chunk<3>(e, sp, sm, gl, 28 + 11 * l, ok);
        err = __dadd_rn(err, s3_div53(e));
    }
    return (err > 10.0) ? 10.0 : err;
}

template <int LAYOUT>
__global__ void __launch_bounds__(S3_THREADS, 2)
abc_score3_tile_kernel(const AbcScoreArgs a, const AbcScore3Tables x) {
    extern __shared__ __align__(128) unsigned char s3_raw[];
    S3Smem& sm = *reinterpret_cast<S3Smem*>(s3_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int T = (int)(blockIdx.x % (unsigned int)x.ntiles);
    const long long k = blockIdx.x / (unsigned int)x.ntiles;
    const long long i0 = k * S3_PB;
    const double qnan = __longlong_as_double(0x7ff8000000000000ll);

    if (tid < S3_PB / 32) {
        const long long wi = k * (S3_PB / 32) + tid;
        const unsigned int lw = (wi < x.W) ? x.live[(long long)T * x.W + wi] : 0u;
        sm.livew[tid] = lw;
        sm.cnt[tid] = __popc(lw);
    }
    if (tid < S3_TG) sm.gidx[tid] = x.gidx[T * S3_TG + tid];
    if (tid == 0) { sm.qcount = 0u; sm.next = 0; sm.warps_done = 0; }
    if (warp == S3_WARPS) {
        for (int j = lane; j < S3_FILL_DOUBLES; j += 32) sm.tens[j] = 10.0;
        if (lane == 0) s3_mbar_init(&sm.mbar, 1);
        s3_fence_proxy_async();                // sm.tens and the barrier become visible to the async proxy
    }
    __syncthreads();

    if (warp == S3_WARPS) {
        // ================= service warp: gene constants in, fill out, both as bulk copies; then it retires.
        // Nothing else writes the matrix in this kernel; stage 3 runs as the next kernel on the stream.
        if (lane == 0) {
            s3_mbar_expect_tx(&sm.mbar, (unsigned int)sizeof(sm.ab));
            s3_bulk_g2s(&sm.ab[0][0], x.ab + (long long)T * S3_NPAIR * S3_TG, (unsigned int)sizeof(sm.ab), &sm.mbar);
        }
        if (LAYOUT != ABC_ERR_NONE) {
            const int rows = (int)min((long long)S3_PB, a.n - i0);
            bool nn = false;
            for (int j = lane; j < S3_PB / 32; j += 32) {
                const long long wi = k * (S3_PB / 32) + j;
                nn = nn || (wi < x.W && x.nanw[wi] != 0u);
            }
            // a particle with a NaN statistic has NaN errors for every gene (all 53 statistics enter every error)
            const bool any_nan = __any_sync(0xffffffffu, nn);
            if (LAYOUT == ABC_ERR_GENE_MAJOR) {
                // this CTA's own 32 gene rows x 2048 particles
                for (int sl = 0; sl < S3_TG; ++sl) {
                    const int g = sm.gidx[sl];
                    if (g < 0) continue;
                    double* p = a.err + (long long)g * a.gm_stride + i0;
                    if (!any_nan) {
                        s3_fill_tma(p, rows, sm.tens, lane);
                    } else {
                        for (int r = lane; r < rows; r += 32)
                            p[r] = ((x.nanw[k * (S3_PB / 32) + (r >> 5)] >> (r & 31)) & 1u) ? qnan : 10.0;
                    }
                }
            } else {
                // the block's rows are contiguous: slice T of ntiles
                const long long L = (long long)rows * a.G;
                long long chunk = (L + x.ntiles - 1) / x.ntiles;
                chunk += chunk & 1;
                const long long lo = min(L, (long long)T * chunk), hi = min(L, lo + chunk);
                double* base = a.err + i0 * (long long)a.G;
                if (!any_nan) {
                    s3_fill_tma(base + lo, hi - lo, sm.tens, lane);
                } else {
                    for (long long j = lo + lane; j < hi; j += 32) {
                        const int r = (int)(j / a.G);
                        base[j] = ((x.nanw[k * (S3_PB / 32) + (r >> 5)] >> (r & 31)) & 1u) ? qnan : 10.0;
                    }
                }
            }
            s3_bulk_commit_wait();             // the source buffer must outlive the reads; the kernel boundary orders the rest
        }
        return;
    }

    // ================= compute warps
    // ---- list of live particles
    if (tid < S3_PB / 32) {
        int off = 0;
        for (int q = 0; q < tid; ++q) off += sm.cnt[q];
        unsigned int w = sm.livew[tid];
        while (w) {
            const int b = __ffs(w) - 1;
            w &= w - 1u;
            sm.list[off++] = (unsigned short)(tid * 32 + b);
        }
        if (tid == S3_PB / 32 - 1) sm.nlist = off;
    }
    asm volatile("bar.sync 1, %0;" ::"n"(S3_WARPS * 32) : "memory");
    const int nl = sm.nlist;
    // pad the list with its last entry so that passes can always read four entries and prefetch one pass ahead
    if (nl > 0 && tid < 2 * S3_WARPS * S3_NP) sm.list[nl + tid] = sm.list[nl - 1];
    asm volatile("bar.sync 1, %0;" ::"n"(S3_WARPS * 32) : "memory");
    s3_mbar_wait(&sm.mbar, 0);                 // gene constants have landed (also: no bulk copy in flight at exit)

    // ---- filter passes: four live particles x 53 terms per pass, the gene constants of the lane in registers.
    //      Passes are handed out by a shared counter; the statistics of a pass arrive through cp.async two passes ahead.
    const bool lane_valid = sm.gidx[lane] >= 0;
    const unsigned int lt_mask = (1u << lane) - 1u;
    unsigned short* seg = x.q2 + (size_t)blockIdx.x * (S3_PB * S3_TG);
    if (nl > 0) {
        for (int j = lane; j < S3_DEPTH * S3_NP * (S3_SROW - ABC_NSTATS); j += 32)       // the padding stays 0
            (&sm.st[warp][0][0][0])[(j / (S3_SROW - ABC_NSTATS)) * S3_SROW + ABC_NSTATS + j % (S3_SROW - ABC_NSTATS)] = 0.f;
        float4 c[S3_NPAIR];
#pragma unroll
        for (int j = 0; j < S3_NPAIR; ++j) c[j] = sm.ab[j][lane];
        int bq[S3_DEPTH];                      // list positions of the passes in flight (slot = pass number mod depth)
        auto request = [&](int slot) {
            int b = 0;
            if (lane == 0) b = atomicAdd(&sm.next, S3_NP);
            b = __shfl_sync(0xffffffffu, b, 0);
            if (b < nl) {
#pragma unroll
                for (int p = 0; p < S3_NP; ++p) {
                    const float* sp = a.fstats + (i0 + sm.list[b + p]) * ABC_NSTATS;
                    float* dst = &sm.st[warp][slot][p][0];
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s3_saddr(dst + lane)), "l"(sp + lane) : "memory");
                    if (lane < ABC_NSTATS - 32)
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s3_saddr(dst + 32 + lane)), "l"(sp + 32 + lane) : "memory");
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            return b;
        };
        __syncwarp();
#pragma unroll
        for (int j = 0; j < S3_DEPTH; ++j) bq[j] = request(j);
        int slot = 0;
        while (true) {
            int base = bq[0];
#pragma unroll
            for (int j = 1; j < S3_DEPTH; ++j) base = (slot == j) ? bq[j] : base;
            if (base >= nl) break;
            const int nv = min(S3_NP, nl - base);
            asm volatile("cp.async.wait_group %0;" ::"n"(S3_DEPTH - 1) : "memory");
            __syncwarp();
            const float* stw = &sm.st[warp][slot][0][0];
            float2 acc[S3_NP];
#pragma unroll
            for (int p = 0; p < S3_NP; ++p) acc[p] = make_float2(0.f, 0.f);
#pragma unroll
            for (int q = 0; q < S3_SROW / 4; ++q) {
#pragma unroll
                for (int p = 0; p < S3_NP; ++p) {
                    const float4 sv = *reinterpret_cast<const float4*>(stw + p * S3_SROW + 4 * q);    // broadcast
                    const float2 t0 = __ffma2_rn(make_float2(c[2 * q].z, c[2 * q].w), make_float2(sv.x, sv.y),
                                                 make_float2(c[2 * q].x, c[2 * q].y));
                    acc[p] = __ffma2_rn(t0, t0, acc[p]);
                    if (2 * q + 1 < S3_NPAIR) {
                        const float2 t1 = __ffma2_rn(make_float2(c[2 * q + 1].z, c[2 * q + 1].w), make_float2(sv.z, sv.w),
                                                     make_float2(c[2 * q + 1].x, c[2 * q + 1].y));
                        acc[p] = __ffma2_rn(t1, t1, acc[p]);
                    }
                }
            }
            // queue the pairs that may be below the threshold for stage 3 in this CTA's segment: particle << 5 | gene slot
            unsigned int bal[S3_NP];
            bool unsure[S3_NP];
            int total = 0;
#pragma unroll
            for (int p = 0; p < S3_NP; ++p) {
                const float pp = __fadd_rn(acc[p].x, acc[p].y);
                unsure[p] = lane_valid && p < nv && !(pp > x.sure);
                bal[p] = __ballot_sync(0xffffffffu, unsure[p]);
                total += __popc(bal[p]);
            }
            if (total > 0) {
                unsigned int pos = 0;
                if (lane == 0) pos = atomicAdd(&sm.qcount, (unsigned int)total);
                pos = __shfl_sync(0xffffffffu, pos, 0);
#pragma unroll
                for (int p = 0; p < S3_NP; ++p) {
                    if (unsure[p]) seg[pos + __popc(bal[p] & lt_mask)] = (unsigned short)((sm.list[base + p] << 5) | lane);
                    pos += __popc(bal[p]);
                }
            }
            __syncwarp();                      // every lane is done with this slot before it is refilled
            const int nb = request(slot);
#pragma unroll
            for (int j = 0; j < S3_DEPTH; ++j) bq[j] = (slot == j) ? nb : bq[j];
            slot = (slot + 1 == S3_DEPTH) ? 0 : slot + 1;
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    // the last warp to finish publishes the fill of the segment
    __syncwarp();
    if (lane == 0) {
        __threadfence_block();
        if (atomicAdd(&sm.warps_done, 1) == S3_WARPS - 1) x.qcnt[blockIdx.x] = *(volatile unsigned int*)&sm.qcount;
    }
}

// Stage 3: one CTA per (tile, particle block) segment, one queued pair per thread: the reference's FP64 arithmetic
// (compute_errors.jl:30-43, 55-64), the value stored over the fill, fused eps-acceptance (accepted_particles.jl:20).
#define S3E_THREADS 256
template <int LAYOUT>
__global__ void __launch_bounds__(S3E_THREADS, 4)
abc_score3_exact_kernel(const AbcScoreArgs a, const AbcScore3Tables x) {
    __shared__ __align__(128) S3ExactSmem sm;
    const unsigned int cnt = x.qcnt[blockIdx.x];
    if (cnt == 0u) return;
    const int tid = threadIdx.x, lane = tid & 31;
    const int T = (int)(blockIdx.x % (unsigned int)x.ntiles);
    const long long i0 = (long long)(blockIdx.x / (unsigned int)x.ntiles) * S3_PB;
    if (tid < S3_TG) sm.gidx[tid] = x.gidx[T * S3_TG + tid];
    if (tid == 0) {
        s3_mbar_init(&sm.mbar, 1);
        s3_fence_proxy_async();
    }
    __syncthreads();
    if (tid == 0) {
        s3_mbar_expect_tx(&sm.mbar, (unsigned int)sizeof(sm.wt));
        s3_bulk_g2s(&sm.wt[0][0][0], x.wt + (long long)T * 6 * ABC_NSTATS * S3_TG, (unsigned int)sizeof(sm.wt), &sm.mbar);
    }
    const unsigned int okmask = x.okmask[T];
    const unsigned short* seg = x.q2 + (size_t)blockIdx.x * (S3_PB * S3_TG);
    const unsigned int lt_mask = (1u << lane) - 1u;
    s3_mbar_wait(&sm.mbar, 0);
    for (unsigned int e0 = 0; e0 < cnt; e0 += S3E_THREADS) {
        const unsigned int e = e0 + tid;
        bool acc = false;
        double err = 0.0;
        long long i = 0;
        int g = 0;
        if (e < cnt) {
            const unsigned int ent = seg[e];
            const int gl = (int)(ent & 31u);
            i = i0 + (long long)(ent >> 5);
            g = sm.gidx[gl];
            err = s3_exact(a.stats + i * ABC_NSTATS, sm, gl, ((okmask >> gl) & 1u) != 0u);
            if (LAYOUT == ABC_ERR_GENE_MAJOR) a.err[(long long)g * a.gm_stride + i] = err;
            else if (LAYOUT == ABC_ERR_PARTICLE_MAJOR) a.err[i * (long long)a.G + g] = err;
            acc = err <= a.eps;                       // NaN is never accepted
        }
        const unsigned int mask = __ballot_sync(0xffffffffu, acc);
        if (mask != 0u) {
            const int leader = __ffs(mask) - 1;
            unsigned long long slot = 0;
            if (lane == leader) slot = atomicAdd(a.acc_count, (unsigned long long)__popc(mask));
            slot = __shfl_sync(0xffffffffu, slot, leader) + (unsigned long long)__popc(mask & lt_mask);
            if (acc) {
                atomicAdd(a.counts + g, 1ull);
                if ((long long)slot < a.acc_capacity) {
                    a.acc_gene[slot] = g;
                    a.acc_particle[slot] = a.particle_offset + i + 1;     // 1-based like Julia
                    a.acc_err[slot] = err;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ tensor-core filter
// The same decision as classification + filter -- "can this pair be <= 10?" -- taken by one TF32 GEMM on the 5th-generation
// tensor cores (tcgen05.mma, accumulators in tensor memory) instead of FP32 FMAs on the SM:
//
//     sum_t w_t (d_t - s_t)^2  =  sum_t a_t^2  -  2 sum_t (a_t b_t) s_t  +  sum_t b_t^2 s_t^2 ,     a = sqrt(w) d, b = sqrt(w),
//
// i.e. V[i][g] = A[i][:] . B[g][:] with  A[i] = (s_0..s_52, s_0^2..s_52^2, 1, 0/1 row-invalid flag, 0...)  per particle and
// B[g] = (-2 a_t b_t, b_t^2, sum a^2 - thr_g, 1, 0...) per gene, all rounded to TF32 beforehand (the low 13 mantissa bits are
// zero, so whatever the hardware does with them does not matter).  A pair goes to stage 3 iff V < 0 (sign bit).
//
// Soundness (true error <= 10  =>  V < 0).  Write R = sum a_t^2 and Q = sum b_t^2 s_t^2.  If the true error is <= 10 then
// sqrt(Q) <= sqrt(10) + sqrt(R) (triangle inequality), sum |2 a_t b_t s_t| <= 2 sqrt(R Q) (Cauchy-Schwarz), every operand
// carries a relative rounding error <= 2^-11 (+2^-24 for the double rounding through binary32), so each product is off by at
// most 2^-10 (1 + 2^-9) of its size, and the FP32 accumulation of K = 128 products adds at most K 2^-21 of the sum of the
// magnitudes (two ulps per addition: covers truncating adders).  Hence V + thr_g - 10 <= slack_g with
//     slack_g = 1.002 2^-10 (2 sqrt(R Qmax) + Qmax) + 2^-14 (R + 2 sqrt(R Qmax) + Qmax),   Qmax = (sqrt(10) + sqrt(R))^2,
// and thr_g = 10 + slack_g + 0.002 (the constant is rounded towards -inf).  R <= 100 because den >= 0.01 d^2, so
// slack_g <= 0.45; on the shipped data the median is 0.22 and 6 % more pairs reach stage 3 than are truly <= 10.
// Elements that are not finite in binary32 (|s| > 1.8e19 or Inf) enter as 0: such a statistic puts the true error of every
// gene with usable constants above 10, so any decision is sound.  Rows with a NaN statistic (all errors NaN) and rows past
// the end of the batch are all-zero with the invalid flag set: V = +1.  Genes whose constants cannot use the fast division
// have B = (0, ..., -1, 1): V = -1 for valid rows.  Padding slots have B = (0, ..., +1, 1): V = +1.
//
// Operand layout: the canonical K-major UMMA layout with the 128-byte swizzle.  A K slice of 32 values is one 128-byte row per
// particle (gene); eight rows form a 1024-byte atom in which the 16-byte chunk c of row r sits at chunk position c ^ (r % 8);
// atoms of consecutive row groups follow each other (SBO = 1024).  One MMA consumes K = 8 = 32 bytes of every row, so the
// descriptor's start address advances by 32 bytes per step inside a slice.  (The un-swizzled "interleaved" layout with
// LBO / SBO strides gives the same results and was the first version.)  Both operands are written in exactly this image to global memory (B once per data set on the host, in K slices of 256
// genes x 32; A by abc_score_mma_prep_kernel, four slices of 128 particles x 32) and arrive in shared memory by plain bulk
// copies -- no tensor maps.
#define MF_M 128                          // particles per work item (= MMA M, one TMEM lane each)
#define MF_N 256                          // genes per accumulator tile: eight tiles of 32.  At N = 256 one MMA (K = 8) reads
                                          // 4 KB of A and 8 KB of B in its 128 cycles (96 B/clk); at N = 64 it would be
                                          // 4 + 2 KB in 32 cycles (192 B/clk, above the shared-memory rate)
#define MF_K 128
#define MF_KSTAGE 32                      // K per ring slot: the gene operand arrives in slices of 256 genes x 32 K (32 KB)
#define MF_NKS (MF_K / MF_KSTAGE)
#define MF_RING 3
#define MF_A_FLOATS (MF_M * MF_K)         // 64 KB
#define MF_B_FLOATS (MF_N * MF_KSTAGE)    // 32 KB
#define MF_EPI_WARPS 8
#define MF_THREADS (32 * (3 + MF_EPI_WARPS)) // warp 0 loads, warp 1 issues the MMAs, warps 2-9 read the accumulators, warp 10 fills
#define MF_MAXT 112                       // tiles of 32 genes per item (14 x 256 genes)
#define MF_COL_ONE 106
#define MF_COL_INVALID 107

static float tf32_rn(double x) {
    float f = (float)x;
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7f800000u) == 0x7f800000u) return f;
    u = (u + 0x0fffu + ((u >> 13) & 1u)) & 0xffffe000u;
    memcpy(&f, &u, 4);
    return f;
}
static float tf32_down(double x) {
    float f = f32_down(x);
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7f800000u) == 0x7f800000u) return f;
    if (u & 0x1fffu) {
        if (f > 0.f) u &= 0xffffe000u;                    // towards zero = down
        else u = (u | 0x1fffu) + 1u;                       // away from zero = down
    }
    memcpy(&f, &u, 4);
    return f;
}
// float index of element (row r, kk) in the image of a K slice (32 K = one 128-byte row per particle / gene): 8-row atoms of
// 1024 B, the 16-byte chunk c of row r stored at chunk position c ^ (r % 8) (128-byte swizzle)
static inline size_t mf_off(int r, int kk) { return (size_t)(r / 8) * 256 + (size_t)(r % 8) * 32 + (size_t)(((kk / 4) ^ (r % 8)) * 4) + (size_t)(kk % 4); }

int abc_score_mma_tiles(int ntiles) { return (ntiles * S3_TG + MF_N - 1) / MF_N; }

void abc_score_mma_build(const double* d, const double* den, const AbcScore3Host& h, std::vector<float>& bblob, double* max_slack) {
    const int ntn = abc_score_mma_tiles(h.ntiles);
    bblob.assign((size_t)ntn * MF_NKS * MF_B_FLOATS, 0.f);
    double worst = 0.0;
    for (int col = 0; col < ntn * MF_N; ++col) {
        const int r = col % MF_N, T = col / S3_TG, l = col % S3_TG;
        // element (column r of tile col / 256, k): K slice k / 32, image of a 256-row operand inside the slice
        auto at = [&](int k) -> float& {
            return bblob[((size_t)(col / MF_N) * MF_NKS + (size_t)(k / MF_KSTAGE)) * MF_B_FLOATS + mf_off(r, k % MF_KSTAGE)];
        };
        const int g = (T < h.ntiles) ? h.gidx[(size_t)T * S3_TG + l] : -1;
        at(MF_COL_INVALID) = 1.f;
        if (g < 0) { at(MF_COL_ONE) = 1.f; continue; }
        if (!((h.okmask[T] >> l) & 1u)) { at(MF_COL_ONE) = -1.f; continue; }
        double R = 0.0;
        for (int t = 0; t < ABC_NSTATS; ++t) {
            const double w = 1.0 / (53.0 * den[(size_t)g * ABC_NSTATS + t]), dv = d[(size_t)g * ABC_NSTATS + t];
            R += w * dv * dv;
            at(t) = tf32_rn(-2.0 * w * dv);
            at(ABC_NSTATS + t) = tf32_rn(w);
        }
        const double sq = std::sqrt(10.0) + std::sqrt(R), Q = sq * sq, X = 2.0 * std::sqrt(R * Q);
        const double slack = 1.002 * std::ldexp(X + Q, -10) + std::ldexp(R + X + Q, -14);
        worst = std::max(worst, slack);
        at(MF_COL_ONE) = tf32_down(R - (10.0 + slack + 0.002));
    }
    if (max_slack) *max_slack = worst;
}

// statistics -> the A operand (one 64 KB image per 128 particles) and the per-particle NaN bits.  One thread per particle.
__global__ void __launch_bounds__(128)
abc_score_mma_prep_kernel(const double* __restrict__ stats, long long n, float* __restrict__ ablob, unsigned int* __restrict__ nanw, long long W) {
    const long long i = (long long)blockIdx.x * 128 + threadIdx.x;
    const int r = threadIdx.x, lane = r & 31;
    const bool in = i < n;
    const double* sp = stats + (in ? i : 0) * ABC_NSTATS;
    bool nn = false;
#pragma unroll 1
    for (int t = 0; t < ABC_NSTATS; ++t) { const double v = sp[t]; nn = nn || (v != v); }
    const unsigned int nb = __ballot_sync(0xffffffffu, in && nn);
    const long long word = i >> 5;
    if (lane == 0 && word < W) nanw[word] = nb;
    const bool valid = in && !nn;
    float4* dst = reinterpret_cast<float4*>(ablob + (size_t)blockIdx.x * MF_A_FLOATS) + (r >> 3) * 64 + (r & 7) * 8;
    auto elem = [&](int k) -> float {
        if (k == MF_COL_ONE) return valid ? 1.f : 0.f;
        if (k == MF_COL_INVALID) return valid ? 0.f : 1.f;
        if (k > MF_COL_INVALID || !valid) return 0.f;
        const double s = sp[k < ABC_NSTATS ? k : k - ABC_NSTATS];
        float f = (float)(k < ABC_NSTATS ? s : s * s);
        unsigned int u;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(f));
        f = __uint_as_float(u);
        return (fabsf(f) <= 3.0e38f) ? f : 0.f;            // Inf / NaN -> 0 (see the header)
    };
#pragma unroll
    for (int c = 0; c < MF_K / 4; ++c)           // K slice c / 8 (16 KB each), chunk c % 8 of the row, swizzled
        dst[(size_t)(c >> 3) * 1024 + (size_t)((c & 7) ^ (r & 7))] = make_float4(elem(4 * c), elem(4 * c + 1), elem(4 * c + 2), elem(4 * c + 3));
}

struct MfSmem {
    alignas(1024) float a[MF_A_FLOATS];
    alignas(1024) float b[MF_RING][MF_B_FLOATS];
    alignas(16) double tens[S3_FILL_DOUBLES];
    unsigned int mask[MF_MAXT][MF_M];          // queue mode: sign bits of the item, [tile of 32 genes][particle]
    unsigned int qbase[MF_MAXT];               // queue mode: reserved position in the segment of every tile
    unsigned long long a_full, a_empty, b_full[MF_RING], b_empty[MF_RING], acc_full[2], acc_empty[2];
    unsigned int tmem_base;
};

__device__ __forceinline__ void mf_mbar_arrive(unsigned long long* mbar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s3_saddr(mbar)) : "memory");
}
__device__ __forceinline__ void mf_tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void mf_tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void mf_commit(unsigned long long* mbar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s3_saddr(mbar)) : "memory");
}
__device__ __forceinline__ unsigned long long mf_desc(unsigned int saddr) {
    // K-major, 128-byte swizzle: start address, LBO = 1 (unused), SBO = 1024 B (8-row atoms), descriptor version 1, layout 2
    return (unsigned long long)((saddr >> 4) & 0x3fffu) | (1ull << 16) | ((unsigned long long)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void mf_mma(unsigned int tmem_d, unsigned long long adesc, unsigned long long bdesc, unsigned int idesc, unsigned int acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mf_tmem_ld32(unsigned int taddr, unsigned int (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr) : "memory");
}

// Persistent: one CTA per SM.  Work item = (128 particles, one of `ks` ranges of gene tiles): V = A B^T tile by tile -- the
// particle operand stays in shared memory for the item (64 KB), the gene operand streams through a three-slot ring of
// 256 genes x 32 K slices (32 KB each) by bulk copies, accumulators are double-buffered in tensor memory (2 x 256 columns) --
// and the sign bits of every (tile of 32 genes, particle) go to gmask as one 32-bit word.  Warp 0 loads, warp 1 (one thread)
// issues the MMAs and commits them to the ring's and the accumulators' mbarriers, warps 2-9 read the accumulators
// (tcgen05.ld, 32 columns per load), warp 10 writes the background of the first fill_b0 particle blocks.
template <int LAYOUT, bool QUEUE>
__global__ void __launch_bounds__(MF_THREADS, 1)
abc_score_mma_filter_kernel(const AbcScoreArgs a, const AbcScore3Tables x, const float* __restrict__ ablob,
                            const float* __restrict__ bblob, int ntn, int ks, float* __restrict__ dbg) {
    extern __shared__ unsigned char mf_raw[];
    // the operand tiles want their natural alignment whatever the base of the dynamic window is
    MfSmem& sm = *reinterpret_cast<MfSmem*>(mf_raw + ((1024u - (s3_saddr(mf_raw) & 1023u)) & 1023u));
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_items = (int)((a.n + MF_M - 1) / MF_M) * ks;
    if (tid == 0) {
        s3_mbar_init(&sm.a_full, 1);
        s3_mbar_init(&sm.a_empty, 1);
        for (int s = 0; s < MF_RING; ++s) {
            s3_mbar_init(&sm.b_full[s], 1);
            s3_mbar_init(&sm.b_empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            s3_mbar_init(&sm.acc_full[s], 1);
            s3_mbar_init(&sm.acc_empty[s], MF_EPI_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        s3_fence_proxy_async();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(s3_saddr(&sm.tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (warp == 2 + MF_EPI_WARPS) {
        for (int j = lane; j < S3_FILL_DOUBLES; j += 32) sm.tens[j] = 10.0;
        s3_fence_proxy_async();
    }
    mf_tc_fence_before();
    __syncthreads();
    mf_tc_fence_after();
    const unsigned int tmem = *(volatile unsigned int*)&sm.tmem_base;

    if (warp == 0) {
        // ================= loads: the particle tile of the item, its gene tiles through the ring
        if (lane == 0) {
            int g = 0, iter = 0;
            for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++iter) {
                const int blk = it / ks, part = it - blk * ks;
                const int t0 = part * ntn / ks, t1 = (part + 1) * ntn / ks;
                if (iter >= 1) s3_mbar_wait(&sm.a_empty, (unsigned int)((iter - 1) & 1));
                s3_mbar_expect_tx(&sm.a_full, MF_A_FLOATS * 4);
                for (int c = 0; c < 4; ++c)
                    s3_bulk_g2s(sm.a + c * (MF_A_FLOATS / 4), ablob + (size_t)blk * MF_A_FLOATS + c * (MF_A_FLOATS / 4), MF_A_FLOATS, &sm.a_full);
                for (int j = t0; j < t1; ++j) {
                    for (int q = 0; q < MF_NKS; ++q, ++g) {
                        const int s = g % MF_RING, u = g / MF_RING;
                        if (u >= 1) s3_mbar_wait(&sm.b_empty[s], (unsigned int)((u - 1) & 1));
                        s3_mbar_expect_tx(&sm.b_full[s], MF_B_FLOATS * 4);
                        const float* src = bblob + ((size_t)j * MF_NKS + (size_t)q) * MF_B_FLOATS;
                        for (int c = 0; c < 2; ++c)
                            s3_bulk_g2s(sm.b[s] + c * (MF_B_FLOATS / 2), src + c * (MF_B_FLOATS / 2), MF_B_FLOATS * 2, &sm.b_full[s]);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issue: one thread
        if (lane == 0) {
            // instruction descriptor: D = F32, A = B = TF32, both K-major, N = 256, M = 128
            const unsigned int idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned int)(MF_N >> 3) << 17) | ((unsigned int)(MF_M >> 4) << 24);
            const unsigned long long adesc = mf_desc(s3_saddr(sm.a));
            int g = 0, iter = 0, tc = 0;
            for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++iter) {
                const int blk = it / ks, part = it - blk * ks;
                const int t0 = part * ntn / ks, t1 = (part + 1) * ntn / ks;
                s3_mbar_wait(&sm.a_full, (unsigned int)(iter & 1));
                for (int j = t0; j < t1; ++j, ++tc) {
                    const int as = tc & 1;
                    if (tc >= 2) s3_mbar_wait(&sm.acc_empty[as], (unsigned int)(((tc >> 1) - 1) & 1));
                    for (int q = 0; q < MF_NKS; ++q, ++g) {
                        const int s = g % MF_RING, u = g / MF_RING;
                        s3_mbar_wait(&sm.b_full[s], (unsigned int)(u & 1));
                        mf_tc_fence_after();
                        const unsigned long long bdesc = mf_desc(s3_saddr(sm.b[s]));
#pragma unroll
                        for (int kk = 0; kk < MF_KSTAGE / 8; ++kk)       // K slice q of A (16 KB each), 32 bytes along the row per step
                            mf_mma(tmem + (unsigned int)(as * MF_N),
                                   adesc + (unsigned long long)((q * (MF_M * MF_KSTAGE * 4) + kk * 32) >> 4),
                                   bdesc + (unsigned long long)((kk * 32) >> 4), idesc, (q | kk) > 0 ? 1u : 0u);
                        mf_commit(&sm.b_empty[s]);
                    }
                    mf_commit(&sm.acc_full[as]);
                }
                mf_commit(&sm.a_empty);
            }
            // every commit has landed before the CTA may exit (they complete in issue order)
            if (iter > 0) s3_mbar_wait(&sm.a_empty, (unsigned int)((iter - 1) & 1));
        }
    } else if (warp < 2 + MF_EPI_WARPS) {
        // ================= accumulators -> sign bits.  Warp w may read TMEM lanes 32 (w % 4) .. + 31; the two warps of a
        // lane group take four of the eight 32-gene tiles each.  The 32 sign bits of (tile of 32 genes, particle) go to
        // gmask[tile][particle] (128 bytes per warp and store); stage 3 builds its work list from them.
        const int lg = warp & 3, half = (warp - 2) >> 2;
        int tc = 0;
        for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
            const int blk = it / ks, part = it - blk * ks;
            const int t0 = part * ntn / ks, t1 = (part + 1) * ntn / ks;
            const long long ib = (long long)blk * MF_M;
            for (int j = t0; j < t1; ++j, ++tc) {
                const int as = tc & 1;
                s3_mbar_wait(&sm.acc_full[as], (unsigned int)((tc >> 1) & 1));
                mf_tc_fence_after();
#pragma unroll 1
                for (int q = 0; q < 4; ++q) {                  // this warp: columns half * 128 + 32 q .. + 31 of the tile
                    unsigned int v[32];
                    const int col = half * 128 + q * 32;
                    mf_tmem_ld32(tmem + ((unsigned int)(lg * 32) << 16) + (unsigned int)(as * MF_N + col), v);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    unsigned int mm = 0u;
#pragma unroll
                    for (int c = 31; c >= 0; --c) mm = __funnelshift_l(v[c], mm, 1);
                    if (QUEUE) sm.mask[8 * (j - t0) + half * 4 + q][lg * 32 + lane] = mm;
                    else x.gmask[(size_t)(8 * j + half * 4 + q) * (size_t)x.n_pad + (size_t)(ib + lg * 32 + lane)] = mm;
                    if (dbg != nullptr) {
#pragma unroll
                        for (int c = 0; c < 32; ++c)
                            dbg[(size_t)(ib + lg * 32 + lane) * (size_t)(ntn * MF_N) + (size_t)(j * MF_N + col + c)] = __uint_as_float(v[c]);
                    }
                }
                mf_tc_fence_before();
                __syncwarp();
                if (lane == 0) mf_mbar_arrive(&sm.acc_empty[as]);
            }
            if (QUEUE) {
                // the item's sign bits -> entries (particle << 5 | gene slot) in the stage-3 queue segments of
                // (tile, block of 2048 particles): epilogue warp ew takes the tiles ew, ew + 8, ...; one round of atomics
                // reserves room in all of them, then every lane writes the entries of its four particles
                asm volatile("bar.sync 1, %0;" ::"n"(MF_EPI_WARPS * 32) : "memory");
                const int ew = warp - 2, nT = 8 * (t1 - t0), Tg0 = 8 * t0;
                const size_t seg0 = (size_t)(ib >> 11) * (size_t)x.ntiles;
                unsigned int mytot = 0u;
                for (int k = 0; k < 32; ++k) {
                    const int tt = ew + MF_EPI_WARPS * k;
                    if (tt >= nT) break;
                    unsigned int c = (unsigned int)(__popc(sm.mask[tt][lane]) + __popc(sm.mask[tt][lane + 32]) +
                                                    __popc(sm.mask[tt][lane + 64]) + __popc(sm.mask[tt][lane + 96]));
                    c = __reduce_add_sync(0xffffffffu, c);
                    if (lane == k) mytot = c;
                }
                const int myT = Tg0 + ew + MF_EPI_WARPS * lane;
                if (mytot > 0u && myT < x.ntiles) sm.qbase[ew + MF_EPI_WARPS * lane] = atomicAdd(&x.qcnt[seg0 + (size_t)myT], mytot);
                __syncwarp();
                for (int k = 0; k < 32; ++k) {
                    const int tt = ew + MF_EPI_WARPS * k;
                    if (tt >= nT) break;
                    if (Tg0 + tt >= x.ntiles) continue;
                    unsigned int mk[4];
                    int c = 0;
#pragma unroll
                    for (int r = 0; r < 4; ++r) { mk[r] = sm.mask[tt][lane + 32 * r]; c += __popc(mk[r]); }
                    if (!__any_sync(0xffffffffu, c != 0)) continue;
                    int incl = c;
#pragma unroll
                    for (int dd = 1; dd < 32; dd <<= 1) {
                        const int t = __shfl_up_sync(0xffffffffu, incl, dd);
                        if (lane >= dd) incl += t;
                    }
                    unsigned short* seg = x.q2 + (seg0 + (size_t)(Tg0 + tt)) * (size_t)(S3_PB * S3_TG);
                    unsigned int pos = sm.qbase[tt] + (unsigned int)(incl - c);
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const unsigned int pl = (unsigned int)((ib + lane + 32 * r) & (S3_PB - 1));
                        unsigned int mm = mk[r];
                        while (mm) {
                            const int b = __ffs(mm) - 1;
                            mm &= mm - 1u;
                            seg[pos++] = (unsigned short)((pl << 5) | (unsigned int)b);
                        }
                    }
                }
                asm volatile("bar.sync 1, %0;" ::"n"(MF_EPI_WARPS * 32) : "memory");
            }
        }
    }
    if (LAYOUT != ABC_ERR_NONE && warp == 2 + MF_EPI_WARPS) {
        // ================= service warp: the background of the first fill_b0 particle blocks (the memory system is idle
        // under the GEMM); the stage-3 kernel writes the rest while it computes
        const double qnan = __longlong_as_double(0x7ff8000000000000ll);
        const long long R0 = min((long long)a.n, (long long)x.fill_b0 * S3_PB);
        if (LAYOUT == ABC_ERR_PARTICLE_MAJOR) {
            const long long L = R0 * a.G;
            long long chunk = (L + gridDim.x - 1) / gridDim.x;
            chunk += chunk & 1;
            const long long lo = min(L, (long long)blockIdx.x * chunk), hi = min(L, lo + chunk);
            if (hi > lo) {
                const long long w0 = (lo / a.G) >> 5, w1 = ((hi - 1) / a.G) >> 5;
                bool nn = false;
                for (long long wi = w0 + lane; wi <= w1; wi += 32) nn = nn || (wi < x.W && x.nanw[wi] != 0u);
                if (!__any_sync(0xffffffffu, nn)) {
                    s3_fill_tma(a.err + lo, hi - lo, sm.tens, lane);
                } else {
                    for (long long j = lo + lane; j < hi; j += 32) {
                        const long long r = j / a.G;
                        a.err[j] = ((x.nanw[r >> 5] >> (int)(r & 31)) & 1u) ? qnan : 10.0;
                    }
                }
            }
        } else if (R0 > 0) {
            bool nn = false;
            for (long long wi = lane; wi < ((R0 + 31) >> 5); wi += 32) nn = nn || (wi < x.W && x.nanw[wi] != 0u);
            const bool any_nan = __any_sync(0xffffffffu, nn);
            const int g0 = (int)((long long)blockIdx.x * a.G / gridDim.x), g1 = (int)((long long)(blockIdx.x + 1) * a.G / gridDim.x);
            for (int g = g0; g < g1; ++g) {
                double* p = a.err + (long long)g * a.gm_stride;
                if (!any_nan) {
                    s3_fill_tma(p, R0, sm.tens, lane);
                } else {
                    for (long long r = lane; r < R0; r += 32) p[r] = ((x.nanw[r >> 5] >> (int)(r & 31)) & 1u) ? qnan : 10.0;
                }
            }
        }
        s3_bulk_commit_wait();
    }
    mf_tc_fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

// Stage 3 behind the tensor-core filter: one CTA per (tile of 32 genes, block of 2048 particles).  The work list is built
// from the filter's sign-bit words, 256 particles at a time: popcount, block-wide exclusive scan, entries
// (particle << 5 | gene slot) into a ring in shared memory; whenever the ring holds >= 256 pairs, one pair per thread goes
// through the reference's FP64 arithmetic (same code as abc_score3_exact_kernel).  A chunk whose pairs would not fit the
// ring (every particle close to every gene of the tile) is appended warp by warp (<= 1024 pairs) with drains in between.
#define MX_CAP 2048
#define MX_FILL_DOUBLES 256                    // 2 KB: with the tables (40.7 KB) and the ring (4 KB) a CTA stays below 48 KB, so four
                                           // of them fit the 196 KB carve-out and leave the SM 60 KB of L1 for the statistics rows
struct MxSmem {
    S3ExactSmem t;
    alignas(16) double tens[MX_FILL_DOUBLES];     // 10.0: source of the background's bulk stores
    unsigned short ring[MX_CAP];
    int wtot[S3E_THREADS / 32];
};

// The matrix's background (10.0; NaN rows for particles with a NaN statistic).  The first fill_b0 particle blocks are written
// by the service warp of the filter kernel (which leaves the memory system idle otherwise); block j >= fill_b0 is written by
// the stage-3 CTAs of block j - fill_d while they compute: CTA (k, T) writes, of block k + fill_d, slice T of its rows
// (particle-major) or the 32 gene rows of tile T (gene-major), and publishes that through a per-block counter when it is
// done.  The CTAs of block j wait for fill_done[j] == ntiles before their first store.  fill_d particle blocks are more CTAs
// than fit the GPU at once and CTAs are dispatched in blockIdx order, so the writers of a block have normally left the GPU
// before its readers arrive; a CTA never waits for a later block, and the wait is bounded (trap).
template <int LAYOUT>
__device__ __forceinline__ void mx_fill_block(const AbcScoreArgs& a, const AbcScore3Tables& x, const int* gidx, const double* tens,
                                              int T, long long kt, int lane) {
    const double qnan = __longlong_as_double(0x7ff8000000000000ll);
    const long long i0 = kt * S3_PB;
    const int rows = (int)max(0ll, min((long long)S3_PB, a.n - i0));
    if (rows <= 0) return;
    bool nn = false;
    for (int j = lane; j < S3_PB / 32; j += 32) {
        const long long wi = (i0 >> 5) + j;
        nn = nn || (wi < x.W && x.nanw[wi] != 0u);
    }
    const bool any_nan = __any_sync(0xffffffffu, nn);
    if (LAYOUT == ABC_ERR_GENE_MAJOR) {
        const int g = gidx[lane];
        if (g >= 0) {
            double* p = a.err + (long long)g * a.gm_stride + i0;
            if (!any_nan) {
                const int head = (int)((reinterpret_cast<unsigned long long>(p) >> 3) & 1ull);
                if (head) p[0] = 10.0;
                const int mid = (rows - head) & ~1;
                for (int c = 0; c < mid; c += MX_FILL_DOUBLES) s3_bulk_s2g(p + head + c, tens, (unsigned int)(min(MX_FILL_DOUBLES, mid - c) * 8));
                if (head + mid < rows) p[rows - 1] = 10.0;
            } else {
                for (int r = 0; r < rows; ++r) p[r] = ((x.nanw[(i0 >> 5) + (r >> 5)] >> (r & 31)) & 1u) ? qnan : 10.0;
            }
        }
    } else if (LAYOUT == ABC_ERR_PARTICLE_MAJOR) {
        const long long L = (long long)rows * a.G;
        long long chunk = (L + x.ntiles - 1) / x.ntiles;
        chunk += chunk & 1;
        const long long lo = min(L, (long long)T * chunk), hi = min(L, lo + chunk);
        double* base = a.err + i0 * (long long)a.G;
        if (!any_nan) {
            s3_fill_tma<MX_FILL_DOUBLES>(base + lo, hi - lo, tens, lane);
        } else {
            for (long long j = lo + lane; j < hi; j += 32) {
                const int r = (int)(j / a.G);
                base[j] = ((x.nanw[(i0 >> 5) + (r >> 5)] >> (r & 31)) & 1u) ? qnan : 10.0;
            }
        }
    }
}

// by the issuing warp: the stores have completed -> publish them
__device__ __forceinline__ void mx_fill_publish(const AbcScore3Tables& x, long long kt, int lane) {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    s3_fence_proxy_async();
    __threadfence();
    __syncwarp();
    if (lane == 0) atomicAdd(&x.fill_done[kt], 1u);
}

// by one thread before the CTA's first store: the background of this block's rows is in place
__device__ __forceinline__ void mx_wait_block(const AbcScore3Tables& x, long long kblk) {
    // relaxed polls (an acquire load invalidates the SM's L1 on every poll: the statistics rows the other warps re-read);
    // one acquire once the count is there
    unsigned int spins = 0u;
    while (s3_ld_relaxed(&x.fill_done[kblk]) < (unsigned int)x.ntiles) {
        __nanosleep(200);
        if (++spins > (1u << 24)) __trap();       // > 3 s: the dispatch-order assumption failed; fail loudly
    }
    (void)s3_ld_acquire(&x.fill_done[kblk]);
}

template <int LAYOUT>
__device__ __forceinline__ void mx_process(const AbcScoreArgs& a, const MxSmem& sm, unsigned int okmask, long long i0, int head, int tail,
                                           int tid, int lane) {
    const int e = head + tid;
    bool acc = false;
    double err = 0.0;
    long long i = 0;
    int g = 0;
    if (e < tail) {
        const unsigned int ent = sm.ring[e & (MX_CAP - 1)];
        const int gl = (int)(ent & 31u);
        i = i0 + (long long)(ent >> 5);
        g = sm.t.gidx[gl];
        err = s3_exact(a.stats + i * ABC_NSTATS, sm.t, gl, ((okmask >> gl) & 1u) != 0u);
        if (LAYOUT == ABC_ERR_GENE_MAJOR) a.err[(long long)g * a.gm_stride + i] = err;
        else if (LAYOUT == ABC_ERR_PARTICLE_MAJOR) a.err[i * (long long)a.G + g] = err;
        acc = err <= a.eps;                       // NaN is never accepted
    }
    const unsigned int mask = __ballot_sync(0xffffffffu, acc);
    if (mask != 0u) {
        const int leader = __ffs(mask) - 1;
        unsigned long long slot = 0;
        if (lane == leader) slot = atomicAdd(a.acc_count, (unsigned long long)__popc(mask));
        slot = __shfl_sync(0xffffffffu, slot, leader) + (unsigned long long)__popc(mask & ((1u << lane) - 1u));
        if (acc) {
            atomicAdd(a.counts + g, 1ull);
            if ((long long)slot < a.acc_capacity) {
                a.acc_gene[slot] = g;
                a.acc_particle[slot] = a.particle_offset + i + 1;     // 1-based like Julia
                a.acc_err[slot] = err;
            }
        }
    }
}

template <int LAYOUT>
__global__ void __launch_bounds__(S3E_THREADS, 4)
abc_score_mask_exact_kernel(const AbcScoreArgs a, const AbcScore3Tables x) {
    __shared__ __align__(128) MxSmem sm;
    constexpr bool FILL = LAYOUT != ABC_ERR_NONE;
    constexpr int FW = S3E_THREADS / 32 - 1;       // the warp that also writes the background
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int T = (int)(blockIdx.x % (unsigned int)x.ntiles);
    const long long kblk = (long long)(blockIdx.x / (unsigned int)x.ntiles);
    const long long i0 = kblk * S3_PB;
    const unsigned int* mrow = x.gmask + (size_t)T * (size_t)x.n_pad + (size_t)i0;
    const int nrows = (int)min((long long)S3_PB, x.n_rows - i0);
    if (tid < S3_TG) sm.t.gidx[tid] = x.gidx[T * S3_TG + tid];
    if (FILL) {
        for (int j = tid; j < MX_FILL_DOUBLES; j += S3E_THREADS) sm.tens[j] = 10.0;
        s3_fence_proxy_async();
    }
    __syncthreads();
    const long long kt = kblk + x.fill_d;          // the block whose background this CTA writes
    const bool writer = FILL && kt >= x.fill_b0 && kt * S3_PB < a.n;
    if (writer && warp == FW) mx_fill_block<LAYOUT>(a, x, sm.t.gidx, sm.tens, T, kt, lane);
    const bool waiter = FILL && kblk >= x.fill_b0;
    // the gene constants of the tile: plain loads (a bulk copy would queue behind the SM's background stores)
    {
        const uint4* src = reinterpret_cast<const uint4*>(x.wt + (long long)T * 6 * ABC_NSTATS * S3_TG);
        uint4* dst = reinterpret_cast<uint4*>(&sm.t.wt[0][0][0]);
        for (int j = tid; j < (int)(sizeof(sm.t.wt) / 16); j += S3E_THREADS) dst[j] = s3_ldg_stream(src + j);
    }
    const unsigned int okmask = x.okmask[T];
    // the whole block at once, eight consecutive particles per thread: popcounts, one block-wide scan
    unsigned int m8[8];
    int cnt = 0;
    const int r0 = 8 * tid;
    if (r0 + 8 <= nrows) {
        const uint4 lo = s3_ldg_stream(reinterpret_cast<const uint4*>(mrow + r0)), hi = s3_ldg_stream(reinterpret_cast<const uint4*>(mrow + r0 + 4));
        m8[0] = lo.x; m8[1] = lo.y; m8[2] = lo.z; m8[3] = lo.w; m8[4] = hi.x; m8[5] = hi.y; m8[6] = hi.z; m8[7] = hi.w;
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) m8[k] = (r0 + k < nrows) ? mrow[r0 + k] : 0u;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) cnt += __popc(m8[k]);
    int incl = cnt;
#pragma unroll
    for (int dd = 1; dd < 32; dd <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, dd);
        if (lane >= dd) incl += t;
    }
    if (lane == 31) sm.wtot[warp] = incl;
    __syncthreads();
    int total = 0, woff = 0;
#pragma unroll
    for (int w = 0; w < S3E_THREADS / 32; ++w) {
        const int t = sm.wtot[w];
        woff += (w < warp) ? t : 0;
        total += t;
    }
    if (total == 0) {
        if (writer && warp == FW) mx_fill_publish(x, kt, lane);
        return;
    }
    if (total <= MX_CAP) {
        // ---- the usual case: all pairs of the block fit the ring
        int pos = woff + incl - cnt;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            unsigned int m = m8[k];
            while (m) {
                const int b = __ffs(m) - 1;
                m &= m - 1u;
                sm.ring[pos++] = (unsigned short)(((unsigned int)(r0 + k) << 5) | (unsigned int)b);
            }
        }
        __syncthreads();
        if (waiter) {
            if (tid == 0) mx_wait_block(x, kblk);
            __syncthreads();
        }
        for (int head = 0; head < total; head += S3E_THREADS) mx_process<LAYOUT>(a, sm, okmask, i0, head, total, tid, lane);
        if (writer && warp == FW) mx_fill_publish(x, kt, lane);
        return;
    }
    // ---- dense block (every particle close to every gene of the tile): the background first, then 256 particles at a time
    if (waiter && tid == 0) mx_wait_block(x, kblk);
    __syncthreads();
    int head = 0, tail = 0;                        // the same in every thread
    for (int c0 = 0; c0 < nrows; c0 += S3E_THREADS) {
        const int r = c0 + tid;
        unsigned int m = (r < nrows) ? mrow[r] : 0u;
        const int c1 = __popc(m);
        int in1 = c1;
#pragma unroll
        for (int dd = 1; dd < 32; dd <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, in1, dd);
            if (lane >= dd) in1 += t;
        }
        if (lane == 31) sm.wtot[warp] = in1;
        __syncthreads();
        int tot1 = 0, wo1 = 0;
#pragma unroll
        for (int w = 0; w < S3E_THREADS / 32; ++w) {
            const int t = sm.wtot[w];
            wo1 += (w < warp) ? t : 0;
            tot1 += t;
        }
        if (tot1 == 0) { __syncthreads(); continue; }
        if (tail - head + tot1 <= MX_CAP) {
            int pos = tail + wo1 + in1 - c1;
            while (m) {
                const int b = __ffs(m) - 1;
                m &= m - 1u;
                sm.ring[(pos++) & (MX_CAP - 1)] = (unsigned short)(((unsigned int)r << 5) | (unsigned int)b);
            }
            tail += tot1;
            __syncthreads();
            while (tail - head >= S3E_THREADS) {
                mx_process<LAYOUT>(a, sm, okmask, i0, head, tail, tid, lane);
                head += S3E_THREADS;
            }
            __syncthreads();
        } else {
            for (int w = 0; w < S3E_THREADS / 32; ++w) {
                const int wt = sm.wtot[w];
                if (warp == w) {
                    int pos = tail + in1 - c1;
                    while (m) {
                        const int b = __ffs(m) - 1;
                        m &= m - 1u;
                        sm.ring[(pos++) & (MX_CAP - 1)] = (unsigned short)(((unsigned int)r << 5) | (unsigned int)b);
                    }
                }
                tail += wt;
                __syncthreads();
                while (tail - head >= S3E_THREADS) {
                    mx_process<LAYOUT>(a, sm, okmask, i0, head, tail, tid, lane);
                    head += S3E_THREADS;
                }
                __syncthreads();
            }
        }
    }
    if (tail > head) mx_process<LAYOUT>(a, sm, okmask, i0, head, tail, tid, lane);
    if (writer && warp == FW) mx_fill_publish(x, kt, lane);
}

// ------------------------------------------------------------------------------------------------ launcher
size_t abc_score3_blocks(int64_t n) { return (size_t)((n + S3_PB - 1) / S3_PB); }
size_t abc_score3_queue_entries(int64_t n, int ntiles) { return abc_score3_blocks(n) * (size_t)ntiles * (size_t)(S3_PB * S3_TG); }

int abc_launch_score3(const AbcScoreArgs& a, const AbcScore3Tables& x_in, cudaStream_t st) {
    if (a.n <= 0 || a.G <= 0) return ABC_OK;
    const long long nblocks = (a.n + S3_PB - 1) / S3_PB;
    if (nblocks * x_in.ntiles > 0x7fffffffll) { abc_set_error("abc_score: batch too large for one launch"); return ABC_ERR_ARG; }
    const int lay = (a.err == nullptr) ? ABC_ERR_NONE : a.err_layout;
    const int smem = (int)sizeof(S3Smem);
    // "surely above" threshold of the FP32 bounds: 10.01 when the matrix is wanted (every value below 10 is needed);
    // without a matrix only pairs that can be accepted matter, i.e. err <= eps (eps < 10 on this path)
    AbcScore3Tables x = x_in;
    x.sure = S3_SURE;
    if (lay == ABC_ERR_NONE && a.eps == a.eps) x.sure = fminf(S3_SURE, f32_up(a.eps + 0.01));
    // per device and cheap: set on every launch
    ABC_CUDA_CHECK(cudaFuncSetAttribute(abc_score3_tile_kernel<ABC_ERR_NONE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    ABC_CUDA_CHECK(cudaFuncSetAttribute(abc_score3_tile_kernel<ABC_ERR_GENE_MAJOR>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    ABC_CUDA_CHECK(cudaFuncSetAttribute(abc_score3_tile_kernel<ABC_ERR_PARTICLE_MAJOR>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const unsigned int cgrid = (unsigned int)((a.n + 2 * S3C_THREADS - 1) / (2 * S3C_THREADS));
    abc_score3_classify_kernel<<<cgrid, S3C_THREADS, 0, st>>>(a.stats, (long long)a.n, x.ntiles, x.tb, const_cast<float*>(a.fstats),
                                                              x.live, x.nanw, (long long)x.W, x.sure);
    ABC_CUDA_CHECK(cudaGetLastError());
    const unsigned int grid = (unsigned int)(nblocks * x.ntiles);
    if (lay == ABC_ERR_NONE) {
        abc_score3_tile_kernel<ABC_ERR_NONE><<<grid, S3_THREADS, smem, st>>>(a, x);
        abc_score3_exact_kernel<ABC_ERR_NONE><<<grid, S3E_THREADS, 0, st>>>(a, x);
    } else if (lay == ABC_ERR_GENE_MAJOR) {
        abc_score3_tile_kernel<ABC_ERR_GENE_MAJOR><<<grid, S3_THREADS, smem, st>>>(a, x);
        abc_score3_exact_kernel<ABC_ERR_GENE_MAJOR><<<grid, S3E_THREADS, 0, st>>>(a, x);
    } else {
        abc_score3_tile_kernel<ABC_ERR_PARTICLE_MAJOR><<<grid, S3_THREADS, smem, st>>>(a, x);
        abc_score3_exact_kernel<ABC_ERR_PARTICLE_MAJOR><<<grid, S3E_THREADS, 0, st>>>(a, x);
    }
    ABC_CUDA_CHECK(cudaGetLastError());
    return ABC_OK;
}

int abc_launch_score_mma(const AbcScoreArgs& a, const AbcScore3Tables& x_in, float* d_ablob, const float* d_bblob, float* d_dbg,
                         int sm_count, cudaStream_t st) {
    if (a.n <= 0 || a.G <= 0) return ABC_OK;
    const long long nblocks = (a.n + S3_PB - 1) / S3_PB;
    if (nblocks * x_in.ntiles > 0x7fffffffll) { abc_set_error("abc_score: batch too large for one launch"); return ABC_ERR_ARG; }
    const int lay = (a.err == nullptr) ? ABC_ERR_NONE : a.err_layout;
    const int smem = (int)sizeof(MfSmem) + 1024;
    const int ntn = abc_score_mma_tiles(x_in.ntiles);
    const int nblk128 = (int)((a.n + MF_M - 1) / MF_M);
    // one persistent CTA per SM; items are split over gene-tile ranges until there are >= ~6 per CTA (wave quantisation)
    int ks = 1;
    while (ks < 8 && ks * 2 <= ntn && (long long)nblk128 * ks < 6ll * sm_count) ks *= 2;
    AbcScore3Tables x = x_in;
    const bool queue = x.q2 != nullptr;            // queue mode: the filter kernel writes the whole background and the stage-3 queue
    while (queue && (ntn + ks - 1) / ks > MF_MAXT / 8) ks *= 2;      // the item's sign bits must fit the shared-memory array
    const unsigned int pgrid = (unsigned int)std::min<long long>((long long)nblk128 * ks, sm_count);
    // background: the first fill_b0 particle blocks by the filter kernel (about what fits its duration, at least the
    // look-ahead), block j >= fill_b0 by the stage-3 CTAs of block j - fill_d (more blocks than are resident at once)
    const long long resident = 4ll * sm_count / std::max(1, x.ntiles) + 2;
    x.fill_d = (int32_t)resident;
    x.fill_b0 = queue ? (int32_t)nblocks : (int32_t)std::min<long long>(nblocks, std::max<long long>(resident, (nblocks * 3 + 5) / 10));
    if (queue) ABC_CUDA_CHECK(cudaMemsetAsync(x.qcnt, 0, (size_t)nblocks * (size_t)x.ntiles * sizeof(uint32_t), st));
    else if (lay != ABC_ERR_NONE) ABC_CUDA_CHECK(cudaMemsetAsync(x.fill_done, 0, (size_t)(nblocks + 1) * sizeof(uint32_t), st));
    abc_score_mma_prep_kernel<<<(unsigned int)nblk128, 128, 0, st>>>(a.stats, (long long)a.n, d_ablob, x.nanw, (long long)x.W);
    ABC_CUDA_CHECK(cudaGetLastError());
#define MF_LAUNCH(LAY, Q)                                                                                                   \
    do {                                                                                                                    \
        ABC_CUDA_CHECK(cudaFuncSetAttribute(abc_score_mma_filter_kernel<LAY, Q>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
        abc_score_mma_filter_kernel<LAY, Q><<<pgrid, MF_THREADS, smem, st>>>(a, x, d_ablob, d_bblob, ntn, ks, d_dbg);       \
    } while (0)
    if (lay == ABC_ERR_NONE) { if (queue) MF_LAUNCH(ABC_ERR_NONE, true); else MF_LAUNCH(ABC_ERR_NONE, false); }
    else if (lay == ABC_ERR_GENE_MAJOR) { if (queue) MF_LAUNCH(ABC_ERR_GENE_MAJOR, true); else MF_LAUNCH(ABC_ERR_GENE_MAJOR, false); }
    else { if (queue) MF_LAUNCH(ABC_ERR_PARTICLE_MAJOR, true); else MF_LAUNCH(ABC_ERR_PARTICLE_MAJOR, false); }
#undef MF_LAUNCH
    ABC_CUDA_CHECK(cudaGetLastError());
    const unsigned int grid = (unsigned int)(nblocks * x.ntiles);
    if (queue) {
        if (lay == ABC_ERR_NONE) abc_score3_exact_kernel<ABC_ERR_NONE><<<grid, S3E_THREADS, 0, st>>>(a, x);
        else if (lay == ABC_ERR_GENE_MAJOR) abc_score3_exact_kernel<ABC_ERR_GENE_MAJOR><<<grid, S3E_THREADS, 0, st>>>(a, x);
        else abc_score3_exact_kernel<ABC_ERR_PARTICLE_MAJOR><<<grid, S3E_THREADS, 0, st>>>(a, x);
    } else {
        if (lay == ABC_ERR_NONE) abc_score_mask_exact_kernel<ABC_ERR_NONE><<<grid, S3E_THREADS, 0, st>>>(a, x);
        else if (lay == ABC_ERR_GENE_MAJOR) abc_score_mask_exact_kernel<ABC_ERR_GENE_MAJOR><<<grid, S3E_THREADS, 0, st>>>(a, x);
        else abc_score_mask_exact_kernel<ABC_ERR_PARTICLE_MAJOR><<<grid, S3E_THREADS, 0, st>>>(a, x);
    }
    ABC_CUDA_CHECK(cudaGetLastError());
    return ABC_OK;
}
