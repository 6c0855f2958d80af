// abc_ssa_dev.cuh -- device helpers shared by the SSA kernels (abc_ssa.cu: six-channel direct method and the hybrid hand-over;
// abc_tele.cu: telegraph SSA + conditional Poisson read-out): explicit-rounding float arithmetic, the Philox word source, exact
// binomials, the Poisson sampler, warp sums.
#pragma once
#include "abc_common.cuh"
#include "abc_internal.h"

// ------------------------------------------------------------------------------------------------
// explicit-rounding float helpers: the file is compiled with -fmad=false, FMAs are explicit
__device__ __forceinline__ float f_add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float f_mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float f_fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }

// deterministic natural log for u in (0, 1]: IEEE ops only (bit-reproducible on the CPU)
__device__ __forceinline__ float log_det(float u) {
    uint32_t ix = __float_as_uint(u) - 0x3f3504f3u;
    int e = (int)ix >> 23;
    float mnt = __uint_as_float((ix & 0x007fffffu) + 0x3f3504f3u);
    float f = f_add(mnt, -1.0f);
    float s = __fdiv_rn(f, f_add(2.0f, f));
    float z = f_mul(s, s);
    float p = f_fma(z, 1.0f / 9.0f, 1.0f / 7.0f);
    p = f_fma(z, p, 1.0f / 5.0f);
    p = f_fma(z, p, 1.0f / 3.0f);
    p = f_fma(z, p, 1.0f);
    float l1p = f_mul(f_add(s, s), p);
    return f_fma((float)e, 0.693147182464599609375f, l1p);
}

template <bool EXACT>
__device__ __forceinline__ float exp_variate(uint32_t w) {
    // u in (0,1]: (w + 0.5) * 2^-32
    float u = f_fma((float)w, 2.3283064365386963e-10f, 1.1641532182693481e-10f);
    if (EXACT) {
        return -log_det(u);
    } else {
        float l;
        asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(u));
        return f_mul(l, -0.693147182464599609375f);
    }
}

template <bool EXACT>
__device__ __forceinline__ float wait_time(float c0, float c1, float E) {
    // solve c0*tau + c1*tau^2/2 = E for tau >= 0 (linear-in-time total propensity)
    float disc = f_fma(f_add(c1, c1), E, f_mul(c0, c0));
    float E2 = f_add(E, E);
    if (EXACT) {
        return __fdiv_rn(E2, f_add(c0, __fsqrt_rn(disc)));
    } else {
        float r, q;
        asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(disc));
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(q) : "f"(f_add(c0, r)));
        return f_mul(E2, q);
    }
}

struct Lineage {
    float U, L;            // molecule counts, exact in binary32 (< 2^24)
    int g;
    uint32_t ctr;          // next Philox block index
    uint32_t c1, c2, c3, k0, k1;
    uint32_t n_events;
};

__device__ __forceinline__ uint4 next_block(Lineage& s) {
    uint4 b = philox4x32_10(s.ctr, s.c1, s.c2, s.c3, s.k0, s.k1);
    s.ctr += 1u;
    return b;
}

__device__ __forceinline__ uint32_t next_word(WordSrc& ws, Lineage& s) {
    if (ws.avail == 0) {
        uint4 b = next_block(s);
        ws.w0 = b.x; ws.w1 = b.y; ws.w2 = b.z; ws.w3 = b.w;
        ws.avail = 4;
    }
    uint32_t r = ws.w0;
    ws.w0 = ws.w1; ws.w1 = ws.w2; ws.w2 = ws.w3;
    ws.avail -= 1;
    return r;
}

// Binomial(n, 1/2): the number of set bits among n fresh random bits
__device__ __forceinline__ uint32_t binhalf(uint32_t n, WordSrc& ws, Lineage& s) {
    uint32_t cnt = 0;
    while (n >= 32u) { cnt += __popc(next_word(ws, s)); n -= 32u; }
    if (n > 0u) cnt += __popc(next_word(ws, s) & ((1u << n) - 1u));
    return cnt;
}

// Binomial(n, B / 2^32), exact: every molecule's uniform is compared with B bit by bit (MSB first);
// at each level the undecided molecules split Bin(m, 1/2).
__device__ __forceinline__ uint32_t binom_q32(uint32_t n, uint32_t B, WordSrc& ws, Lineage& s) {
    uint32_t m = n, acc = 0;
    for (int bit = 31; bit >= 0 && m > 0u; --bit) {
        uint32_t h = binhalf(m, ws, s);
        if ((B >> bit) & 1u) { acc += h; m -= h; }
        else m = h;
    }
    return acc;
}

// exact "g ? x : 0" / "g ? b : a" on the FMA pipe (integer multiply-add on the bit patterns; g in {0,1}):
// the ALU pipe (half rate on sm_100) is the busiest pipe of this kernel, selects would add to it
__device__ __forceinline__ float gate(int g, float x) { return __uint_as_float((uint32_t)g * __float_as_uint(x)); }
__device__ __forceinline__ float pick(int g, float a, float b) {
    return __uint_as_float(__float_as_uint(a) + (uint32_t)g * (__float_as_uint(b) - __float_as_uint(a)));
}

// Poisson(lam): inversion by sequential search below 12 (binary32: the cumulative sum carries ~2^-22 of rounding, i.e. the
// law is exact to ~1e-6 in total variation), Hoermann's PTRS (1993) above; its rarely taken exact acceptance test runs in FP64.
static __device__ __noinline__ float poisson_draw(float lam, WordSrc& ws, Lineage& s) {
    if (!(lam > 0.0f)) return 0.0f;
    if (lam < 12.0f) {
        const float u = f_fma((float)next_word(ws, s), 2.3283064365386963e-10f, 1.1641532182693481e-10f);
        float p;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p) : "f"(f_mul(lam, -1.4426950408889634f)));
        float c = p, k = 0.0f;
        while (u > c && k < 96.0f) {
            k += 1.0f;
            float rk;
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rk) : "f"(k));
            p = f_mul(p, f_mul(lam, rk));
            c = f_add(c, p);
        }
        return k;
    }
    const float slam = sqrtf(lam);
    const float b = 0.931f + 2.53f * slam;
    const float a = -0.059f + 0.02483f * b;
    const float inv_alpha = 1.1239f + 1.1328f / (b - 3.4f);
    const float vr = 0.9277f - 3.6224f / (b - 2.0f);
    for (int it = 0; it < 64; ++it) {
        const float U = f_fma((float)next_word(ws, s), 2.3283064365386963e-10f, -0.5f);
        const float V = f_fma((float)next_word(ws, s), 2.3283064365386963e-10f, 1.1641532182693481e-10f);
        const float us = 0.5f - fabsf(U);
        const float k = floorf((2.0f * a / us + b) * U + lam + 0.43f);
        if (us >= 0.07f && V <= vr) return k;
        if (k < 0.0f || (us < 0.013f && V > us)) continue;
        const double lhs = log((double)V) + log((double)inv_alpha) - log((double)a / ((double)us * (double)us) + (double)b);
        const double rhs = -(double)lam + (double)k * log((double)lam) - lgamma((double)k + 1.0);
        if (lhs <= rhs) return k;
    }
    return floorf(lam + 0.5f);
}

__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

