// abc_common.cuh -- shared device helpers: Philox4x32-10, deterministic exp10, exact binomial thinning.
//
// Everything here is specified to the bit so that the CPU test oracle
// (oracle/oracle_ssa.c, an independent restatement) reproduces it draw for draw.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#define ABC_NAGE 5
#define ABC_NCOND 11
#define ABC_NREAD 55
#define ABC_NSTATS 53

#define ABC_DOM_SSA 0u
#define ABC_DOM_PRIOR 1u

// ---------------------------------------------------------------- Philox4x32-10 (Salmon et al. 2011)
// counter = (block index, particle lo, particle hi, tag), key = seed
__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                               uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
        uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0;
        uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += W0; k1 += W1;
    }
    return make_uint4(c0, c1, c2, c3);
}

// lineage tag: cell (20 bits) | read-out (cond*5+age, 6 bits) << 20 | (m-1) << 26 | domain << 29
__device__ __forceinline__ uint32_t abc_tag(uint32_t cell, uint32_t readout, uint32_t m, uint32_t dom) {
    return (cell & 0xFFFFFu) | (readout << 20) | ((m - 1u) << 26) | (dom << 29);
}

// sequential word source on top of one Philox stream (block counter ctr advances by one per 4 words)
struct WordSrc {
    uint32_t w0, w1, w2, w3;
    int      avail;
};

// ---------------------------------------------------------------- deterministic exp10 (IEEE ops only)
// 10^x = 2^n * 2^r,  n = rint(x*log2(10)), r in [-0.5,0.5]; 2^r by a degree-13 Taylor polynomial in
// r*ln2 evaluated with explicit FMAs.  Only correctly rounded basic operations are used, so CPU
// (fma()) and GPU (__fma_rn) agree to the bit.  |rel err| < 4e-16 on the prior range.
__host__ __device__ __forceinline__ double abc_fma(double a, double b, double c) {
#ifdef __CUDA_ARCH__
    return __fma_rn(a, b, c);
#else
    return __builtin_fma(a, b, c);
#endif
}

__host__ __device__ __forceinline__ double abc_exp10_det(double x) {
    const double L2_10_HI = 3.321928094887362181708567732130177319049835205078125;   // RN(log2(10))
    const double L2_10_LO = 1.66146163114990303432e-16;                                // log2(10) - HI
    const double LN2 = 0.6931471805599453094172321214581765680755001343602552;
    if (!(x > -300.0)) return (x != x) ? x : 0.0;
    if (x > 300.0) return 1.0 / 0.0;
    double t = x * L2_10_HI;
    double n = rint(t);
    double r = abc_fma(x, L2_10_HI, -n);      // exact product minus n
    r = abc_fma(x, L2_10_LO, r);
    double z = r * LN2;
    double p = 1.0 / 6227020800.0;             // 1/13!
    p = abc_fma(p, z, 1.0 / 479001600.0);
    p = abc_fma(p, z, 1.0 / 39916800.0);
    p = abc_fma(p, z, 1.0 / 3628800.0);
    p = abc_fma(p, z, 1.0 / 362880.0);
    p = abc_fma(p, z, 1.0 / 40320.0);
    p = abc_fma(p, z, 1.0 / 5040.0);
    p = abc_fma(p, z, 1.0 / 720.0);
    p = abc_fma(p, z, 1.0 / 120.0);
    p = abc_fma(p, z, 1.0 / 24.0);
    p = abc_fma(p, z, 1.0 / 6.0);
    p = abc_fma(p, z, 0.5);
    p = abc_fma(p, z, 1.0);
    p = abc_fma(p, z, 1.0);
    // scale by 2^n exactly (n in [-1000, 1000])
    int ni = (int)n;
    union { uint64_t u; double d; } s;
    s.u = (uint64_t)(ni + 1023) << 52;
    return p * s.d;
}

#define ABC_CUDA_CHECK(expr)                                                            \
    do {                                                                                \
        cudaError_t _e = (expr);                                                        \
        if (_e != cudaSuccess) {                                                        \
            abc_set_error("CUDA error %s at %s:%d: %s", #expr, __FILE__, __LINE__,      \
                          cudaGetErrorString(_e));                                      \
            return ABC_ERR_CUDA;                                                        \
        }                                                                               \
    } while (0)
