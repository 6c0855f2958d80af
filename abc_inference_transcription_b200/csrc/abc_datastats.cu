// abc_datastats.cu -- the data side of the scoring kernel's inputs (second half): the 53 summary statistics
// of every gene from the per-cell UMI counts and their bootstrap standard errors, scripts/data_summary_statistics.jl:
//   get_ccp_data / get_ccp_se          :2-58     per age cluster mean and Fano factor of u + l over the pulse (chase) cells
//   get_ratios / get_ratio_se          :61-97    per condition mean(l) / (mean(u) + mean(l))
//   get_correlations / get_correlation_se :100-177  per condition mean_corr, corr_mean (weighted_cov :179-181)
//   get_summary_stats                  :183-194  the 14 vectors per gene = d[53], se[53] of abc_set_data
// The reference resamples the cells with replacement 100 times per gene and statistic family; here the four families
// (pulse cells, chase cells, all cells for the ratios, all cells for the correlations) draw their resamples once from
// Philox -- as per-cell multiplicities -- and every gene is evaluated under the same resamples (the marginal law of every
// gene's bootstrap is the reference's).  Counts are integers: all sums are exact, the statistics are FP64 in the reference's
// operation order, so a numpy restatement on the same Philox draws agrees bit for bit.  sm_100a.
#include "abc_common.cuh"
#include "abc_internal.h"

#include <algorithm>
#include <cstring>
#include <vector>

#define ABC_DOM_DATA 3u
#define DS_WARPS 8
#define DS_MAXB 128

struct DsFamily {          // one resampling population, its cells sorted by bin
    int n_pop;             // cells in the population
    int n_bins;
    int off[56];           // bin b = sorted positions off[b] .. off[b+1]; cells outside every bin follow off[n_bins]
};
struct DsArgs {
    DsFamily fam[4];       // 0 pulse (5 age clusters), 1 chase (5), 2 ratios (11 conditions), 3 correlations (11 x 5)
    const int* cell[4];    // [n_pop] cell index at a sorted position
    const unsigned short* mult[4];   // [B][n_pop] multiplicity of the cell at a sorted position in bootstrap b
    const unsigned int* u;           // [G][n_cells]
    const unsigned int* l;
    const double* age_dist;          // 5 x 11 column-major, used as given for the point estimates
    int n_cells, G, B;
    double* d;             // [G][53]
    double* se;            // [G][53]
};

// multiplicities of one resample: n_pop draws with replacement, index = floor(u64 * n_pop / 2^64)
__global__ void ds_mult_kernel(int family, int n_pop, const int* __restrict__ sorted_pos, uint32_t k0, uint32_t k1,
                               unsigned short* __restrict__ mult) {
    extern __shared__ unsigned int cnt[];
    const int b = blockIdx.x;
    for (int i = threadIdx.x; i < n_pop; i += blockDim.x) cnt[i] = 0u;
    __syncthreads();
    const int nblk = (n_pop + 1) >> 1;
    for (int j = threadIdx.x; j < nblk; j += blockDim.x) {
        const uint4 r = philox4x32_10((uint32_t)j, (uint32_t)b, (uint32_t)family, ABC_DOM_DATA << 29, k0, k1);
        const unsigned long long u0 = ((unsigned long long)r.x << 32) | r.y, u1 = ((unsigned long long)r.z << 32) | r.w;
        atomicAdd(&cnt[sorted_pos[(int)__umul64hi(u0, (unsigned long long)n_pop)]], 1u);
        if (2 * j + 1 < n_pop) atomicAdd(&cnt[sorted_pos[(int)__umul64hi(u1, (unsigned long long)n_pop)]], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n_pop; i += blockDim.x) mult[(size_t)b * n_pop + i] = (unsigned short)cnt[i];
}

struct BinSums { unsigned long long n, su, sl, suu, sll, sul; };

__device__ __forceinline__ unsigned long long ds_wsum(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// sums over one bin; mult == nullptr: every cell once (the point estimate)
__device__ __forceinline__ BinSums ds_bin(const DsFamily& f, int bin, const int* __restrict__ cell,
                                          const unsigned short* __restrict__ mult, const unsigned int* su_, const unsigned int* sl_, int lane) {
    BinSums s = {0, 0, 0, 0, 0, 0};
    for (int i = f.off[bin] + lane; i < f.off[bin + 1]; i += 32) {
        const unsigned long long m = mult ? (unsigned long long)mult[i] : 1ull;
        const int c = cell[i];
        const unsigned long long u = su_[c], l = sl_[c];
        s.n += m; s.su += m * u; s.sl += m * l; s.suu += m * u * u; s.sll += m * l * l; s.sul += m * u * l;
    }
    s.n = ds_wsum(s.n); s.su = ds_wsum(s.su); s.sl = ds_wsum(s.sl);
    s.suu = ds_wsum(s.suu); s.sll = ds_wsum(s.sll); s.sul = ds_wsum(s.sul);
    return s;
}

__device__ __forceinline__ double ds_u128(unsigned __int128 v) {
    return __dadd_rn(__dmul_rn((double)(unsigned long long)(v >> 64), 18446744073709551616.0), (double)(unsigned long long)v);
}
// mean, corrected variance / covariance from exact integer sums (var of a single value is NaN like Julia's var)
__device__ __forceinline__ double ds_mean(unsigned long long s, unsigned long long n) { return __ddiv_rn((double)s, (double)n); }
__device__ __forceinline__ double ds_cov(unsigned long long n, unsigned long long sxy, unsigned long long sx, unsigned long long sy) {
    if (n < 2) return __longlong_as_double(0x7ff8000000000000ll);
    const unsigned __int128 a = (unsigned __int128)n * sxy, b = (unsigned __int128)sx * sy;
    const double den = __dmul_rn((double)n, (double)(n - 1));
    return (a >= b) ? __ddiv_rn(ds_u128(a - b), den) : -__ddiv_rn(ds_u128(b - a), den);
}

__device__ __forceinline__ double ds_add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double ds_mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double ds_wsum5(const double* w, const double* x) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 5; ++i) s = ds_add(s, ds_mul(w[i], x[i]));
    return s;
}
__device__ __forceinline__ double ds_wcov5(const double* x, const double* y, const double* w) {      // weighted_cov, :179-181
    const double mx = ds_wsum5(w, x), my = ds_wsum5(w, y);
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 5; ++i) s = ds_add(s, ds_mul(w[i], ds_mul(ds_add(x[i], -mx), ds_add(y[i], -my))));
    return s;
}

// the 53 statistics of one gene under one resample (mult pointers null: the data themselves); out[53]
__device__ void ds_stats(const DsArgs& a, const unsigned short* const mult[4], const unsigned int* su_, const unsigned int* sl_,
                         int lane, double* out) {
    // pulse / chase: mean and Fano factor of u + l per age cluster (:2-38); a cluster without cells gives 0
    for (int f = 0; f < 2; ++f)
        for (int c = 0; c < 5; ++c) {
            const BinSums s = ds_bin(a.fam[f], c, a.cell[f], mult[f], su_, sl_, lane);
            const unsigned long long st = s.su + s.sl, stt = s.suu + 2ull * s.sul + s.sll;
            double mean = 0.0, ff = 0.0;
            if (s.n > 0) {
                mean = ds_mean(st, s.n);
                const double eps = (mean == 0.0) ? 0.0001 : 0.0;
                ff = __ddiv_rn(ds_cov(s.n, stt, st, st), ds_add(mean, eps));
            }
            if (lane == 0) { out[10 * f + c] = mean; out[10 * f + 5 + c] = ff; }
        }
    // ratios per condition (:61-79)
    for (int j = 0; j < 11; ++j) {
        const BinSums s = ds_bin(a.fam[2], j, a.cell[2], mult[2], su_, sl_, lane);
        double r = 0.0;
        if (s.n > 0) {
            const double mu = ds_mean(s.su, s.n), ml = ds_mean(s.sl, s.n);
            if (ds_add(mu, ml) > 0.0) r = __ddiv_rn(ml, ds_add(mu, ml));
        }
        if (lane == 0) out[20 + j] = r;
    }
    // correlations per condition (:100-147); weights: age_id_dist as given for the data, the resample's own per-condition
    // age distribution for a bootstrap (:160-164)
    for (int j = 0; j < 11; ++j) {
        double m1[5], m2[5], v1[5], v2[5], c12[5], w[5];
        unsigned long long ntot = 0, nc[5];
        for (int c = 0; c < 5; ++c) {
            const BinSums s = ds_bin(a.fam[3], j * 5 + c, a.cell[3], mult[3], su_, sl_, lane);
            nc[c] = s.n; ntot += s.n;
            if (s.n > 0) {
                m1[c] = ds_mean(s.su, s.n); m2[c] = ds_mean(s.sl, s.n);
                v1[c] = ds_cov(s.n, s.suu, s.su, s.su); v2[c] = ds_cov(s.n, s.sll, s.sl, s.sl);
                c12[c] = ds_cov(s.n, s.sul, s.su, s.sl);
            } else {
                m1[c] = m2[c] = v1[c] = v2[c] = c12[c] = 0.0;
            }
        }
        for (int c = 0; c < 5; ++c)
            w[c] = mult[3] ? (ntot > 0 ? __ddiv_rn((double)nc[c], (double)ntot) : 0.0) : a.age_dist[j * 5 + c];
        const double tv1 = ds_add(ds_wsum5(w, v1), ds_wcov5(m1, m1, w)), tv2 = ds_add(ds_wsum5(w, v2), ds_wcov5(m2, m2, w));
        bool cov_zero = true;
        for (int c = 0; c < 5; ++c) cov_zero = cov_zero && (c12[c] == 0.0);
        const double sd = __dsqrt_rn(fabs(ds_mul(tv1, tv2)));
        double mc = 0.0, cm = 0.0;
        if (!cov_zero && tv1 != 0.0 && tv2 != 0.0) mc = __ddiv_rn(ds_wsum5(w, c12), sd);
        if (tv1 != 0.0 && tv2 != 0.0) cm = __ddiv_rn(ds_wcov5(m1, m2, w), sd);
        if (lane == 0) { out[31 + j] = mc; out[42 + j] = cm; }
    }
}

// one CTA per gene: the gene's counts in shared memory, warp w takes the data (b = -1) and the bootstraps b = w, w + 8, ...
__global__ void __launch_bounds__(DS_WARPS * 32)
ds_stats_kernel(const DsArgs a) {
    extern __shared__ __align__(16) unsigned char ds_raw[];
    unsigned int* su_ = reinterpret_cast<unsigned int*>(ds_raw);
    unsigned int* sl_ = su_ + a.n_cells;
    double* boot = reinterpret_cast<double*>(sl_ + a.n_cells + (a.n_cells & 1));       // [B][53]
    const int g = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < a.n_cells; i += blockDim.x) {
        su_[i] = a.u[(size_t)g * a.n_cells + i];
        sl_[i] = a.l[(size_t)g * a.n_cells + i];
    }
    __syncthreads();
    for (int b = warp - 1; b < a.B; b += DS_WARPS) {
        if (b < 0) {
            const unsigned short* none[4] = {nullptr, nullptr, nullptr, nullptr};
            ds_stats(a, none, su_, sl_, lane, a.d + (size_t)g * ABC_NSTATS);
        } else {
            const unsigned short* m[4];
            for (int f = 0; f < 4; ++f) m[f] = a.mult[f] + (size_t)b * a.fam[f].n_pop;
            ds_stats(a, m, su_, sl_, lane, boot + (size_t)b * ABC_NSTATS);
        }
    }
    __syncthreads();
    // standard error over the bootstraps (:50-56): sqrt(1/(B-1) sum (x - mean(x))^2), sums in bootstrap order
    for (int t = threadIdx.x; t < ABC_NSTATS; t += blockDim.x) {
        double s = 0.0;
        for (int b = 0; b < a.B; ++b) s = ds_add(s, boot[(size_t)b * ABC_NSTATS + t]);
        const double mean = __ddiv_rn(s, (double)a.B);
        double q = 0.0;
        for (int b = 0; b < a.B; ++b) { const double dlt = ds_add(boot[(size_t)b * ABC_NSTATS + t], -mean); q = ds_add(q, ds_mul(dlt, dlt)); }
        a.se[(size_t)g * ABC_NSTATS + t] = __dsqrt_rn(ds_mul(__ddiv_rn(1.0, (double)(a.B - 1)), q));
    }
}

// host side: sort every family's population by bin, upload, draw the multiplicities, run
int abc_run_data_summary_stats(const double* u, const double* l, int n_cells, int G, const int32_t* age, const int32_t* experiment,
                               const int32_t* cond_vec, const int32_t* pulse_idx, int n_pulse, const int32_t* chase_idx, int n_chase,
                               const double* age_id_dist, int B, uint64_t seed, double* d_out, double* se_out, int64_t* launches,
                               cudaStream_t st) {
    if (!u || !l || !age || !experiment || !cond_vec || !pulse_idx || !chase_idx || !age_id_dist || !d_out || !se_out ||
        n_cells < 2 || n_cells > 60000 || G < 1 || B < 2 || B > DS_MAXB || n_pulse < 1 || n_chase < 1) {
        abc_set_error("abc_data_summary_stats: bad arguments (2 <= n_cells <= 60000, 2 <= n_bootstraps <= %d)", DS_MAXB);
        return ABC_ERR_ARG;
    }
    // ---- populations and bins
    std::vector<int> pop[4], bin[4];
    for (int f = 0; f < 2; ++f) {
        const int32_t* idx = f ? chase_idx : pulse_idx;
        const int n = f ? n_chase : n_pulse;
        for (int i = 0; i < n; ++i) {
            const int c = idx[i] - 1;
            if (c < 0 || c >= n_cells) { abc_set_error("abc_data_summary_stats: cell index %d out of range (1-based)", idx[i]); return ABC_ERR_ARG; }
            const int a = age[c];
            pop[f].push_back(c); bin[f].push_back((a >= 1 && a <= 5) ? a - 1 : -1);
        }
    }
    for (int c = 0; c < n_cells; ++c) {
        int j = -1;
        for (int q = 0; q < ABC_NCOND; ++q) if (experiment[c] == cond_vec[q]) { j = q; break; }
        const int a = age[c];
        pop[2].push_back(c); bin[2].push_back(j);
        pop[3].push_back(c); bin[3].push_back((j >= 0 && a >= 1 && a <= 5) ? j * 5 + (a - 1) : -1);
    }
    const int nbins[4] = {5, 5, 11, 55};
    DsArgs a;
    memset(&a, 0, sizeof(a));
    std::vector<int> cell_sorted[4], pos_of[4];
    for (int f = 0; f < 4; ++f) {
        const int n = (int)pop[f].size();
        std::vector<int> order((size_t)n);
        for (int i = 0; i < n; ++i) order[i] = i;
        auto key = [&](int i) { return bin[f][i] < 0 ? nbins[f] : bin[f][i]; };
        std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return key(x) < key(y); });
        cell_sorted[f].resize(n); pos_of[f].resize(n);
        a.fam[f].n_pop = n; a.fam[f].n_bins = nbins[f];
        for (int b = 0; b <= nbins[f]; ++b) a.fam[f].off[b] = 0;
        for (int k = 0; k < n; ++k) {
            cell_sorted[f][k] = pop[f][order[k]];
            pos_of[f][order[k]] = k;
            const int b = key(order[k]);
            if (b < nbins[f]) a.fam[f].off[b + 1] = k + 1;
        }
        for (int b = 1; b <= nbins[f]; ++b) a.fam[f].off[b] = std::max(a.fam[f].off[b], a.fam[f].off[b - 1]);
    }
    // ---- counts as integers
    std::vector<unsigned int> hu((size_t)G * n_cells), hl((size_t)G * n_cells);
    for (size_t i = 0; i < hu.size(); ++i) {
        const double x = u[i], y = l[i];
        if (!(x >= 0.0 && x <= 1.0e6 && x == (double)(unsigned int)x) || !(y >= 0.0 && y <= 1.0e6 && y == (double)(unsigned int)y)) {
            abc_set_error("abc_data_summary_stats: counts must be integers in [0, 1e6]");
            return ABC_ERR_ARG;
        }
        hu[i] = (unsigned int)x; hl[i] = (unsigned int)y;
    }
    // ---- device buffers (freed on every path)
    struct Bufs {
        void* p[16]; int n = 0;
        ~Bufs() { for (int i = 0; i < n; ++i) cudaFree(p[i]); }
        void* get(size_t bytes) { void* q = nullptr; if (cudaMalloc(&q, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); return nullptr; } p[n++] = q; return q; }
    } bufs;
    unsigned int* d_u = (unsigned int*)bufs.get(hu.size() * 4);
    unsigned int* d_l = (unsigned int*)bufs.get(hl.size() * 4);
    double* d_age = (double*)bufs.get(55 * 8);
    double* d_d = (double*)bufs.get((size_t)G * ABC_NSTATS * 8);
    double* d_se = (double*)bufs.get((size_t)G * ABC_NSTATS * 8);
    int* d_cell[4]; int* d_pos[4]; unsigned short* d_mult[4];
    bool ok = d_u && d_l && d_age && d_d && d_se;
    for (int f = 0; f < 4 && ok; ++f) {
        d_cell[f] = (int*)bufs.get(cell_sorted[f].size() * 4);
        d_pos[f] = (int*)bufs.get(pos_of[f].size() * 4);
        d_mult[f] = (unsigned short*)bufs.get((size_t)B * cell_sorted[f].size() * 2);
        ok = d_cell[f] && d_pos[f] && d_mult[f];
    }
    if (!ok) { abc_set_error("abc_data_summary_stats: out of device memory"); return ABC_ERR_NOMEM; }
    ABC_CUDA_CHECK(cudaMemcpyAsync(d_u, hu.data(), hu.size() * 4, cudaMemcpyHostToDevice, st));
    ABC_CUDA_CHECK(cudaMemcpyAsync(d_l, hl.data(), hl.size() * 4, cudaMemcpyHostToDevice, st));
    ABC_CUDA_CHECK(cudaMemcpyAsync(d_age, age_id_dist, 55 * 8, cudaMemcpyHostToDevice, st));
    for (int f = 0; f < 4; ++f) {
        ABC_CUDA_CHECK(cudaMemcpyAsync(d_cell[f], cell_sorted[f].data(), cell_sorted[f].size() * 4, cudaMemcpyHostToDevice, st));
        ABC_CUDA_CHECK(cudaMemcpyAsync(d_pos[f], pos_of[f].data(), pos_of[f].size() * 4, cudaMemcpyHostToDevice, st));
        ds_mult_kernel<<<B, 256, (size_t)a.fam[f].n_pop * 4, st>>>(f, a.fam[f].n_pop, d_pos[f], (uint32_t)seed, (uint32_t)(seed >> 32), d_mult[f]);
        a.cell[f] = d_cell[f]; a.mult[f] = d_mult[f];
    }
    ABC_CUDA_CHECK(cudaGetLastError());
    a.u = d_u; a.l = d_l; a.age_dist = d_age; a.n_cells = n_cells; a.G = G; a.B = B; a.d = d_d; a.se = d_se;
    const size_t smem = ((size_t)2 * n_cells + (n_cells & 1)) * 4 + (size_t)B * ABC_NSTATS * 8;
    if (smem > 220 * 1024) { abc_set_error("abc_data_summary_stats: %d cells x %d bootstraps exceed the shared memory of one CTA", n_cells, B); return ABC_ERR_ARG; }
    ABC_CUDA_CHECK(cudaFuncSetAttribute(ds_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ds_stats_kernel<<<G, DS_WARPS * 32, smem, st>>>(a);
    ABC_CUDA_CHECK(cudaGetLastError());
    ABC_CUDA_CHECK(cudaMemcpyAsync(d_out, d_d, (size_t)G * ABC_NSTATS * 8, cudaMemcpyDeviceToHost, st));
    ABC_CUDA_CHECK(cudaMemcpyAsync(se_out, d_se, (size_t)G * ABC_NSTATS * 8, cudaMemcpyDeviceToHost, st));
    ABC_CUDA_CHECK(cudaStreamSynchronize(st));
    if (launches) *launches += 5;
    return ABC_OK;
}
