// abc_internal.h -- launcher declarations shared by the .cu files of libabcb200.so (not installed)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/abc_b200.h"

void abc_set_error(const char* fmt, ...);

// per-particle linear-scale rates of the CME; 24 words = 96 B
struct AbcRates {
    float kon[5], koff[5], alpha[5], gamma[5];
    float lam;          // labelling efficiency 10^theta_lambda clamped to [0,1]
    uint32_t pon_thr;   // floor(P_on * 2^32): initial gene state threshold
    float pad0, pad1;   // pad0 = predicted work (scheduling hint only); pad1 != 0: particle refused (abc_window_kernel)
};

struct AbcSsaParams {
    int64_t  n_particles;      // particles in this launch
    int64_t  particle_offset;  // global index of the first one (Philox key part)
    uint32_t seed_lo, seed_hi;
    int32_t  m;                // model 1..5
    int32_t  scaling;          // m != 2
    int32_t  n_cells;
    int32_t  chunks;           // ceil(n_cells / 32)
    int32_t  n_pre;            // complete cycles before the read-out cycle
    int32_t  downsampling;
    int32_t  single_readout;   // >= 0: debug mode, simulate only this read-out of particle 0
    int32_t  hybrid;           // 1: exact telegraph + Poisson burn-in before the label window; 2: to the read-out
                               // (fast-math kernel only)
    int32_t  adaptive;         // 1: per-particle burn-in in whole cycles with the truncation bias bound of n_pre cycles (modes 1, 2)
                               // 2: mode 2 only, start time per (particle, read-out) from the exact mean contributions
    int32_t  pad_;
    double   cycle;
    double   agevec[5];
    double   pulse[11];
    double   chase[11];
    int32_t  beta_off[11];     // offsets into beta_q32: pulse clusters 1..5, chase clusters 1..5
};

// device buffers produced by the simulate stage
//   sums:     [particle][55][5] uint64  (sum u, sum l, sum u^2, sum u*l, sum l^2 over the cells)
//   counters: [4] uint64 (lineages, events, draws, spare)
int abc_launch_rates(const double* d_theta, int m, int64_t n, AbcRates* d_rates, int ssa_hybrid, int n_pre_adaptive,
                     double cycle, cudaStream_t st);
int abc_launch_prior(double* d_theta, int m, int64_t n, int64_t offset, uint64_t seed, cudaStream_t st);
int abc_launch_ssa(const AbcRates* d_rates, const AbcSsaParams& prm, const uint32_t* d_beta_q32,
                   unsigned long long* d_sums, unsigned long long* d_counters, unsigned int* d_work,
                   uint32_t* d_cells_out, const int* d_order, int exact_math, int sm_count, cudaStream_t st);
// mode 2 (abc_tele.cu): start time of every (particle, read-out) + scheduling cost, then the telegraph kernel
int abc_launch_window(AbcRates* d_rates, const AbcSsaParams& prm, float* d_win, int64_t n, cudaStream_t st);
int abc_launch_tele(const AbcRates* d_rates, const AbcSsaParams& prm, const float* d_win, const uint32_t* d_beta_q32,
                    unsigned long long* d_sums, unsigned long long* d_counters, unsigned int* d_work,
                    uint32_t* d_cells_out, const int* d_order, int sm_count, cudaStream_t st);
size_t abc_order_temp_bytes(int n);
int abc_launch_order(const AbcRates* d_rates, int n, unsigned int* d_keys_in, unsigned int* d_keys_out, int* d_idx_in,
                     int* d_order, void* d_temp, size_t temp_bytes, cudaStream_t st);
int abc_launch_moments_from_sums(const unsigned long long* d_sums, int64_t n, int n_cells, double* d_moments,
                                 cudaStream_t st);
int abc_launch_summary_stats(const double* d_moments, const double* d_age_dist, int64_t n, double* d_stats,
                             int sample_guards, cudaStream_t st);

// moment-ODE simulator (abc_ode.cu); counters[4] accumulates accepted integrator steps
int abc_launch_ode(const double* d_theta, const abc_design_t& des, int m, int64_t n, const double* d_beta_mom,
                   double* d_ss_iv, double* d_prefix, double* d_moments, unsigned long long* d_counters, cudaStream_t st);

// scoring
int abc_launch_prepare_data(const double* d_d, const double* d_se, int G, double* d_den, float2* d_fbw, float2* d_fa,
                            cudaStream_t st);
struct AbcScoreArgs {
    const double* stats;   // [n][53]
    const double* d;       // [G][53]
    const double* den;     // [G][53]
    const float2* fbw;     // [G][53] FP32 pre-filter constants (2 w d, w), w = 1/(53 den)
    const float2* fa;      // [G] (sum_{t<15} w d^2, sum_{t>=15} w d^2)
    const float* fstats;   // [n][53] FP32 copy of stats (abc_launch_score_prep)
    const unsigned char* rnan;   // [n] 1 if the particle has a NaN statistic
    int32_t force_reference_kernel;   // 1: always use the plain FP64 kernel (abc_score_kernel)
    int64_t n;
    int32_t G;
    int64_t particle_offset;
    double  eps;
    int32_t err_layout;
    int64_t gm_stride;     // gene-major layout: doubles between consecutive gene rows (the whole batch's n)
    double* err;           // nullable
    unsigned long long* counts;       // [G]
    unsigned long long* acc_count;    // [1] running number of accepted tuples
    int64_t acc_capacity;
    int32_t* acc_gene; long long* acc_particle; double* acc_err;
};
int abc_launch_score(const AbcScoreArgs& a, int sm_count, cudaStream_t st);

// tile-pruned scoring (abc_score3.cu): per-data-set tables built on the host by abc_score3_build
struct AbcScore3Tables {
    int32_t ntiles;
    const float4* tb;        // [ntiles][31]: (amin, -bmax, bmin, -amax) of a = sqrt(w) d, b = sqrt(w), w = 1/(53 den)
    const float4* ab;        // [ntiles][27][32]: (a_2j, a_2j+1, -b_2j, -b_2j+1) per (term pair, gene slot)
    const uint32_t* wt;      // [ntiles][6][53][32]: d, den, RN(1/den) as (high, low) words in tile order
    const int32_t* gidx;     // [ntiles][32] original gene index of a slot, -1 = padding
    const uint32_t* okmask;  // [ntiles] bit l: slot l may divide through the stored reciprocal
    uint32_t* live;          // [ntiles][W] work: bit = (tile, particle) needs stages 1-3
    uint32_t* nanw;          // [W] work: bit = particle has a NaN statistic
    uint16_t* q2;            // [blocks*ntiles][2048*32] work: pairs queued for stage 3 (particle << 5 | gene slot)
    uint32_t* qcnt;          // [blocks*ntiles] work: fill of each segment
    int64_t W;               // ceil(n / 32)
    float sure;              // FP32 bounds above this prove that a pair is not needed (set by the launcher)
    uint32_t* gmask;         // [8 * mma tiles][n_pad] work (tensor-core filter): bit l = (particle, slot l of the tile) reaches stage 3
    int64_t n_pad;           // its row pitch: particles of the launch rounded up to 128
    int64_t n_rows;          // rows readable from gmask (stage 3 may be launched on a row range)
    uint32_t* fill_done;     // [blocks + 1] work: CTAs that have written their slice of a particle block's background
    int32_t fill_b0, fill_d; // background: blocks < fill_b0 by the filter kernel, block j >= fill_b0 by stage 3 of block j - fill_d
};
#ifdef __cplusplus
#include <vector>
struct AbcScore3Host {
    int ntiles = 0;
    std::vector<float> tb, ab;
    std::vector<uint32_t> wt;
    std::vector<int32_t> gidx;
    std::vector<uint32_t> okmask;
};
void abc_score3_build(const double* d, const double* den, int G, AbcScore3Host& out);
#endif
size_t abc_score3_blocks(int64_t n);
size_t abc_score3_queue_entries(int64_t n, int ntiles);
int abc_launch_score3(const AbcScoreArgs& a, const AbcScore3Tables& x, cudaStream_t st);
// tensor-core filter (TF32 tcgen05 GEMM) in front of the same stage 3: abc_score3.cu
#ifdef __cplusplus
void abc_score_mma_build(const double* d, const double* den, const AbcScore3Host& h, std::vector<float>& bblob, double* max_slack);
#endif
int abc_score_mma_tiles(int ntiles);
int abc_launch_score_mma(const AbcScoreArgs& a, const AbcScore3Tables& x, float* d_ablob, const float* d_bblob, float* d_dbg,
                         int sm_count, cudaStream_t st);
int abc_launch_score_prep(const double* d_stats, int64_t n, float* d_fstats, unsigned char* d_rnan, cudaStream_t st);

// A1 ordering on the device (abc_accept.cu)
size_t abc_accept_sort_temp_bytes(size_t total);
int abc_launch_accept_sort(const int32_t* d_gene, const long long* d_particle, const double* d_err, size_t total, int G,
                           unsigned long long* d_key64[2], uint32_t* d_key32[2], uint32_t* d_perm[2], void* d_temp,
                           size_t temp_bytes, long long* d_out_idx, double* d_out_err, int* n_launches, cudaStream_t st);

// SURVEY 8f-3 posterior summaries over the ordered lists (abc_accept.cu)
size_t abc_posterior_temp_bytes(size_t total, int G);
int abc_launch_posterior(const long long* d_idx, const long long* d_offsets, size_t total, int G, const double* d_theta,
                         long long n, int P, long long particle_offset, double q, double* d_vals[2], void* d_temp,
                         size_t temp_bytes, int* d_bad, double* d_map, double* d_mean, double* d_lo, double* d_hi,
                         int* n_launches, cudaStream_t st);

// SURVEY 8f-2 model-probability bootstrap (abc_accept.cu)
int abc_launch_model_probs(const long long* d_counts, int K, int G, int B, double alpha, uint64_t seed, double* d_stats,
                           double* d_prob, double* d_lb, double* d_ub, int* n_launches, cudaStream_t st);

// data-side summary statistics with bootstrap standard errors (abc_datastats.cu)
int abc_run_data_summary_stats(const double* u, const double* l, int n_cells, int G, const int32_t* age, const int32_t* experiment,
                               const int32_t* cond_vec, const int32_t* pulse_idx, int n_pulse, const int32_t* chase_idx, int n_chase,
                               const double* age_id_dist, int B, uint64_t seed, double* d_out, double* se_out, int64_t* launches,
                               cudaStream_t st);
