// abc_ctx.h -- the context behind abc_ctx_t and the helpers shared by abc_capi.cu and abc_multi.cu (not installed)
#pragma once
#include <cstdint>
#include <utility>
#include <vector>

#include "abc_common.cuh"
#include "abc_internal.h"

// owning device buffer: freed with its owner (a context member or a local of an entry point, whatever the exit path)
template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    DevBuf(DevBuf&& o) noexcept : p(o.p), cap(o.cap) { o.p = nullptr; o.cap = 0; }
    DevBuf& operator=(DevBuf&& o) noexcept {
        if (this != &o) { release(); p = o.p; cap = o.cap; o.p = nullptr; o.cap = 0; }
        return *this;
    }
    ~DevBuf() { release(); }
    // grows geometrically (a work buffer whose demand creeps up call by call -- accepted tuples, sort keys -- must not pay a
    // cudaFree + cudaMalloc, i.e. a device synchronisation, every time)
    int ensure(size_t n) {
        if (n <= cap) return ABC_OK;
        if (cap > 0 && n < cap + cap / 4) n = cap + cap / 4;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc((void**)&p, n * sizeof(T));
        if (e != cudaSuccess) {
            abc_set_error("cudaMalloc(%zu bytes) failed: %s", n * sizeof(T), cudaGetErrorString(e));
            cudaGetLastError();
            return ABC_ERR_NOMEM;
        }
        cap = n;
        return ABC_OK;
    }
    void release() { if (p) { cudaFree(p); p = nullptr; } cap = 0; }
};

struct abc_ctx {
    int device = 0, sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    // design
    bool has_design = false;
    abc_design_t design;
    std::vector<uint32_t> beta_q32;
    int32_t beta_off[11];
    double beta_mean[10], beta_m2[10], beta_var[10];
    DevBuf<uint32_t> d_beta;
    DevBuf<double> d_age_dist;
    DevBuf<double> d_beta_mom;   // [30]: mean, m2, var for the 10 groups
    // data statistics
    bool has_data = false;
    int32_t G = 0;
    DevBuf<double> d_d, d_den;
    DevBuf<float2> d_fbw, d_fa;
    DevBuf<float> d_fstats;
    DevBuf<unsigned char> d_rnan;
    int force_reference_score = 0;
    int score_tile_kernel = 1;   // 1: tile-pruned scoring (abc_score3.cu); 0: three-stage kernel of abc_score.cu
    // tile-pruned scoring tables (per data set) and work buffers
    int32_t s3_ntiles = 0;
    DevBuf<float4> d_s3_tb, d_s3_ab;
    DevBuf<uint32_t> d_s3_wt;
    DevBuf<int32_t> d_s3_gidx;
    DevBuf<uint32_t> d_s3_ok;
    // two lanes of work buffers: sub-batches alternate between two internal streams so that the filter kernel of one
    // (FP32 pipe + bulk stores) overlaps the stage-3 kernel of the other (FP64 pipe + shared memory)
    DevBuf<uint32_t> d_s3_live[2], d_s3_nanw[2], d_s3_qcnt[2];
    DevBuf<uint16_t> d_s3_q2[2];
    DevBuf<float> d_s3_fstats[2];
    DevBuf<float> d_mf_b, d_mf_a[2];   // tensor-core filter: gene operand (per data set), particle operand (per sub-batch)
    DevBuf<uint32_t> d_mf_mask[2];     // its output: sign-bit words [tile of 32 genes][particle]
    DevBuf<uint32_t> d_mf_done;        // per particle block: CTAs that have written their slice of the background
    int score_mma_filter = 0;    // 1: TF32 tcgen05 GEMM decides which pairs reach stage 3 (abc_score3.cu, tensor-core filter)
    double mf_max_slack = 0.0;
    cudaStream_t s3_stream[2] = {nullptr, nullptr};
    cudaEvent_t s3_ev_begin = nullptr, s3_ev_end[2] = {nullptr, nullptr};
    int score_overlap = 1;
    int score_sub_batches = 0;   // sub-batches per call when overlapping; 0 = 2 below 256k particles, else 4
    int stats_guards = -1;       // -1: sample guards iff sim_kind == SSA; 0 / 1 force
    int ssa_hybrid = 2;          // exact telegraph + conditional-Poisson sampling: 1 = burn-in only, 2 = to the read-out
    int ssa_adaptive = 2;        // burn-in from the decay of the discarded history: 1 = whole cycles per particle (modes 1, 2),
                                 // 2 = start time per (particle, read-out) from the exact mean contributions (mode 2; mode 1 uses 1)
    int64_t simscore_sub_min = 8192;   // abc_simulate_score: smallest sub-batch worth pipelining
    // simulate work buffers
    DevBuf<double> d_theta, d_stats, d_moments, d_ss_iv, d_prefix;
    DevBuf<AbcRates> d_rates;
    DevBuf<float> d_win;         // mode 2: start time of every (particle, read-out)
    DevBuf<unsigned long long> d_sums, d_counters;
    DevBuf<unsigned int> d_work;
    DevBuf<uint32_t> d_cells;
    DevBuf<unsigned int> d_keys_in, d_keys_out;
    DevBuf<int> d_idx_in, d_order;
    DevBuf<unsigned char> d_sort_tmp;
    // score work buffers
    DevBuf<double> d_sstats, d_err;
    // abc_simulate_score: two sets of output buffers, the copy stream that drains them, per-set events and a page-locked
    // landing zone for the per-sub-batch device counters
    DevBuf<double> d_p_theta[2], d_p_stats[2], d_p_err[2];
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t p_done[2] = {nullptr, nullptr}, p_copied[2] = {nullptr, nullptr}, p_t0[2] = {nullptr, nullptr},
                p_t1[2] = {nullptr, nullptr}, p_t2[2] = {nullptr, nullptr};
    unsigned long long* h_p_counters = nullptr;      // [2][8]
    // abc_simulate_score_async: batches in flight per output buffer set (0 = none), the set the next call uses, and the
    // device times accumulated for abc_wait
    int64_t async_nb[2] = {0, 0};
    int async_next = 0;
    double async_ms_sim = 0.0, async_ms_score = 0.0;
    bool async_open = false;
    DevBuf<unsigned long long> d_counts, d_acc_count;
    DevBuf<int32_t> d_acc_gene;
    DevBuf<long long> d_acc_particle;
    DevBuf<double> d_acc_err;
    // abc_accept_fetch: work buffers of the device sort (abc_accept.cu)
    DevBuf<unsigned long long> d_as_k64[2];
    DevBuf<uint32_t> d_as_k32[2], d_as_perm[2];
    DevBuf<long long> d_as_idx;
    DevBuf<double> d_as_err;
    DevBuf<unsigned char> d_as_tmp;
    int64_t acc_capacity = 0, acc_budget = 0, acc_min_capacity = 0;
    int64_t launches = 0;
    abc_counters_t last;
    // the *_dev entry points enqueue on the caller's stream: an event recorded there after every such call orders the
    // accept_* / posterior entry points (which read d_acc_count / d_counts on the host) behind that work
    cudaEvent_t ev_user = nullptr;
    bool user_pending = false;
    // multi-GPU: communicator and exchange buffers (abc_multi.cu, AbcComm); nullptr until abc_comm_init_rank / abc_multi_create
    void* comm = nullptr;
};
void abc_comm_free(abc_ctx* c);

// all work this context has enqueued -- on its own stream and on caller streams of the *_dev entry points -- is complete
inline int sync_ctx(abc_ctx* c) {
    ABC_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    if (c->user_pending) {
        ABC_CUDA_CHECK(cudaEventSynchronize(c->ev_user));
        c->user_pending = false;
    }
    return ABC_OK;
}
inline int mark_user_stream(abc_ctx* c, cudaStream_t st) {
    ABC_CUDA_CHECK(cudaEventRecord(c->ev_user, st));
    c->user_pending = true;
    return ABC_OK;
}

#define CTX_GUARD(ctx)                                                       \
    if (!(ctx)) { abc_set_error("null context"); return ABC_ERR_ARG; }       \
    ABC_CUDA_CHECK(cudaSetDevice((ctx)->device))


// abc_capi.cu internals used by the multi-GPU layer
int abc_build_accepted_lists(abc_ctx* c, int64_t* offsets, unsigned long long* total_out, bool want_lists);
int abc_simulate_score_impl(abc_ctx* c, int m, int64_t n, int64_t offset, uint64_t seed, int prior_supplied, double* theta,
                            double* stats, double eps, int layout, double* err, int64_t gm_pitch, int64_t* counts,
                            abc_counters_t* counters);
