// abc_capi.cu -- the C ABI of libabcb200.so (include/abc_b200.h): context, buffers, host<->device
// staging and kernel sequencing.  No CPU fallback anywhere: every compute entry point needs a device.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <utility>
#include <vector>

#include "abc_ctx.h"

#define ABC_VERSION 200

static thread_local char g_err[1024] = "";

void abc_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* abc_last_error(void) { return g_err; }
extern "C" int abc_version(void) { return ABC_VERSION; }

extern "C" int abc_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" int abc_n_params(int m) { return (m >= 1 && m <= 2) ? 5 : (m >= 3 && m <= 5) ? 9 : -1; }

extern "C" const char* abc_model_name(int m) {
    static const char* names[5] = {"const", "const_const", "kon", "alpha", "gamma"};
    return (m >= 1 && m <= 5) ? names[m - 1] : nullptr;
}

extern "C" int abc_create(int device, abc_ctx_t** out) {
    if (!out) { abc_set_error("abc_create: out is NULL"); return ABC_ERR_ARG; }
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        abc_set_error("no CUDA device available (%s); libabcb200 has no CPU fallback",
                      e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
        cudaGetLastError();
        return ABC_ERR_CUDA;
    }
    if (device < 0 || device >= n) { abc_set_error("device %d out of range [0,%d)", device, n); return ABC_ERR_ARG; }
    ABC_CUDA_CHECK(cudaSetDevice(device));
    abc_ctx* c = new (std::nothrow) abc_ctx();
    if (!c) { abc_set_error("out of host memory"); return ABC_ERR_NOMEM; }
    c->device = device;
    cudaDeviceProp prop;
    ABC_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    ABC_CUDA_CHECK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    for (int i = 0; i < 5; ++i) ABC_CUDA_CHECK(cudaEventCreate(&c->ev[i]));
    for (int l = 0; l < 2; ++l) {
        ABC_CUDA_CHECK(cudaStreamCreateWithFlags(&c->s3_stream[l], cudaStreamNonBlocking));
        ABC_CUDA_CHECK(cudaEventCreateWithFlags(&c->s3_ev_end[l], cudaEventDisableTiming));
    }
    ABC_CUDA_CHECK(cudaEventCreateWithFlags(&c->s3_ev_begin, cudaEventDisableTiming));
    ABC_CUDA_CHECK(cudaEventCreateWithFlags(&c->ev_user, cudaEventDisableTiming));
    ABC_CUDA_CHECK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    for (int l = 0; l < 2; ++l) {
        ABC_CUDA_CHECK(cudaEventCreateWithFlags(&c->p_done[l], cudaEventDisableTiming));
        ABC_CUDA_CHECK(cudaEventCreateWithFlags(&c->p_copied[l], cudaEventDisableTiming));
        ABC_CUDA_CHECK(cudaEventCreate(&c->p_t0[l]));
        ABC_CUDA_CHECK(cudaEventCreate(&c->p_t1[l]));
        ABC_CUDA_CHECK(cudaEventCreate(&c->p_t2[l]));
    }
    ABC_CUDA_CHECK(cudaHostAlloc((void**)&c->h_p_counters, 16 * sizeof(unsigned long long), cudaHostAllocDefault));
    memset(&c->last, 0, sizeof(c->last));
    memset(&c->design, 0, sizeof(c->design));
    int rc = c->d_counters.ensure(8);
    if (rc == ABC_OK) rc = c->d_work.ensure(1);
    if (rc == ABC_OK) rc = c->d_acc_count.ensure(1);
    if (rc != ABC_OK) { delete c; return rc; }
    ABC_CUDA_CHECK(cudaMemset(c->d_acc_count.p, 0, sizeof(unsigned long long)));
    *out = c;
    return ABC_OK;
}

extern "C" int abc_destroy(abc_ctx_t* c) {
    if (!c) return ABC_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    if (c->user_pending) cudaEventSynchronize(c->ev_user);
    if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);      // asynchronous batches still travelling to the host
    abc_comm_free(c);
    c->d_beta.release(); c->d_age_dist.release(); c->d_beta_mom.release(); c->d_d.release(); c->d_den.release(); c->d_fbw.release(); c->d_fa.release(); c->d_fstats.release(); c->d_rnan.release();
    c->d_s3_tb.release(); c->d_s3_ab.release(); c->d_s3_wt.release();
    c->d_s3_gidx.release(); c->d_s3_ok.release(); for (int l = 0; l < 2; ++l) { c->d_s3_live[l].release(); c->d_s3_nanw[l].release(); c->d_s3_qcnt[l].release(); c->d_s3_q2[l].release(); c->d_s3_fstats[l].release(); }
    c->d_mf_b.release(); c->d_mf_done.release(); for (int l = 0; l < 2; ++l) { c->d_mf_a[l].release(); c->d_mf_mask[l].release(); }
    c->d_theta.release(); c->d_stats.release(); c->d_moments.release(); c->d_ss_iv.release(); c->d_prefix.release(); c->d_rates.release(); c->d_win.release();
    c->d_keys_in.release(); c->d_keys_out.release(); c->d_idx_in.release(); c->d_order.release(); c->d_sort_tmp.release();
    c->d_sums.release(); c->d_counters.release(); c->d_work.release(); c->d_cells.release();
    c->d_sstats.release(); c->d_err.release(); c->d_counts.release(); c->d_acc_count.release();
    c->d_acc_gene.release(); c->d_acc_particle.release(); c->d_acc_err.release();
    for (int l = 0; l < 2; ++l) { c->d_as_k64[l].release(); c->d_as_k32[l].release(); c->d_as_perm[l].release(); }
    c->d_as_idx.release(); c->d_as_err.release(); c->d_as_tmp.release();
    for (int i = 0; i < 5; ++i) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    if (c->stream) cudaStreamDestroy(c->stream);
    for (int l = 0; l < 2; ++l) {
        if (c->s3_stream[l]) cudaStreamDestroy(c->s3_stream[l]);
        if (c->s3_ev_end[l]) cudaEventDestroy(c->s3_ev_end[l]);
    }
    if (c->s3_ev_begin) cudaEventDestroy(c->s3_ev_begin);
    if (c->ev_user) cudaEventDestroy(c->ev_user);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    for (int l = 0; l < 2; ++l) {
        if (c->p_done[l]) cudaEventDestroy(c->p_done[l]);
        if (c->p_copied[l]) cudaEventDestroy(c->p_copied[l]);
        if (c->p_t0[l]) cudaEventDestroy(c->p_t0[l]);
        if (c->p_t1[l]) cudaEventDestroy(c->p_t1[l]);
        if (c->p_t2[l]) cudaEventDestroy(c->p_t2[l]);
        c->d_p_theta[l].release(); c->d_p_stats[l].release(); c->d_p_err[l].release();
    }
    if (c->h_p_counters) cudaFreeHost(c->h_p_counters);
    delete c;
    return ABC_OK;
}

extern "C" int abc_host_alloc(size_t bytes, void** ptr) {
    if (!ptr) { abc_set_error("abc_host_alloc: ptr is NULL"); return ABC_ERR_ARG; }
    *ptr = nullptr;
    if (bytes == 0) return ABC_OK;
    cudaError_t e = cudaHostAlloc(ptr, bytes, cudaHostAllocPortable);
    if (e != cudaSuccess) {
        abc_set_error("cudaHostAlloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
        cudaGetLastError();
        *ptr = nullptr;
        return ABC_ERR_NOMEM;
    }
    return ABC_OK;
}

extern "C" int abc_host_free(void* ptr) {
    if (!ptr) return ABC_OK;
    ABC_CUDA_CHECK(cudaFreeHost(ptr));
    return ABC_OK;
}

// ------------------------------------------------------------------------------------------------
extern "C" int abc_set_design(abc_ctx_t* c, const abc_design_t* d) {
    CTX_GUARD(c);
    if (!d) { abc_set_error("abc_set_design: design is NULL"); return ABC_ERR_ARG; }
    if (!(d->cycle > 0.0)) { abc_set_error("cycle must be > 0"); return ABC_ERR_ARG; }
    for (int a = 0; a < ABC_NAGE; ++a)
        if (!(d->agevec[a] > 0.0 && d->agevec[a] < d->cycle)) {
            abc_set_error("agevec[%d] = %g must lie in (0, cycle)", a, d->agevec[a]);
            return ABC_ERR_ARG;
        }
    for (int j = 0; j < ABC_NCOND; ++j)
        if (!(d->pulse[j] >= 0.0 && d->chase[j] >= 0.0)) { abc_set_error("pulse/chase must be >= 0"); return ABC_ERR_ARG; }
    if (d->sim_kind != ABC_SIM_SSA && d->sim_kind != ABC_SIM_ODE) { abc_set_error("unknown sim_kind %d", d->sim_kind); return ABC_ERR_ARG; }
    if (d->sim_kind == ABC_SIM_SSA) {
        if (d->n_cells < 2 || d->n_cells > (1 << 20)) { abc_set_error("n_cells must be in [2, 2^20]"); return ABC_ERR_ARG; }
        if (d->n_pre_cycles < 0 || d->n_pre_cycles > 13) { abc_set_error("n_pre_cycles must be in [0, 13]"); return ABC_ERR_ARG; }
        // the label window must start inside the simulated time span
        for (int j = 0; j < ABC_NCOND; ++j)
            for (int a = 0; a < ABC_NAGE; ++a)
                if (d->agevec[a] - d->pulse[j] - d->chase[j] < -(double)d->n_pre_cycles * d->cycle) {
                    abc_set_error("n_pre_cycles too small for pulse+chase of condition %d", j + 1);
                    return ABC_ERR_ARG;
                }
    }
    c->design = *d;
    c->beta_q32.clear();
    for (int g = 0; g < 11; ++g) c->beta_off[g] = 0;
    if (d->downsampling) {
        const double* bl[2] = {d->betas_pulse, d->betas_chase};
        const int32_t* cl[2] = {d->cluster_pulse, d->cluster_chase};
        const int32_t nn[2] = {d->n_pulse, d->n_chase};
        for (int s = 0; s < 2; ++s) {
            if (!bl[s] || !cl[s] || nn[s] <= 0) { abc_set_error("downsampling requires beta lists and cluster ids"); return ABC_ERR_ARG; }
            for (int k = 1; k <= ABC_NAGE; ++k) {
                const int grp = s * ABC_NAGE + (k - 1);
                c->beta_off[grp] = (int32_t)c->beta_q32.size();
                double sum = 0.0, sum2 = 0.0;
                long cnt = 0;
                for (int i = 0; i < nn[s]; ++i) {
                    if (cl[s][i] != k) continue;
                    const double b = bl[s][i];
                    if (!(b >= 0.0 && b <= 1.0)) { abc_set_error("capture efficiency %g outside [0,1]", b); return ABC_ERR_ARG; }
                    double q = std::floor(b * 4294967296.0 + 0.5);
                    c->beta_q32.push_back(q >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)q);
                    sum += b; sum2 += b * b; cnt++;
                }
                if (cnt < 2) { abc_set_error("age cluster %d of the %s cells has < 2 capture efficiencies", k, s ? "chase" : "pulse"); return ABC_ERR_ARG; }
                // model.jl:229-234: mean(beta), mean(beta.^2), var(beta) (corrected)
                const double mu = sum / (double)cnt;
                double ss = 0.0;
                for (int i = 0; i < nn[s]; ++i) if (cl[s][i] == k) ss += (bl[s][i] - mu) * (bl[s][i] - mu);
                c->beta_mean[grp] = mu; c->beta_m2[grp] = sum2 / (double)cnt; c->beta_var[grp] = ss / (double)(cnt - 1);
            }
        }
        c->beta_off[10] = (int32_t)c->beta_q32.size();
        int rc = c->d_beta.ensure(c->beta_q32.size());
        if (rc != ABC_OK) return rc;
        ABC_CUDA_CHECK(cudaMemcpy(c->d_beta.p, c->beta_q32.data(), c->beta_q32.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
        double bm[30];
        for (int g = 0; g < 10; ++g) { bm[g] = c->beta_mean[g]; bm[10 + g] = c->beta_m2[g]; bm[20 + g] = c->beta_var[g]; }
        rc = c->d_beta_mom.ensure(30);
        if (rc != ABC_OK) return rc;
        ABC_CUDA_CHECK(cudaMemcpy(c->d_beta_mom.p, bm, sizeof(bm), cudaMemcpyHostToDevice));
    } else {
        int rc = c->d_beta.ensure(1);
        if (rc != ABC_OK) return rc;
    }
    c->design.betas_pulse = c->design.betas_chase = nullptr;   // not retained
    c->design.cluster_pulse = c->design.cluster_chase = nullptr;
    int rc = c->d_age_dist.ensure(ABC_NAGE * ABC_NCOND);
    if (rc != ABC_OK) return rc;
    ABC_CUDA_CHECK(cudaMemcpy(c->d_age_dist.p, d->age_dist, sizeof(double) * ABC_NAGE * ABC_NCOND, cudaMemcpyHostToDevice));
    c->has_design = true;
    return ABC_OK;
}

extern "C" int abc_set_data(abc_ctx_t* c, const double* d, const double* se, int32_t G) {
    CTX_GUARD(c);
    if (!d || !se || G <= 0) { abc_set_error("abc_set_data: bad arguments"); return ABC_ERR_ARG; }
    const size_t n = (size_t)G * ABC_NSTATS;
    int rc = c->d_d.ensure(n);
    if (rc == ABC_OK) rc = c->d_den.ensure(n);
    if (rc == ABC_OK) rc = c->d_fbw.ensure(n);
    if (rc == ABC_OK) rc = c->d_fa.ensure((size_t)G);
    DevBuf<double> d_se;
    if (rc == ABC_OK) rc = d_se.ensure(n);
    if (rc == ABC_OK) rc = c->d_counts.ensure((size_t)G);
    if (rc != ABC_OK) return rc;
    ABC_CUDA_CHECK(cudaMemcpy(c->d_d.p, d, n * sizeof(double), cudaMemcpyHostToDevice));
    ABC_CUDA_CHECK(cudaMemcpy(d_se.p, se, n * sizeof(double), cudaMemcpyHostToDevice));
    rc = abc_launch_prepare_data(c->d_d.p, d_se.p, G, c->d_den.p, c->d_fbw.p, c->d_fa.p, c->stream);
    c->launches += 2;
    if (rc == ABC_OK) {
        cudaError_t e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) { abc_set_error("prepare_data failed: %s", cudaGetErrorString(e)); rc = ABC_ERR_CUDA; }
    }
    d_se.release();
    if (rc != ABC_OK) return rc;
    {
        // tables of the tile-pruned scoring path, built from the denominators exactly as the device computed them
        std::vector<double> h_den(n);
        ABC_CUDA_CHECK(cudaMemcpy(h_den.data(), c->d_den.p, n * sizeof(double), cudaMemcpyDeviceToHost));
        AbcScore3Host h;
        abc_score3_build(d, h_den.data(), G, h);
        if ((rc = c->d_s3_tb.ensure(h.tb.size() / 4)) != ABC_OK) return rc;
        if ((rc = c->d_s3_ab.ensure(h.ab.size() / 4)) != ABC_OK) return rc;
        if ((rc = c->d_s3_wt.ensure(h.wt.size())) != ABC_OK) return rc;
        if ((rc = c->d_s3_gidx.ensure(h.gidx.size())) != ABC_OK) return rc;
        if ((rc = c->d_s3_ok.ensure(h.okmask.size())) != ABC_OK) return rc;
        ABC_CUDA_CHECK(cudaMemcpy(c->d_s3_tb.p, h.tb.data(), h.tb.size() * sizeof(float), cudaMemcpyHostToDevice));
        ABC_CUDA_CHECK(cudaMemcpy(c->d_s3_ab.p, h.ab.data(), h.ab.size() * sizeof(float), cudaMemcpyHostToDevice));
        ABC_CUDA_CHECK(cudaMemcpy(c->d_s3_wt.p, h.wt.data(), h.wt.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
        ABC_CUDA_CHECK(cudaMemcpy(c->d_s3_gidx.p, h.gidx.data(), h.gidx.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
        ABC_CUDA_CHECK(cudaMemcpy(c->d_s3_ok.p, h.okmask.data(), h.okmask.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
        c->s3_ntiles = h.ntiles;
        std::vector<float> bblob;
        abc_score_mma_build(d, h_den.data(), h, bblob, &c->mf_max_slack);
        if ((rc = c->d_mf_b.ensure(bblob.size())) != ABC_OK) return rc;
        ABC_CUDA_CHECK(cudaMemcpy(c->d_mf_b.p, bblob.data(), bblob.size() * sizeof(float), cudaMemcpyHostToDevice));
    }
    ABC_CUDA_CHECK(cudaMemset(c->d_counts.p, 0, (size_t)G * sizeof(unsigned long long)));
    ABC_CUDA_CHECK(cudaMemset(c->d_acc_count.p, 0, sizeof(unsigned long long)));
    c->G = G;
    c->has_data = true;
    c->acc_budget = 0;
    return ABC_OK;
}

// ------------------------------------------------------------------------------------------------
static int check_model(int m) {
    if (m < 1 || m > 5) { abc_set_error("model index m = %d must be in 1..5", m); return ABC_ERR_ARG; }
    return ABC_OK;
}

extern "C" int abc_fix_params(abc_ctx_t* c, int m, int64_t n, int64_t offset, uint64_t seed, double* theta) {
    CTX_GUARD(c);
    int rc = check_model(m);
    if (rc != ABC_OK) return rc;
    if (n < 0 || (n > 0 && !theta)) { abc_set_error("abc_fix_params: bad arguments"); return ABC_ERR_ARG; }
    if (n == 0) return ABC_OK;
    const int P = abc_n_params(m);
    rc = c->d_theta.ensure((size_t)n * P);
    if (rc != ABC_OK) return rc;
    rc = abc_launch_prior(c->d_theta.p, m, n, offset, seed, c->stream);
    c->launches++;
    if (rc != ABC_OK) return rc;
    ABC_CUDA_CHECK(cudaMemcpyAsync(theta, c->d_theta.p, (size_t)n * P * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    ABC_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    return ABC_OK;
}

static AbcSsaParams make_ssa_params(const abc_ctx* c, int m, int64_t n, int64_t offset, uint64_t seed) {
    AbcSsaParams p;
    memset(&p, 0, sizeof(p));
    p.n_particles = n; p.particle_offset = offset;
    p.seed_lo = (uint32_t)seed; p.seed_hi = (uint32_t)(seed >> 32);
    p.m = m; p.scaling = (m != 2) ? 1 : 0;
    p.n_cells = c->design.n_cells;
    p.chunks = (c->design.n_cells + 31) / 32;
    p.n_pre = c->design.n_pre_cycles;
    p.downsampling = c->design.downsampling;
    p.single_readout = -1;
    p.hybrid = c->ssa_hybrid;
    p.adaptive = c->ssa_adaptive;
    p.cycle = c->design.cycle;
    for (int a = 0; a < 5; ++a) p.agevec[a] = c->design.agevec[a];
    for (int j = 0; j < 11; ++j) { p.pulse[j] = c->design.pulse[j]; p.chase[j] = c->design.chase[j]; p.beta_off[j] = c->beta_off[j]; }
    return p;
}

// device-side pipeline: (prior) -> rates -> SSA -> moments -> statistics, all on `st`
static int simulate_device(abc_ctx* c, int m, int64_t n, int64_t offset, uint64_t seed, int prior_supplied,
                           double* d_theta, double* d_stats, double* d_moments_out, cudaStream_t st) {
    int rc = ABC_OK;
    if (!c->has_design) { abc_set_error("abc_set_design has not been called"); return ABC_ERR_STATE; }
    if (!prior_supplied) {
        rc = abc_launch_prior(d_theta, m, n, offset, seed, st);
        c->launches++;
        if (rc != ABC_OK) return rc;
    }
    if (c->design.sim_kind == ABC_SIM_ODE) {
        // what scripts/model.jl integrates: moment ODEs -> (downsample) -> moments -> statistics
        double* d_mom_ode = d_moments_out;
        if (!d_mom_ode) {
            if ((rc = c->d_moments.ensure((size_t)n * ABC_NREAD * 5)) != ABC_OK) return rc;
            d_mom_ode = c->d_moments.p;
        }
        if ((rc = c->d_ss_iv.ensure((size_t)n * 9)) != ABC_OK) return rc;
        if ((rc = c->d_prefix.ensure((size_t)n * ABC_NREAD * 9)) != ABC_OK) return rc;
        ABC_CUDA_CHECK(cudaMemsetAsync(c->d_counters.p, 0, 8 * sizeof(unsigned long long), st));
        ABC_CUDA_CHECK(cudaEventRecord(c->ev[0], st));
        if ((rc = abc_launch_ode(d_theta, c->design, m, n, c->d_beta_mom.p, c->d_ss_iv.p, c->d_prefix.p, d_mom_ode, c->d_counters.p, st)) != ABC_OK) return rc;
        c->launches += 3;
        ABC_CUDA_CHECK(cudaEventRecord(c->ev[1], st));
        if (d_stats) {
            if ((rc = abc_launch_summary_stats(d_mom_ode, c->d_age_dist.p, n, d_stats, c->stats_guards == 1, st)) != ABC_OK) return rc;
            c->launches++;
        }
        ABC_CUDA_CHECK(cudaEventRecord(c->ev[2], st));
        return ABC_OK;
    }
    if ((rc = c->d_rates.ensure((size_t)n)) != ABC_OK) return rc;
    if ((rc = c->d_sums.ensure((size_t)n * ABC_NREAD * 5)) != ABC_OK) return rc;
    double* d_mom = d_moments_out;
    if (!d_mom) {
        if ((rc = c->d_moments.ensure((size_t)n * ABC_NREAD * 5)) != ABC_OK) return rc;
        d_mom = c->d_moments.p;
    }
    if ((rc = abc_launch_rates(d_theta, m, n, c->d_rates.p, c->ssa_hybrid, (c->ssa_adaptive && c->ssa_hybrid) ? c->design.n_pre_cycles : 0,
                               c->design.cycle, st)) != ABC_OK) return rc;
    c->launches++;
    AbcSsaParams prm = make_ssa_params(c, m, n, offset, seed);
    if (c->ssa_hybrid == 2) {
        // where every (particle, read-out) starts, and the expected work per particle (replaces the rate-based hint)
        if ((rc = c->d_win.ensure((size_t)n * ABC_NREAD)) != ABC_OK) return rc;
        if ((rc = abc_launch_window(c->d_rates.p, prm, c->d_win.p, n, st)) != ABC_OK) return rc;
        c->launches++;
    }
    // longest-processing-time-first order of the particles (scheduling only, results are order independent)
    const int* d_order = nullptr;
    if (n >= 64 && n < (1ll << 31)) {
        const size_t tmp = abc_order_temp_bytes((int)n);
        if ((rc = c->d_keys_in.ensure((size_t)n)) != ABC_OK) return rc;
        if ((rc = c->d_keys_out.ensure((size_t)n)) != ABC_OK) return rc;
        if ((rc = c->d_idx_in.ensure((size_t)n)) != ABC_OK) return rc;
        if ((rc = c->d_order.ensure((size_t)n)) != ABC_OK) return rc;
        if ((rc = c->d_sort_tmp.ensure(tmp)) != ABC_OK) return rc;
        if ((rc = abc_launch_order(c->d_rates.p, (int)n, c->d_keys_in.p, c->d_keys_out.p, c->d_idx_in.p, c->d_order.p,
                                   c->d_sort_tmp.p, tmp, st)) != ABC_OK) return rc;
        c->launches += 2;
        d_order = c->d_order.p;
    }
    ABC_CUDA_CHECK(cudaMemsetAsync(c->d_sums.p, 0, (size_t)n * ABC_NREAD * 5 * sizeof(unsigned long long), st));
    ABC_CUDA_CHECK(cudaMemsetAsync(c->d_counters.p, 0, 8 * sizeof(unsigned long long), st));
    ABC_CUDA_CHECK(cudaEventRecord(c->ev[0], st));
    if (c->ssa_hybrid == 2) {
        if ((rc = abc_launch_tele(c->d_rates.p, prm, c->d_win.p, c->d_beta.p, c->d_sums.p, c->d_counters.p, c->d_work.p, nullptr,
                                  d_order, c->sm_count, st)) != ABC_OK) return rc;
    } else {
        if ((rc = abc_launch_ssa(c->d_rates.p, prm, c->d_beta.p, c->d_sums.p, c->d_counters.p, c->d_work.p, nullptr, d_order, 0,
                                 c->sm_count, st)) != ABC_OK) return rc;
    }
    c->launches++;
    ABC_CUDA_CHECK(cudaEventRecord(c->ev[1], st));
    if ((rc = abc_launch_moments_from_sums(c->d_sums.p, n, c->design.n_cells, d_mom, st)) != ABC_OK) return rc;
    c->launches++;
    if (d_stats) {
        if ((rc = abc_launch_summary_stats(d_mom, c->d_age_dist.p, n, d_stats, c->stats_guards != 0, st)) != ABC_OK) return rc;
        c->launches++;
    }
    ABC_CUDA_CHECK(cudaEventRecord(c->ev[2], st));
    return ABC_OK;
}

static int read_counters(abc_ctx* c, int64_t n, abc_counters_t* out, bool accumulate) {
    unsigned long long h[8];
    ABC_CUDA_CHECK(cudaMemcpy(h, c->d_counters.p, sizeof(h), cudaMemcpyDeviceToHost));
    float ms_sim = 0.f, ms_st = 0.f;
    if (cudaEventElapsedTime(&ms_sim, c->ev[0], c->ev[1]) != cudaSuccess) cudaGetLastError();
    if (cudaEventElapsedTime(&ms_st, c->ev[1], c->ev[2]) != cudaSuccess) cudaGetLastError();
    abc_counters_t r;
    memset(&r, 0, sizeof(r));
    r.n_particles = (uint64_t)n; r.n_lineages = h[0]; r.n_events = h[1]; r.n_draws = h[2]; r.n_ode_steps = h[4];
    r.ms_simulate = ms_sim; r.ms_stats = ms_st;
    if (accumulate) {
        c->last.n_particles += r.n_particles; c->last.n_lineages += r.n_lineages; c->last.n_events += r.n_events;
        c->last.n_draws += r.n_draws; c->last.n_ode_steps += r.n_ode_steps; c->last.ms_simulate += r.ms_simulate; c->last.ms_stats += r.ms_stats;
    } else {
        double keep = c->last.ms_score;
        c->last = r;
        c->last.ms_score = keep;
    }
    if (out) *out = c->last;
    return ABC_OK;
}

#define SIM_CHUNK (1 << 17)

// particles per device launch: bounded by the work buffers and by the 32-bit work-item counter of the SSA kernel
static int64_t sim_chunk(const abc_ctx* c) {
    int64_t ch = SIM_CHUNK;
    if (c->has_design && c->design.sim_kind == ABC_SIM_SSA) {
        const int64_t items_per_particle = (int64_t)ABC_NREAD * ((c->design.n_cells + 31) / 32);
        ch = std::min<int64_t>(ch, std::max<int64_t>(1, 0xE0000000ll / items_per_particle));
    }
    return ch;
}

extern "C" int abc_simulate(abc_ctx_t* c, int m, int64_t n, int64_t offset, uint64_t seed, int prior_supplied,
                            double* theta, double* stats, abc_counters_t* counters) {
    CTX_GUARD(c);
    int rc = check_model(m);
    if (rc != ABC_OK) return rc;
    if (n < 0 || (n > 0 && (!theta || !stats))) { abc_set_error("abc_simulate: bad arguments"); return ABC_ERR_ARG; }
    const int P = abc_n_params(m);
    memset(&c->last, 0, sizeof(c->last));
    const int64_t chunk = sim_chunk(c);
    for (int64_t b0 = 0; b0 < n; b0 += chunk) {
        const int64_t nb = std::min<int64_t>(chunk, n - b0);
        if ((rc = c->d_theta.ensure((size_t)nb * P)) != ABC_OK) return rc;
        if ((rc = c->d_stats.ensure((size_t)nb * ABC_NSTATS)) != ABC_OK) return rc;
        if (prior_supplied)
            ABC_CUDA_CHECK(cudaMemcpyAsync(c->d_theta.p, theta + b0 * P, (size_t)nb * P * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        rc = simulate_device(c, m, nb, offset + b0, seed, prior_supplied, c->d_theta.p, c->d_stats.p, nullptr, c->stream);
        if (rc != ABC_OK) return rc;
        if (!prior_supplied)
            ABC_CUDA_CHECK(cudaMemcpyAsync(theta + b0 * P, c->d_theta.p, (size_t)nb * P * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        ABC_CUDA_CHECK(cudaMemcpyAsync(stats + b0 * ABC_NSTATS, c->d_stats.p, (size_t)nb * ABC_NSTATS * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        ABC_CUDA_CHECK(cudaStreamSynchronize(c->stream));
        if ((rc = read_counters(c, nb, nullptr, true)) != ABC_OK) return rc;
    }
    if (counters) *counters = c->last;
    return ABC_OK;
}

extern "C" int abc_simulate_moments(abc_ctx_t* c, int m, int64_t n, int64_t offset, uint64_t seed,
                                    const double* theta, double* moments, abc_counters_t* counters) {
    CTX_GUARD(c);
    int rc = check_model(m);
    if (rc != ABC_OK) return rc;
    if (n < 0 || (n > 0 && (!theta || !moments))) { abc_set_error("abc_simulate_moments: bad arguments"); return ABC_ERR_ARG; }
    const int P = abc_n_params(m);
    memset(&c->last, 0, sizeof(c->last));
    const int64_t chunk = sim_chunk(c);
    for (int64_t b0 = 0; b0 < n; b0 += chunk) {
        const int64_t nb = std::min<int64_t>(chunk, n - b0);
        if ((rc = c->d_theta.ensure((size_t)nb * P)) != ABC_OK) return rc;
        if ((rc = c->d_moments.ensure((size_t)nb * ABC_NREAD * 5)) != ABC_OK) return rc;
        ABC_CUDA_CHECK(cudaMemcpyAsync(c->d_theta.p, theta + b0 * P, (size_t)nb * P * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        rc = simulate_device(c, m, nb, offset + b0, seed, 1, c->d_theta.p, nullptr, c->d_moments.p, c->stream);
        if (rc != ABC_OK) return rc;
        ABC_CUDA_CHECK(cudaMemcpyAsync(moments + b0 * ABC_NREAD * 5, c->d_moments.p, (size_t)nb * ABC_NREAD * 5 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        ABC_CUDA_CHECK(cudaStreamSynchronize(c->stream));
        if ((rc = read_counters(c, nb, nullptr, true)) != ABC_OK) return rc;
    }
    if (counters) *counters = c->last;
    return ABC_OK;
}

extern "C" int abc_ssa_cells(abc_ctx_t* c, int m, const double* theta, int64_t particle_index, uint64_t seed,
                             int cond, int age, int exact_math, uint32_t* counts) {
    CTX_GUARD(c);
    int rc = check_model(m);
    if (rc != ABC_OK) return rc;
    if (!c->has_design) { abc_set_error("abc_set_design has not been called"); return ABC_ERR_STATE; }
    if (c->design.sim_kind != ABC_SIM_SSA) { abc_set_error("abc_ssa_cells needs sim_kind == ABC_SIM_SSA"); return ABC_ERR_STATE; }
    if (!theta || !counts || cond < 0 || cond >= ABC_NCOND || age < 0 || age >= ABC_NAGE) { abc_set_error("abc_ssa_cells: bad arguments"); return ABC_ERR_ARG; }
    const int P = abc_n_params(m);
    const int nc = c->design.n_cells;
    if ((rc = c->d_theta.ensure((size_t)P)) != ABC_OK) return rc;
    if ((rc = c->d_rates.ensure(1)) != ABC_OK) return rc;
    if ((rc = c->d_cells.ensure((size_t)4 * nc)) != ABC_OK) return rc;
    ABC_CUDA_CHECK(cudaMemcpyAsync(c->d_theta.p, theta, P * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    if ((rc = abc_launch_rates(c->d_theta.p, m, 1, c->d_rates.p, c->ssa_hybrid, 0, c->design.cycle, c->stream)) != ABC_OK) return rc;
    ABC_CUDA_CHECK(cudaMemsetAsync(c->d_counters.p, 0, 8 * sizeof(unsigned long long), c->stream));
    AbcSsaParams prm = make_ssa_params(c, m, 1, particle_index, seed);
    prm.single_readout = cond * ABC_NAGE + age;
    if (c->ssa_hybrid == 2 && !exact_math) {
        if ((rc = c->d_win.ensure((size_t)ABC_NREAD)) != ABC_OK) return rc;
        if ((rc = abc_launch_window(c->d_rates.p, prm, c->d_win.p, 1, c->stream)) != ABC_OK) return rc;
        rc = abc_launch_tele(c->d_rates.p, prm, c->d_win.p, c->d_beta.p, nullptr, c->d_counters.p, c->d_work.p, c->d_cells.p, nullptr,
                             c->sm_count, c->stream);
        c->launches++;
    } else {
        rc = abc_launch_ssa(c->d_rates.p, prm, c->d_beta.p, nullptr, c->d_counters.p, c->d_work.p, c->d_cells.p, nullptr, exact_math,
                            c->sm_count, c->stream);
    }
    c->launches += 2;
    if (rc != ABC_OK) return rc;
    ABC_CUDA_CHECK(cudaMemcpyAsync(counts, c->d_cells.p, (size_t)4 * nc * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    ABC_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    return ABC_OK;
}

extern "C" int abc_ssa_window(abc_ctx_t* c, int m, const double* theta, float* starts, double* expected_draws) {
    CTX_GUARD(c);
    int rc = check_model(m);
    if (rc != ABC_OK) return rc;
    if (!c->has_design) { abc_set_error("abc_set_design has not been called"); return ABC_ERR_STATE; }
    if (c->design.sim_kind != ABC_SIM_SSA) { abc_set_error("abc_ssa_window needs sim_kind == ABC_SIM_SSA"); return ABC_ERR_STATE; }
    if (!theta || !starts) { abc_set_error("abc_ssa_window: bad arguments"); return ABC_ERR_ARG; }
    const int P = abc_n_params(m);
    if ((rc = c->d_theta.ensure((size_t)P)) != ABC_OK) return rc;
    if ((rc = c->d_rates.ensure(1)) != ABC_OK) return rc;
    if ((rc = c->d_win.ensure((size_t)ABC_NREAD)) != ABC_OK) return rc;
    ABC_CUDA_CHECK(cudaMemcpyAsync(c->d_theta.p, theta, P * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    if ((rc = abc_launch_rates(c->d_theta.p, m, 1, c->d_rates.p, c->ssa_hybrid, 0, c->design.cycle, c->stream)) != ABC_OK) return rc;
    AbcSsaParams prm = make_ssa_params(c, m, 1, 0, 0);
    if ((rc = abc_launch_window(c->d_rates.p, prm, c->d_win.p, 1, c->stream)) != ABC_OK) return rc;
    c->launches += 2;
    AbcRates r;
    ABC_CUDA_CHECK(cudaMemcpyAsync(starts, c->d_win.p, ABC_NREAD * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    ABC_CUDA_CHECK(cudaMemcpyAsync(&r, c->d_rates.p, sizeof(r), cudaMemcpyDeviceToHost, c->stream));
    ABC_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    if (expected_draws) *expected_draws = (r.pad1 != 0.0f) ? NAN : (double)r.pad0;
    return ABC_OK;
}

extern "C" int abc_summary_stats(abc_ctx_t* c, const double* moments, int64_t n, double* stats) {
    CTX_GUARD(c);
    if (!c->has_design) { abc_set_error("abc_set_design has not been called"); return ABC_ERR_STATE; }
    if (n < 0 || (n > 0 && (!moments || !stats))) { abc_set_error("abc_summary_stats: bad arguments"); return ABC_ERR_ARG; }
    if (n == 0) return ABC_OK;
    int rc;
    if ((rc = c->d_moments.ensure((size_t)n * ABC_NREAD * 5)) != ABC_OK) return rc;
    if ((rc = c->d_stats.ensure((size_t)n * ABC_NSTATS)) != ABC_OK) return rc;
    ABC_CUDA_CHECK(cudaMemcpyAsync(c->d_moments.p, moments, (size_t)n * ABC_NREAD * 5 * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    const int guards = (c->stats_guards < 0) ? (c->design.sim_kind == ABC_SIM_SSA) : c->stats_guards;
    if ((rc = abc_launch_summary_stats(c->d_moments.p, c->d_age_dist.p, n, c->d_stats.p, guards, c->stream)) != ABC_OK) return rc;
    c->launches++;
    ABC_CUDA_CHECK(cudaMemcpyAsync(stats, c->d_stats.p, (size_t)n * ABC_NSTATS * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    ABC_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    return ABC_OK;
}

// ------------------------------------------------------------------------------------------------
// make room for `want` accepted tuples in total, keeping the ones already stored (device-to-device copy)
static int ensure_accept(abc_ctx* c, int64_t want, cudaStream_t st) {
    if (c->acc_capacity >= want) return ABC_OK;
    unsigned long long cnt = 0;
    if (c->acc_capacity > 0) {
        ABC_CUDA_CHECK(cudaStreamSynchronize(st));
        ABC_CUDA_CHECK(cudaMemcpy(&cnt, c->d_acc_count.p, sizeof(cnt), cudaMemcpyDeviceToHost));
        if ((int64_t)cnt > c->acc_capacity) { abc_set_error("accepted-tuple buffer overflowed"); return ABC_ERR_NOMEM; }
    }
    DevBuf<int32_t> ng; DevBuf<long long> np_; DevBuf<double> ne;
    int rc;
    if ((rc = ng.ensure((size_t)want)) != ABC_OK) return rc;
    if ((rc = np_.ensure((size_t)want)) != ABC_OK) return rc;
    if ((rc = ne.ensure((size_t)want)) != ABC_OK) return rc;
    if (cnt) {
        ABC_CUDA_CHECK(cudaMemcpy(ng.p, c->d_acc_gene.p, cnt * sizeof(int32_t), cudaMemcpyDeviceToDevice));
        ABC_CUDA_CHECK(cudaMemcpy(np_.p, c->d_acc_particle.p, cnt * sizeof(long long), cudaMemcpyDeviceToDevice));
        ABC_CUDA_CHECK(cudaMemcpy(ne.p, c->d_acc_err.p, cnt * sizeof(double), cudaMemcpyDeviceToDevice));
    }
    c->d_acc_gene = std::move(ng); c->d_acc_particle = std::move(np_); c->d_acc_err = std::move(ne);
    c->acc_capacity = want;
    return ABC_OK;
}

// more pairs were accepted than the tuple buffer holds: the surplus tuples were dropped (counts[] still include them).  The
// stored count is clamped back to the capacity so that the context stays usable; the caller must abc_accept_reset and
// rescore with a larger "accept_capacity" (or a lower eps / smaller batches).
static int accept_overflow(abc_ctx* c, unsigned long long total) {
    const unsigned long long cap = (unsigned long long)c->acc_capacity;
    cudaMemcpy(c->d_acc_count.p, &cap, sizeof(cap), cudaMemcpyHostToDevice);
    abc_set_error("accepted-tuple buffer overflow: %llu accepted pairs > capacity %lld (surplus dropped; per-gene counts include "
                  "them): call abc_accept_reset, then raise the option accept_capacity, lower eps or score in smaller batches",
                  total, (long long)c->acc_capacity);
    return ABC_ERR_NOMEM;
}

static int64_t default_accept_capacity(int64_t n, int G) {
    // generous: 2 % of the pairs of this call (measured: 0.16 % for prior draws at eps = 4.8), at least 1 Mi tuples
    double w = 0.02 * (double)n * (double)G;
    int64_t cap = (int64_t)w;
    if (cap < (1 << 20)) cap = 1 << 20;
    return cap;
}

static int score_device(abc_ctx* c, const double* d_stats, int64_t n, int64_t offset, double eps, int layout,
                        double* d_err, cudaStream_t st) {
    if (!c->has_data) { abc_set_error("abc_set_data has not been called"); return ABC_ERR_STATE; }
    if (layout != ABC_ERR_NONE && layout != ABC_ERR_GENE_MAJOR && layout != ABC_ERR_PARTICLE_MAJOR) {
        abc_set_error("unknown err_layout %d", layout);
        return ABC_ERR_ARG;
    }
    // room for this call: up to 2 % of its pairs on top of the budget already promised to earlier calls since
    // the last reset (a conservative host-side running bound; the exact count is only read when growing)
    // eps >= 10 accepts every finite pair (errors are clipped at 10.0): room for all of them, or a clear refusal
    if (eps >= 10.0) {
        if ((double)n * (double)c->G > 1.0e9) {
            abc_set_error("eps = %g >= 10 accepts every (particle, gene) pair: %lld x %d tuples do not fit; score in batches of "
                          "<= %lld particles or use eps < 10", eps, (long long)n, c->G, (long long)(1000000000ll / c->G));
            return ABC_ERR_ARG;
        }
        c->acc_budget += n * (int64_t)c->G;
    } else {
        c->acc_budget += default_accept_capacity(n, c->G);
    }
    int rc = ensure_accept(c, std::max<int64_t>(c->acc_budget, c->acc_min_capacity), st);
    if (rc != ABC_OK) return rc;
    AbcScoreArgs a;
    a.stats = d_stats; a.d = c->d_d.p; a.den = c->d_den.p; a.fbw = c->d_fbw.p; a.fa = c->d_fa.p;
    a.force_reference_kernel = c->force_reference_score;
    a.n = n; a.G = c->G; a.particle_offset = offset; a.eps = eps;
    a.err_layout = layout; a.err = (layout == ABC_ERR_NONE) ? nullptr : d_err;
    a.counts = c->d_counts.p; a.acc_count = c->d_acc_count.p; a.acc_capacity = c->acc_capacity;
    a.acc_gene = c->d_acc_gene.p; a.acc_particle = c->d_acc_particle.p; a.acc_err = c->d_acc_err.p;
    a.fstats = nullptr; a.rnan = nullptr;
    ABC_CUDA_CHECK(cudaEventRecord(c->ev[3], st));
    a.gm_stride = n;
    if (!c->force_reference_score && c->score_tile_kernel && eps < 10.0 && n > 0) {
        // tile-pruned path: classification, filter (one CTA per gene tile x particle block; also fills the matrix),
        // stage 3 on the queued pairs.  The stage-3 queue has room for every pair of a sub-batch (2 bytes each),
        // so a sub-batch is at most ~2 GB of queue.
        AbcScore3Tables x;
        x.ntiles = c->s3_ntiles;
        x.tb = c->d_s3_tb.p; x.ab = c->d_s3_ab.p; x.wt = c->d_s3_wt.p;
        x.gidx = c->d_s3_gidx.p; x.okmask = c->d_s3_ok.p;
        if (c->score_mma_filter && layout != ABC_ERR_NONE) {
            // tensor-core filter: statistics -> TF32 operand, GEMM -> sign-bit words, stage 3 on the flagged pairs; the matrix's
            // background is written on a side stream meanwhile.  Chunks of <= 2^19 particles (1 KB of work space each).
            // mode 1: sign-bit words + mask-driven stage 3 (background split between the filter kernel and stage 3);
            // mode 2: the filter kernel writes the whole background and the stage-3 queue, then the queue-driven stage 3
            const bool queue = c->score_mma_filter == 2;
            const int64_t per_particle_q = (int64_t)x.ntiles * 32 * 2;
            int64_t chunk = std::min<int64_t>((n + 2047) / 2048 * 2048, 1ll << 19);
            if (queue) chunk = std::min<int64_t>(chunk, std::max<int64_t>(2048, ((int64_t)2000000000 / per_particle_q) / 2048 * 2048));
            if ((rc = c->d_s3_nanw[0].ensure((size_t)((chunk + 31) / 32))) != ABC_OK) return rc;
            if ((rc = c->d_mf_a[0].ensure((size_t)chunk * 128)) != ABC_OK) return rc;
            if (queue) {
                if ((rc = c->d_s3_qcnt[0].ensure(abc_score3_blocks(chunk) * (size_t)x.ntiles)) != ABC_OK) return rc;
                if ((rc = c->d_s3_q2[0].ensure(abc_score3_queue_entries(chunk, x.ntiles))) != ABC_OK) return rc;
            } else {
                if ((rc = c->d_mf_mask[0].ensure((size_t)abc_score_mma_tiles(x.ntiles) * 8 * (size_t)chunk)) != ABC_OK) return rc;
                if ((rc = c->d_mf_done.ensure((size_t)(chunk / 2048 + 2))) != ABC_OK) return rc;
            }
            for (int64_t s0 = 0; s0 < n && rc == ABC_OK; s0 += chunk) {
                AbcScoreArgs b = a;
                b.n = std::min<int64_t>(chunk, n - s0);
                b.stats = d_stats + s0 * ABC_NSTATS;
                b.particle_offset = offset + s0;
                if (b.err != nullptr) b.err = (layout == ABC_ERR_GENE_MAJOR) ? d_err + s0 : d_err + s0 * (int64_t)c->G;
                x.nanw = c->d_s3_nanw[0].p; x.W = (b.n + 31) / 32;
                x.gmask = c->d_mf_mask[0].p; x.n_pad = (b.n + 127) / 128 * 128; x.n_rows = x.n_pad; x.fill_done = c->d_mf_done.p;
                x.q2 = queue ? c->d_s3_q2[0].p : nullptr; x.qcnt = queue ? c->d_s3_qcnt[0].p : nullptr;
                rc = abc_launch_score_mma(b, x, c->d_mf_a[0].p, c->d_mf_b.p, nullptr, c->sm_count, st);
                c->launches += 3;
            }
            ABC_CUDA_CHECK(cudaEventRecord(c->ev[4], st));
            return rc;
        }
        const int64_t per_particle = (int64_t)x.ntiles * 32 * 2;
        int64_t sub = std::max<int64_t>(2048, ((int64_t)2000000000 / per_particle) / 2048 * 2048);
        // at least four sub-batches of >= 16 particle blocks when the batch is large enough to pipeline
        const int nsb = c->score_sub_batches >= 2 ? c->score_sub_batches : (n < 262144 ? 2 : 4);
        const bool overlap = c->score_overlap && n >= (int64_t)nsb * 4 * 2048;
        if (overlap) sub = std::min<int64_t>(sub, ((n + nsb - 1) / nsb + 2047) / 2048 * 2048);
        sub = std::min<int64_t>(sub, (n + 2047) / 2048 * 2048);
        const size_t nblocks = abc_score3_blocks(sub);
        const int lanes = overlap ? 2 : 1;
        for (int l = 0; l < lanes; ++l) {
            if ((rc = c->d_s3_fstats[l].ensure((size_t)sub * ABC_NSTATS)) != ABC_OK) return rc;
            if ((rc = c->d_s3_live[l].ensure((size_t)x.ntiles * (size_t)((sub + 31) / 32))) != ABC_OK) return rc;
            if ((rc = c->d_s3_nanw[l].ensure((size_t)((sub + 31) / 32))) != ABC_OK) return rc;
            if ((rc = c->d_s3_qcnt[l].ensure(nblocks * (size_t)x.ntiles)) != ABC_OK) return rc;
            if ((rc = c->d_s3_q2[l].ensure(abc_score3_queue_entries(sub, x.ntiles))) != ABC_OK) return rc;
        }
        if (overlap) {
            ABC_CUDA_CHECK(cudaEventRecord(c->s3_ev_begin, st));
            for (int l = 0; l < 2; ++l) ABC_CUDA_CHECK(cudaStreamWaitEvent(c->s3_stream[l], c->s3_ev_begin, 0));
        }
        int j = 0;
        for (int64_t s0 = 0; s0 < n && rc == ABC_OK; s0 += sub, ++j) {
            const int l = overlap ? (j & 1) : 0;
            AbcScoreArgs b = a;
            b.n = std::min<int64_t>(sub, n - s0);
            b.stats = d_stats + s0 * ABC_NSTATS;
            b.particle_offset = offset + s0;
            b.fstats = c->d_s3_fstats[l].p;
            if (b.err != nullptr) b.err = (layout == ABC_ERR_GENE_MAJOR) ? d_err + s0 : d_err + s0 * (int64_t)c->G;
            x.live = c->d_s3_live[l].p; x.nanw = c->d_s3_nanw[l].p; x.q2 = c->d_s3_q2[l].p; x.qcnt = c->d_s3_qcnt[l].p;
            x.W = (b.n + 31) / 32;
            rc = abc_launch_score3(b, x, overlap ? c->s3_stream[l] : st);
            c->launches += 3;
        }
        if (overlap) {
            for (int l = 0; l < 2; ++l) {
                ABC_CUDA_CHECK(cudaEventRecord(c->s3_ev_end[l], c->s3_stream[l]));
                ABC_CUDA_CHECK(cudaStreamWaitEvent(st, c->s3_ev_end[l], 0));
            }
        }
        ABC_CUDA_CHECK(cudaEventRecord(c->ev[4], st));
        return rc;
    }
    if (!c->force_reference_score && eps < 10.0) {
        if ((rc = c->d_fstats.ensure((size_t)n * ABC_NSTATS)) != ABC_OK) return rc;
        if ((rc = c->d_rnan.ensure((size_t)n)) != ABC_OK) return rc;
        if ((rc = abc_launch_score_prep(d_stats, n, c->d_fstats.p, c->d_rnan.p, st)) != ABC_OK) return rc;
        c->launches++;
        a.fstats = c->d_fstats.p; a.rnan = c->d_rnan.p;
    }
    rc = abc_launch_score(a, c->sm_count, st);
    c->launches++;
    ABC_CUDA_CHECK(cudaEventRecord(c->ev[4], st));
    return rc;
}

// diagnostic: the raw accumulators of the tensor-core filter, V[i][column] (column = tile * 32 + slot), for device-resident
// statistics; d_out has ceil(n / 128) * 128 rows of abc_score_mma_columns() floats.  No matrix, no acceptance.
extern "C" int abc_score_mma_columns(abc_ctx_t* c) {
    if (!c || !c->has_data) return 0;
    return abc_score_mma_tiles(c->s3_ntiles) * 256;
}
extern "C" int abc_score_mma_debug(abc_ctx_t* c, const double* d_stats, int64_t n, float* d_out, int32_t* gene_of_column) {
    CTX_GUARD(c);
    if (!c->has_data) { abc_set_error("abc_set_data has not been called"); return ABC_ERR_STATE; }
    if (n <= 0 || !d_stats || !d_out) { abc_set_error("abc_score_mma_debug: bad arguments"); return ABC_ERR_ARG; }
    int rc;
    AbcScoreArgs a{};
    a.stats = d_stats; a.n = n; a.G = c->G; a.eps = 0.0; a.err = nullptr; a.err_layout = ABC_ERR_NONE; a.gm_stride = n;
    a.counts = c->d_counts.p; a.acc_count = c->d_acc_count.p; a.acc_capacity = 0;
    AbcScore3Tables x{};
    x.ntiles = c->s3_ntiles; x.tb = c->d_s3_tb.p; x.ab = c->d_s3_ab.p; x.wt = c->d_s3_wt.p; x.gidx = c->d_s3_gidx.p; x.okmask = c->d_s3_ok.p;
    if ((rc = c->d_s3_nanw[0].ensure((size_t)((n + 31) / 32))) != ABC_OK) return rc;
    const size_t n_pad = (size_t)((n + 127) / 128) * 128;
    if ((rc = c->d_mf_a[0].ensure(n_pad * 128)) != ABC_OK) return rc;
    if ((rc = c->d_mf_mask[0].ensure((size_t)abc_score_mma_tiles(x.ntiles) * 8 * n_pad)) != ABC_OK) return rc;
    if ((rc = c->d_mf_done.ensure(n_pad / 2048 + 2)) != ABC_OK) return rc;
    x.fill_done = c->d_mf_done.p;
    x.nanw = c->d_s3_nanw[0].p; x.W = (n + 31) / 32; x.gmask = c->d_mf_mask[0].p; x.n_pad = (int64_t)n_pad;
    a.eps = -1.0;                                        // nothing is accepted
    x.n_rows = x.n_pad; x.q2 = nullptr; x.qcnt = nullptr;
    rc = abc_launch_score_mma(a, x, c->d_mf_a[0].p, c->d_mf_b.p, d_out, c->sm_count, c->stream);
    if (rc != ABC_OK) return rc;
    ABC_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    if (gene_of_column) {
        const int cols = abc_score_mma_columns(c);
        for (int k = 0; k < cols; ++k) gene_of_column[k] = -1;
        ABC_CUDA_CHECK(cudaMemcpy(gene_of_column, c->d_s3_gidx.p, (size_t)x.ntiles * 32 * sizeof(int32_t), cudaMemcpyDeviceToHost));
    }
    return ABC_OK;
}

extern "C" int abc_score(abc_ctx_t* c, const double* stats, int64_t n, int64_t offset, double eps, int layout,
                         double* err, int64_t* counts, abc_counters_t* counters) {
    CTX_GUARD(c);
    if (!c->has_data) { abc_set_error("abc_set_data has not been called"); return ABC_ERR_STATE; }
    if (n < 0 || (n > 0 && !stats) || (layout != ABC_ERR_NONE && n > 0 && !err)) { abc_set_error("abc_score: bad arguments"); return ABC_ERR_ARG; }
    const int G = c->G;
    int rc;
    // chunk so that the device copy of the error matrix stays below ~8 GB
    int64_t chunk = (layout == ABC_ERR_NONE) ? (1ll << 22) : std::max<int64_t>(1024, (int64_t)(8.0e9 / (8.0 * G)));
    double ms_total = 0.0;
    for (int64_t b0 = 0; b0 < n; b0 += chunk) {
        const int64_t nb = std::min<int64_t>(chunk, n - b0);
        if ((rc = c->d_sstats.ensure((size_t)nb * ABC_NSTATS)) != ABC_OK) return rc;
        if (layout != ABC_ERR_NONE && (rc = c->d_err.ensure((size_t)nb * G)) != ABC_OK) return rc;
        ABC_CUDA_CHECK(cudaMemcpyAsync(c->d_sstats.p, stats + b0 * ABC_NSTATS, (size_t)nb * ABC_NSTATS * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        rc = score_device(c, c->d_sstats.p, nb, offset + b0, eps, layout, c->d_err.p, c->stream);
        if (rc != ABC_OK) return rc;
        if (layout == ABC_ERR_PARTICLE_MAJOR) {
            ABC_CUDA_CHECK(cudaMemcpyAsync(err + b0 * G, c->d_err.p, (size_t)nb * G * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        } else if (layout == ABC_ERR_GENE_MAJOR) {
            ABC_CUDA_CHECK(cudaMemcpy2DAsync(err + b0, (size_t)n * sizeof(double), c->d_err.p, (size_t)nb * sizeof(double),
                                             (size_t)nb * sizeof(double), (size_t)G, cudaMemcpyDeviceToHost, c->stream));
        }
        ABC_CUDA_CHECK(cudaStreamSynchronize(c->stream));
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, c->ev[3], c->ev[4]) != cudaSuccess) cudaGetLastError();
        ms_total += ms;
    }
    c->last.ms_score = ms_total;
    unsigned long long total = 0;
    ABC_CUDA_CHECK(cudaMemcpy(&total, c->d_acc_count.p, sizeof(total), cudaMemcpyDeviceToHost));
    if ((int64_t)total > c->acc_capacity) return accept_overflow(c, total);
    if (counts) {
        std::vector<unsigned long long> h((size_t)G);
        ABC_CUDA_CHECK(cudaMemcpy(h.data(), c->d_counts.p, (size_t)G * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        for (int g = 0; g < G; ++g) counts[g] = (int64_t)h[g];
    }
    if (counters) *counters = c->last;
    return ABC_OK;
}

// wrapper.jl sections 2 and 3 for one batch in one call: simulate -> statistics -> errors -> acceptance, pipelined over
// sub-batches so that the device-to-host copy of one sub-batch's outputs (the error matrix is 27 KB per particle) runs
// on a second stream under the simulation of the next one.  Results are identical to abc_simulate followed by abc_score.
extern "C" int abc_simulate_score(abc_ctx_t* c, int m, int64_t n, int64_t offset, uint64_t seed, int prior_supplied,
                                  double* theta, double* stats, double eps, int layout, double* err, int64_t* counts,
                                  abc_counters_t* counters) {
    return abc_simulate_score_impl(c, m, n, offset, seed, prior_supplied, theta, stats, eps, layout, err, n, counts, counters);
}

// ---- asynchronous batches: abc_simulate_score_async enqueues one batch and returns; abc_wait completes all of them ---------
static int async_drain_set(abc_ctx* c, int l) {
    if (c->async_nb[l] == 0) return ABC_OK;
    ABC_CUDA_CHECK(cudaEventSynchronize(c->p_copied[l]));
    const unsigned long long* h = c->h_p_counters + 8 * l;
    c->last.n_particles += (uint64_t)c->async_nb[l]; c->last.n_lineages += h[0]; c->last.n_events += h[1];
    c->last.n_draws += h[2]; c->last.n_ode_steps += h[4];
    float a = 0.f, b = 0.f;
    if (cudaEventElapsedTime(&a, c->p_t0[l], c->p_t1[l]) != cudaSuccess) cudaGetLastError();
    if (cudaEventElapsedTime(&b, c->p_t1[l], c->p_t2[l]) != cudaSuccess) cudaGetLastError();
    c->async_ms_sim += a; c->async_ms_score += b;
    c->async_nb[l] = 0;
    return ABC_OK;
}

extern "C" int abc_wait(abc_ctx_t* c, int64_t* counts, abc_counters_t* counters) {
    CTX_GUARD(c);
    int rc;
    for (int k = 0; k < 2; ++k)          // oldest first
        if ((rc = async_drain_set(c, (c->async_next + k) & 1)) != ABC_OK) return rc;
    ABC_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    if (c->async_open) {
        c->last.ms_simulate = c->async_ms_sim; c->last.ms_score = c->async_ms_score;
        c->async_open = false;
    }
    if (c->has_data) {
        unsigned long long total = 0;
        ABC_CUDA_CHECK(cudaMemcpy(&total, c->d_acc_count.p, sizeof(total), cudaMemcpyDeviceToHost));
        if ((int64_t)total > c->acc_capacity) return accept_overflow(c, total);
        if (counts) {
            std::vector<unsigned long long> h((size_t)c->G);
            ABC_CUDA_CHECK(cudaMemcpy(h.data(), c->d_counts.p, (size_t)c->G * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
            for (int g = 0; g < c->G; ++g) counts[g] = (int64_t)h[g];
        }
    }
    if (counters) *counters = c->last;
    return ABC_OK;
}

// One batch, enqueued: the same work and results as abc_simulate_score.  Up to two batches may be in flight (two sets of
// device output buffers): the outputs of one travel to the host while the next one is simulated; a third call first waits
// for the oldest.  theta / stats / err belong to the library until abc_wait returns (page-locked memory from abc_host_alloc
// keeps the copies asynchronous).  A batch too large for one launch is run by the blocking path.
extern "C" int abc_simulate_score_async(abc_ctx_t* c, int m, int64_t n, int64_t offset, uint64_t seed, int prior_supplied,
                                        double* theta, double* stats, double eps, int layout, double* err) {
    CTX_GUARD(c);
    int rc = check_model(m);
    if (rc != ABC_OK) return rc;
    if (!c->has_data) { abc_set_error("abc_set_data has not been called"); return ABC_ERR_STATE; }
    if (n < 0 || (n > 0 && (!theta || !stats)) || (layout != ABC_ERR_NONE && n > 0 && !err)) {
        abc_set_error("abc_simulate_score_async: bad arguments");
        return ABC_ERR_ARG;
    }
    if (n == 0) return ABC_OK;
    const int P = abc_n_params(m), G = c->G;
    const int64_t limit = std::min<int64_t>(sim_chunk(c), layout != ABC_ERR_NONE ? std::max<int64_t>(1024, (int64_t)(2.0e9 / (8.0 * G))) : (1ll << 40));
    if (n > limit) {
        if ((rc = abc_wait(c, nullptr, nullptr)) != ABC_OK) return rc;
        return abc_simulate_score_impl(c, m, n, offset, seed, prior_supplied, theta, stats, eps, layout, err, n, nullptr, nullptr);
    }
    if (!c->async_open) {
        memset(&c->last, 0, sizeof(c->last));
        c->async_ms_sim = c->async_ms_score = 0.0;
        c->async_open = true;
    }
    const int l = c->async_next;
    if ((rc = async_drain_set(c, l)) != ABC_OK) return rc;      // the batch that used this set two calls ago
    if ((rc = c->d_p_theta[l].ensure((size_t)n * P)) != ABC_OK) return rc;
    if ((rc = c->d_p_stats[l].ensure((size_t)n * ABC_NSTATS)) != ABC_OK) return rc;
    if (layout != ABC_ERR_NONE && (rc = c->d_p_err[l].ensure((size_t)n * G)) != ABC_OK) return rc;
    if (prior_supplied)
        ABC_CUDA_CHECK(cudaMemcpyAsync(c->d_p_theta[l].p, theta, (size_t)n * P * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    ABC_CUDA_CHECK(cudaEventRecord(c->p_t0[l], c->stream));
    if ((rc = simulate_device(c, m, n, offset, seed, prior_supplied, c->d_p_theta[l].p, c->d_p_stats[l].p, nullptr, c->stream)) != ABC_OK) return rc;
    ABC_CUDA_CHECK(cudaMemcpyAsync(c->h_p_counters + 8 * l, c->d_counters.p, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    ABC_CUDA_CHECK(cudaEventRecord(c->p_t1[l], c->stream));
    if ((rc = score_device(c, c->d_p_stats[l].p, n, offset, eps, layout, c->d_p_err[l].p, c->stream)) != ABC_OK) return rc;
    ABC_CUDA_CHECK(cudaEventRecord(c->p_t2[l], c->stream));
    ABC_CUDA_CHECK(cudaEventRecord(c->p_done[l], c->stream));
    ABC_CUDA_CHECK(cudaStreamWaitEvent(c->copy_stream, c->p_done[l], 0));
    if (!prior_supplied)
        ABC_CUDA_CHECK(cudaMemcpyAsync(theta, c->d_p_theta[l].p, (size_t)n * P * sizeof(double), cudaMemcpyDeviceToHost, c->copy_stream));
    ABC_CUDA_CHECK(cudaMemcpyAsync(stats, c->d_p_stats[l].p, (size_t)n * ABC_NSTATS * sizeof(double), cudaMemcpyDeviceToHost, c->copy_stream));
    if (layout == ABC_ERR_PARTICLE_MAJOR) {
        ABC_CUDA_CHECK(cudaMemcpyAsync(err, c->d_p_err[l].p, (size_t)n * G * sizeof(double), cudaMemcpyDeviceToHost, c->copy_stream));
    } else if (layout == ABC_ERR_GENE_MAJOR) {
        ABC_CUDA_CHECK(cudaMemcpyAsync(err, c->d_p_err[l].p, (size_t)n * G * sizeof(double), cudaMemcpyDeviceToHost, c->copy_stream));
    }
    ABC_CUDA_CHECK(cudaEventRecord(c->p_copied[l], c->copy_stream));
    c->async_nb[l] = n;
    c->async_next = l ^ 1;
    return ABC_OK;
}

// gm_pitch: doubles between consecutive gene rows of a gene-major `err` (n for a stand-alone call; the whole batch when this
// context computes one shard of a multi-GPU call)
int abc_simulate_score_impl(abc_ctx* c, int m, int64_t n, int64_t offset, uint64_t seed, int prior_supplied, double* theta,
                            double* stats, double eps, int layout, double* err, int64_t gm_pitch, int64_t* counts,
                            abc_counters_t* counters) {
    CTX_GUARD(c);
    int rc = check_model(m);
    if (rc != ABC_OK) return rc;
    if (!c->has_data) { abc_set_error("abc_set_data has not been called"); return ABC_ERR_STATE; }
    if (n < 0 || (n > 0 && (!theta || !stats)) || (layout != ABC_ERR_NONE && n > 0 && !err)) {
        abc_set_error("abc_simulate_score: bad arguments");
        return ABC_ERR_ARG;
    }
    if (c->async_nb[0] != 0 || c->async_nb[1] != 0 || c->async_open) {       // asynchronous batches share the buffer sets
        if ((rc = abc_wait(c, nullptr, nullptr)) != ABC_OK) return rc;
    }
    const int P = abc_n_params(m), G = c->G;
    memset(&c->last, 0, sizeof(c->last));
    // up to four sub-batches per call, none below `simulate_score_sub_batch` particles (the SSA's longest lineages take ~25 ms
    // whatever the batch, so shallow launches waste their tail), bounded by the simulate chunk and by ~2 GB of device error
    // matrix per set
    const int64_t sub_min = std::max<int64_t>(c->simscore_sub_min, 256);
    const int64_t nsb = std::min<int64_t>(4, n / sub_min);           // equal sub-batches, no short tail
    int64_t sub = (nsb >= 2) ? (n + nsb - 1) / nsb : std::max<int64_t>(n, 1);
    sub = std::min<int64_t>(sub, sim_chunk(c));
    if (layout != ABC_ERR_NONE) sub = std::min<int64_t>(sub, std::max<int64_t>(1024, (int64_t)(2.0e9 / (8.0 * G))));
    double ms_sim = 0.0, ms_score = 0.0;       // simulate incl. prior draw and statistics; scoring
    auto drain = [&](int l, int64_t nb) -> int {          // outputs of set l are on the host; account its counters
        ABC_CUDA_CHECK(cudaEventSynchronize(c->p_copied[l]));
        const unsigned long long* h = c->h_p_counters + 8 * l;
        c->last.n_particles += (uint64_t)nb; c->last.n_lineages += h[0]; c->last.n_events += h[1];
        c->last.n_draws += h[2]; c->last.n_ode_steps += h[4];
        float a = 0.f, b = 0.f;
        if (cudaEventElapsedTime(&a, c->p_t0[l], c->p_t1[l]) != cudaSuccess) cudaGetLastError();
        if (cudaEventElapsedTime(&b, c->p_t1[l], c->p_t2[l]) != cudaSuccess) cudaGetLastError();
        ms_sim += a; ms_score += b;
        return ABC_OK;
    };
    int64_t nb_of[2] = {0, 0};
    int j = 0;
    for (int64_t b0 = 0; b0 < n; b0 += sub, ++j) {
        const int l = j & 1;
        const int64_t nb = std::min<int64_t>(sub, n - b0);
        if (j >= 2 && (rc = drain(l, nb_of[l])) != ABC_OK) return rc;
        nb_of[l] = nb;
        if ((rc = c->d_p_theta[l].ensure((size_t)nb * P)) != ABC_OK) return rc;
        if ((rc = c->d_p_stats[l].ensure((size_t)nb * ABC_NSTATS)) != ABC_OK) return rc;
        if (layout != ABC_ERR_NONE && (rc = c->d_p_err[l].ensure((size_t)nb * G)) != ABC_OK) return rc;
        if (prior_supplied)
            ABC_CUDA_CHECK(cudaMemcpyAsync(c->d_p_theta[l].p, theta + b0 * P, (size_t)nb * P * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        ABC_CUDA_CHECK(cudaEventRecord(c->p_t0[l], c->stream));
        rc = simulate_device(c, m, nb, offset + b0, seed, prior_supplied, c->d_p_theta[l].p, c->d_p_stats[l].p, nullptr, c->stream);
        if (rc != ABC_OK) return rc;
        // the device counters are reset by the next sub-batch: park this one's in page-locked memory (stream ordered)
        ABC_CUDA_CHECK(cudaMemcpyAsync(c->h_p_counters + 8 * l, c->d_counters.p, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
        ABC_CUDA_CHECK(cudaEventRecord(c->p_t1[l], c->stream));
        rc = score_device(c, c->d_p_stats[l].p, nb, offset + b0, eps, layout, c->d_p_err[l].p, c->stream);
        if (rc != ABC_OK) return rc;
        ABC_CUDA_CHECK(cudaEventRecord(c->p_t2[l], c->stream));
        ABC_CUDA_CHECK(cudaEventRecord(c->p_done[l], c->stream));
        ABC_CUDA_CHECK(cudaStreamWaitEvent(c->copy_stream, c->p_done[l], 0));
        if (!prior_supplied)
            ABC_CUDA_CHECK(cudaMemcpyAsync(theta + b0 * P, c->d_p_theta[l].p, (size_t)nb * P * sizeof(double), cudaMemcpyDeviceToHost, c->copy_stream));
        ABC_CUDA_CHECK(cudaMemcpyAsync(stats + b0 * ABC_NSTATS, c->d_p_stats[l].p, (size_t)nb * ABC_NSTATS * sizeof(double), cudaMemcpyDeviceToHost, c->copy_stream));
        if (layout == ABC_ERR_PARTICLE_MAJOR) {
            ABC_CUDA_CHECK(cudaMemcpyAsync(err + b0 * G, c->d_p_err[l].p, (size_t)nb * G * sizeof(double), cudaMemcpyDeviceToHost, c->copy_stream));
        } else if (layout == ABC_ERR_GENE_MAJOR) {
            ABC_CUDA_CHECK(cudaMemcpy2DAsync(err + b0, (size_t)gm_pitch * sizeof(double), c->d_p_err[l].p, (size_t)nb * sizeof(double),
                                             (size_t)nb * sizeof(double), (size_t)G, cudaMemcpyDeviceToHost, c->copy_stream));
        }
        ABC_CUDA_CHECK(cudaEventRecord(c->p_copied[l], c->copy_stream));
        // the next sub-batch reuses the simulate work buffers in stream order; this set's output buffers are only
        // rewritten after drain(l)
    }
    // the last (up to) two sets
    for (int k = std::max(0, j - 2); k < j; ++k)
        if ((rc = drain(k & 1, nb_of[k & 1])) != ABC_OK) return rc;
    ABC_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    c->last.ms_simulate = ms_sim; c->last.ms_score = ms_score;
    unsigned long long total = 0;
    ABC_CUDA_CHECK(cudaMemcpy(&total, c->d_acc_count.p, sizeof(total), cudaMemcpyDeviceToHost));
    if ((int64_t)total > c->acc_capacity) return accept_overflow(c, total);
    if (counts) {
        std::vector<unsigned long long> h((size_t)G);
        ABC_CUDA_CHECK(cudaMemcpy(h.data(), c->d_counts.p, (size_t)G * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        for (int g = 0; g < G; ++g) counts[g] = (int64_t)h[g];
    }
    if (counters) *counters = c->last;
    return ABC_OK;
}

extern "C" int64_t abc_accept_total(abc_ctx_t* c) {
    if (!c) return -1;
    if (cudaSetDevice(c->device) != cudaSuccess) return -1;
    unsigned long long total = 0;
    if (sync_ctx(c) != ABC_OK) return -1;
    if (cudaMemcpy(&total, c->d_acc_count.p, sizeof(total), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return (int64_t)total;
}

extern "C" int abc_accept_reset(abc_ctx_t* c) {
    CTX_GUARD(c);
    int rcs = sync_ctx(c);
    if (rcs != ABC_OK) return rcs;
    ABC_CUDA_CHECK(cudaMemset(c->d_acc_count.p, 0, sizeof(unsigned long long)));
    if (c->G > 0) ABC_CUDA_CHECK(cudaMemset(c->d_counts.p, 0, (size_t)c->G * sizeof(unsigned long long)));
    c->acc_budget = 0;
    return ABC_OK;
}

static int fetch_tuples(abc_ctx* c, std::vector<int32_t>& g, std::vector<long long>& p, std::vector<double>& e) {
    int rcs = sync_ctx(c);
    if (rcs != ABC_OK) return rcs;
    unsigned long long total = 0;
    ABC_CUDA_CHECK(cudaMemcpy(&total, c->d_acc_count.p, sizeof(total), cudaMemcpyDeviceToHost));
    if ((int64_t)total > c->acc_capacity) { abc_set_error("accepted-tuple buffer overflowed"); return ABC_ERR_NOMEM; }
    g.resize(total); p.resize(total); e.resize(total);
    if (total) {
        ABC_CUDA_CHECK(cudaMemcpy(g.data(), c->d_acc_gene.p, total * sizeof(int32_t), cudaMemcpyDeviceToHost));
        ABC_CUDA_CHECK(cudaMemcpy(p.data(), c->d_acc_particle.p, total * sizeof(long long), cudaMemcpyDeviceToHost));
        ABC_CUDA_CHECK(cudaMemcpy(e.data(), c->d_acc_err.p, total * sizeof(double), cudaMemcpyDeviceToHost));
    }
    return ABC_OK;
}

extern "C" int abc_accept_tuples(abc_ctx_t* c, int32_t* gene, int64_t* particle, double* err) {
    CTX_GUARD(c);
    std::vector<int32_t> g; std::vector<long long> p; std::vector<double> e;
    int rc = fetch_tuples(c, g, p, e);
    if (rc != ABC_OK) return rc;
    for (size_t k = 0; k < g.size(); ++k) {
        if (gene) gene[k] = g[k];
        if (particle) particle[k] = (int64_t)p[k];
        if (err) err[k] = e[k];
    }
    return ABC_OK;
}

// offsets[G+1] (host) from the per-gene counts, and -- if want_lists -- the per-gene ordered lists in c->d_as_idx / c->d_as_err:
// ascending error, ties by ascending particle index == v[sortperm(err[v])] (stable), three stable radix passes on the device
// (abc_accept.cu).  Work is enqueued on c->stream; the caller synchronises.
int abc_build_accepted_lists(abc_ctx* c, int64_t* offsets, unsigned long long* total_out, bool want_lists) {
    int rcs = sync_ctx(c);
    if (rcs != ABC_OK) return rcs;
    unsigned long long total = 0;
    ABC_CUDA_CHECK(cudaMemcpy(&total, c->d_acc_count.p, sizeof(total), cudaMemcpyDeviceToHost));
    if ((int64_t)total > c->acc_capacity) { abc_set_error("accepted-tuple buffer overflowed"); return ABC_ERR_NOMEM; }
    std::vector<unsigned long long> h((size_t)c->G);
    ABC_CUDA_CHECK(cudaMemcpy(h.data(), c->d_counts.p, (size_t)c->G * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    offsets[0] = 0;
    for (int g = 0; g < c->G; ++g) offsets[g + 1] = offsets[g] + (int64_t)h[g];
    if ((unsigned long long)offsets[c->G] != total) {
        abc_set_error("per-gene counts (%lld) and stored tuples (%llu) disagree", (long long)offsets[c->G], total);
        return ABC_ERR_STATE;
    }
    *total_out = total;
    if (total == 0 || !want_lists) return ABC_OK;
    int rc = ABC_OK;
    const size_t tmp_bytes = abc_accept_sort_temp_bytes((size_t)total);
    for (int l = 0; l < 2; ++l) {
        if ((rc = c->d_as_k64[l].ensure((size_t)total)) != ABC_OK) return rc;
        if ((rc = c->d_as_k32[l].ensure((size_t)total)) != ABC_OK) return rc;
        if ((rc = c->d_as_perm[l].ensure((size_t)total)) != ABC_OK) return rc;
    }
    if ((rc = c->d_as_idx.ensure((size_t)total)) != ABC_OK) return rc;
    if ((rc = c->d_as_err.ensure((size_t)total)) != ABC_OK) return rc;
    if ((rc = c->d_as_tmp.ensure(tmp_bytes)) != ABC_OK) return rc;
    unsigned long long* pk64[2] = {c->d_as_k64[0].p, c->d_as_k64[1].p};
    uint32_t* pk32[2] = {c->d_as_k32[0].p, c->d_as_k32[1].p};
    uint32_t* pperm[2] = {c->d_as_perm[0].p, c->d_as_perm[1].p};
    int nl = 0;
    rc = abc_launch_accept_sort(c->d_acc_gene.p, c->d_acc_particle.p, c->d_acc_err.p, (size_t)total, c->G, pk64, pk32, pperm,
                                c->d_as_tmp.p, tmp_bytes, c->d_as_idx.p, c->d_as_err.p, &nl, c->stream);
    c->launches += nl;
    return rc;
}

extern "C" int abc_accept_fetch(abc_ctx_t* c, int64_t* offsets, int64_t* idx, double* errs) {
    CTX_GUARD(c);
    if (!c->has_data || !offsets) { abc_set_error("abc_accept_fetch: bad state/arguments"); return ABC_ERR_ARG; }
    unsigned long long total = 0;
    int rc = abc_build_accepted_lists(c, offsets, &total, idx || errs);
    if (rc != ABC_OK) return rc;
    if (total == 0 || (!idx && !errs)) return ABC_OK;
    if (idx) ABC_CUDA_CHECK(cudaMemcpyAsync(idx, c->d_as_idx.p, (size_t)total * sizeof(long long), cudaMemcpyDeviceToHost, c->stream));
    if (errs) ABC_CUDA_CHECK(cudaMemcpyAsync(errs, c->d_as_err.p, (size_t)total * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    ABC_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    return ABC_OK;
}

// SURVEY 8f-3 (posterior_kinetics.jl:10-33): MAP, mean and the (1-q, q) quantiles of the accepted parameter rows per gene
extern "C" int abc_posterior_summary(abc_ctx_t* c, const double* theta, int64_t n, int32_t P, int64_t particle_offset, double q,
                                     double* map, double* mean, double* lo, double* hi, int64_t* n_acc) {
    CTX_GUARD(c);
    if (!c->has_data) { abc_set_error("abc_set_data has not been called"); return ABC_ERR_STATE; }
    if (!theta || n <= 0 || P < 1 || P > ABC_MAXP || !(q >= 0.0 && q <= 1.0)) { abc_set_error("abc_posterior_summary: bad arguments"); return ABC_ERR_ARG; }
    const int G = c->G;
    std::vector<int64_t> offsets((size_t)G + 1);
    unsigned long long total = 0;
    int rc = abc_build_accepted_lists(c, offsets.data(), &total, true);
    if (rc != ABC_OK) return rc;
    if (n_acc) for (int g = 0; g < G; ++g) n_acc[g] = offsets[g + 1] - offsets[g];
    DevBuf<long long> d_off;
    DevBuf<double> d_theta, d_vals[2], d_out[4];
    DevBuf<unsigned char> d_tmp;
    DevBuf<int> d_bad;
    auto release = [&]() {
        d_off.release(); d_theta.release(); d_vals[0].release(); d_vals[1].release(); d_tmp.release(); d_bad.release();
        for (int k = 0; k < 4; ++k) d_out[k].release();
    };
    const size_t tmp_bytes = abc_posterior_temp_bytes((size_t)std::max<unsigned long long>(total, 1), G);
    rc = d_off.ensure((size_t)G + 1);
    if (rc == ABC_OK) rc = d_theta.ensure((size_t)n * P);
    if (rc == ABC_OK) rc = d_vals[0].ensure((size_t)std::max<unsigned long long>(total, 1));
    if (rc == ABC_OK) rc = d_vals[1].ensure((size_t)std::max<unsigned long long>(total, 1));
    if (rc == ABC_OK) rc = d_tmp.ensure(std::max<size_t>(tmp_bytes, 16));
    if (rc == ABC_OK) rc = d_bad.ensure(1);
    double* host_out[4] = {map, mean, lo, hi};
    for (int k = 0; k < 4 && rc == ABC_OK; ++k) if (host_out[k]) rc = d_out[k].ensure((size_t)G * P);
    if (rc != ABC_OK) { release(); return rc; }
    cudaError_t e = cudaMemcpyAsync(d_off.p, offsets.data(), ((size_t)G + 1) * sizeof(long long), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_theta.p, theta, (size_t)n * P * sizeof(double), cudaMemcpyHostToDevice, c->stream);
    int nl = 0, bad = 0;
    if (e == cudaSuccess) {
        double* pv[2] = {d_vals[0].p, d_vals[1].p};
        rc = abc_launch_posterior(c->d_as_idx.p, d_off.p, (size_t)total, G, d_theta.p, n, P, particle_offset, q, pv, d_tmp.p,
                                  std::max<size_t>(tmp_bytes, 16), d_bad.p, d_out[0].p, d_out[1].p, d_out[2].p, d_out[3].p, &nl, c->stream);
        c->launches += nl;
    }
    for (int k = 0; k < 4 && rc == ABC_OK && e == cudaSuccess; ++k)
        if (host_out[k]) e = cudaMemcpyAsync(host_out[k], d_out[k].p, (size_t)G * P * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
    if (rc == ABC_OK && e == cudaSuccess) e = cudaMemcpyAsync(&bad, d_bad.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    release();
    if (rc != ABC_OK) return rc;
    if (e != cudaSuccess) { abc_set_error("CUDA error in abc_posterior_summary: %s", cudaGetErrorString(e)); cudaGetLastError(); return ABC_ERR_CUDA; }
    if (bad) { abc_set_error("abc_posterior_summary: an accepted particle index lies outside [particle_offset+1, particle_offset+n]"); return ABC_ERR_ARG; }
    return ABC_OK;
}

// SURVEY 8f-4 (data_summary_statistics.jl:183-194 get_summary_stats for every gene)
extern "C" int abc_data_summary_stats(abc_ctx_t* c, const double* u, const double* l, int32_t n_cells, int32_t n_genes,
                                      const int32_t* age, const int32_t* experiment, const int32_t* cond_vec,
                                      const int32_t* pulse_idx, int32_t n_pulse, const int32_t* chase_idx, int32_t n_chase,
                                      const double* age_id_dist, int32_t n_bootstraps, uint64_t seed, double* d, double* se) {
    CTX_GUARD(c);
    return abc_run_data_summary_stats(u, l, n_cells, n_genes, age, experiment, cond_vec, pulse_idx, n_pulse, chase_idx, n_chase,
                                      age_id_dist, n_bootstraps, seed, d, se, &c->launches, c->stream);
}

// SURVEY 8f-2 (model_probs.jl:1-54): per-gene model probabilities = acceptance-count ratios + bootstrap percentile bounds
extern "C" int abc_model_probs(abc_ctx_t* c, const int64_t* counts, int32_t K, int32_t G, int32_t n_bootstraps, double alpha,
                               uint64_t seed, double* prob, double* lb, double* ub) {
    CTX_GUARD(c);
    if (!counts || !prob || !lb || !ub || K < 1 || G < 1 || !(alpha >= 0.0 && alpha <= 1.0)) { abc_set_error("abc_model_probs: bad arguments"); return ABC_ERR_ARG; }
    for (int g = 0; g < G; ++g) {
        unsigned long long tot = 0;
        for (int k = 0; k < K; ++k) {
            if (counts[(size_t)k * G + g] < 0) { abc_set_error("abc_model_probs: negative count"); return ABC_ERR_ARG; }
            tot += (unsigned long long)counts[(size_t)k * G + g];
        }
        if (tot >= 0xFFFFFFFFull) { abc_set_error("abc_model_probs: more than 2^32 accepted particles for gene %d", g + 1); return ABC_ERR_ARG; }
    }
    DevBuf<long long> d_counts;
    DevBuf<double> d_stats, d_out;
    int rc = d_counts.ensure((size_t)K * G);
    if (rc == ABC_OK) rc = d_stats.ensure((size_t)G * K * (size_t)std::max(n_bootstraps, 1));
    if (rc == ABC_OK) rc = d_out.ensure((size_t)3 * G * K);
    if (rc != ABC_OK) return rc;
    ABC_CUDA_CHECK(cudaMemcpyAsync(d_counts.p, counts, (size_t)K * G * sizeof(long long), cudaMemcpyHostToDevice, c->stream));
    int nl = 0;
    const size_t gk = (size_t)G * K;
    rc = abc_launch_model_probs(d_counts.p, K, G, n_bootstraps, alpha, seed, d_stats.p, d_out.p, d_out.p + gk, d_out.p + 2 * gk, &nl, c->stream);
    if (rc != ABC_OK) return rc;
    c->launches += nl;
    ABC_CUDA_CHECK(cudaMemcpyAsync(prob, d_out.p, gk * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    ABC_CUDA_CHECK(cudaMemcpyAsync(lb, d_out.p + gk, gk * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    ABC_CUDA_CHECK(cudaMemcpyAsync(ub, d_out.p + 2 * gk, gk * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    ABC_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    return ABC_OK;
}

// ------------------------------------------------------------------------------------------------
extern "C" int abc_simulate_dev(abc_ctx_t* c, int m, int64_t n, int64_t offset, uint64_t seed, int prior_supplied,
                                double* d_theta, double* d_stats, void* stream) {
    CTX_GUARD(c);
    int rc = check_model(m);
    if (rc != ABC_OK) return rc;
    if (n <= 0 || !d_theta || !d_stats) { abc_set_error("abc_simulate_dev: bad arguments"); return ABC_ERR_ARG; }
    cudaStream_t st = (cudaStream_t)stream;   // NULL = the legacy default stream (torch's default stream)
    rc = simulate_device(c, m, n, offset, seed, prior_supplied, d_theta, d_stats, nullptr, st);
    if (rc != ABC_OK) return rc;
    return mark_user_stream(c, st);
}

extern "C" int abc_score_dev(abc_ctx_t* c, const double* d_stats, int64_t n, int64_t offset, double eps, int layout,
                             double* d_err, void* stream) {
    CTX_GUARD(c);
    if (n <= 0 || !d_stats || (layout != ABC_ERR_NONE && !d_err)) { abc_set_error("abc_score_dev: bad arguments"); return ABC_ERR_ARG; }
    cudaStream_t st = (cudaStream_t)stream;   // NULL = the legacy default stream (torch's default stream)
    int rc = score_device(c, d_stats, n, offset, eps, layout, d_err, st);
    if (rc != ABC_OK) return rc;
    return mark_user_stream(c, st);
}

extern "C" int abc_counts_dev(abc_ctx_t* c, int64_t* d_counts, void* stream) {
    CTX_GUARD(c);
    if (!c->has_data || !d_counts) { abc_set_error("abc_counts_dev: bad state/arguments"); return ABC_ERR_ARG; }
    cudaStream_t st = (cudaStream_t)stream;   // NULL = the legacy default stream (torch's default stream)
    if (c->user_pending) ABC_CUDA_CHECK(cudaStreamWaitEvent(st, c->ev_user, 0));   // scoring enqueued on another caller stream
    ABC_CUDA_CHECK(cudaMemcpyAsync(d_counts, c->d_counts.p, (size_t)c->G * sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
    return mark_user_stream(c, st);
}

extern "C" int abc_accept_tuples_dev(abc_ctx_t* c, int32_t* d_gene, int64_t* d_particle, double* d_err,
                                     int64_t capacity, void* stream) {
    CTX_GUARD(c);
    if (!d_gene || !d_particle || !d_err || capacity < 0) { abc_set_error("abc_accept_tuples_dev: bad arguments"); return ABC_ERR_ARG; }
    cudaStream_t st = (cudaStream_t)stream;   // NULL = the legacy default stream (torch's default stream)
    int rcs = sync_ctx(c);                    // every scoring call so far, whichever stream it ran on
    if (rcs != ABC_OK) return rcs;
    ABC_CUDA_CHECK(cudaStreamSynchronize(st));
    unsigned long long total = 0;
    ABC_CUDA_CHECK(cudaMemcpy(&total, c->d_acc_count.p, sizeof(total), cudaMemcpyDeviceToHost));
    if ((int64_t)total > c->acc_capacity) { abc_set_error("accepted-tuple buffer overflowed"); return ABC_ERR_NOMEM; }
    if ((int64_t)total > capacity) {
        abc_set_error("abc_accept_tuples_dev: %llu accepted tuples do not fit the caller's capacity %lld", total, (long long)capacity);
        return ABC_ERR_ARG;
    }
    size_t k = (size_t)total;
    if (k) {
        ABC_CUDA_CHECK(cudaMemcpyAsync(d_gene, c->d_acc_gene.p, k * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
        ABC_CUDA_CHECK(cudaMemcpyAsync(d_particle, c->d_acc_particle.p, k * sizeof(long long), cudaMemcpyDeviceToDevice, st));
        ABC_CUDA_CHECK(cudaMemcpyAsync(d_err, c->d_acc_err.p, k * sizeof(double), cudaMemcpyDeviceToDevice, st));
    }
    return ABC_OK;
}

extern "C" int abc_set_option(abc_ctx_t* c, const char* name, int64_t value) {
    CTX_GUARD(c);
    if (!name) { abc_set_error("abc_set_option: name is NULL"); return ABC_ERR_ARG; }
    if (strcmp(name, "score_reference_kernel") == 0) { c->force_reference_score = value ? 1 : 0; return ABC_OK; }
    if (strcmp(name, "score_tile_kernel") == 0) { c->score_tile_kernel = value ? 1 : 0; return ABC_OK; }
    if (strcmp(name, "score_overlap") == 0) { c->score_overlap = value ? 1 : 0; return ABC_OK; }
    if (strcmp(name, "score_mma_filter") == 0) { c->score_mma_filter = (value == 2) ? 2 : (value ? 1 : 0); return ABC_OK; }
    if (strcmp(name, "score_sub_batches") == 0) { c->score_sub_batches = (int)std::min<int64_t>(std::max<int64_t>(value, 0), 64); return ABC_OK; }
    if (strcmp(name, "accept_capacity") == 0) { c->acc_min_capacity = value > 0 ? value : 0; return ABC_OK; }
    if (strcmp(name, "stats_sample_guards") == 0) { c->stats_guards = value < 0 ? -1 : (value ? 1 : 0); return ABC_OK; }
    if (strcmp(name, "simulate_score_sub_batch") == 0) { c->simscore_sub_min = value > 0 ? value : 8192; return ABC_OK; }
    if (strcmp(name, "ssa_adaptive_burnin") == 0) { c->ssa_adaptive = value <= 0 ? 0 : (value == 1 ? 1 : 2); return ABC_OK; }
    if (strcmp(name, "ssa_hybrid_burnin") == 0) { c->ssa_hybrid = value <= 0 ? 0 : (value == 1 ? 1 : 2); return ABC_OK; }
    abc_set_error("abc_set_option: unknown option '%s'", name);
    return ABC_ERR_ARG;
}

extern "C" int abc_counters(abc_ctx_t* c, abc_counters_t* out) {
    CTX_GUARD(c);
    if (!out) { abc_set_error("abc_counters: out is NULL"); return ABC_ERR_ARG; }
    int rc = sync_ctx(c);          // this context's own work only (not other streams of the process, e.g. a collective in flight)
    if (rc != ABC_OK) return rc;
    rc = read_counters(c, 0, nullptr, false);
    if (rc != ABC_OK) return rc;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, c->ev[3], c->ev[4]) == cudaSuccess) c->last.ms_score = ms; else cudaGetLastError();
    *out = c->last;
    return ABC_OK;
}

extern "C" int64_t abc_launch_count(abc_ctx_t* c) { return c ? c->launches : -1; }
