/*
 * abc_b200.h -- C ABI of libabcb200.so: the B200-native ABC simulate -> summarise -> score -> accept
 * hot path of pthomaslab/abc_inference_transcription.
 *
 * The reference has no FFI; its boundary is script level (Julia globals m, n_trials, submit and
 * `include`d scripts).  Each entry point below names the reference code it replaces.  A Julia host
 * binds these with  ccall((:abc_simulate, "libabcb200"), Cint, (...), ...)  -- see INTEGRATION.md.
 *
 * Conventions
 *   - every function returns 0 on success, a negative abc_status_t otherwise; the message is in
 *     abc_last_error() (thread local).  No C++ exception crosses this boundary.
 *   - all host buffers are caller owned, column-major Julia arrays map as documented per argument,
 *     and must stay alive for the duration of the (blocking) call.
 *   - the library owns all device memory.  There is NO CPU fallback: without a CUDA device every
 *     compute entry point fails with ABC_ERR_CUDA.
 *   - particle indices in outputs are 1-based (Julia), everything else is 0-based.
 *   - calls on one context are not re-entrant; use one context per GPU / per host thread.
 */
#ifndef ABC_B200_H
#define ABC_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ABC_NAGE   5    /* n_age_clusters            scripts/load_process_data.jl:68 */
#define ABC_NCOND  11   /* labelling conditions      scripts/abc_simulation.jl:65    */
#define ABC_NSTATS 53   /* 4*5 + 3*11                scripts/compute_errors.jl:51    */
#define ABC_NMODELS 5   /* const, const_const, kon, alpha, gamma   abc_simulation.jl:82 */
#define ABC_MAXP   9    /* parameters per particle: 5 (m=1,2) or 9 (m=3,4,5) */

typedef enum {
    ABC_OK = 0,
    ABC_ERR_ARG = -1,      /* bad argument (NULL, out of range model index, ...) */
    ABC_ERR_CUDA = -2,     /* CUDA runtime error or no device */
    ABC_ERR_STATE = -3,    /* design / data statistics not set yet */
    ABC_ERR_NOMEM = -4
} abc_status_t;

typedef struct abc_ctx abc_ctx_t;

/* simulator selection for abc_simulate */
#define ABC_SIM_SSA 0      /* Gillespie direct-method SSA (the north-star path) */
#define ABC_SIM_ODE 1      /* moment ODEs, what scripts/model.jl actually integrates (model.jl:74-96) */

/* Experimental design: the globals scripts/abc_simulation.jl:65-79 builds plus the data-derived
 * globals of scripts/load_process_data.jl:59-83 (age, pulse_idx, chase_idx, age_id_distribution)
 * and data/capture_efficiencies.txt, which abc_sim() receives as arguments (abc_simulation.jl:13-14). */
typedef struct {
    double  cycle;                       /* 20.0                                abc_simulation.jl:72 */
    double  t0;                          /* -3*cycle (moment-ODE path only)     abc_simulation.jl:73 */
    double  agevec[ABC_NAGE];            /* tau_ .* cycle                       abc_simulation.jl:74 */
    double  pulse[ABC_NCOND];            /* condition_id[:,1]                   abc_simulation.jl:65 */
    double  chase[ABC_NCOND];            /* condition_id[:,2] */
    double  age_dist[ABC_NAGE * ABC_NCOND]; /* age_id_distribution, Julia 5x11 column-major, used AS GIVEN */
    double  iv[9];                       /* ODE path initial moments            abc_simulation.jl:70-71 */
    int32_t downsampling;                /* 1: apply capture efficiencies       abc_simulation.jl:79 */
    int32_t n_cells;                     /* SSA: cells per (condition, age) read-out (>= 2) */
    int32_t n_pre_cycles;                /* SSA: complete cell cycles simulated before the read-out cycle (<= 13; the maximum
                                          * when ssa_adaptive_burnin = 1, see abc_set_option) */
    int32_t sim_kind;                    /* ABC_SIM_SSA or ABC_SIM_ODE */
    /* capture efficiencies betas[pulse_idx], age[pulse_idx] and betas[chase_idx], age[chase_idx]
     * (abc_simulation.jl:24,36); cluster ids are 1..5 */
    const double*  betas_pulse;
    const int32_t* cluster_pulse;
    int32_t        n_pulse;
    const double*  betas_chase;
    const int32_t* cluster_chase;
    int32_t        n_chase;
    /* ODE path integrator tolerances (CVODE defaults of the reference: 1e-3 / 1e-6) */
    double  ode_rtol, ode_atol;
} abc_design_t;

/* run counters (metrics row) */
typedef struct {
    uint64_t n_particles;
    uint64_t n_lineages;     /* SSA cell lineages simulated */
    uint64_t n_events;       /* SSA reaction events fired */
    uint64_t n_draws;        /* SSA event draws (events + discarded draws at breakpoints) */
    uint64_t n_ode_steps;    /* ODE path: accepted integrator steps */
    double   ms_simulate;    /* device time of the simulate kernels (CUDA events) */
    double   ms_stats;
    double   ms_score;
} abc_counters_t;

/* ---- lifetime ---------------------------------------------------------------------------- */
int  abc_create(int device, abc_ctx_t** ctx);
int  abc_destroy(abc_ctx_t* ctx);
const char* abc_last_error(void);
int  abc_version(void);
int  abc_device_count(void);
/* number of parameters of model m in 1..5: length(vary_map) flattened; model.jl:30-43 */
int  abc_n_params(int m);
/* the reference's model_name table, abc_simulation.jl:82 */
const char* abc_model_name(int m);

/* page-locked host memory for the large outputs (error matrix, statistics): copies from the device into such
 * buffers run at full PCIe rate and asynchronously.  A Julia host wraps the pointer with unsafe_wrap(Array, ...).
 * Any ordinary (pageable) host array is accepted by every entry point as well. */
int  abc_host_alloc(size_t bytes, void** ptr);
int  abc_host_free(void* ptr);

/* ---- configuration ----------------------------------------------------------------------- */
/* replaces the globals consumed by abc_sim(), abc_simulation.jl:13-14, 65-79 */
int  abc_set_design(abc_ctx_t* ctx, const abc_design_t* design);
/* data-side summary statistics and bootstrap SEs: the 14 matrices compute_trunc_errors receives
 * (compute_errors.jl:45-49), concatenated per gene in the order pulse_mean[5], pulse_ff[5],
 * chase_mean[5], chase_ff[5], ratio[11], mean_corr[11], corr_mean[11].
 * d, se: Julia 53 x G column-major (gene contiguous). */
int  abc_set_data(abc_ctx_t* ctx, const double* d, const double* se, int32_t n_genes);

/* ---- P1: prior draws, fix_params(vary_map, N)  abc_simulation.jl:3-11 ---------------------- */
/* theta: Julia P x n column-major (one particle's P values contiguous), log10 units, column order
 * [kon..., koff, alpha..., gamma..., lambda].  Counter-based: draw i depends only on
 * (seed, m, particle_offset + i). */
int  abc_fix_params(abc_ctx_t* ctx, int m, int64_t n, int64_t particle_offset, uint64_t seed, double* theta);

/* ---- M1-M10 + S1: abc_sim() over a batch  abc_simulation.jl:13-46, 88-97 -------------------- */
/* prior_supplied == 0: theta is an output (drawn as abc_fix_params would); != 0: theta is an input.
 * stats: Julia 53 x n column-major, order as in abc_set_data.  counters may be NULL. */
int  abc_simulate(abc_ctx_t* ctx, int m, int64_t n_trials, int64_t particle_offset, uint64_t seed,
                  int prior_supplied, double* theta, double* stats, abc_counters_t* counters);
/* per (condition, age) moments behind the statistics (after optional downsampling):
 * moments: Julia 5 x 5 x 11 x n column-major = [particle][cond][age][mean_u,mean_l,var_u,cov_ul,var_l];
 * with downsampling == 0 and sim_kind == ABC_SIM_ODE this is run_part_sim(), recover_statistics.jl:1-11 */
int  abc_simulate_moments(abc_ctx_t* ctx, int m, int64_t n, int64_t particle_offset, uint64_t seed,
                          const double* theta, double* moments, abc_counters_t* counters);
/* SSA debug/parity hook: the per-cell counts of one read-out.  counts: 4 x n_cells uint32
 * (U, L before thinning, U', L' after), for particle_index (global), condition j, age a (0-based). */
int  abc_ssa_cells(abc_ctx_t* ctx, int m, const double* theta, int64_t particle_index, uint64_t seed,
                   int cond, int age, int exact_math, uint32_t* counts);

/* SSA parity hook (ssa_hybrid_burnin = 2): where the lineages of each of the 55 read-outs of one particle start, in hours
 * relative to the start of the read-out cycle (starts[cond*5 + age]; <= 0: before the read-out cycle), and the expected
 * number of switch draws of the particle (nullable).  Transcripts born before the start are not simulated: their expected
 * share of either Poisson mean at the read-out is below 2^-n_pre_cycles (option ssa_adaptive_burnin). */
int  abc_ssa_window(abc_ctx_t* ctx, int m, const double* theta, float* starts, double* expected_draws);

/* ---- S1 alone: the 53 statistics from moments  abc_simulation.jl:23-46 --------------------- */
int  abc_summary_stats(abc_ctx_t* ctx, const double* moments, int64_t n, double* stats);

/* ---- E2/E3 + A1: compute_trunc_errors and the eps-acceptance  compute_errors.jl:30-70,
 *      accepted_particles.jl:10-32 ----------------------------------------------------------- */
#define ABC_ERR_NONE         0  /* do not materialise the error matrix (fused acceptance only) */
#define ABC_ERR_GENE_MAJOR   1  /* err[g*n + i]: one column per gene like the .jdf (process_error_files.jl:3-7) */
#define ABC_ERR_PARTICLE_MAJOR 2 /* err[i*G + g]: one row per particle like error_<model>.txt (compute_errors.jl:66-68) */
/* stats: 53 x n column-major.  eps: acceptance threshold (4.8, accepted_particles.jl:10).
 * err: NULL or n*G doubles in err_layout.  counts: NULL or G int64 (accepted per gene).
 * The accepted set is kept in the context for abc_accept_fetch. */
int  abc_score(abc_ctx_t* ctx, const double* stats, int64_t n, int64_t particle_offset, double eps,
               int err_layout, double* err, int64_t* counts, abc_counters_t* counters);
/* Diagnostics of the tensor-core filter (option "score_mma_filter": a TF32 tcgen05 GEMM decides which pairs can be <= 10 and
 * reach the FP64 stage; results are identical either way).  abc_score_mma_columns: padded gene columns of the GEMM (multiple of
 * 64).  abc_score_mma_debug: V[i][column] for DEVICE-resident statistics (53 x n column-major) into d_out, a device array of
 * ceil(n/128)*128 rows x columns floats; V < 0 <=> the pair is queued.  gene_of_column: NULL or columns int32 (host), the
 * gene index of every column, -1 for padding.  No reference counterpart. */
int  abc_score_mma_columns(abc_ctx_t* ctx);
int  abc_score_mma_debug(abc_ctx_t* ctx, const double* d_stats, int64_t n, float* d_out, int32_t* gene_of_column);
/* wrapper.jl:59-66 followed by wrapper.jl:72-78 for one batch: abc_simulate and abc_score in one call (same arguments,
 * same results).  Batches of >= 16384 particles are pipelined in sub-batches: the device-to-host copy of theta, stats and
 * the error matrix of one sub-batch runs under the simulation of the next one (page-locked outputs: abc_host_alloc). */
int  abc_simulate_score(abc_ctx_t* ctx, int m, int64_t n_trials, int64_t particle_offset, uint64_t seed,
                        int prior_supplied, double* theta, double* stats, double eps, int err_layout, double* err,
                        int64_t* counts, abc_counters_t* counters);
/* The same batch, enqueued: abc_simulate_score_async returns as soon as the work is queued, abc_wait returns when every batch
 * in flight is complete (outputs on the host; counts = running per-gene counts since abc_accept_reset, counters = sums over the
 * batches since the previous abc_wait; both nullable).  Up to two batches are in flight: while the 27 KB per particle of one
 * travel to the host, the next one is simulated -- a host that loops over the five models (wrapper.jl:59-81) pays for the
 * copies of the last batch only.  theta / stats / err must stay untouched until abc_wait; use abc_host_alloc memory (pageable
 * memory makes the copies synchronous).  Results are identical to abc_simulate_score. */
int  abc_simulate_score_async(abc_ctx_t* ctx, int m, int64_t n_trials, int64_t particle_offset, uint64_t seed,
                              int prior_supplied, double* theta, double* stats, double eps, int err_layout, double* err);
int  abc_wait(abc_ctx_t* ctx, int64_t* counts, abc_counters_t* counters);
/* number of accepted (gene, particle) pairs of the last abc_score call(s) since abc_accept_reset */
int64_t abc_accept_total(abc_ctx_t* ctx);
int  abc_accept_reset(abc_ctx_t* ctx);
/* CSR of accepted particles: offsets[G+1]; idx[total] 1-based global particle indices sorted per gene
 * by (error ascending, index ascending) == v[sortperm(err[v])], accepted_particles.jl:20-24;
 * errs (nullable) the matching error values.  A gene with offsets[g+1]==offsets[g] is the
 * reference's "0" line (accepted_particles.jl:27-29). */
int  abc_accept_fetch(abc_ctx_t* ctx, int64_t* offsets, int64_t* idx, double* errs);
/* ---- next row: get_posterior_estimate / get_posterior_ci, posterior_kinetics.jl:10-33 -----------------
 * Over the accepted lists of the abc_score calls since the last abc_accept_reset, for every gene g:
 *   map[g][P]  = theta of the first accepted index = smallest error            (posterior_kinetics.jl:14)
 *   mean[g][P] = mean of the accepted theta rows, summed in list order          (posterior_kinetics.jl:18-22)
 *   lo[g][P], hi[g][P] = quantile(1-q), quantile(q) with Julia's default definition (Statistics.jl, alpha = beta = 1):
 *                        aleph = n p + (1-p), j = clamp(trunc(aleph), 1, n-1), v[j] + clamp(aleph-j, 0, 1)(v[j+1] - v[j])
 *                        on the sorted values  (posterior_kinetics.jl:26-33)
 * theta: host, n x P row-major (Julia P x n): the parameter sets of the particles with global 1-based indices
 * particle_offset+1 .. particle_offset+n (rows of sets_<model>.txt).  n_acc[g]: accepted particles; genes without any get NaN
 * rows.  Any output may be NULL.  Lists are ordered, values gathered, sorted per gene and reduced on the device. */
int  abc_posterior_summary(abc_ctx_t* ctx, const double* theta, int64_t n, int32_t P, int64_t particle_offset, double q,
                           double* map, double* mean, double* lo, double* hi, int64_t* n_acc);
/* ---- next row: get_model_probs + the case split around it, model_probs.jl:1-54,
 *      constant_model_probs.jl:1-28, non_constant_model_probs.jl:1-30 ------------------------------------------------
 * counts: host, K x n_genes row-major: accepted particles of hypothesis k (a model, or a pooled group of models) for gene g
 * -- the per-gene counts abc_score / abc_simulate_score return.  Per gene: no hypothesis accepted -> zeros; exactly one ->
 * probability and both bounds 1 for it; otherwise prob = l / sum(l) and the (1 - alpha, alpha) quantiles (Julia's default
 * quantile) over n_bootstraps resamplings of the sum(l) model labels with replacement (Philox, counter = (block, gene,
 * bootstrap), key = seed: reproducible).  prob, lb, ub: host, n_genes x K row-major.  K <= 8, n_bootstraps <= 256. */
int  abc_model_probs(abc_ctx_t* ctx, const int64_t* counts, int32_t K, int32_t n_genes, int32_t n_bootstraps, double alpha,
                     uint64_t seed, double* prob, double* lb, double* ub);
/* unsorted accepted tuples (for multi-GPU gathers): gene int32, particle int64 (1-based global), err double */
int  abc_accept_tuples(abc_ctx_t* ctx, int32_t* gene, int64_t* particle, double* err);

/* ---- device-resident variants (inputs/outputs are device pointers on the context's device;
 *      stream is a cudaStream_t passed as void*, NULL = the legacy default stream; asynchronous) ---------- */
int  abc_simulate_dev(abc_ctx_t* ctx, int m, int64_t n, int64_t particle_offset, uint64_t seed,
                      int prior_supplied, double* d_theta, double* d_stats, void* stream);
int  abc_score_dev(abc_ctx_t* ctx, const double* d_stats, int64_t n, int64_t particle_offset, double eps,
                   int err_layout, double* d_err, void* stream);
/* device-side exports for multi-GPU gathers (torch.distributed / NCCL work on the caller's tensors):
 * copy the per-gene accepted counts (G int64) / the unsorted accepted tuples into caller device
 * buffers on `stream`.  abc_accept_tuples_dev fails with ABC_ERR_ARG when the
 * accepted tuples do not fit `capacity` (nothing is truncated silently).  All accept_* entry points wait for the *_dev work
 * enqueued on caller streams (an event recorded after every *_dev call). */
int  abc_counts_dev(abc_ctx_t* ctx, int64_t* d_counts, void* stream);
int  abc_accept_tuples_dev(abc_ctx_t* ctx, int32_t* d_gene, int64_t* d_particle, double* d_err, int64_t capacity,
                           void* stream);
/* tuning / diagnostics switches.  "score_reference_kernel" = 1: score with the plain FP64 kernel (every pair
 * evaluated in full) instead of the three-stage kernel; results are bit-identical either way.
 * "score_mma_filter" = 1 or 2 (default 0): for err_layout != ABC_ERR_NONE, decide which pairs can be <= 10 with a TF32 GEMM on
 * the tensor cores (tcgen05) instead of the FP32 tile filter; 1 hands sign-bit words to the FP64 stage, 2 writes its queue;
 * results are bit-identical in all three cases.
 * "accept_capacity" = N: reserve room for at least N accepted tuples (20 bytes each); by default the buffer grows by
 * 2 % of the pairs of every abc_score call since the last abc_accept_reset (a call whose acceptance rate exceeds
 * that fails with ABC_ERR_NOMEM instead of dropping tuples).
 * "stats_sample_guards" = -1 (default): when the moments come from SSA cells, degenerate samples follow the
 * reference's data-side conventions (ratio = 0 without counts, correlations = 0 when a total variance is 0,
 * scripts/data_summary_statistics.jl:64-71,138-147); the ODE path follows abc_simulation.jl:23-46 verbatim.  0 / 1 force.
 * "ssa_hybrid_burnin" = 0: run the full six-channel direct method from the first simulated cycle; 1: before the
 * label window opens simulate only the gene switch (Gillespie on the telegraph process) and draw
 * U ~ Poisson(Lam | gene path) at the window start, then run the six-channel direct method to the read-out;
 * 2 (default): simulate the gene switch to the read-out and draw U ~ Poisson(Lam_U | gene path),
 * L ~ Poisson(Lam_L | gene path) there.  All three sample the same law of (g, U, L) at the read-out;
 * the exact_math variant of abc_ssa_cells always uses 0.
 * "ssa_adaptive_burnin": n_pre_cycles is the maximum burn-in and 2^-n_pre_cycles its bias bound (dilution halves the
 * memory of the initial condition every cycle).  2 (default, mode 2): the lineages of a (particle, read-out) start at the
 * latest time s0 for which the transcripts born before s0 contribute, in expectation, less than 2^-n_pre_cycles of the
 * unlabelled AND of the labelled Poisson mean at the read-out -- evaluated from the exact mean contribution of every piece
 * of the rate schedule (decay, dilution, label window), see abc_ssa_window.  1 (modes 1 and 2): whole cycles per particle
 * by the worst-case rule k (1 + log2(e) sum gamma_s cycle/5) >= n_pre_cycles.  0 = always n_pre_cycles cycles. */
int  abc_set_option(abc_ctx_t* ctx, const char* name, int64_t value);
/* device counters of the last *_dev launches (synchronises the stream) */
int  abc_counters(abc_ctx_t* ctx, abc_counters_t* counters);
/* how many kernels this library has launched on the context since creation */
int64_t abc_launch_count(abc_ctx_t* ctx);

/* ---- data side: get_summary_stats of scripts/data_summary_statistics.jl:183-194 for every gene -----------------
 * The 53 data statistics d and their bootstrap standard errors se -- the inputs of abc_set_data -- from the per-cell UMI counts.
 * u, l: host, Julia n_cells x n_genes column-major (one gene's cells contiguous), integer-valued (gene_selection.jl:36-38);
 * age[n_cells]: age cluster 1..5 (load_process_data.jl:68-69); experiment[n_cells]: labelling condition id of the cell;
 * cond_vec[11]: the ids of the 11 conditions in statistic order; pulse_idx / chase_idx: 1-based cell indices
 * (load_process_data.jl:72-73); age_id_dist: 5 x 11 column-major, used as given for the data correlations (R12).
 * Point estimates: :2-38 (mean, Fano of u+l per age cluster over the pulse / chase cells), :61-79 (ratios), :100-147
 * (mean_corr, corr_mean).  Standard errors: :40-58, :81-97, :149-177 -- n_bootstraps (<= 128) resamples of the cells with
 * replacement per statistic family, drawn from Philox (counter = (block, bootstrap, family), key = seed) and shared by all
 * genes; the correlation bootstrap uses the resample's own per-condition age distribution (:160-164).  d, se: 53 x n_genes. */
int  abc_data_summary_stats(abc_ctx_t* ctx, const double* u, const double* l, int32_t n_cells, int32_t n_genes,
                            const int32_t* age, const int32_t* experiment, const int32_t* cond_vec,
                            const int32_t* pulse_idx, int32_t n_pulse, const int32_t* chase_idx, int32_t n_chase,
                            const double* age_id_dist, int32_t n_bootstraps, uint64_t seed, double* d, double* se);

/* ---- on-disk layouts written by the library; host-side, no device needed ------------------------------
 * abc_format_float64: Julia's print(io, ::Float64) -- what writedlm puts into every file of the pipeline (shortest
 * round-trip digits, fixed notation for 1e-5 <= |x| < 1e6, else d.ddde[-]x, NaN / Inf / -Inf); buf >= 32 bytes, returns
 * the length.  abc_writedlm: writedlm(io, A) of a row-major rows x cols matrix (compute_errors.jl:66-68: one 1 x G row
 * per particle), formatted on all host threads.  abc_write_simulation: the seven files of abc_simulation.jl:47-61, 89-95
 * for n trials (theta n x P, stats n x 53 row-major) under <dir>/<model>/.  abc_write_accepted:
 * data/posteriors/particles_<model>.txt (accepted_particles.jl:19-30, a "0" line for a gene without accepted particles).
 * abc_write_error_columns / abc_read_error_column: one raw little-endian Float64 file x<g>.f64 per gene column (the
 * column-per-file layout idea of the .jdf store, process_error_files.jl:3-7; JDF.jl's own bytes are third-party). */
int  abc_format_float64(double x, char* buf, size_t cap);
int  abc_writedlm(const char* path, const double* a, int64_t rows, int64_t cols, int append);
int  abc_write_simulation(const char* dir, int m, int32_t submit, const double* theta, const double* stats, int64_t n,
                          int64_t first_trial);
int  abc_write_accepted(const char* path, const int64_t* offsets, const int64_t* idx, int32_t n_genes, int append);
int  abc_write_error_columns(const char* dir, const double* err_gene_major, int64_t n, int64_t pitch, int32_t n_genes, int append);
int  abc_read_error_column(const char* dir, int32_t g, double* out, int64_t cap, int64_t* n);

/* ---- multi-GPU (the reference: independent `submit` processes + files concatenated by hand, wrapper.jl:62-63) ----------
 * Particles shard across GPUs by contiguous ranges of the global particle index (identical bits for any partition: Philox
 * is keyed by the global index).  NCCL is used only after a batch: per-gene counts are summed, and the accepted tuples are
 * exchanged by gene range so that every GPU orders the lists of G / n genes (equal tuple mass).
 *
 * (a) one host process, n GPUs -- what a Julia host calls: one context and one host thread per device inside the library,
 *     ncclCommInitAll.  devices == NULL: 0 .. n_dev-1.  abc_multi_simulate_score / abc_multi_accept_fetch return exactly what
 *     abc_simulate_score / abc_accept_fetch return on one device for the same arguments. */
typedef struct abc_multi abc_multi_t;
int  abc_multi_create(const int32_t* devices, int32_t n_dev, abc_multi_t** out);
int  abc_multi_destroy(abc_multi_t* mg);
int  abc_multi_n_devices(abc_multi_t* mg);
abc_ctx_t* abc_multi_ctx(abc_multi_t* mg, int32_t i);          /* the i-th device's context (options, counters) */
int  abc_multi_set_design(abc_multi_t* mg, const abc_design_t* design);
int  abc_multi_set_data(abc_multi_t* mg, const double* d, const double* se, int32_t n_genes);
int  abc_multi_set_option(abc_multi_t* mg, const char* name, int64_t value);
int  abc_multi_simulate_score(abc_multi_t* mg, int m, int64_t n_trials, int64_t particle_offset, uint64_t seed,
                              int prior_supplied, double* theta, double* stats, double eps, int err_layout, double* err,
                              int64_t* counts, abc_counters_t* counters);
int  abc_multi_accept_reset(abc_multi_t* mg);
int64_t abc_multi_accept_total(abc_multi_t* mg);
int  abc_multi_accept_fetch(abc_multi_t* mg, int64_t* offsets, int64_t* idx, double* errs);
/* (b) one process per GPU (torchrun / MPI style): rank 0 obtains a unique id (>= 128 bytes), the host distributes it, every
 *     rank attaches a communicator to its context.  abc_comm_counts: per-gene counts summed over the ranks.
 *     abc_comm_accept_fetch is collective: offsets[G+1] are global on every rank; root >= 0: that rank receives the complete
 *     ordered lists; root < 0: every rank receives the lists of its own gene range gene_range[0] .. gene_range[1]-1 at their
 *     global positions in idx / errs. */
int  abc_comm_unique_id(void* id, size_t bytes);
int  abc_comm_init_rank(abc_ctx_t* ctx, const void* id, size_t bytes, int32_t n_ranks, int32_t rank);
int  abc_comm_rank(abc_ctx_t* ctx, int32_t* n_ranks, int32_t* rank);
int  abc_comm_counts(abc_ctx_t* ctx, int64_t* counts);
int  abc_comm_accept_fetch(abc_ctx_t* ctx, int32_t root, int64_t* offsets, int64_t* idx, double* errs, int64_t* gene_range);
/* the cut used by the exchange: bounds[n_ranks + 1], rank k orders genes bounds[k] .. bounds[k+1]-1 (equal tuple mass) */
int  abc_gene_ranges(const int64_t* counts, int32_t n_genes, int32_t n_ranks, int64_t* bounds);

#ifdef __cplusplus
}
#endif
#endif
