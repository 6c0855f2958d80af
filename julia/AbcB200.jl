# AbcB200.jl -- thin ccall binding of libabcb200.so (include/abc_b200.h) for the reference's Julia host.
#
# NOT EXECUTED IN THE BUILD CONTAINER (no Julia toolchain there, SURVEY R7); the same symbols are exercised
# through ctypes by abc_inference_transcription_b200/_lib.py and tests/.  No CUDA.jl, no codegen: every call is
# a plain ccall with Julia-owned column-major arrays.
module AbcB200

const LIB = get(ENV, "ABCB200_LIB", joinpath(@__DIR__, "..", "abc_inference_transcription_b200", "libabcb200.so"))

const SIM_SSA = Cint(0); const SIM_ODE = Cint(1)
const ERR_NONE = Cint(0); const ERR_GENE_MAJOR = Cint(1); const ERR_PARTICLE_MAJOR = Cint(2)

# mirrors abc_design_t field by field (isbits: passed by Ref)
struct Design
    cycle::Cdouble; t0::Cdouble
    agevec::NTuple{5,Cdouble}
    pulse::NTuple{11,Cdouble}; chase::NTuple{11,Cdouble}
    age_dist::NTuple{55,Cdouble}          # vec(age_dist::Matrix 5x11), column-major, used as given
    iv::NTuple{9,Cdouble}
    downsampling::Cint; n_cells::Cint; n_pre_cycles::Cint; sim_kind::Cint
    betas_pulse::Ptr{Cdouble}; cluster_pulse::Ptr{Cint}; n_pulse::Cint
    betas_chase::Ptr{Cdouble}; cluster_chase::Ptr{Cint}; n_chase::Cint
    ode_rtol::Cdouble; ode_atol::Cdouble
end

struct Counters
    n_particles::UInt64; n_lineages::UInt64; n_events::UInt64; n_draws::UInt64; n_ode_steps::UInt64
    ms_simulate::Cdouble; ms_stats::Cdouble; ms_score::Cdouble
end

mutable struct Context
    ptr::Ptr{Cvoid}
    n_genes::Int
end

last_error() = unsafe_string(ccall((:abc_last_error, LIB), Cstring, ()))
check(rc) = rc == 0 ? nothing : error("libabcb200 error $rc: $(last_error())")

function Context(device::Integer=0)
    p = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:abc_create, LIB), Cint, (Cint, Ref{Ptr{Cvoid}}), device, p))
    ctx = Context(p[], 0)
    finalizer(c -> ccall((:abc_destroy, LIB), Cint, (Ptr{Cvoid},), c.ptr), ctx)
    return ctx
end

n_params(m) = Int(ccall((:abc_n_params, LIB), Cint, (Cint,), m))

"""set_design(ctx; ...) -- the globals abc_sim receives (abc_simulation.jl:13-14, 65-79)"""
function set_design(ctx::Context; cycle=20.0, t0=-3cycle, agevec, pulsevec, chasevec, age_dist::Matrix{Float64},
                    iv=[0.0, 0.5, 0, 0, 0, 0, 0, 0, 0], downsampling=true, betas::Vector{Float64}, age::Vector{<:Integer},
                    pulse_idx::Vector{Int}, chase_idx::Vector{Int}, n_cells=96, n_pre_cycles=10, sim_kind=SIM_SSA,
                    ode_rtol=1e-6, ode_atol=1e-9)
    bp = betas[pulse_idx]; ap = Cint.(age[pulse_idx]); bc = betas[chase_idx]; ac = Cint.(age[chase_idx])
    GC.@preserve bp ap bc ac begin
        d = Design(cycle, t0, Tuple(agevec), Tuple(pulsevec), Tuple(chasevec), Tuple(vec(age_dist)), Tuple(iv),
                   downsampling, n_cells, n_pre_cycles, sim_kind,
                   pointer(bp), pointer(ap), length(bp), pointer(bc), pointer(ac), length(bc), ode_rtol, ode_atol)
        check(ccall((:abc_set_design, LIB), Cint, (Ptr{Cvoid}, Ref{Design}), ctx.ptr, d))
    end
end

"""set_data(ctx, d, se): d, se are 53 x G (rows: pulse_mean, pulse_ff, chase_mean, chase_ff, ratio, mean_corr, corr_mean)"""
function set_data(ctx::Context, d::Matrix{Float64}, se::Matrix{Float64})
    @assert size(d) == size(se) && size(d, 1) == 53
    check(ccall((:abc_set_data, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Cint), ctx.ptr, d, se, size(d, 2)))
    ctx.n_genes = size(d, 2)
end

"""fix_params(ctx, m, N; particle_offset, seed) -> P x N matrix (transpose it to get the reference's N x P)"""
function fix_params(ctx::Context, m, N; particle_offset=0, seed=UInt64(20240229))
    θ = Matrix{Float64}(undef, n_params(m), N)
    check(ccall((:abc_fix_params, LIB), Cint, (Ptr{Cvoid}, Cint, Int64, Int64, UInt64, Ptr{Cdouble}),
                ctx.ptr, m, N, particle_offset, seed, θ))
    return θ
end

"""simulate(ctx, m, n_trials) -> (θ P x n, stats 53 x n, counters): the trial loop of abc_simulation.jl:88-97"""
function simulate(ctx::Context, m, n_trials; particle_offset=0, seed=UInt64(20240229), theta=nothing)
    θ = theta === nothing ? Matrix{Float64}(undef, n_params(m), n_trials) : theta
    stats = Matrix{Float64}(undef, 53, size(θ, 2))
    c = Ref{Counters}()
    check(ccall((:abc_simulate, LIB), Cint,
                (Ptr{Cvoid}, Cint, Int64, Int64, UInt64, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Ref{Counters}),
                ctx.ptr, m, size(θ, 2), particle_offset, seed, theta === nothing ? 0 : 1, θ, stats, c))
    return θ, stats, c[]
end

"""simulate_moments(ctx, m, θ) -> 5 x 5 x 11 x n array [moment, age, condition, particle] (mean_u, mean_l, var_u, cov_ul,
var_l): run_part_sim of recover_statistics.jl:1-11 when the design has downsampling = false and sim_kind = SIM_ODE"""
function simulate_moments(ctx::Context, m, θ::Matrix{Float64}; particle_offset=0, seed=UInt64(20240229))
    n = size(θ, 2)
    mom = Array{Float64}(undef, 5, 5, 11, n)
    check(ccall((:abc_simulate_moments, LIB), Cint,
                (Ptr{Cvoid}, Cint, Int64, Int64, UInt64, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cvoid}),
                ctx.ptr, m, n, particle_offset, seed, θ, mom, C_NULL))
    return mom
end

"""set_option(ctx, "ssa_hybrid_burnin" | "ssa_adaptive_burnin" | "stats_sample_guards" | "score_reference_kernel" |
"accept_capacity", value)"""
set_option(ctx::Context, name::AbstractString, value::Integer) =
    check(ccall((:abc_set_option, LIB), Cint, (Ptr{Cvoid}, Cstring, Int64), ctx.ptr, name, value))

"""pinned_matrix(rows, cols) -> (A::Matrix{Float64}, ptr): page-locked host memory for the error matrix; release with
host_free(ptr) once A is no longer used"""
function pinned_matrix(rows::Integer, cols::Integer)
    p = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:abc_host_alloc, LIB), Cint, (Csize_t, Ref{Ptr{Cvoid}}), rows * cols * sizeof(Float64), p))
    return unsafe_wrap(Array, Ptr{Float64}(p[]), (rows, cols)), p[]
end
host_free(ptr::Ptr{Cvoid}) = check(ccall((:abc_host_free, LIB), Cint, (Ptr{Cvoid},), ptr))

accept_reset(ctx::Context) = check(ccall((:abc_accept_reset, LIB), Cint, (Ptr{Cvoid},), ctx.ptr))

"""score(ctx, stats; eps, layout) -> (err, counts).  ERR_PARTICLE_MAJOR: err is G x n (column i = row i of
error_<model>.txt); ERR_GENE_MAJOR: n x G (column g = JDF column x<g>)."""
function score(ctx::Context, stats::Matrix{Float64}; eps=4.8, particle_offset=0, layout=ERR_PARTICLE_MAJOR)
    n, G = size(stats, 2), ctx.n_genes
    err = layout == ERR_NONE ? Matrix{Float64}(undef, 0, 0) :
          layout == ERR_PARTICLE_MAJOR ? Matrix{Float64}(undef, G, n) : Matrix{Float64}(undef, n, G)
    counts = Vector{Int64}(undef, G)
    check(ccall((:abc_score, LIB), Cint,
                (Ptr{Cvoid}, Ptr{Cdouble}, Int64, Int64, Cdouble, Cint, Ptr{Cdouble}, Ptr{Int64}, Ptr{Cvoid}),
                ctx.ptr, stats, n, particle_offset, eps, layout, layout == ERR_NONE ? C_NULL : pointer(err), counts, C_NULL))
    return err, counts
end

"""simulate_score(ctx, m, n_trials; eps, layout, err) -> (θ, stats, err, counts, counters): wrapper.jl sections 2 and 3 for
one batch in one ccall (abc_simulate_score); the device-to-host copies are pipelined under the simulation.  Pass a
`pinned_matrix` as `err` (G x n for ERR_PARTICLE_MAJOR, n x G for ERR_GENE_MAJOR) to receive the error matrix at PCIe rate."""
function simulate_score(ctx::Context, m, n_trials; particle_offset=0, seed=UInt64(20240229), theta=nothing, eps=4.8,
                        layout=ERR_PARTICLE_MAJOR, err=nothing)
    θ = theta === nothing ? Matrix{Float64}(undef, n_params(m), n_trials) : theta
    n, G = size(θ, 2), ctx.n_genes
    stats = Matrix{Float64}(undef, 53, n)
    e = err !== nothing ? err : layout == ERR_NONE ? Matrix{Float64}(undef, 0, 0) :
        layout == ERR_PARTICLE_MAJOR ? Matrix{Float64}(undef, G, n) : Matrix{Float64}(undef, n, G)
    counts = Vector{Int64}(undef, G)
    c = Ref{Counters}()
    check(ccall((:abc_simulate_score, LIB), Cint,
                (Ptr{Cvoid}, Cint, Int64, Int64, UInt64, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Cdouble, Cint, Ptr{Cdouble},
                 Ptr{Int64}, Ref{Counters}),
                ctx.ptr, m, n, particle_offset, seed, theta === nothing ? 0 : 1, θ, stats, eps, layout,
                layout == ERR_NONE ? C_NULL : pointer(e), counts, c))
    return θ, stats, e, counts, c[]
end

"""simulate_score_async!(ctx, m, θ, stats, err; ...): enqueue one batch (abc_simulate_score_async) and return; θ (P x n, input
when prior_supplied), stats (53 x n) and err are caller-owned page-locked arrays (pinned_matrix) that stay untouched until
wait(ctx).  Two batches may be in flight: the copies of one overlap the simulation of the next."""
function simulate_score_async!(ctx::Context, m, θ::Matrix{Float64}, stats::Matrix{Float64}, err::Matrix{Float64};
                               prior_supplied=false, particle_offset=0, seed=UInt64(20240229), eps=4.8, layout=ERR_PARTICLE_MAJOR)
    check(ccall((:abc_simulate_score_async, LIB), Cint,
                (Ptr{Cvoid}, Cint, Int64, Int64, UInt64, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Cdouble, Cint, Ptr{Cdouble}),
                ctx.ptr, m, size(θ, 2), particle_offset, seed, prior_supplied ? 1 : 0, θ, stats, eps, layout,
                layout == ERR_NONE ? C_NULL : pointer(err)))
end

"""wait(ctx) -> (counts, counters): every asynchronous batch is complete"""
function wait(ctx::Context)
    counts = Vector{Int64}(undef, ctx.n_genes); c = Ref{Counters}()
    check(ccall((:abc_wait, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ref{Counters}), ctx.ptr, counts, c))
    return counts, c[]
end

"""accept_fetch(ctx) -> (offsets G+1, idx): idx[offsets[g]+1 : offsets[g+1]] == v[sortperm(err[v])] of gene g"""
function accept_fetch(ctx::Context)
    total = ccall((:abc_accept_total, LIB), Int64, (Ptr{Cvoid},), ctx.ptr)
    offsets = Vector{Int64}(undef, ctx.n_genes + 1); idx = Vector{Int64}(undef, total); errs = Vector{Float64}(undef, total)
    check(ccall((:abc_accept_fetch, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Cdouble}), ctx.ptr, offsets, idx, errs))
    return offsets, idx, errs
end

"""posterior_summary(ctx, θ; particle_offset, q) -> (map, mean, lo, hi, n): get_posterior_estimate / get_posterior_ci of
posterior_kinetics.jl:10-33 for every gene over the lists accepted since accept_reset (P x G matrices; NaN columns for genes
without accepted particles).  θ is P x n: the parameter sets of particles particle_offset+1 .. particle_offset+n."""
function posterior_summary(ctx::Context, θ::Matrix{Float64}; particle_offset=0, q=0.95)
    P, n, G = size(θ, 1), size(θ, 2), ctx.n_genes
    out = [Matrix{Float64}(undef, P, G) for _ in 1:4]
    nacc = Vector{Int64}(undef, G)
    check(ccall((:abc_posterior_summary, LIB), Cint,
                (Ptr{Cvoid}, Ptr{Cdouble}, Int64, Int32, Int64, Cdouble, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Int64}),
                ctx.ptr, θ, n, P, particle_offset, q, out[1], out[2], out[3], out[4], nacc))
    return out[1], out[2], out[3], out[4], nacc
end

"""model_probs(ctx, counts; n_bootstraps, alpha, seed) -> (prob, l_bound, u_bound), each K x G: get_model_probs and the case
split of model_probs.jl:1-54 on the device.  counts is G x K (column k = accepted particles per gene of hypothesis k, e.g.
hcat(counts_const .+ counts_const_const, counts_kon .+ counts_alpha .+ counts_gamma))."""
function model_probs(ctx::Context, counts::Matrix{Int64}; n_bootstraps=100, alpha=0.95, seed=UInt64(20240229))
    G, K = size(counts)
    out = [Matrix{Float64}(undef, K, G) for _ in 1:3]
    check(ccall((:abc_model_probs, LIB), Cint,
                (Ptr{Cvoid}, Ptr{Int64}, Int32, Int32, Int32, Cdouble, UInt64, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
                ctx.ptr, counts, K, G, n_bootstraps, alpha, seed, out[1], out[2], out[3]))
    return out[1], out[2], out[3]
end

"""data_summary_stats(ctx, u_data, l_data, age, experiment, cond_vec, pulse_idx, chase_idx, age_id_distribution) -> (d, se), each
53 x G: get_summary_stats of scripts/data_summary_statistics.jl:183-194 for every gene on the device (bootstrap SEs from seeded
Philox resamples).  u_data, l_data: n_cells x G count matrices (gene_selection.jl:36-38)."""
function data_summary_stats(ctx::Context, u_data::Matrix{Float64}, l_data::Matrix{Float64}, age::Vector{<:Integer},
                            experiment::Vector{<:Integer}, cond_vec::Vector{<:Integer}, pulse_idx::Vector{<:Integer},
                            chase_idx::Vector{<:Integer}, age_id_dist::Matrix{Float64}; n_bootstraps=100, seed=UInt64(20240229))
    n_cells, G = size(u_data)
    d = Matrix{Float64}(undef, 53, G); se = Matrix{Float64}(undef, 53, G)
    a, e, cv, pi, ci = Int32.(age), Int32.(experiment), Int32.(cond_vec), Int32.(pulse_idx), Int32.(chase_idx)
    check(ccall((:abc_data_summary_stats, LIB), Cint,
                (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Int32, Int32, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Int32,
                 Ptr{Int32}, Int32, Ptr{Cdouble}, Int32, UInt64, Ptr{Cdouble}, Ptr{Cdouble}),
                ctx.ptr, u_data, l_data, n_cells, G, a, e, cv, pi, length(pi), ci, length(ci), age_id_dist, n_bootstraps, seed, d, se))
    return d, se
end

# ---- the reference's on-disk layouts written by the library (no 65 KB of text per particle formatted in Julia) ----------
"""writedlm_lib(path, A; append): writedlm(io, transpose(A)) -- A is cols x rows (Julia column-major = the library's rows)"""
writedlm_lib(path::AbstractString, A::Matrix{Float64}; append=true) =
    check(ccall((:abc_writedlm, LIB), Cint, (Cstring, Ptr{Cdouble}, Int64, Int64, Cint), path, A, size(A, 2), size(A, 1), append))

"""write_simulation(dir, m, submit, θ, stats; first_trial): the seven appends of abc_simulation.jl:47-61, 89-95 for a batch"""
write_simulation(dir::AbstractString, m, submit, θ::Matrix{Float64}, stats::Matrix{Float64}; first_trial=1) =
    check(ccall((:abc_write_simulation, LIB), Cint, (Cstring, Cint, Int32, Ptr{Cdouble}, Ptr{Cdouble}, Int64, Int64),
                dir, m, submit, θ, stats, size(θ, 2), first_trial))

"""write_accepted(path, offsets, idx; append): particles_<model>.txt (accepted_particles.jl:19-30)"""
write_accepted(path::AbstractString, offsets::Vector{Int64}, idx::Vector{Int64}; append=true) =
    check(ccall((:abc_write_accepted, LIB), Cint, (Cstring, Ptr{Int64}, Ptr{Int64}, Int32, Cint), path, offsets, idx, length(offsets) - 1, append))

"""write_error_columns(dir, err; append): err is n x G (ERR_GENE_MAJOR): one raw Float64 file x<g>.f64 per gene"""
write_error_columns(dir::AbstractString, err::Matrix{Float64}; append=true) =
    check(ccall((:abc_write_error_columns, LIB), Cint, (Cstring, Ptr{Cdouble}, Int64, Int64, Int32, Cint),
                dir, err, size(err, 1), size(err, 1), size(err, 2), append))

"""read_error_column(dir, g) -> Vector{Float64}: f["x\$g"] of the reference's JDFFile (accepted_particles.jl:14-18)"""
function read_error_column(dir::AbstractString, g::Integer)
    n = Ref{Int64}(0)
    check(ccall((:abc_read_error_column, LIB), Cint, (Cstring, Int32, Ptr{Cdouble}, Int64, Ref{Int64}), dir, g, C_NULL, 0, n))
    v = Vector{Float64}(undef, n[])
    check(ccall((:abc_read_error_column, LIB), Cint, (Cstring, Int32, Ptr{Cdouble}, Int64, Ref{Int64}), dir, g, v, n[], n))
    return v
end

# ---- all GPUs of the box from this one Julia process (abc_multi_*): the library runs one host thread per device -------
mutable struct MultiContext
    ptr::Ptr{Cvoid}
    n_genes::Int
end

"""MultiContext(; devices): devices = nothing uses every visible GPU"""
function MultiContext(; devices::Union{Nothing,Vector{Int32}}=nothing)
    p = Ref{Ptr{Cvoid}}(C_NULL)
    n = devices === nothing ? ccall((:abc_device_count, LIB), Cint, ()) : Cint(length(devices))
    check(ccall((:abc_multi_create, LIB), Cint, (Ptr{Int32}, Int32, Ref{Ptr{Cvoid}}), devices === nothing ? C_NULL : pointer(devices), n, p))
    mg = MultiContext(p[], 0)
    finalizer(c -> ccall((:abc_multi_destroy, LIB), Cint, (Ptr{Cvoid},), c.ptr), mg)
    return mg
end

n_devices(mg::MultiContext) = Int(ccall((:abc_multi_n_devices, LIB), Cint, (Ptr{Cvoid},), mg.ptr))

function set_design(mg::MultiContext; cycle=20.0, t0=-3cycle, agevec, pulsevec, chasevec, age_dist::Matrix{Float64},
                    iv=[0.0, 0.5, 0, 0, 0, 0, 0, 0, 0], downsampling=true, betas::Vector{Float64}, age::Vector{<:Integer},
                    pulse_idx::Vector{Int}, chase_idx::Vector{Int}, n_cells=96, n_pre_cycles=10, sim_kind=SIM_SSA,
                    ode_rtol=1e-6, ode_atol=1e-9)
    bp = betas[pulse_idx]; ap = Cint.(age[pulse_idx]); bc = betas[chase_idx]; ac = Cint.(age[chase_idx])
    GC.@preserve bp ap bc ac begin
        d = Design(cycle, t0, Tuple(agevec), Tuple(pulsevec), Tuple(chasevec), Tuple(vec(age_dist)), Tuple(iv),
                   downsampling, n_cells, n_pre_cycles, sim_kind,
                   pointer(bp), pointer(ap), length(bp), pointer(bc), pointer(ac), length(bc), ode_rtol, ode_atol)
        check(ccall((:abc_multi_set_design, LIB), Cint, (Ptr{Cvoid}, Ref{Design}), mg.ptr, d))
    end
end

function set_data(mg::MultiContext, d::Matrix{Float64}, se::Matrix{Float64})
    @assert size(d) == size(se) && size(d, 1) == 53
    check(ccall((:abc_multi_set_data, LIB), Cint, (Ptr{Cvoid}, Ptr{Cdouble}, Ptr{Cdouble}, Cint), mg.ptr, d, se, size(d, 2)))
    mg.n_genes = size(d, 2)
end

set_option(mg::MultiContext, name::AbstractString, value::Integer) =
    check(ccall((:abc_multi_set_option, LIB), Cint, (Ptr{Cvoid}, Cstring, Int64), mg.ptr, name, value))
accept_reset(mg::MultiContext) = check(ccall((:abc_multi_accept_reset, LIB), Cint, (Ptr{Cvoid},), mg.ptr))

"""simulate_score(mg, m, n_trials; ...): the particles are sharded over the GPUs by contiguous ranges; same arguments and the
same results (bit for bit) as on one Context"""
function simulate_score(mg::MultiContext, m, n_trials; particle_offset=0, seed=UInt64(20240229), theta=nothing, eps=4.8,
                        layout=ERR_NONE, err=nothing)
    θ = theta === nothing ? Matrix{Float64}(undef, n_params(m), n_trials) : theta
    n, G = size(θ, 2), mg.n_genes
    stats = Matrix{Float64}(undef, 53, n)
    e = err !== nothing ? err : layout == ERR_NONE ? Matrix{Float64}(undef, 0, 0) :
        layout == ERR_PARTICLE_MAJOR ? Matrix{Float64}(undef, G, n) : Matrix{Float64}(undef, n, G)
    counts = Vector{Int64}(undef, G)
    c = Ref{Counters}()
    check(ccall((:abc_multi_simulate_score, LIB), Cint,
                (Ptr{Cvoid}, Cint, Int64, Int64, UInt64, Cint, Ptr{Cdouble}, Ptr{Cdouble}, Cdouble, Cint, Ptr{Cdouble},
                 Ptr{Int64}, Ref{Counters}),
                mg.ptr, m, n, particle_offset, seed, theta === nothing ? 0 : 1, θ, stats, eps, layout,
                layout == ERR_NONE ? C_NULL : pointer(e), counts, c))
    return θ, stats, e, counts, c[]
end

"""accept_fetch(mg) -> (offsets, idx, errs): the merged per-gene lists of all GPUs (gene-range exchange over NCCL)"""
function accept_fetch(mg::MultiContext)
    total = ccall((:abc_multi_accept_total, LIB), Int64, (Ptr{Cvoid},), mg.ptr)
    offsets = Vector{Int64}(undef, mg.n_genes + 1); idx = Vector{Int64}(undef, total); errs = Vector{Float64}(undef, total)
    check(ccall((:abc_multi_accept_fetch, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Cdouble}), mg.ptr, offsets, idx, errs))
    return offsets, idx, errs
end

end # module
