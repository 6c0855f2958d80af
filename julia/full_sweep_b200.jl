# full_sweep_b200.jl -- wrapper.jl sections 2 and 3 for all five models on every GPU of the box from ONE Julia process
# (README.md:37: "M = 5*10^6 parameter sets ... simulate our 5 models"; scripts/accepted_particles.jl:13: 10^6 rows per
# model).  Keeps the entry-point globals n_trials (particles per model) and the file layouts; no error matrix is
# materialised (fused eps-acceptance), the per-gene accepted lists and counts come back merged over the GPUs.
# Expects the globals of section 1 of wrapper.jl (τ_, betas, age, pulse_idx, chase_idx, age_id_distribution) and the 14
# data matrices.  NOT EXECUTED IN THE BUILD CONTAINER (no Julia there).
using DelimitedFiles
include(joinpath(@__DIR__, "AbcB200.jl"))

condition_id = hcat([1/4,1/2,3/4,1,2,3,22,22,22,22,22], [0,0,0,0,0,0,0,1,2,4,6])   # abc_simulation.jl:65
cycle = 20.0
age_dist = age_id_distribution isa Matrix ? age_id_distribution : hcat(age_id_distribution...)
d  = permutedims(hcat(pulse_mean, pulse_ff, chase_mean, chase_ff, ratio_data, mean_corr_data, corr_mean_data))       # 53 x G
se = permutedims(hcat(pulse_mean_se, pulse_ff_se, chase_mean_se, chase_ff_se, ratio_se, mean_corr_se, corr_mean_se))

mg = AbcB200.MultiContext()                                   # every visible GPU; NCCL communicators owned by the library
AbcB200.set_design(mg; cycle=cycle, t0=-3cycle, agevec=τ_ .* cycle, pulsevec=condition_id[:,1], chasevec=condition_id[:,2],
                   age_dist=age_dist, downsampling=true, betas=betas, age=age, pulse_idx=pulse_idx, chase_idx=chase_idx)
AbcB200.set_data(mg, d, se)
ε = 4.8                                                       # accepted_particles.jl:10
batch = 262144
all_counts = Matrix{Int64}(undef, size(d, 2), 5)
mkpath("data/posteriors")
@time for m in 1:5
    model_name = ["const","const_const","kon","alpha","gamma"][m]
    AbcB200.accept_reset(mg)
    counts = zeros(Int64, size(d, 2))
    for b0 in 0:batch:n_trials-1
        nb = min(batch, n_trials - b0)
        θ, stats, _, c, _ = AbcB200.simulate_score(mg, m, nb; particle_offset=b0, eps=ε, layout=AbcB200.ERR_NONE)
        AbcB200.write_simulation("data/simulations", m, 1, θ, stats; first_trial=b0 + 1)
        counts .= c                                            # running per-gene counts since accept_reset
    end
    offsets, idx, _ = AbcB200.accept_fetch(mg)                 # v[sortperm(err[v])] per gene, merged over the GPUs
    AbcB200.write_accepted("data/posteriors/particles_"*model_name*".txt", offsets, idx; append=false)
    all_counts[:, m] = diff(offsets)
end
# model_probs.jl: constant (models 1, 2) vs non-constant (3, 4, 5) with bootstrap bounds, on the device of context 1
ctx = AbcB200.Context(0)
prob, l_bound, u_bound = AbcB200.model_probs(ctx, hcat(all_counts[:,1] .+ all_counts[:,2], all_counts[:,3] .+ all_counts[:,4] .+ all_counts[:,5]))
mkpath("data/model_selection/all")
AbcB200.writedlm_lib("data/model_selection/all/model_prob.txt", prob; append=false)
AbcB200.writedlm_lib("data/model_selection/all/l_bound.txt", l_bound; append=false)
AbcB200.writedlm_lib("data/model_selection/all/u_bound.txt", u_bound; append=false)
