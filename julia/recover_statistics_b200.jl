# recover_statistics_b200.jl -- drop-in for `include("scripts/recover_statistics.jl")` (wrapper.jl:109-111): the moment
# equations of scripts/model.jl without downsampling on every MAP parameter set, on the device (sim_kind = SIM_ODE), written
# to data/recovered_statistics/<model>/<condition>/{mean_u,mean_l,var_u,cov_ul,var_l}.txt (recover_statistics.jl:49-68).
# Keeps the entry-point global m.  Expects τ_, betas, age, pulse_idx, chase_idx, age_id_distribution of section 1.
# NOT EXECUTED IN THE BUILD CONTAINER (no Julia there).
using DelimitedFiles
include(joinpath(@__DIR__, "AbcB200.jl"))

model_name = ["const","const_const","kon","alpha","gamma"][m]
maps = readdlm("data/posterior_estimates/map_sets_"*model_name*".txt")              # recover_statistics.jl:28
condition_id = hcat([1/4,1/2,3/4,1,2,3,22,22,22,22,22], [0,0,0,0,0,0,0,1,2,4,6])
cycle = 20.0
iv = zeros(9); iv[1] = 1/2                                                          # recover_statistics.jl:33-34
age_dist = age_id_distribution isa Matrix ? age_id_distribution : hcat(age_id_distribution...)
ctx = AbcB200.Context(0)
AbcB200.set_design(ctx; cycle=cycle, t0=-3cycle, agevec=τ_ .* cycle, pulsevec=condition_id[:,1], chasevec=condition_id[:,2],
                   age_dist=age_dist, iv=iv, downsampling=false, betas=betas, age=age, pulse_idx=pulse_idx, chase_idx=chase_idx,
                   sim_kind=AbcB200.SIM_ODE, ode_rtol=1e-6, ode_atol=1e-9)
id_labels = ["pulse_15", "pulse_30", "pulse_45", "pulse_60", "pulse_120", "pulse_180", "chase_0", "chase_60", "chase_120", "chase_240", "chase_360"]
@time mom = AbcB200.simulate_moments(ctx, m, permutedims(maps))                     # 5 x 5 x 11 x n
for (k, id) in enumerate(id_labels)
    dir = "data/recovered_statistics/"*model_name*"/"*id*"/"
    mkpath(dir)
    for (q, stem) in enumerate(["mean_u", "mean_l", "var_u", "cov_ul", "var_l"])
        AbcB200.writedlm_lib(dir*stem*".txt", mom[q, :, k, :]; append=true)        # ages x n -> one row of 5 ages per MAP row
    end
end
