# abc_simulation_b200.jl -- drop-in for `include("scripts/abc_simulation.jl")` (wrapper.jl:59-66).
# Keeps the entry-point globals m, n_trials, submit and writes the same files under
# data/simulations/<model>/ (abc_simulation.jl:47-61, 89-95); the per-particle work runs in libabcb200.
# Expects the globals of section 1 of wrapper.jl: τ_, betas, age, pulse_idx, chase_idx, age_id_distribution
# (a 5x11 Matrix{Float64}: hcat(age_id_distribution...) if it is the Vector{Vector} of load_process_data.jl:81).
# NOT EXECUTED IN THE BUILD CONTAINER (no Julia there).
using DelimitedFiles
include(joinpath(@__DIR__, "AbcB200.jl"))

condition_id = hcat([1/4,1/2,3/4,1,2,3,22,22,22,22,22], [0,0,0,0,0,0,0,1,2,4,6])   # abc_simulation.jl:65
cycle = 20.0
agevec = τ_ .* cycle
model_name = ["const","const_const","kon","alpha","gamma"][m]                     # abc_simulation.jl:82
age_dist = age_id_distribution isa Matrix ? age_id_distribution : hcat(age_id_distribution...)

ctx = AbcB200.Context(get(ENV, "ABCB200_DEVICE", "0") |> x -> parse(Int, x))
AbcB200.set_design(ctx; cycle=cycle, t0=-3cycle, agevec=agevec, pulsevec=condition_id[:,1], chasevec=condition_id[:,2],
                   age_dist=age_dist, downsampling=true, betas=betas, age=age, pulse_idx=pulse_idx, chase_idx=chase_idx)

batch = 65536
first_particle = (submit - 1) * n_trials          # distinct Philox streams per submit (wrapper.jl:62-63)
@time for b0 in 0:batch:n_trials-1
    nb = min(batch, n_trials - b0)
    θ, stats, _ = AbcB200.simulate(ctx, m, nb; particle_offset=first_particle + b0)
    # the seven appends of abc_simulation.jl:47-61, 89-95 (sets_, s_pulse_ / s_chase_ as 2 rows x 5 per trial, s_ratios_,
    # s_mean_corr_, s_corr_mean_, progress_), formatted like writedlm by the library: byte-identical files
    AbcB200.write_simulation("data/simulations", m, submit, θ, stats; first_trial=b0 + 1)
end
