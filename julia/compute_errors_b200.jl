# compute_errors_b200.jl -- drop-in for compute_errors.jl + process_error_files.jl + accepted_particles.jl
# (wrapper.jl:72-81).  Keeps `m` / `model_name`, reads the simulation files (own restatement of the three loaders of
# compute_errors.jl:1-28 -- the reference script itself is NOT included: its tail, compute_errors.jl:72-81, runs the CPU
# scoring loop), writes the JDF column store and data/posteriors/particles_<model>.txt.
# NOT EXECUTED IN THE BUILD CONTAINER (no Julia there).
using DelimitedFiles
include(joinpath(@__DIR__, "AbcB200.jl"))

# s_pulse / s_chase hold two rows per particle: odd rows the means, even rows the Fano factors (abc_simulation.jl:47-52)
get_mean_subset(data::Matrix{Float64}) = data[1:2:end, :]                                    # compute_errors.jl:1-7
get_ff_subset(data::Matrix{Float64}) = data[2:2:end, :]                                      # compute_errors.jl:9-15
function load_s_data(path::String, model_name::String, ext::String)                          # compute_errors.jl:17-28
    rd(stem) = convert(Matrix{Float64}, readdlm(path*model_name*"/"*stem*"_"*model_name*ext))
    s_pulse, s_chase = rd("s_pulse"), rd("s_chase")
    return get_mean_subset(s_pulse), get_ff_subset(s_pulse), get_mean_subset(s_chase), get_ff_subset(s_chase),
           rd("s_ratios"), rd("s_mean_corr"), rd("s_corr_mean")
end

d  = permutedims(hcat(pulse_mean, pulse_ff, chase_mean, chase_ff, ratio_data, mean_corr_data, corr_mean_data))       # 53 x G
se = permutedims(hcat(pulse_mean_se, pulse_ff_se, chase_mean_se, chase_ff_se, ratio_se, mean_corr_se, corr_mean_se))
ctx = AbcB200.Context(0)
AbcB200.set_data(ctx, d, se)

s_pulse_mean,s_pulse_ff,s_chase_mean,s_chase_ff,s_ratios,s_mean_corr,s_corr_mean = load_s_data("data/simulations/",model_name,".txt")
stats = permutedims(hcat(s_pulse_mean,s_pulse_ff,s_chase_mean,s_chase_ff,s_ratios,s_mean_corr,s_corr_mean))           # 53 x M

ε = 4.8                                                                                      # accepted_particles.jl:10
AbcB200.accept_reset(ctx)
@time err, counts = AbcB200.score(ctx, stats; eps=ε, layout=AbcB200.ERR_GENE_MAJOR)          # M x G, column g == x<g>
# process_error_files.jl:5-6: one column per gene.  With JDF.jl installed:  JDF.save("data/errors/error_"*model_name*".jdf",
# DataFrame(err, :auto));  without it the library's own column store (one raw Float64 file per gene, AbcB200.read_error_column):
AbcB200.write_error_columns("data/errors/error_"*model_name*".cols", err; append=false)
# compute_errors.jl:66-68 (only if the text rows are wanted; 65 KB per particle):
# AbcB200.writedlm_lib("data/errors/error_"*model_name*".txt", permutedims(err); append=true)
offsets, idx, _ = AbcB200.accept_fetch(ctx)
mkpath("data/posteriors")
AbcB200.write_accepted("data/posteriors/particles_"*model_name*".txt", offsets, idx; append=true)   # accepted_particles.jl:23-30
