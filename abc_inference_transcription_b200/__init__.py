"""abc_inference_transcription_b200 -- B200-native ABC simulate -> summarise -> score -> accept hot path
of pthomaslab/abc_inference_transcription behind a C ABI (include/abc_b200.h).

Host-side mirror of the reference's script interface (same names, argument meaning, file layouts):
  model.py               model tables, get_vary_map           scripts/model.jl:30-43, abc_simulation.jl:82-85
  abc_simulation.py      fix_params, abc_sim, run_sim, run     scripts/abc_simulation.jl
  compute_errors.py      load_s_data, compute_trunc_errors     scripts/compute_errors.jl, process_error_files.jl
  accepted_particles.py  accepted_particles                    scripts/accepted_particles.jl
All arithmetic runs in libabcb200.so (hand-written CUDA, sm_100a); there is no CPU fallback.
"""
from . import _lib, io
from ._lib import AbcError, ERR_GENE_MAJOR, ERR_NONE, ERR_PARTICLE_MAJOR, SIM_ODE, SIM_SSA
from .design import Design, split_betas, synthetic_design
from .engine import AbcEngine, AbcMulti, PinnedArray, comm_unique_id, gene_ranges
from .model import (CONDITION_ID, ID_LABELS, MODEL_NAMES, get_vary_map, model_name, n_params, prior_bounds,
                    scaling_for, vary_map_for)

__all__ = ["AbcEngine", "AbcMulti", "PinnedArray", "comm_unique_id", "gene_ranges", "AbcError", "Design", "synthetic_design", "split_betas", "MODEL_NAMES", "CONDITION_ID",
           "ID_LABELS", "get_vary_map", "model_name", "n_params", "prior_bounds", "scaling_for", "vary_map_for",
           "ERR_NONE", "ERR_GENE_MAJOR", "ERR_PARTICLE_MAJOR", "SIM_SSA", "SIM_ODE"]
