"""Model tables of the reference (scripts/model.jl:30-43, scripts/abc_simulation.jl:82-85).

Host-side mirror: same names, same argument meaning, 1-based model index ``m`` like the Julia
driver (wrapper.jl:59).
"""
import numpy as np

MODEL_NAMES = ["const", "const_const", "kon", "alpha", "gamma"]          # abc_simulation.jl:82
VARY_FLAGS = [[0, 0, 0, 0], [0, 0, 0, 0], [1, 0, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]]  # abc_simulation.jl:83
N_AGE_CLUSTERS = 5                                                         # load_process_data.jl:68
# experimental conditions: 1st column pulse, 2nd column chase (hours)        abc_simulation.jl:65
CONDITION_ID = np.array([[1 / 4, 0], [1 / 2, 0], [3 / 4, 0], [1, 0], [2, 0], [3, 0],
                         [22, 0], [22, 1], [22, 2], [22, 4], [22, 6]], dtype=np.float64)
ID_LABELS = ["pulse_15", "pulse_30", "pulse_45", "pulse_60", "pulse_120", "pulse_180",
             "chase_0", "chase_60", "chase_120", "chase_240", "chase_360"]   # recover_statistics.jl:47
# prior box of fix_params (abc_simulation.jl:4-9), log10 units, per rate (kon, koff, alpha, gamma) + lambda
PRIOR_BOX = {"kon": (-3.0, 3.0), "koff": (-3.0, 3.0), "alpha": (-3.0, 3.0), "gamma": (-3.0, 2.0),
             "lambda": (-0.7, 0.0)}


def model_name(m):
    """model_name = [...][m]   (abc_simulation.jl:82)"""
    return MODEL_NAMES[_check_m(m) - 1]


def _check_m(m):
    if not (isinstance(m, (int, np.integer)) and 1 <= m <= 5):
        raise ValueError(f"model index m = {m!r} must be an integer in 1..5")
    return int(m)


def get_vary_map(vary_flag, n_steps=N_AGE_CLUSTERS):
    """get_vary_map (model.jl:30-43): 1-based parameter indices per rate; a varying rate owns n_steps."""
    keys, k = [], 1
    for f in vary_flag:
        if f == 0:
            keys.append(k)
            k += 1
        else:
            keys.append(list(range(k, k + n_steps)))
            k += n_steps
    return keys


def vary_map_for(m):
    return get_vary_map(VARY_FLAGS[_check_m(m) - 1], N_AGE_CLUSTERS)


def scaling_for(m):
    """scaling = 1 * (m != 2)   (abc_simulation.jl:85)"""
    return int(_check_m(m) != 2)


def n_params(m):
    """columns of fix_params' output: kon.., koff, alpha.., gamma.., lambda"""
    vm = vary_map_for(m)
    return sum(len(v) if isinstance(v, list) else 1 for v in vm) + 1


def prior_bounds(m):
    """(lo, hi) arrays in the column order of fix_params (abc_simulation.jl:3-11)"""
    vm = vary_map_for(m)
    lo, hi = [], []
    for name, v in zip(["kon", "koff", "alpha", "gamma"], vm):
        k = len(v) if isinstance(v, list) else 1
        lo += [PRIOR_BOX[name][0]] * k
        hi += [PRIOR_BOX[name][1]] * k
    lo.append(PRIOR_BOX["lambda"][0])
    hi.append(PRIOR_BOX["lambda"][1])
    return np.array(lo), np.array(hi)
