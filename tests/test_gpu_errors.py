"""Error behaviour of the C ABI on a GPU box: negative status + message, never a crash or a silent fallback."""
import ctypes
import os

import numpy as np
import pytest

from abc_inference_transcription_b200 import AbcEngine, AbcError, Design, synthetic_design, _lib

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_state_and_argument_errors():
    betas = np.load(os.path.join(GOLD, "ref_betas.npy"))
    z = np.load(os.path.join(GOLD, "ref_summary_stats.npz"))
    with AbcEngine(0) as eng:
        with pytest.raises(AbcError, match="abc_set_design"):
            eng.simulate(1, n_trials=4)                                  # ABC_ERR_STATE
        with pytest.raises(AbcError, match="abc_set_data"):
            eng.score(np.zeros((2, 53)))
        des = synthetic_design(betas)
        eng.set_design(des)
        lib = _lib.load()
        th, st = np.zeros((2, 5)), np.zeros((2, 53))
        assert lib.abc_simulate(eng._ctx, 6, 2, 0, 1, 0, _lib.ptr(th), _lib.ptr(st), None) == -1   # bad model index
        assert b"1..5" in lib.abc_last_error()
        assert lib.abc_simulate(eng._ctx, 1, 2, 0, 1, 0, None, _lib.ptr(st), None) == -1            # NULL buffer
        assert lib.abc_fix_params(eng._ctx, 1, -1, 0, 1, _lib.ptr(th)) == -1
        bad = synthetic_design(betas, n_pre_cycles=1)                      # chase 6 h + pulse 22 h needs >= 2 cycles
        with pytest.raises(AbcError, match="n_pre_cycles too small"):
            eng.set_design(bad)
        bad = synthetic_design(betas, agevec=np.array([2.0, 6.0, 10.0, 14.0, 20.0]))
        with pytest.raises(AbcError, match="agevec"):
            eng.set_design(bad)
        bad = synthetic_design(betas, n_cells=1)
        with pytest.raises(AbcError, match="n_cells"):
            eng.set_design(bad)
        with pytest.raises(AbcError, match="downsampling requires"):
            eng.set_design(Design(downsampling=True, betas_pulse=np.zeros(0), age_pulse=np.zeros(0, dtype=np.int32),
                                  betas_chase=np.zeros(0), age_chase=np.zeros(0, dtype=np.int32)))
        # the context is still usable after errors
        eng.set_design(des)
        eng.set_data(z["d"][:10], z["se"][:10])
        _, stats, _ = eng.simulate(2, n_trials=3)
        err, counts, _ = eng.score(stats)
        assert err.shape == (3, 10)
        with pytest.raises(AbcError, match="unknown option"):
            eng.set_option("no_such_option", 1)
        with pytest.raises(AbcError, match="bad arguments|err_layout"):
            eng.score(stats, err_layout=7)
    ctx = ctypes.c_void_p()
    assert _lib.load().abc_create(99, ctypes.byref(ctx)) == -1 and not ctx          # device out of range
    assert _lib.load().abc_destroy(None) == 0


def test_no_downsampling_design_runs():
    """downsampling = false (recover_statistics.jl:44): U', L' equal U, L"""
    betas = np.load(os.path.join(GOLD, "ref_betas.npy"))
    with AbcEngine(0) as eng:
        eng.set_design(synthetic_design(betas, downsampling=False, n_cells=64))
        c = eng.ssa_cells(1, np.log10([0.5, 1.0, 20.0, 0.1, 0.7]), particle_index=0, cond=6, age=2)
        assert np.array_equal(c[0], c[2]) and np.array_equal(c[1], c[3]) and c[1].sum() > 0
