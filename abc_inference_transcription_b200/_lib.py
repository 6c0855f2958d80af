"""ctypes binding of libabcb200.so (include/abc_b200.h).

This is the stand-in for Julia's ``ccall`` in this container (no Julia toolchain): the same
symbols, the same argument meaning.  There is deliberately no fallback: if the shared library is
missing or no CUDA device is present every compute call raises.
"""
import ctypes
import os


_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ABCB200_LIB", os.path.join(_HERE, "libabcb200.so"))   # same override as julia/AbcB200.jl

NAGE, NCOND, NSTATS, NREAD = 5, 11, 53, 55
SIM_SSA, SIM_ODE = 0, 1
ERR_NONE, ERR_GENE_MAJOR, ERR_PARTICLE_MAJOR = 0, 1, 2

c_double_p = ctypes.POINTER(ctypes.c_double)
c_int32_p = ctypes.POINTER(ctypes.c_int32)
c_int64_p = ctypes.POINTER(ctypes.c_int64)
c_uint32_p = ctypes.POINTER(ctypes.c_uint32)


class AbcDesign(ctypes.Structure):
    """abc_design_t"""
    _fields_ = [
        ("cycle", ctypes.c_double), ("t0", ctypes.c_double),
        ("agevec", ctypes.c_double * NAGE),
        ("pulse", ctypes.c_double * NCOND), ("chase", ctypes.c_double * NCOND),
        ("age_dist", ctypes.c_double * (NAGE * NCOND)),
        ("iv", ctypes.c_double * 9),
        ("downsampling", ctypes.c_int32), ("n_cells", ctypes.c_int32),
        ("n_pre_cycles", ctypes.c_int32), ("sim_kind", ctypes.c_int32),
        ("betas_pulse", c_double_p), ("cluster_pulse", c_int32_p), ("n_pulse", ctypes.c_int32),
        ("betas_chase", c_double_p), ("cluster_chase", c_int32_p), ("n_chase", ctypes.c_int32),
        ("ode_rtol", ctypes.c_double), ("ode_atol", ctypes.c_double),
    ]


class AbcCounters(ctypes.Structure):
    """abc_counters_t"""
    _fields_ = [
        ("n_particles", ctypes.c_uint64), ("n_lineages", ctypes.c_uint64), ("n_events", ctypes.c_uint64),
        ("n_draws", ctypes.c_uint64), ("n_ode_steps", ctypes.c_uint64),
        ("ms_simulate", ctypes.c_double), ("ms_stats", ctypes.c_double), ("ms_score", ctypes.c_double),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


# every symbol include/abc_b200.h declares: name -> (restype, argtypes)
_vp = ctypes.c_void_p
SYMBOLS = {
    "abc_create": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(_vp)]),
    "abc_destroy": (ctypes.c_int, [_vp]),
    "abc_last_error": (ctypes.c_char_p, []),
    "abc_version": (ctypes.c_int, []),
    "abc_device_count": (ctypes.c_int, []),
    "abc_n_params": (ctypes.c_int, [ctypes.c_int]),
    "abc_model_name": (ctypes.c_char_p, [ctypes.c_int]),
    "abc_host_alloc": (ctypes.c_int, [ctypes.c_size_t, ctypes.POINTER(_vp)]),
    "abc_host_free": (ctypes.c_int, [_vp]),
    "abc_set_design": (ctypes.c_int, [_vp, ctypes.POINTER(AbcDesign)]),
    "abc_set_data": (ctypes.c_int, [_vp, _vp, _vp, ctypes.c_int32]),
    "abc_fix_params": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int64, ctypes.c_int64, ctypes.c_uint64, _vp]),
    "abc_simulate": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int64, ctypes.c_int64, ctypes.c_uint64, ctypes.c_int,
                                    _vp, _vp, ctypes.POINTER(AbcCounters)]),
    "abc_simulate_moments": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int64, ctypes.c_int64, ctypes.c_uint64,
                                            _vp, _vp, ctypes.POINTER(AbcCounters)]),
    "abc_ssa_cells": (ctypes.c_int, [_vp, ctypes.c_int, _vp, ctypes.c_int64, ctypes.c_uint64, ctypes.c_int,
                                     ctypes.c_int, ctypes.c_int, _vp]),
    "abc_ssa_window": (ctypes.c_int, [_vp, ctypes.c_int, _vp, _vp, ctypes.POINTER(ctypes.c_double)]),
    "abc_summary_stats": (ctypes.c_int, [_vp, _vp, ctypes.c_int64, _vp]),
    "abc_score": (ctypes.c_int, [_vp, _vp, ctypes.c_int64, ctypes.c_int64, ctypes.c_double, ctypes.c_int, _vp, _vp,
                                 ctypes.POINTER(AbcCounters)]),
    "abc_score_mma_columns": (ctypes.c_int, [_vp]),
    "abc_score_mma_debug": (ctypes.c_int, [_vp, _vp, ctypes.c_int64, _vp, _vp]),
    "abc_simulate_score": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int64, ctypes.c_int64, ctypes.c_uint64, ctypes.c_int,
                                          _vp, _vp, ctypes.c_double, ctypes.c_int, _vp, _vp, ctypes.POINTER(AbcCounters)]),
    "abc_simulate_score_async": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int64, ctypes.c_int64, ctypes.c_uint64, ctypes.c_int,
                                                _vp, _vp, ctypes.c_double, ctypes.c_int, _vp]),
    "abc_wait": (ctypes.c_int, [_vp, _vp, ctypes.POINTER(AbcCounters)]),
    "abc_accept_total": (ctypes.c_int64, [_vp]),
    "abc_accept_reset": (ctypes.c_int, [_vp]),
    "abc_accept_fetch": (ctypes.c_int, [_vp, _vp, _vp, _vp]),
    "abc_accept_tuples": (ctypes.c_int, [_vp, _vp, _vp, _vp]),
    "abc_posterior_summary": (ctypes.c_int, [_vp, _vp, ctypes.c_int64, ctypes.c_int32, ctypes.c_int64, ctypes.c_double,
                                             _vp, _vp, _vp, _vp, _vp]),
    "abc_simulate_dev": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int64, ctypes.c_int64, ctypes.c_uint64,
                                        ctypes.c_int, _vp, _vp, _vp]),
    "abc_score_dev": (ctypes.c_int, [_vp, _vp, ctypes.c_int64, ctypes.c_int64, ctypes.c_double, ctypes.c_int, _vp, _vp]),
    "abc_counts_dev": (ctypes.c_int, [_vp, _vp, _vp]),
    "abc_accept_tuples_dev": (ctypes.c_int, [_vp, _vp, _vp, _vp, ctypes.c_int64, _vp]),
    "abc_set_option": (ctypes.c_int, [_vp, ctypes.c_char_p, ctypes.c_int64]),
    "abc_counters": (ctypes.c_int, [_vp, ctypes.POINTER(AbcCounters)]),
    "abc_launch_count": (ctypes.c_int64, [_vp]),
    "abc_model_probs": (ctypes.c_int, [_vp, _vp, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_double, ctypes.c_uint64,
                                       _vp, _vp, _vp]),
    "abc_data_summary_stats": (ctypes.c_int, [_vp, _vp, _vp, ctypes.c_int32, ctypes.c_int32, _vp, _vp, _vp, _vp, ctypes.c_int32, _vp,
                                              ctypes.c_int32, _vp, ctypes.c_int32, ctypes.c_uint64, _vp, _vp]),
    "abc_format_float64": (ctypes.c_int, [ctypes.c_double, ctypes.c_char_p, ctypes.c_size_t]),
    "abc_writedlm": (ctypes.c_int, [ctypes.c_char_p, _vp, ctypes.c_int64, ctypes.c_int64, ctypes.c_int]),
    "abc_write_simulation": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_int, ctypes.c_int32, _vp, _vp, ctypes.c_int64, ctypes.c_int64]),
    "abc_write_accepted": (ctypes.c_int, [ctypes.c_char_p, _vp, _vp, ctypes.c_int32, ctypes.c_int]),
    "abc_write_error_columns": (ctypes.c_int, [ctypes.c_char_p, _vp, ctypes.c_int64, ctypes.c_int64, ctypes.c_int32, ctypes.c_int]),
    "abc_read_error_column": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_int32, _vp, ctypes.c_int64, c_int64_p]),
    "abc_multi_create": (ctypes.c_int, [_vp, ctypes.c_int32, ctypes.POINTER(_vp)]),
    "abc_multi_destroy": (ctypes.c_int, [_vp]),
    "abc_multi_n_devices": (ctypes.c_int, [_vp]),
    "abc_multi_ctx": (_vp, [_vp, ctypes.c_int32]),
    "abc_multi_set_design": (ctypes.c_int, [_vp, ctypes.POINTER(AbcDesign)]),
    "abc_multi_set_data": (ctypes.c_int, [_vp, _vp, _vp, ctypes.c_int32]),
    "abc_multi_set_option": (ctypes.c_int, [_vp, ctypes.c_char_p, ctypes.c_int64]),
    "abc_multi_simulate_score": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int64, ctypes.c_int64, ctypes.c_uint64, ctypes.c_int,
                                                _vp, _vp, ctypes.c_double, ctypes.c_int, _vp, _vp, ctypes.POINTER(AbcCounters)]),
    "abc_multi_accept_reset": (ctypes.c_int, [_vp]),
    "abc_multi_accept_total": (ctypes.c_int64, [_vp]),
    "abc_multi_accept_fetch": (ctypes.c_int, [_vp, _vp, _vp, _vp]),
    "abc_comm_unique_id": (ctypes.c_int, [_vp, ctypes.c_size_t]),
    "abc_comm_init_rank": (ctypes.c_int, [_vp, _vp, ctypes.c_size_t, ctypes.c_int32, ctypes.c_int32]),
    "abc_comm_rank": (ctypes.c_int, [_vp, c_int32_p, c_int32_p]),
    "abc_comm_counts": (ctypes.c_int, [_vp, _vp]),
    "abc_comm_accept_fetch": (ctypes.c_int, [_vp, ctypes.c_int32, _vp, _vp, _vp, _vp]),
    "abc_gene_ranges": (ctypes.c_int, [_vp, ctypes.c_int32, ctypes.c_int32, _vp]),
}

_lib = None


class AbcError(RuntimeError):
    pass


def load():
    """dlopen libabcb200.so and declare every prototype.  Raises if the library is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise AbcError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                       "(nvcc, sm_100a).  There is no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)   # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().abc_last_error()
        raise AbcError(f"libabcb200 error {rc}: {msg.decode() if msg else ''}")


def ptr(a):
    """raw pointer of a C-contiguous numpy array (or None)"""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(ctypes.c_void_p)
