"""Experimental design handed to the library (abc_design_t).

The reference builds these as Julia globals: scripts/abc_simulation.jl:65-79 (conditions, cycle, t0,
iv, agevec, downsampling) and scripts/load_process_data.jl:59-83 (age, pulse_idx, chase_idx,
age_id_distribution) from raw data that is not shipped.  ``synthetic_design`` reproduces
the SURVEY section 8d stand-ins: uniform age weights, the shipped capture efficiencies split into
chase rows 1-2364 / pulse rows 2365-5422 with round-robin age clusters.
"""
from dataclasses import dataclass, field

import numpy as np

from . import _lib
from .model import CONDITION_ID, N_AGE_CLUSTERS


@dataclass
class Design:
    cycle: float = 20.0                                                     # abc_simulation.jl:72
    t0: float = -60.0                                                       # abc_simulation.jl:73
    agevec: np.ndarray = field(default_factory=lambda: np.array([2.0, 6.0, 10.0, 14.0, 18.0]))  # tau_ .* cycle
    pulsevec: np.ndarray = field(default_factory=lambda: CONDITION_ID[:, 0].copy())
    chasevec: np.ndarray = field(default_factory=lambda: CONDITION_ID[:, 1].copy())
    age_dist: np.ndarray = field(default_factory=lambda: np.full((N_AGE_CLUSTERS, 11), 0.2))  # 5 x 11, as given
    iv: np.ndarray = field(default_factory=lambda: np.array([0, 0.5, 0, 0, 0, 0, 0, 0, 0], dtype=np.float64))
    downsampling: bool = True                                               # abc_simulation.jl:79
    betas_pulse: np.ndarray = None        # betas[pulse_idx]
    age_pulse: np.ndarray = None          # age[pulse_idx]  (cluster ids 1..5)
    betas_chase: np.ndarray = None        # betas[chase_idx]
    age_chase: np.ndarray = None          # age[chase_idx]
    n_cells: int = 96                     # SSA cells per (condition, age) read-out
    n_pre_cycles: int = 10                # SSA complete cycles before the read-out cycle
    sim_kind: int = _lib.SIM_SSA
    ode_rtol: float = 1e-6                # device ODE path; the reference's CVODE runs at 1e-3 / 1e-6
    ode_atol: float = 1e-9

    def to_c(self):
        """abc_design_t plus the numpy arrays that must stay alive while it is used"""
        d = _lib.AbcDesign()
        d.cycle, d.t0 = float(self.cycle), float(self.t0)
        d.agevec[:] = [float(x) for x in self.agevec]
        d.pulse[:] = [float(x) for x in self.pulsevec]
        d.chase[:] = [float(x) for x in self.chasevec]
        ad = np.asarray(self.age_dist, dtype=np.float64)
        assert ad.shape == (5, 11), "age_dist must be 5 x 11 (age cluster x condition)"
        d.age_dist[:] = [float(x) for x in ad.T.reshape(-1)]     # Julia column-major: condition j contiguous
        d.iv[:] = [float(x) for x in self.iv]
        d.downsampling = int(bool(self.downsampling))
        d.n_cells, d.n_pre_cycles, d.sim_kind = int(self.n_cells), int(self.n_pre_cycles), int(self.sim_kind)
        d.ode_rtol, d.ode_atol = float(self.ode_rtol), float(self.ode_atol)
        keep = []
        if self.downsampling:
            bp = np.ascontiguousarray(self.betas_pulse, dtype=np.float64)
            ap = np.ascontiguousarray(self.age_pulse, dtype=np.int32)
            bc = np.ascontiguousarray(self.betas_chase, dtype=np.float64)
            ac = np.ascontiguousarray(self.age_chase, dtype=np.int32)
            assert len(bp) == len(ap) and len(bc) == len(ac)
            d.betas_pulse = bp.ctypes.data_as(_lib.c_double_p)
            d.cluster_pulse = ap.ctypes.data_as(_lib.c_int32_p)
            d.n_pulse = len(bp)
            d.betas_chase = bc.ctypes.data_as(_lib.c_double_p)
            d.cluster_chase = ac.ctypes.data_as(_lib.c_int32_p)
            d.n_chase = len(bc)
            keep = [bp, ap, bc, ac]
        return d, keep


def split_betas(betas):
    """data/capture_efficiencies.txt: rows 1-2364 are chase cells, 2365-5422 pulse cells;
    age clusters are not shipped: round-robin 1 + (row mod 5) as in SURVEY 8d."""
    betas = np.asarray(betas, dtype=np.float64)
    chase, pulse = betas[:2364], betas[2364:]
    age_chase = (1 + np.arange(len(chase)) % 5).astype(np.int32)
    age_pulse = (1 + np.arange(len(pulse)) % 5).astype(np.int32)
    return pulse, age_pulse, chase, age_chase


def synthetic_design(betas, **kw):
    bp, ap, bc, ac = split_betas(betas)
    return Design(betas_pulse=bp, age_pulse=ap, betas_chase=bc, age_chase=ac, **kw)
