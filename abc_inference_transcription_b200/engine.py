"""AbcEngine: one libabcb200 context on one GPU, numpy in / numpy out.

Thin wrapper over the C ABI (include/abc_b200.h) -- what the Julia host does with ``ccall``.  All
arithmetic happens in the CUDA library; this module only allocates host arrays and passes pointers.
Array conventions follow the Julia host: a Julia ``P x n`` column-major matrix is a numpy ``(n, P)``
C-contiguous array.
"""
import ctypes

import numpy as np

from . import _lib
from .design import Design
from .model import _check_m, n_params


class PinnedArray:
    """numpy view of page-locked host memory from abc_host_alloc (freed with the object)"""

    def __init__(self, shape, dtype=np.float64):
        self._lib = _lib.load()
        self._ptr = ctypes.c_void_p()
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        _lib.check(self._lib.abc_host_alloc(max(nbytes, 1), ctypes.byref(self._ptr)))
        buf = (ctypes.c_char * max(nbytes, 1)).from_address(self._ptr.value)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def __del__(self):
        try:
            if self._ptr:
                self.array = None
                self._lib.abc_host_free(self._ptr)
                self._ptr = None
        except Exception:
            pass


class AbcEngine:
    def __init__(self, device=0):
        self._lib = _lib.load()
        self._ctx = ctypes.c_void_p()
        _lib.check(self._lib.abc_create(int(device), ctypes.byref(self._ctx)))
        self.device = int(device)
        self.design = None
        self.n_genes = 0

    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx:
            self._lib.abc_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ---- configuration ---------------------------------------------------------------------
    def set_design(self, design: Design):
        cd, keep = design.to_c()
        _lib.check(self._lib.abc_set_design(self._ctx, ctypes.byref(cd)))
        del keep
        self.design = design

    def set_data(self, d, se):
        """d, se: (G, 53) in the order pulse_mean, pulse_ff, chase_mean, chase_ff, ratio, mean_corr, corr_mean"""
        d = np.ascontiguousarray(d, dtype=np.float64)
        se = np.ascontiguousarray(se, dtype=np.float64)
        assert d.ndim == 2 and d.shape[1] == _lib.NSTATS and d.shape == se.shape
        _lib.check(self._lib.abc_set_data(self._ctx, _lib.ptr(d), _lib.ptr(se), d.shape[0]))
        self.n_genes = d.shape[0]

    # ---- P1 ----------------------------------------------------------------------------------
    def fix_params(self, m, n, particle_offset=0, seed=20240229, out=None):
        """fix_params(vary_map, N) (abc_simulation.jl:3-11) -> (n, P) log10 parameters (written into `out` if given)"""
        P = n_params(_check_m(m))
        theta = out if out is not None else np.empty((int(n), P), dtype=np.float64)
        assert theta.shape == (int(n), P) and theta.dtype == np.float64 and theta.flags["C_CONTIGUOUS"]
        _lib.check(self._lib.abc_fix_params(self._ctx, m, int(n), int(particle_offset), int(seed), _lib.ptr(theta)))
        return theta

    # ---- M1-M10 + S1 ---------------------------------------------------------------------------
    def simulate(self, m, n_trials=None, theta=None, particle_offset=0, seed=20240229):
        """abc_sim over a batch.  theta=None draws the prior.  Returns (theta (n,P), stats (n,53), counters)"""
        P = n_params(_check_m(m))
        if theta is None:
            n = int(n_trials)
            theta = np.empty((n, P), dtype=np.float64)
            supplied = 0
        else:
            theta = np.ascontiguousarray(theta, dtype=np.float64).reshape(-1, P)
            n = theta.shape[0]
            supplied = 1
        stats = np.empty((n, _lib.NSTATS), dtype=np.float64)
        cnt = _lib.AbcCounters()
        _lib.check(self._lib.abc_simulate(self._ctx, m, n, int(particle_offset), int(seed), supplied,
                                          _lib.ptr(theta), _lib.ptr(stats), ctypes.byref(cnt)))
        return theta, stats, cnt.as_dict()

    def simulate_moments(self, m, theta, particle_offset=0, seed=20240229):
        """per (condition, age) moments: (n, 11, 5, 5) [mean_u, mean_l, var_u, cov_ul, var_l]"""
        P = n_params(_check_m(m))
        theta = np.ascontiguousarray(theta, dtype=np.float64).reshape(-1, P)
        n = theta.shape[0]
        mom = np.empty((n, _lib.NCOND, _lib.NAGE, 5), dtype=np.float64)
        cnt = _lib.AbcCounters()
        _lib.check(self._lib.abc_simulate_moments(self._ctx, m, n, int(particle_offset), int(seed),
                                                  _lib.ptr(theta), _lib.ptr(mom), ctypes.byref(cnt)))
        return mom, cnt.as_dict()

    def ssa_cells(self, m, theta, particle_index, cond, age, seed=20240229, exact_math=False):
        """per-cell counts of one read-out: (4, n_cells) uint32 rows U, L (before thinning), U', L'"""
        P = n_params(_check_m(m))
        theta = np.ascontiguousarray(theta, dtype=np.float64).reshape(P)
        out = np.empty((4, self.design.n_cells), dtype=np.uint32)
        _lib.check(self._lib.abc_ssa_cells(self._ctx, m, _lib.ptr(theta), int(particle_index), int(seed),
                                           int(cond), int(age), int(bool(exact_math)), _lib.ptr(out)))
        return out

    def ssa_window(self, m, theta):
        """start time (hours, 0 = start of the read-out cycle) of the lineages of each read-out, (11, 5) float32, and the
        expected switch draws of the particle (ssa_hybrid_burnin = 2)"""
        P = n_params(_check_m(m))
        theta = np.ascontiguousarray(theta, dtype=np.float64).reshape(P)
        starts = np.empty((_lib.NCOND, _lib.NAGE), dtype=np.float32)
        draws = ctypes.c_double()
        _lib.check(self._lib.abc_ssa_window(self._ctx, m, _lib.ptr(theta), _lib.ptr(starts), ctypes.byref(draws)))
        return starts, draws.value

    def summary_stats(self, moments):
        """S1 (abc_simulation.jl:23-46): (n, 11, 5, 5) moments -> (n, 53)"""
        mom = np.ascontiguousarray(moments, dtype=np.float64).reshape(-1, _lib.NCOND, _lib.NAGE, 5)
        stats = np.empty((mom.shape[0], _lib.NSTATS), dtype=np.float64)
        _lib.check(self._lib.abc_summary_stats(self._ctx, _lib.ptr(mom), mom.shape[0], _lib.ptr(stats)))
        return stats

    # ---- E2/E3 + A1 ------------------------------------------------------------------------------
    def accept_reset(self):
        _lib.check(self._lib.abc_accept_reset(self._ctx))

    def score(self, stats, eps=4.8, particle_offset=0, err_layout=_lib.ERR_PARTICLE_MAJOR, want_counts=True, out=None):
        """compute_trunc_errors + eps-acceptance.  Returns (err or None, counts or None, counters).
        err is (n, G) for ERR_PARTICLE_MAJOR (rows of error_<model>.txt) or (G, n) for ERR_GENE_MAJOR.
        out: optional preallocated C-contiguous float64 array of that shape (e.g. PinnedArray(...).array)."""
        stats = np.ascontiguousarray(stats, dtype=np.float64).reshape(-1, _lib.NSTATS)
        n, G = stats.shape[0], self.n_genes
        err = None
        shape = (n, G) if err_layout == _lib.ERR_PARTICLE_MAJOR else (G, n) if err_layout == _lib.ERR_GENE_MAJOR else None
        if shape is not None:
            if out is not None:
                assert out.shape == shape and out.dtype == np.float64 and out.flags["C_CONTIGUOUS"]
                err = out
            else:
                err = np.empty(shape, dtype=np.float64)
        counts = np.zeros(G, dtype=np.int64) if want_counts else None
        cnt = _lib.AbcCounters()
        _lib.check(self._lib.abc_score(self._ctx, _lib.ptr(stats), n, int(particle_offset), float(eps), int(err_layout),
                                       _lib.ptr(err), _lib.ptr(counts), ctypes.byref(cnt)))
        return err, counts, cnt.as_dict()

    def simulate_score(self, m, n_trials=None, theta=None, particle_offset=0, seed=20240229, eps=4.8,
                       err_layout=_lib.ERR_PARTICLE_MAJOR, want_counts=True, out=None, theta_out=None, stats_out=None):
        """wrapper.jl sections 2 + 3 for one batch in one pipelined call (abc_simulate_score): the same results as
        simulate() followed by score().  Returns (theta, stats, err or None, counts or None, counters).
        out / theta_out / stats_out: optional preallocated (ideally page-locked) output arrays."""
        P = n_params(_check_m(m))
        if theta is None:
            n = int(n_trials)
            theta = theta_out if theta_out is not None else np.empty((n, P), dtype=np.float64)
            supplied = 0
        else:
            theta = np.ascontiguousarray(theta, dtype=np.float64).reshape(-1, P)
            n = theta.shape[0]
            supplied = 1
        assert theta.shape == (n, P) and theta.dtype == np.float64 and theta.flags["C_CONTIGUOUS"]
        stats = stats_out if stats_out is not None else np.empty((n, _lib.NSTATS), dtype=np.float64)
        assert stats.shape == (n, _lib.NSTATS) and stats.dtype == np.float64 and stats.flags["C_CONTIGUOUS"]
        G = self.n_genes
        err = None
        shape = (n, G) if err_layout == _lib.ERR_PARTICLE_MAJOR else (G, n) if err_layout == _lib.ERR_GENE_MAJOR else None
        if shape is not None:
            if out is not None:
                assert out.shape == shape and out.dtype == np.float64 and out.flags["C_CONTIGUOUS"]
                err = out
            else:
                err = np.empty(shape, dtype=np.float64)
        counts = np.zeros(G, dtype=np.int64) if want_counts else None
        cnt = _lib.AbcCounters()
        _lib.check(self._lib.abc_simulate_score(self._ctx, m, n, int(particle_offset), int(seed), supplied, _lib.ptr(theta),
                                                _lib.ptr(stats), float(eps), int(err_layout), _lib.ptr(err), _lib.ptr(counts),
                                                ctypes.byref(cnt)))
        return theta, stats, err, counts, cnt.as_dict()

    def simulate_score_async(self, m, theta, stats, err=None, prior_supplied=True, particle_offset=0, seed=20240229, eps=4.8,
                             err_layout=_lib.ERR_PARTICLE_MAJOR):
        """abc_simulate_score_async: enqueue one batch and return.  theta (n, P), stats (n, 53) and err are the caller's
        (page-locked) output arrays -- theta is the input when prior_supplied -- and must stay alive and untouched until
        wait().  Up to two batches are in flight."""
        P = n_params(_check_m(m))
        n = theta.shape[0]
        assert theta.shape == (n, P) and theta.dtype == np.float64 and theta.flags["C_CONTIGUOUS"]
        assert stats.shape == (n, _lib.NSTATS) and stats.dtype == np.float64 and stats.flags["C_CONTIGUOUS"]
        if err_layout != _lib.ERR_NONE:
            shape = (n, self.n_genes) if err_layout == _lib.ERR_PARTICLE_MAJOR else (self.n_genes, n)
            assert err is not None and err.shape == shape and err.dtype == np.float64 and err.flags["C_CONTIGUOUS"]
        _lib.check(self._lib.abc_simulate_score_async(self._ctx, m, n, int(particle_offset), int(seed), int(bool(prior_supplied)),
                                                      _lib.ptr(theta), _lib.ptr(stats), float(eps), int(err_layout), _lib.ptr(err)))

    def wait(self, want_counts=True):
        """abc_wait: every asynchronous batch is complete.  Returns (counts or None, counters)."""
        counts = np.zeros(self.n_genes, dtype=np.int64) if want_counts and self.n_genes else None
        cnt = _lib.AbcCounters()
        _lib.check(self._lib.abc_wait(self._ctx, _lib.ptr(counts), ctypes.byref(cnt)))
        return counts, cnt.as_dict()

    def accept_total(self):
        t = self._lib.abc_accept_total(self._ctx)
        if t < 0:
            raise _lib.AbcError("abc_accept_total failed")
        return int(t)

    def accept_fetch(self):
        """CSR (offsets (G+1,), idx (total,) 1-based sorted by (err, idx) per gene, errs (total,))"""
        total = self.accept_total()
        offsets = np.zeros(self.n_genes + 1, dtype=np.int64)
        idx = np.empty(total, dtype=np.int64)
        errs = np.empty(total, dtype=np.float64)
        _lib.check(self._lib.abc_accept_fetch(self._ctx, _lib.ptr(offsets), _lib.ptr(idx), _lib.ptr(errs)))
        return offsets, idx, errs

    def posterior_summary(self, theta, particle_offset=0, q=0.95):
        """SURVEY 8f-3 on the device (posterior_kinetics.jl:10-33) over the lists accepted since the last accept_reset:
        {"map", "mean", "lo", "hi": (G, P) arrays (NaN rows for genes without accepted particles), "n": (G,) counts}.
        theta: (n, P) parameter sets of the particles particle_offset+1 .. particle_offset+n."""
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        n, P = theta.shape
        G = self.n_genes
        out = {k: np.empty((G, P), dtype=np.float64) for k in ("map", "mean", "lo", "hi")}
        out["n"] = np.zeros(G, dtype=np.int64)
        _lib.check(self._lib.abc_posterior_summary(self._ctx, _lib.ptr(theta), n, P, int(particle_offset), float(q),
                                                   _lib.ptr(out["map"]), _lib.ptr(out["mean"]), _lib.ptr(out["lo"]),
                                                   _lib.ptr(out["hi"]), _lib.ptr(out["n"])))
        return out

    def model_probs(self, counts, n_bootstraps=100, alpha=0.95, seed=20240229):
        """SURVEY 8f-2 on the device (model_probs.jl:1-54): counts (K, G) accepted particles per hypothesis and gene ->
        (prob, l_bound, u_bound), each (G, K)"""
        counts = np.ascontiguousarray(counts, dtype=np.int64)
        K, G = counts.shape
        out = [np.empty((G, K), dtype=np.float64) for _ in range(3)]
        _lib.check(self._lib.abc_model_probs(self._ctx, _lib.ptr(counts), K, G, int(n_bootstraps), float(alpha), int(seed),
                                             _lib.ptr(out[0]), _lib.ptr(out[1]), _lib.ptr(out[2])))
        return tuple(out)

    def data_summary_stats(self, u, l, age, experiment, cond_vec, pulse_idx, chase_idx, age_id_dist, n_bootstraps=100,
                           seed=20240229):
        """get_summary_stats of scripts/data_summary_statistics.jl:183-194 for every gene on the device: u, l (G, n_cells)
        integer-valued counts; age (n_cells,) clusters 1..5; experiment (n_cells,) condition ids; cond_vec (11,); pulse_idx /
        chase_idx 1-based cell indices; age_id_dist (5, 11).  Returns (d, se), each (G, 53): the inputs of set_data."""
        u = np.ascontiguousarray(u, dtype=np.float64)
        l = np.ascontiguousarray(l, dtype=np.float64)
        G, n_cells = u.shape
        assert l.shape == u.shape
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        age, experiment, cond_vec, pulse_idx, chase_idx = i32(age), i32(experiment), i32(cond_vec), i32(pulse_idx), i32(chase_idx)
        assert len(age) == n_cells == len(experiment) and len(cond_vec) == _lib.NCOND
        ad = np.ascontiguousarray(np.asarray(age_id_dist, dtype=np.float64).T.reshape(-1))     # 5 x 11 column-major
        d = np.empty((G, _lib.NSTATS), dtype=np.float64)
        se = np.empty((G, _lib.NSTATS), dtype=np.float64)
        _lib.check(self._lib.abc_data_summary_stats(self._ctx, _lib.ptr(u), _lib.ptr(l), n_cells, G, _lib.ptr(age), _lib.ptr(experiment),
                                                    _lib.ptr(cond_vec), _lib.ptr(pulse_idx), len(pulse_idx), _lib.ptr(chase_idx),
                                                    len(chase_idx), _lib.ptr(ad), int(n_bootstraps), int(seed), _lib.ptr(d), _lib.ptr(se)))
        return d, se

    def accept_tuples(self):
        total = self.accept_total()
        gene = np.empty(total, dtype=np.int32)
        part = np.empty(total, dtype=np.int64)
        errs = np.empty(total, dtype=np.float64)
        _lib.check(self._lib.abc_accept_tuples(self._ctx, _lib.ptr(gene), _lib.ptr(part), _lib.ptr(errs)))
        return gene, part, errs

    # ---- device-resident variants (raw device pointers, e.g. torch tensors' data_ptr()) -----------
    def simulate_dev(self, m, n, d_theta_ptr, d_stats_ptr, particle_offset=0, seed=20240229, prior_supplied=False,
                     stream=None):
        _lib.check(self._lib.abc_simulate_dev(self._ctx, _check_m(m), int(n), int(particle_offset), int(seed),
                                              int(bool(prior_supplied)), ctypes.c_void_p(d_theta_ptr),
                                              ctypes.c_void_p(d_stats_ptr), ctypes.c_void_p(stream or 0)))

    def score_dev(self, d_stats_ptr, n, eps=4.8, particle_offset=0, err_layout=_lib.ERR_NONE, d_err_ptr=0, stream=None):
        _lib.check(self._lib.abc_score_dev(self._ctx, ctypes.c_void_p(d_stats_ptr), int(n), int(particle_offset),
                                           float(eps), int(err_layout), ctypes.c_void_p(d_err_ptr or 0),
                                           ctypes.c_void_p(stream or 0)))

    def score_mma_debug(self, d_stats_ptr, n, d_out_ptr):
        """diagnostic: raw accumulators of the tensor-core filter into a device float array of ceil(n/128)*128 rows x
        score_mma_columns(); returns the gene index of every column (-1 = padding)"""
        cols = int(self._lib.abc_score_mma_columns(self._ctx))
        gene = np.empty(cols, dtype=np.int32)
        _lib.check(self._lib.abc_score_mma_debug(self._ctx, ctypes.c_void_p(d_stats_ptr), int(n), ctypes.c_void_p(d_out_ptr),
                                                 gene.ctypes.data_as(ctypes.c_void_p)))
        return gene

    def score_mma_columns(self):
        return int(self._lib.abc_score_mma_columns(self._ctx))

    def counts_dev(self, d_counts_ptr, stream=None):
        _lib.check(self._lib.abc_counts_dev(self._ctx, ctypes.c_void_p(d_counts_ptr), ctypes.c_void_p(stream or 0)))

    def accept_tuples_dev(self, d_gene_ptr, d_particle_ptr, d_err_ptr, capacity, stream=None):
        _lib.check(self._lib.abc_accept_tuples_dev(self._ctx, ctypes.c_void_p(d_gene_ptr), ctypes.c_void_p(d_particle_ptr),
                                                   ctypes.c_void_p(d_err_ptr), int(capacity), ctypes.c_void_p(stream or 0)))

    def set_option(self, name, value):
        _lib.check(self._lib.abc_set_option(self._ctx, name.encode(), int(value)))

    def counters(self):
        cnt = _lib.AbcCounters()
        _lib.check(self._lib.abc_counters(self._ctx, ctypes.byref(cnt)))
        return cnt.as_dict()

    def launch_count(self):
        return int(self._lib.abc_launch_count(self._ctx))


    # ---- one process per GPU: NCCL communicator owned by the library (abc_comm_*) ---------------------------------
    def comm_init(self, unique_id, n_ranks, rank):
        """attach this context to a communicator of n_ranks processes; unique_id: the 128 bytes rank 0 obtained from
        comm_unique_id() and the host distributed (torch.distributed / MPI / a file)"""
        buf = (ctypes.c_char * 128).from_buffer_copy(bytes(unique_id).ljust(128, b"\0")[:128])
        _lib.check(self._lib.abc_comm_init_rank(self._ctx, buf, 128, int(n_ranks), int(rank)))

    def comm_counts(self):
        """per-gene acceptance counts summed over the ranks (NCCL all-reduce)"""
        counts = np.zeros(self.n_genes, dtype=np.int64)
        _lib.check(self._lib.abc_comm_counts(self._ctx, _lib.ptr(counts)))
        return counts

    def comm_accept_fetch(self, root=0, total_hint=None):
        """collective: gene-range exchange + per-range ordering on every rank.  Returns (offsets, idx, errs, gene_range);
        idx / errs are complete on `root` (None elsewhere); root < 0: every rank gets its own gene range filled in."""
        offsets = np.zeros(self.n_genes + 1, dtype=np.int64)
        grange = np.zeros(2, dtype=np.int64)
        n_ranks, rank = ctypes.c_int32(), ctypes.c_int32()
        _lib.check(self._lib.abc_comm_rank(self._ctx, ctypes.byref(n_ranks), ctypes.byref(rank)))
        total = int(self.comm_counts().sum())
        want = root < 0 or rank.value == root
        idx = np.zeros(total, dtype=np.int64) if want else None
        errs = np.zeros(total, dtype=np.float64) if want else None
        _lib.check(self._lib.abc_comm_accept_fetch(self._ctx, int(root), _lib.ptr(offsets), _lib.ptr(idx), _lib.ptr(errs),
                                                   _lib.ptr(grange)))
        return offsets, idx, errs, (int(grange[0]), int(grange[1]))


def comm_unique_id():
    """128-byte NCCL unique id (rank 0 calls this, the host distributes it)"""
    buf = ctypes.create_string_buffer(128)
    _lib.check(_lib.load().abc_comm_unique_id(buf, 128))
    return buf.raw


def gene_ranges(counts, n_ranks):
    """the exchange's cut: n_ranks contiguous gene ranges of equal accepted-tuple mass -> bounds (n_ranks + 1,)"""
    counts = np.ascontiguousarray(counts, dtype=np.int64)
    bounds = np.zeros(int(n_ranks) + 1, dtype=np.int64)
    _lib.check(_lib.load().abc_gene_ranges(_lib.ptr(counts), len(counts), int(n_ranks), _lib.ptr(bounds)))
    return bounds


class AbcMulti:
    """All GPUs of one box behind one handle in ONE host process (abc_multi_*): what the Julia host calls.  The library
    runs one context and one host thread per device and owns the NCCL communicators."""

    def __init__(self, devices=None, n_dev=None):
        self._lib = _lib.load()
        self._mg = ctypes.c_void_p()
        if devices is None:
            n = int(n_dev if n_dev is not None else self._lib.abc_device_count())
            _lib.check(self._lib.abc_multi_create(None, n, ctypes.byref(self._mg)))
        else:
            dv = np.ascontiguousarray(devices, dtype=np.int32)
            _lib.check(self._lib.abc_multi_create(_lib.ptr(dv), len(dv), ctypes.byref(self._mg)))
        self.n_devices = int(self._lib.abc_multi_n_devices(self._mg))
        self.design, self.n_genes = None, 0

    def close(self):
        if getattr(self, "_mg", None) is not None and self._mg:
            self._lib.abc_multi_destroy(self._mg)
            self._mg = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def set_design(self, design: Design):
        cd, keep = design.to_c()
        _lib.check(self._lib.abc_multi_set_design(self._mg, ctypes.byref(cd)))
        del keep
        self.design = design

    def set_data(self, d, se):
        d = np.ascontiguousarray(d, dtype=np.float64)
        se = np.ascontiguousarray(se, dtype=np.float64)
        assert d.ndim == 2 and d.shape[1] == _lib.NSTATS and d.shape == se.shape
        _lib.check(self._lib.abc_multi_set_data(self._mg, _lib.ptr(d), _lib.ptr(se), d.shape[0]))
        self.n_genes = d.shape[0]

    def set_option(self, name, value):
        _lib.check(self._lib.abc_multi_set_option(self._mg, name.encode(), int(value)))

    def accept_reset(self):
        _lib.check(self._lib.abc_multi_accept_reset(self._mg))

    def accept_total(self):
        t = self._lib.abc_multi_accept_total(self._mg)
        if t < 0:
            raise _lib.AbcError("abc_multi_accept_total failed")
        return int(t)

    def simulate_score(self, m, n_trials=None, theta=None, particle_offset=0, seed=20240229, eps=4.8,
                       err_layout=_lib.ERR_PARTICLE_MAJOR, want_counts=True, out=None, theta_out=None, stats_out=None):
        """AbcEngine.simulate_score sharded over the devices: the same arguments, the same results"""
        P = n_params(_check_m(m))
        if theta is None:
            n = int(n_trials)
            theta = theta_out if theta_out is not None else np.empty((n, P), dtype=np.float64)
            supplied = 0
        else:
            theta = np.ascontiguousarray(theta, dtype=np.float64).reshape(-1, P)
            n = theta.shape[0]
            supplied = 1
        stats = stats_out if stats_out is not None else np.empty((n, _lib.NSTATS), dtype=np.float64)
        G = self.n_genes
        shape = (n, G) if err_layout == _lib.ERR_PARTICLE_MAJOR else (G, n) if err_layout == _lib.ERR_GENE_MAJOR else None
        err = None
        if shape is not None:
            err = out if out is not None else np.empty(shape, dtype=np.float64)
            assert err.shape == shape and err.dtype == np.float64 and err.flags["C_CONTIGUOUS"]
        counts = np.zeros(G, dtype=np.int64) if want_counts else None
        cnt = _lib.AbcCounters()
        _lib.check(self._lib.abc_multi_simulate_score(self._mg, m, n, int(particle_offset), int(seed), supplied, _lib.ptr(theta),
                                                      _lib.ptr(stats), float(eps), int(err_layout), _lib.ptr(err),
                                                      _lib.ptr(counts), ctypes.byref(cnt)))
        return theta, stats, err, counts, cnt.as_dict()

    def accept_fetch(self):
        total = self.accept_total()
        offsets = np.zeros(self.n_genes + 1, dtype=np.int64)
        idx = np.empty(total, dtype=np.int64)
        errs = np.empty(total, dtype=np.float64)
        _lib.check(self._lib.abc_multi_accept_fetch(self._mg, _lib.ptr(offsets), _lib.ptr(idx), _lib.ptr(errs)))
        return offsets, idx, errs
