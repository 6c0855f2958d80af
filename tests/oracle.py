"""ctypes access to the CPU test oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY: imported by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never by the product."""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "liboracle.so")

NAGE, NCOND, NSTATS = 5, 11, 53
PULSE = [0.25, 0.5, 0.75, 1, 2, 3, 22, 22, 22, 22, 22]
CHASE = [0, 0, 0, 0, 0, 0, 0, 1, 2, 4, 6]
MATH_LIBM, MATH_DET = 0, 1


class OrcDesign(ctypes.Structure):
    _fields_ = [("cycle", ctypes.c_double), ("t0", ctypes.c_double), ("agevec", ctypes.c_double * 5),
                ("pulse", ctypes.c_double * 11), ("chase", ctypes.c_double * 11), ("age_dist", ctypes.c_double * 55),
                ("iv", ctypes.c_double * 9), ("downsampling", ctypes.c_int),
                ("beta_mean", ctypes.c_double * 10), ("beta_m2", ctypes.c_double * 10), ("beta_var", ctypes.c_double * 10),
                ("rtol", ctypes.c_double), ("atol", ctypes.c_double)]


class OrcSsaDesign(ctypes.Structure):
    _fields_ = [("cycle", ctypes.c_double), ("agevec", ctypes.c_double * 5), ("pulse", ctypes.c_double * 11),
                ("chase", ctypes.c_double * 11), ("n_cells", ctypes.c_int), ("n_pre", ctypes.c_int),
                ("downsampling", ctypes.c_int), ("beta_q32", ctypes.c_void_p), ("beta_off", ctypes.c_int * 11)]


_lib = None


def build(force=False):
    if force or not os.path.exists(LIB) or any(
            os.path.getmtime(os.path.join(ORACLE_DIR, f)) > os.path.getmtime(LIB)
            for f in os.listdir(ORACLE_DIR) if f.endswith((".c", ".h"))):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(LIB)
        _lib.orc_run_sim.restype = ctypes.c_int
        _lib.orc_run_part_sim.restype = ctypes.c_int
        _lib.orc_transient_phase.restype = ctypes.c_int
        _lib.orc_model.restype = ctypes.c_long
        _lib.orc_accept_gene.restype = ctypes.c_int64
        _lib.orc_nlsqerror_part.restype = ctypes.c_double
        _lib.orc_weighted_cov.restype = ctypes.c_double
        _lib.orc_size_scaling.restype = ctypes.c_double
        _lib.orc_labelling.restype = ctypes.c_double
        _lib.orc_exp10_det.restype = ctypes.c_double
        _lib.orc_exp10_det.argtypes = [ctypes.c_double]
        _lib.orc_size_scaling.argtypes = [ctypes.c_double, ctypes.c_double]
        _lib.orc_labelling.argtypes = [ctypes.c_double] * 4
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def beta_moments(betas, clusters):
    betas = np.ascontiguousarray(betas, dtype=np.float64)
    clusters = np.ascontiguousarray(clusters, dtype=np.int32)
    bm, b2, bv = np.zeros(5), np.zeros(5), np.zeros(5)
    lib().orc_beta_moments(_p(betas), _p(clusters), len(betas), _p(bm), _p(b2), _p(bv))
    return bm, b2, bv


def make_design(age_dist=None, iv_index=1, downsampling=False, betas=None, rtol=1e-8, atol=None):
    """betas = (betas_pulse, age_pulse, betas_chase, age_chase) when downsampling"""
    d = OrcDesign()
    d.cycle, d.t0 = 20.0, -60.0
    d.agevec[:] = [2, 6, 10, 14, 18]
    d.pulse[:] = PULSE
    d.chase[:] = CHASE
    ad = np.full((5, 11), 0.2) if age_dist is None else np.asarray(age_dist, dtype=np.float64)
    d.age_dist[:] = list(ad.T.reshape(-1))
    iv = [0.0] * 9
    iv[iv_index] = 0.5
    d.iv[:] = iv
    d.downsampling = int(downsampling)
    d.rtol = rtol
    d.atol = rtol * 1e-3 if atol is None else atol
    if downsampling:
        bp, ap, bc, ac = betas
        for s, (b, cl) in enumerate([(bp, ap), (bc, ac)]):
            bm, b2, bv = beta_moments(b, cl)
            for k in range(5):
                d.beta_mean[s * 5 + k], d.beta_m2[s * 5 + k], d.beta_var[s * 5 + k] = bm[k], b2[k], bv[k]
    return d


def make_ssa_design(n_cells, n_pre, downsampling, betas=None):
    """returns (OrcSsaDesign, keepalive)"""
    sd = OrcSsaDesign()
    sd.cycle = 20.0
    sd.agevec[:] = [2, 6, 10, 14, 18]
    sd.pulse[:] = PULSE
    sd.chase[:] = CHASE
    sd.n_cells, sd.n_pre, sd.downsampling = int(n_cells), int(n_pre), int(downsampling)
    keep = None
    if downsampling:
        bp, ap, bc, ac = betas
        q = np.zeros(len(bp) + len(bc), dtype=np.uint32)
        off = np.zeros(6, dtype=np.int32)
        b = np.ascontiguousarray(bp, dtype=np.float64)
        c = np.ascontiguousarray(ap, dtype=np.int32)
        lib().orc_quantise_betas(_p(b), _p(c), len(b), _p(q), _p(off))
        for k in range(6):
            sd.beta_off[k] = int(off[k])
        q2 = q[len(bp):]
        b = np.ascontiguousarray(bc, dtype=np.float64)
        c = np.ascontiguousarray(ac, dtype=np.int32)
        lib().orc_quantise_betas(_p(b), _p(c), len(b), _p(q2), _p(off))
        for k in range(6):
            sd.beta_off[5 + k] = int(off[k]) + len(bp)
        sd.beta_q32 = q.ctypes.data
        keep = q
    return sd, keep


def run_part_sim(theta, m, design):
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    out = np.zeros(275)
    k = lib().orc_run_part_sim(_p(theta), int(m), ctypes.byref(design), _p(out))
    return out.reshape(11, 5, 5), k


def run_sim(theta, m, design):
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    stats, mom = np.zeros(53), np.zeros(275)
    lib().orc_run_sim(_p(theta), int(m), ctypes.byref(design), _p(stats), _p(mom))
    return stats, mom.reshape(11, 5, 5)


def summary_stats(moments, age_dist, sample_guards=False):
    """S1; sample_guards=True: finite-sample conventions of the data side (what the SSA path uses)"""
    mom = np.ascontiguousarray(moments, dtype=np.float64).reshape(-1, 275)
    ad = np.ascontiguousarray(np.asarray(age_dist, dtype=np.float64).T.reshape(-1))
    out = np.zeros((mom.shape[0], 53))
    fn = lib().orc_summary_stats_sample if sample_guards else lib().orc_summary_stats
    for i in range(mom.shape[0]):
        fn(_p(mom[i]), _p(ad), _p(out[i]))
    return out


def compute_trunc_errors(stats, d, se):
    stats = np.ascontiguousarray(stats, dtype=np.float64).reshape(-1, 53)
    d = np.ascontiguousarray(d, dtype=np.float64)
    se = np.ascontiguousarray(se, dtype=np.float64)
    err = np.zeros((stats.shape[0], d.shape[0]))
    lib().orc_compute_trunc_errors(_p(stats), ctypes.c_int64(stats.shape[0]), _p(d), _p(se), d.shape[0], _p(err))
    return err


def accept_gene(err_col, eps):
    err_col = np.ascontiguousarray(err_col, dtype=np.float64)
    idx = np.zeros(len(err_col), dtype=np.int64)
    n = lib().orc_accept_gene(_p(err_col), ctypes.c_int64(len(err_col)), ctypes.c_int64(1), ctypes.c_double(eps), _p(idx))
    return idx[:n]


def prior(m, particle, seed, P):
    th = np.zeros(P)
    lib().orc_prior(int(m), ctypes.c_int64(particle), ctypes.c_uint64(seed), _p(th))
    return th


def ssa_readout(theta, m, sd, particle, seed, cond, age, math_mode=MATH_DET):
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    counts = np.zeros((4, sd.n_cells), dtype=np.uint32)
    ev = ctypes.c_uint64(0)
    lib().orc_ssa_readout(_p(theta), int(m), ctypes.byref(sd), ctypes.c_int64(particle), ctypes.c_uint64(seed),
                          int(cond), int(age), int(math_mode), _p(counts), ctypes.byref(ev))
    return counts, ev.value


def ssa_moments(theta, m, sd, particle, seed, math_mode=MATH_DET):
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    mom = np.zeros(275)
    ev = ctypes.c_uint64(0)
    lib().orc_ssa_moments(_p(theta), int(m), ctypes.byref(sd), ctypes.c_int64(particle), ctypes.c_uint64(seed),
                          int(math_mode), _p(mom), ctypes.byref(ev))
    return mom.reshape(11, 5, 5), ev.value


def philox(ctr, key):
    c = np.array(ctr, dtype=np.uint32)
    k = np.array(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    # static inline in the header: exercised through orc_prior / the SSA; a tiny shim is compiled on demand
    shim = os.path.join(ORACLE_DIR, "_philox_shim.so")
    if not os.path.exists(shim):
        src = '#include "oracle_philox.h"\nvoid shim(const uint32_t* c, const uint32_t* k, uint32_t* o){orc_philox4x32_10(c,k,o);}\n'
        subprocess.run(["gcc", "-O2", "-fPIC", "-shared", "-I", ORACLE_DIR, "-x", "c", "-", "-o", shim],
                       input=src.encode(), check=True)
    ctypes.CDLL(shim).shim(_p(c), _p(k), _p(out))
    return out


def same_bits(a, b):
    """bit-exact equality of float64 arrays; NaNs must sit in the same places (payload/sign of a NaN is
    not defined by IEEE-754 arithmetic and differs between x86 and the GPU)"""
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    if a.shape != b.shape:
        return False
    na, nb = np.isnan(a), np.isnan(b)
    if not np.array_equal(na, nb):
        return False
    return np.array_equal(a.view(np.uint64)[~na], b.view(np.uint64)[~nb])


def philox_np(c0, c1, c2, c3, key):
    """Philox4x32-10 (Salmon et al. 2011) vectorised over numpy arrays of counters; returns four uint32 arrays.
    Checked against the C restatement (philox) and Random123's known answers in tests/test_oracle_cpu.py."""
    M0, M1, W0, W1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), 0x9E3779B9, 0xBB67AE85
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & np.uint64(0xFFFFFFFF) for c in np.broadcast_arrays(c0, c1, c2, c3))
    k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    mask, sh = np.uint64(0xFFFFFFFF), np.uint64(32)
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        hi0, lo0, hi1, lo1 = p0 >> sh, p0 & mask, p1 >> sh, p1 & mask
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return tuple(c.astype(np.uint32) for c in (c0, c1, c2, c3))


def julia_quantile(sorted_v, p):
    """Statistics.jl quantile(v, p), default alpha = beta = 1, on sorted data"""
    n = len(sorted_v)
    aleph = n * p + (1.0 - p)
    j = int(min(max(np.trunc(aleph), 1), n - 1))
    g = min(max(aleph - j, 0.0), 1.0)
    a, b = sorted_v[j - 1], sorted_v[j]
    return a + g * (b - a)


def model_probs(counts, n_bootstraps, alpha, seed):
    """numpy restatement of abc_model_probs (model_probs.jl:1-54 with the label resampling drawn from Philox: block j of
    (gene g, bootstrap b) = counter (j, g, b, 2 << 29), two 64-bit words -> two labels by floor(u * sum(l) / 2^64))"""
    counts = np.asarray(counts, dtype=np.int64)
    K, G = counts.shape
    prob, lb, ub = (np.zeros((G, K)) for _ in range(3))
    key = (seed & 0xFFFFFFFF, seed >> 32)
    for g in range(G):
        l = counts[:, g]
        tot, nz = int(l.sum()), int((l > 0).sum())
        if nz == 0:
            continue
        if nz == 1:
            k = int(np.nonzero(l)[0][0])
            prob[g, k] = lb[g, k] = ub[g, k] = 1.0
            continue
        prob[g] = l / tot
        cum = np.cumsum(l)
        nblk = (tot + 1) // 2
        j = np.arange(nblk, dtype=np.uint64)
        stats = np.zeros((n_bootstraps, K))
        for b in range(n_bootstraps):
            x, y, z, w = philox_np(j, g, b, 2 << 29, key)
            u0 = (x.astype(object) << 32) | y.astype(object)
            u1 = (z.astype(object) << 32) | w.astype(object)
            i0 = np.array([(int(v) * tot) >> 64 for v in u0], dtype=np.int64)
            i1 = np.array([(int(v) * tot) >> 64 for v in u1], dtype=np.int64)
            if tot % 2 == 1:
                i1 = i1[:-1]
            lab = np.searchsorted(cum, np.concatenate([i0, i1]), side="right")
            stats[b] = np.bincount(lab, minlength=K) / tot
        s = np.sort(stats, axis=0)
        for k in range(K):
            lb[g, k] = julia_quantile(s[:, k], 1.0 - alpha)
            ub[g, k] = julia_quantile(s[:, k], alpha)
    return prob, lb, ub


def data_summary_stats(u, l, age, experiment, cond_vec, pulse_idx, chase_idx, age_id_dist, n_bootstraps, seed):
    """numpy restatement of abc_data_summary_stats = get_summary_stats (scripts/data_summary_statistics.jl:2-194) for every
    gene, with the resamples drawn from Philox as the library draws them (counter (block j, bootstrap b, family f, 3 << 29):
    two 64-bit words -> two cell positions floor(u64 * n_pop / 2^64) of the family's population: 0 pulse_idx, 1 chase_idx,
    2 and 3 all cells).  u, l: (G, n_cells) integer counts.  Exact integer sums, then the reference's FP64 formulas."""
    u = np.asarray(u).astype(np.int64); l = np.asarray(l).astype(np.int64)
    G, n_cells = u.shape
    age = np.asarray(age); experiment = np.asarray(experiment)
    ad = np.asarray(age_id_dist, dtype=np.float64)                      # (5, 11)
    key = (seed & 0xFFFFFFFF, seed >> 32)
    pops = [np.asarray(pulse_idx) - 1, np.asarray(chase_idx) - 1, np.arange(n_cells), np.arange(n_cells)]
    cond_of = np.full(n_cells, -1)
    for j, e in enumerate(cond_vec):
        cond_of[experiment == e] = j

    def resample(f, b):
        n = len(pops[f])
        j = np.arange((n + 1) // 2, dtype=np.uint64)
        x, y, z, w = philox_np(j, b, f, 3 << 29, key)
        u0 = [(int(a) << 32 | int(c)) * n >> 64 for a, c in zip(x, y)]
        u1 = [(int(a) << 32 | int(c)) * n >> 64 for a, c in zip(z, w)]
        if n % 2 == 1:
            u1 = u1[:-1]
        return pops[f][np.array(u0 + u1, dtype=np.int64)]

    def cov(n, sxy, sx, sy):
        if n < 2:
            return np.nan
        num = int(n) * int(sxy) - int(sx) * int(sy)
        return float(abs(num)) / (float(n) * float(n - 1)) * (1.0 if num >= 0 else -1.0)

    def wsum(w, x):
        s = 0.0
        for a, b in zip(w, x):
            s = s + a * b
        return s

    def wcov(x, y, w):
        mx, my = wsum(w, x), wsum(w, y)
        s = 0.0
        for i in range(5):
            s = s + w[i] * ((x[i] - mx) * (y[i] - my))
        return s

    def stats(g, cells4, boot):
        out = np.zeros(53)
        for f in range(2):
            c_ = cells4[f]
            t = u[g, c_] + l[g, c_]
            a_ = age[c_]
            for c in range(5):
                x = t[a_ == c + 1]
                n = len(x)
                if n > 0:
                    st, stt = int(x.sum()), int((x * x).sum())
                    mean = float(st) / float(n)
                    out[10 * f + c] = mean
                    out[10 * f + 5 + c] = cov(n, stt, st, st) / (mean + (0.0001 if mean == 0.0 else 0.0))
        c_ = cells4[2]
        cj = cond_of[c_]
        for j in range(11):
            sel = c_[cj == j]
            if len(sel) > 0:
                mu, ml = float(int(u[g, sel].sum())) / len(sel), float(int(l[g, sel].sum())) / len(sel)
                if mu + ml > 0.0:
                    out[20 + j] = ml / (mu + ml)
        c_ = cells4[3]
        cj, ca = cond_of[c_], age[c_]
        for j in range(11):
            m1, m2, v1, v2, c12, nc = ([0.0] * 5 for _ in range(6))
            for c in range(5):
                sel = c_[(cj == j) & (ca == c + 1)]
                n = len(sel)
                nc[c] = n
                if n > 0:
                    x, y = u[g, sel], l[g, sel]
                    sx, sy = int(x.sum()), int(y.sum())
                    m1[c], m2[c] = float(sx) / n, float(sy) / n
                    v1[c], v2[c] = cov(n, int((x * x).sum()), sx, sx), cov(n, int((y * y).sum()), sy, sy)
                    c12[c] = cov(n, int((x * y).sum()), sx, sy)
            ntot = sum(nc)
            w = [(float(k) / float(ntot) if ntot > 0 else 0.0) for k in nc] if boot else list(ad[:, j])
            tv1, tv2 = wsum(w, v1) + wcov(m1, m1, w), wsum(w, v2) + wcov(m2, m2, w)
            with np.errstate(invalid="ignore", divide="ignore"):
                sd = np.sqrt(np.abs(np.float64(tv1) * np.float64(tv2)))
                if any(c != 0.0 for c in c12) and tv1 != 0.0 and tv2 != 0.0:
                    out[31 + j] = np.float64(wsum(w, c12)) / sd
                if tv1 != 0.0 and tv2 != 0.0:
                    out[42 + j] = np.float64(wcov(m1, m2, w)) / sd
        return out

    d = np.array([stats(g, pops, False) for g in range(G)])
    samples = [[resample(f, b) for f in range(4)] for b in range(n_bootstraps)]
    boots = np.array([[stats(g, samples[b], True) for b in range(n_bootstraps)] for g in range(G)])     # (G, B, 53)
    with np.errstate(invalid="ignore"):
        s = np.zeros((G, 53))
        for b in range(n_bootstraps):
            s = s + boots[:, b]
        mean = s / float(n_bootstraps)
        q = np.zeros((G, 53))
        for b in range(n_bootstraps):
            dl = boots[:, b] - mean
            q = q + dl * dl
        se = np.sqrt((1.0 / float(n_bootstraps - 1)) * q)
    return d, se
