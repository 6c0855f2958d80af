"""Host-side "next" rows of SURVEY 8f (O(G) work on the by-products of the scoring kernel).

Reference                                   here
get_model_probs (model_probs.jl:1-28,       get_model_probs(counts): acceptance-count ratios + 100-bootstrap
  constant_model_probs.jl:1-28,               percentile bounds.  The reference resamples a vector of model labels
  non_constant_model_probs.jl:1-30)           of length sum(l); drawing the class counts from a multinomial is the
                                              same distribution without materialising the vector.
get_n_particles (posterior_kinetics.jl:1-8) get_n_particles(offsets)
get_posterior_estimate "map"/"mean"          get_posterior_estimate(sets, offsets, idx, gene_vec, estimate)
  (posterior_kinetics.jl:10-24)               MAP = the FIRST accepted index = smallest error (posterior_kinetics.jl:14)
get_posterior_ci (posterior_kinetics.jl:26-33) get_posterior_ci(sets, offsets, idx, gene_vec, q); Julia quantile == numpy default
Particle indices are 1-based like the files the reference reads (data/posteriors/particles_<model>.txt).
The device version of the estimate / CI rows is AbcEngine.posterior_summary (abc_posterior_summary, csrc/abc_accept.cu);
the functions below are the host-side mirror used for file-based workflows and as its cross-check.
"""
import numpy as np


def get_model_probs(counts, n_bootstraps=100, alpha=0.95, rng=None):
    """counts: accepted particles per model for ONE gene (length K).  Returns (model_prob, l_b, u_b)."""
    l = np.asarray(counts, dtype=np.int64)
    K, tot = len(l), int(l.sum())
    if tot == 0:
        return np.zeros(K), np.zeros(K), np.zeros(K)
    rng = np.random.default_rng() if rng is None else rng
    prob = l / tot
    stats = rng.multinomial(tot, prob, size=n_bootstraps) / tot
    return prob, np.quantile(stats, 1.0 - alpha, axis=0), np.quantile(stats, alpha, axis=0)


def model_probs_for_genes(counts_by_model, groups, **kw):
    """counts_by_model: (n_models, G) acceptance counts (abc_score `counts`).  groups: list of lists of model rows
    pooled into one hypothesis, e.g. [[0, 1], [2, 3, 4]] = constant vs non-constant (model_probs.jl:37-41).
    Follows model_probs.jl:42-54: no accepted particle -> zeros; exactly one group accepted -> probability 1."""
    c = np.asarray(counts_by_model, dtype=np.int64)
    G = c.shape[1]
    pooled = np.stack([c[g].sum(0) for g in groups])            # (K, G)
    prob, lb, ub = (np.zeros((G, len(groups))) for _ in range(3))
    for j in range(G):
        which = np.nonzero(pooled[:, j] > 0)[0]
        if len(which) == 1:
            prob[j, which[0]] = lb[j, which[0]] = ub[j, which[0]] = 1.0
        elif len(which) > 1:
            prob[j], lb[j], ub[j] = get_model_probs(pooled[:, j], **kw)
    return prob, lb, ub


def get_n_particles(offsets):
    """accepted particles per gene (the "0" sentinel line counts as 0, posterior_kinetics.jl:1-8)"""
    return np.diff(np.asarray(offsets))


def _rows(sets, offsets, idx, g):
    v = idx[offsets[g - 1]:offsets[g]]
    if len(v) == 0:
        raise ValueError(f"gene {g} has no accepted particle")
    return sets[v - 1]


def get_posterior_estimate(sets, offsets, idx, gene_vec, estimate):
    """sets: (M, P) parameter sets (rows of sets_<model>.txt); gene_vec 1-based gene indices"""
    out = np.empty((len(gene_vec), sets.shape[1]))
    for i, g in enumerate(gene_vec):
        rows = _rows(sets, offsets, idx, g)
        out[i] = rows[0] if estimate == "map" else rows.mean(axis=0)
    return out


def get_posterior_ci(sets, offsets, idx, gene_vec, q):
    lbs = np.empty((len(gene_vec), sets.shape[1]))
    ubs = np.empty_like(lbs)
    for i, g in enumerate(gene_vec):
        rows = _rows(sets, offsets, idx, g)
        lbs[i] = np.quantile(rows, 1 - q, axis=0)
        ubs[i] = np.quantile(rows, q, axis=0)
    return [lbs, ubs]
