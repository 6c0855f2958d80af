"""Independent numpy restatement of the start-time rule of the product sampler (csrc/abc_tele.cu abc_window_kernel,
ssa_adaptive_burnin = 2) by quadrature on a fine time grid -- the kernel uses closed forms per piece and a
bisection.  Times are hours, 0 = start of the read-out cycle, read-out at t* = age.

A transcript born at time s is alive at t* with probability 2^-B(s), B(s) = log2(e) int_s^t* gamma + #divisions in (s, t*]
(scripts/model.jl:74-86 decay terms, :98-111 binomial halving).  With the gene in its stationary law the density of the
Poisson means at the read-out over birth times is
    w_U(s) = alpha(s) P_on(s) (1 - lam(s)) 2^-B(s),   w_L(s) = alpha(s) P_on(s) lam(s) 2^-B(s),
lam(s) = 10^theta_lambda inside the label window [age - pulse - chase, age - chase] (model.jl:58-64, :183), else 0.
The lineages start at the latest s0 with  int_{-inf}^{s0} w_X <= max(2^-n_pre int_{s0}^{t*} w_X, 2^-30)  for X = U and L
(history before -n_pre cycles is never simulated).  Model 3 (kon varies): the gene starts in the stationary law of its
step, which is approximate, so s0 moves back until exp(-int (kon + koff)) <= 2^-6 over [s0, s0_Lambda]."""
import numpy as np

CYCLE = 20.0
AGES = [2.0, 6.0, 10.0, 14.0, 18.0]
PULSE = [0.25, 0.5, 0.75, 1, 2, 3, 22, 22, 22, 22, 22]          # scripts/abc_simulation.jl:65
CHASE = [0, 0, 0, 0, 0, 0, 0, 1, 2, 4, 6]
LOG2E = 1.4426950408889634


def rates_of(theta, m):
    """scripts/model.jl:1-22, 30-43: per-step linear rates (kon, koff, alpha, gamma)[5] and lambda"""
    vary = {3: 0, 4: 2, 5: 3}.get(m, -1)
    th, k, out = 10.0 ** np.asarray(theta, dtype=np.float64), 0, []
    for q in range(4):
        if q == vary:
            out.append(th[k:k + 5].copy()); k += 5
        else:
            out.append(np.full(5, th[k])); k += 1
    return out[0], out[1], out[2], out[3], min(th[k], 1.0)


def contribution_densities(theta, m, cond, age_i, n_pre=10, dt=0.002):
    """grid mid-points t (from -n_pre cycles to the read-out) and w_U dt, w_L dt, (kon + koff) dt on it"""
    kon, koff, al, ga, lam = rates_of(theta, m)
    age = AGES[age_i]
    tl0, tl1 = age - PULSE[cond] - CHASE[cond], age - CHASE[cond]
    n = int(round((n_pre * CYCLE + age) / dt))
    t = -n_pre * CYCLE + (np.arange(n) + 0.5) * dt
    x = np.mod(t, CYCLE)
    step = np.minimum((x / (CYCLE / 5)).astype(int), 4)
    alpha = al[step] * (1.0 + (0.0 if m == 2 else 1.0) * x / CYCLE)
    pon = kon[step] / (kon[step] + koff[step])
    g = ga[step]
    bits_decay = np.cumsum((g * dt * LOG2E)[::-1])[::-1] - 0.5 * g * dt * LOG2E      # from the mid-point to t*
    bits_div = -np.floor(t / CYCLE)                                                   # divisions in (t, t*], 0 < t* < cycle
    w = alpha * pon * np.exp2(-np.minimum(bits_decay + bits_div, 1000.0)) * dt
    inw = (t >= tl0) & (t <= tl1) & (PULSE[cond] > 0)
    return t, w * np.where(inw, 1.0 - lam, 1.0), w * np.where(inw, lam, 0.0), (kon[step] + koff[step]) * dt


def missing_share(theta, m, cond, age_i, s0, n_pre=10, dt=0.002):
    """(missing_U, kept_U, missing_L, kept_L) for a start at s0"""
    t, wu, wl, _ = contribution_densities(theta, m, cond, age_i, n_pre, dt)
    before = t < s0
    return wu[before].sum(), wu[~before].sum(), wl[before].sum(), wl[~before].sum()


def burnin_window(theta, m, cond, age_i, n_pre=10, dt=0.002):
    t, wu, wl, kk = contribution_densities(theta, m, cond, age_i, n_pre, dt)
    eps, floor_abs = 2.0 ** -n_pre, 2.0 ** -30
    cu, cl = np.cumsum(wu), np.cumsum(wl)          # cu[i]: births up to the END of cell i
    ok = (cu <= np.maximum(eps * (cu[-1] - cu), floor_abs)) & (cl <= np.maximum(eps * (cl[-1] - cl), floor_abs))
    bad = np.nonzero(~ok)[0]
    i0 = bad[0] if len(bad) else len(t)            # cells [0, i0) can be dropped
    t_lo = -n_pre * CYCLE
    s0 = t_lo + i0 * dt
    if m == 3 and i0 < len(t):
        bits = np.cumsum((kk * LOG2E)[:i0][::-1])  # gene-memory bits going back from s0
        j = np.nonzero(bits >= 6.0)[0]
        s0 = s0 - (j[0] + 1) * dt if len(j) else t_lo
    return max(s0, t_lo)
