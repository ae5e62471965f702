"""Chain statistics behind the samplers' equilibration / convergence decisions.

Small, array-valued helpers (every argument may carry a leading replica axis) that the
sampler classes call; they take the SUMS the kernels accumulate on the device
(include/cemc_b200.h ``cemc_acc_slot``, ``cemc_energy_autocorrelation``), so no per-step
data ever has to reach the host.  Behaviour follows the reference's samplers; each function
names the lines whose result it reproduces (paths relative to /root/reference).
"""
from __future__ import annotations

import math

import numpy as np

LN2 = math.log(2.0)


def normal_quantile(p):
    """Inverse CDF of the standard normal distribution."""
    from scipy.special import ndtri
    return float(ndtri(p))


def correlation_time(first_half_lag, window_length):
    """Correlation time from the first lag at which the normalised energy autocorrelation
    drops below 1/2 (Van de Walle & Asta, MSMSE 10 (2002) 521: the ACF is modelled as
    rho^k, so rho = 2^(-1/k) and tau = -1/ln(rho) = k / ln 2).  A lag of -1 means that no
    lag of the window went below 1/2: the window is too short and its length stands in.
    Same numbers as cemc/mcmc/montecarlo.py:487-511.

    Returns ``(tau, found)``."""
    k = np.asarray(first_half_lag, dtype=np.float64)
    found = k > 0
    with np.errstate(divide="ignore", invalid="ignore"):
        rho = np.power(2.0, -1.0 / np.where(found, k, 1.0))
        tau = np.where(found, -1.0 / np.log(rho), float(window_length))
    return tau, found


def variance_of_mean(mean, mean_sq, n, tau=None, fold_negative=True):
    """Variance of the mean of n correlated samples: (<x^2> - <x>^2) / n, times 2 tau when a
    correlation time is known (tau is clamped to >= 1).  A slightly negative sample variance
    from rounding is folded back (always for the energy, montecarlo.py:1076-1100; for the
    singlets only when a correlation time is used, sgc_montecarlo.py:94-132)."""
    n = np.asarray(n, dtype=np.float64)
    mean = np.asarray(mean, dtype=np.float64)
    var = np.asarray(mean_sq, dtype=np.float64) - mean * mean
    if tau is None:
        return (np.abs(var) if fold_negative else var) / n
    return 2.0 * np.abs(var) * np.maximum(np.asarray(tau, dtype=np.float64), 1.0) / n


def z_score(mean_new, var_new, mean_old, var_old, floor=1e-6):
    """z value of the difference of two window means under the null hypothesis that both
    windows sample the same distribution; windows with (numerically) no variance give
    z = 0 and ``frozen`` = True -- the system does not move (montecarlo.py:645-652)."""
    var_diff = np.asarray(var_new, dtype=np.float64) + np.asarray(var_old, dtype=np.float64)
    frozen = var_diff < floor
    with np.errstate(divide="ignore", invalid="ignore"):
        z = np.where(frozen, 0.0, (np.asarray(mean_new) - np.asarray(mean_old)) / np.sqrt(np.where(frozen, 1.0, var_diff)))
    return z, frozen


def inside(z, confidence_level):
    """Two-sided acceptance region of the window test: the probability of an even larger
    difference exceeds ``confidence_level`` (montecarlo.py:670)."""
    lo, hi = normal_quantile(confidence_level), normal_quantile(1.0 - confidence_level)
    return (np.asarray(z) > lo) & (np.asarray(z) < hi)


def singlets_agree(singlets, var, prev_singlets, prev_var, confidence_level):
    """Largest z value over the singlets whose variance is non-zero, and whether it lies inside
    the acceptance region; no moving singlet at all counts as agreement
    (sgc_montecarlo.py:172-210).  Returns ``(agree, z_max)``."""
    var = np.maximum(np.asarray(var, dtype=np.float64), 0.0)
    var_diff = var + np.asarray(prev_var, dtype=np.float64)
    moving = var_diff > 0.0
    if not np.any(moving):
        return True, 0.0
    diff = np.abs(np.asarray(singlets) - np.asarray(prev_singlets))[moving]
    z = float(np.max(diff / np.sqrt(var_diff[moving])))
    return bool(inside(z, confidence_level)), z


def concentrations_from_singlets(basis_functions, symbols, singlets):
    """Concentrations x_s from the average singlets <phi_d> = sum_s x_s phi_d(s) together with
    sum_s x_s = 1 (sgc_montecarlo.py:380-396).  ``basis_functions``: list of dicts
    symbol -> value; returns a dict ``{symbol + "_conc": x}``."""
    S = len(symbols)
    A = np.ones((S, S))
    for d in range(S - 1):
        A[d, :] = [basis_functions[d][s] for s in symbols]
    rhs = np.append(np.asarray(singlets, dtype=np.float64)[:S - 1], 1.0)
    x = np.linalg.solve(A, rhs)
    return {s + "_conc": float(v) for s, v in zip(symbols, x)}


def concentrations_from_named_singlets(basis_functions, singlets, eps=1e-6):
    """The same linear problem in the form ``CE.singlet2comp`` exposes it
    (ce_calculator.py:450-518): singlets come as ``{"c1_<d>": value}``, the first species of
    the basis functions is eliminated through the closure relation, concentrations within
    ``eps`` below zero are clipped and anything outside [0, 1] is an error."""
    D = len(basis_functions)
    if len(singlets) != D:
        raise ValueError("The number singlet terms specified is different "
                         "from the number of basis functions")
    elements = list(basis_functions[0].keys())
    first, others = elements[0], elements[1:]
    phi = np.array([[bf[e] for e in elements] for bf in basis_functions], dtype=np.float64)   # [D, S]
    target = np.zeros(D)
    for name, value in singlets.items():
        target[int(name[-1])] = value
    # <phi_d> - phi_d(first) = sum_{s != first} x_s (phi_d(s) - phi_d(first))
    x = np.linalg.solve(phi[:, 1:] - phi[:, :1], target - phi[:, 0])
    x[(x < 0.0) & (x > -eps)] = 0.0
    x_first = 1.0 - float(np.sum(x))
    if -eps < x_first < 0.0:
        x_first = 0.0
    if not (0.0 <= x_first <= 1.0) or np.any(x > 1.0) or np.any(x < 0.0):
        raise RuntimeError("Something went wrong when converting "
                           "singlets to composition")
    out = {first: x_first}
    out.update({e: float(v) for e, v in zip(others, x)})
    return {e: out[e] for e in elements}
