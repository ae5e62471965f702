"""TEST INFRASTRUCTURE -- CPU restatement of the reference samplers' equilibration logic,
driven by oracle chains (oracle/ce_oracle.py).  Only tests/ import this.

Restates, with the line each step follows (paths relative to /root/reference):

  * Montecarlo._estimate_correlation_time   cemc/mcmc/montecarlo.py:461-511
  * Montecarlo._get_var_average_energy      cemc/mcmc/montecarlo.py:1076-1100
  * Montecarlo._equillibriate               cemc/mcmc/montecarlo.py:541-697

The per-step sums (mean_energy += E, energy_squared += E**2, :622-624) are the oracle
chain's Averager accumulators, which the C oracle adds in the reference's order.
"""
import numpy as np
from scipy import stats as sp_stats


def correlation_info(energies, window_length):
    """montecarlo.py:471-511 on a recorded energy window."""
    energies = np.asarray(energies, dtype=np.float64)
    mean = np.mean(energies)                                        # :471
    dev = energies - mean                                           # :472
    var = np.var(dev)                                               # :473
    acf = np.correlate(dev, dev, mode="full")                       # :474
    acf = acf[int(len(acf) / 2):]                                   # :475
    info = {"correlation_time_found": False, "correlation_time": 0.0}
    if var == 0.0:                                                  # :482-488
        info["correlation_time_found"] = True
        info["correlation_time"] = window_length
        return info
    acf = acf / (window_length * var)                               # :490
    if np.min(acf) > 0.5:                                           # :491-496
        info["correlation_time"] = window_length
        return info
    indx = 0
    for i in range(len(acf)):                                       # :505-508
        if acf[i] < 0.5:
            indx = i
            break
    rho = 2.0 ** (-1.0 / indx)                                      # :509
    info["correlation_time"] = -1.0 / np.log(rho)                   # :510
    info["correlation_time_found"] = True
    return info


def var_average_energy(mean, mean_sq, n_steps, info):
    """montecarlo.py:1076-1100."""
    var = mean_sq - mean ** 2
    if var < 0.0:
        var = np.abs(var)
    if info is None or not info["correlation_time_found"]:
        return var / n_steps
    tau = info["correlation_time"]
    if tau < 1.0:
        tau = 1.0
    return 2.0 * var * tau / n_steps


def equilibrate(chain, run, window_length, confidence_level, info, ref, maxiter=1000):
    """montecarlo.py:541-697, mode "stat_equiv", fixed composition: returns the list of
    (E_new, var_E_new, z_diff or None) per window; the last entry is the accepted one."""
    lo = sp_stats.norm.ppf(confidence_level)                        # :588
    hi = sp_stats.norm.ppf(1.0 - confidence_level)                  # :589
    history = []
    E_prev = var_prev = None
    energy_conv = False
    chain.set_ref(ref)
    for _ in range(maxiter):
        chain.reset_acc()                                           # self.reset(), :618
        run(window_length)                                          # :619-624
        n, s1, s2 = chain.acc[0], chain.acc[1], chain.acc[2]
        E_new = (s1 / n) * ref                                      # Averager.mean, averager.py:60-66
        E_sq = (s2 / n) * ref
        var_new = var_average_energy(E_new, E_sq, window_length, info)   # :627
        if E_prev is None:                                          # :631-634
            E_prev, var_prev = E_new, var_new
            history.append((E_new, var_new, None))
            continue
        var_diff = var_new + var_prev                               # :636
        diff = E_new - E_prev                                       # :637
        z = 0.0 if var_diff < 1e-6 else diff / np.sqrt(var_diff)    # :638-643
        history.append((E_new, var_new, z))
        if lo < z < hi:                                             # :658-659
            energy_conv = True
        if energy_conv:                                             # :661 (composition: always True)
            return history
        E_prev, var_prev = E_new, var_new                           # :694-695
    raise RuntimeError("Did not manage to reach equillibrium!")
