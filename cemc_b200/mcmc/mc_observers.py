"""Observers.

``MCObserver`` keeps the reference's callable interface
(/root/reference/cemc/mcmc/mc_observers.py:12-30).  On the GPU path observers
are called at launch boundaries (every ``interval`` steps) with the net
system changes of the chunk; per-step Python callbacks would force one
launch per move ("Python per-step hooks").

``SGCObserver`` (mc_observers.py:185-290) is the one observer the samplers
need every step; its sums are accumulated inside the kernels
(include/cemc_b200.h: cemc_acc_slot) and mirrored here.
"""
import numpy as np

from .averager import Averager


class MCObserver(object):
    def __init__(self):
        self.name = "GenericObserver"

    def __call__(self, system_changes):
        pass

    def reset(self):
        pass

    def get_averages(self):
        return {}


class SGCObserver(MCObserver):
    """Device-backed mirror of the reference SGCObserver."""

    def __init__(self, ce_calc, mc_obj, n_singlets):
        super(SGCObserver, self).__init__()
        self.name = "SGCObersver"          # sic, mc_observers.py:197
        self.ce_calc = ce_calc
        self.mc = mc_obj
        self.recycle_waste = getattr(mc_obj, "recycle_waste", False)
        e0 = ce_calc.get_energy()
        self.quantities = {
            "singlets": np.zeros(n_singlets, dtype=np.float64),
            "singlets_sq": np.zeros(n_singlets, dtype=np.float64),
            "energy": Averager(ref_value=e0),
            "energy_sq": Averager(ref_value=e0),
            "singl_eng": np.zeros(n_singlets, dtype=np.float64),
            "counter": 0,
        }

    def reset(self):
        self.quantities["singlets"][:] = 0.0
        self.quantities["singlets_sq"][:] = 0.0
        self.quantities["energy"].clear()
        self.quantities["energy_sq"].clear()
        self.quantities["singl_eng"][:] = 0.0
        self.quantities["counter"] = 0

    def load_device_sums(self, acc):
        """acc = one replica's accumulator row (CEMC_ACC_* layout)."""
        n = len(self.quantities["singlets"])
        self.quantities["counter"] = int(acc[0])
        self.quantities["energy"].set_sums(acc[1], acc[0])
        self.quantities["energy_sq"].set_sums(acc[2], acc[0])
        for d in range(n):
            self.quantities["singlets"][d] = acc[3 + 3 * d]
            self.quantities["singlets_sq"][d] = acc[4 + 3 * d]
            self.quantities["singl_eng"][d] = acc[5 + 3 * d]

    @property
    def energy(self):
        return self.quantities["energy"]

    @property
    def energy_sq(self):
        return self.quantities["energy_sq"]

    @property
    def singlets(self):
        return self.quantities["singlets"]

    @property
    def singl_eng(self):
        return self.quantities["singl_eng"]

    @property
    def counter(self):
        return self.quantities["counter"]


# ---- observers of the chain state ---------------------------------
# All of them only look at the state on their boundary (energy, CFs, symbols), which is what
# the reference's versions see at that step too.  Two ways of running them:
#   * on the DEVICE (``device_flag``): when every attached observer is one of these and they
#     share one interval, the kernels fold them every ``interval`` steps
#     (include/cemc_b200.h, cemc_set_device_observers) and ``load_device`` mirrors the sums
#     back once per launch chunk -- the device loop never stops for an observer;
#   * on the HOST (``__call__``), like in the reference: the samplers stop the device loop on
#     the boundary, mirror the occupations into ``atoms`` and pass the NET changes since the
#     observer's previous call (``Montecarlo._steps``); used for mixed intervals or when a
#     user-defined observer is attached as well.

class PairCorrelationObserver(MCObserver):
    """Thermal average and spread of the pair correlation functions
    (reference: cemc/mcmc/mc_observers.py:81-136; every ``c2_*`` name of the ECI set)."""

    def __init__(self, ce_calc):
        super(PairCorrelationObserver, self).__init__()
        self.name = "PairCorrelationObserver"
        self.ce_calc = ce_calc
        if getattr(ce_calc, "updater", None) is None:
            raise RuntimeError("This observer needs the CE calculator's updater")
        self._names = [k for k in ce_calc.eci.keys() if k.startswith("c2_")]
        self.reset()

    def reset(self):
        self.cf = {k: 0.0 for k in self._names}
        self.cf_squared = {k: 0.0 for k in self._names}
        self.n_entries = 0

    def __call__(self, system_changes):
        now = self.ce_calc.updater.get_cf()
        self.n_entries += 1
        for k in self._names:
            v = now[k]
            self.cf[k] += v
            self.cf_squared[k] += v * v

    device_flag = 1            # CEMC_OBS_CF_SUMS

    def load_device(self, mc, blk, n, new_energies):
        idx = self.ce_calc.updater.tables.eci_index
        self.n_entries = n
        for k in self._names:
            self.cf[k] = float(blk["cf_sum"][0, idx[k]])
            self.cf_squared[k] = float(blk["cf_sq"][0, idx[k]])

    def get_averages(self):
        return {k: v / self.n_entries for k, v in self.cf.items()}

    def get_std(self):
        n = float(self.n_entries)
        return {k: float(np.sqrt(max(self.cf_squared[k] / n - (self.cf[k] / n) ** 2, 0.0)))
                for k in self._names}


class LowestEnergyStructure(MCObserver):
    """Keeps the lowest-energy state seen on the observer's boundaries
    (reference: mc_observers.py:138-183: first call stores the state, later calls replace
    it when ``mc_obj.current_energy`` is strictly lower)."""

    def __init__(self, ce_calc, mc_obj, verbose=False):
        super(LowestEnergyStructure, self).__init__()
        self.name = "LowestEnergyStructure"
        self.ce_calc = ce_calc
        self.mc_obj = mc_obj
        self.verbose = verbose
        self.reset()

    def reset(self):
        self.lowest_energy = np.inf
        self.lowest_energy_cf = None
        self.atoms = None
        self.lowest_energy_atoms = None        # alias kept by the reference

    def _store(self):
        self.lowest_energy = self.mc_obj.current_energy
        self.lowest_energy_cf = self.ce_calc.get_cf()
        self.atoms = self.mc_obj.atoms.copy()
        self.lowest_energy_atoms = self.atoms

    device_flag = 2            # CEMC_OBS_LOWEST

    def load_device(self, mc, blk, n, new_energies):
        e = float(blk["best_energy"][0])
        if n == 0 or not np.isfinite(e) or (self.atoms is not None and e >= self.lowest_energy):
            return
        tables = self.ce_calc.updater.tables
        self.lowest_energy = e
        self.lowest_energy_cf = {name: float(v) for name, v in zip(tables.eci_names, blk["best_cf"][0])}
        self.atoms = self.mc_obj.atoms.copy()
        for atom, sym in zip(self.atoms, tables.symbols_of(blk["best_occ"][0])):
            atom.symbol = sym
        self.lowest_energy_atoms = self.atoms

    def __call__(self, system_changes):
        if self.atoms is None or self.lowest_energy_cf is None:
            self._store()
        elif self.mc_obj.current_energy < self.lowest_energy:
            dE = self.mc_obj.current_energy - self.lowest_energy
            self._store()
            if self.verbose:
                print("Found new low energy structure. New energy: {} eV. Change: {} eV".format(
                    self.lowest_energy, dE))


class SiteOrderParameter(MCObserver):
    """Number of sites whose species differs from the initial configuration, averaged over
    the calls (reference: mc_observers.py:614-686; symbols stand in for atomic numbers)."""

    def __init__(self, atoms):
        super(SiteOrderParameter, self).__init__()
        self.name = "SiteOrderParameter"
        self.atoms = atoms
        self.orig_symbols = np.array([a.symbol for a in atoms])
        self.reset()

    def _check_all_sites(self):
        now = np.array([a.symbol for a in self.atoms])
        self.site_changed = now != self.orig_symbols
        self.current_num_changed = int(np.count_nonzero(self.site_changed))

    def reset(self):
        self.avg_num_changed = 0
        self.avg_num_changed_sq = 0
        self.num_calls = 0
        self._check_all_sites()

    def __call__(self, system_changes):
        self.num_calls += 1
        for change in system_changes:            # atoms already hold the new symbols
            i = change[0]
            differs = self.atoms[i].symbol != self.orig_symbols[i]
            if differs != bool(self.site_changed[i]):
                self.current_num_changed += 1 if differs else -1
                self.site_changed[i] = differs
        self.avg_num_changed += self.current_num_changed
        self.avg_num_changed_sq += self.current_num_changed ** 2

    device_flag = 8            # CEMC_OBS_SITE_ORDER (reference configuration: ``orig_symbols``)

    def load_device(self, mc, blk, n, new_energies):
        self.num_calls = n
        self.avg_num_changed = float(blk["site_order"][0, 0])
        self.avg_num_changed_sq = float(blk["site_order"][0, 1])

    def get_averages(self):
        avg = float(self.avg_num_changed) / self.num_calls
        var = max(float(self.avg_num_changed_sq) / self.num_calls - avg ** 2, 0.0)
        return {"site_order_average": avg, "site_order_std": float(np.sqrt(var))}


class EnergyEvolution(MCObserver):
    """Energy on every boundary (reference: mc_observers.py:689-706)."""

    def __init__(self, mc_obj):
        super(EnergyEvolution, self).__init__()
        self.name = "EnergyEvolution"
        self.mc = mc_obj
        self.energies = []

    device_flag = 4            # CEMC_OBS_ENERGY

    def load_device(self, mc, blk, n, new_energies):
        self.energies.extend(new_energies)

    def __call__(self, system_changes):
        self.energies.append(self.mc.current_energy_without_vib())

    def reset(self):
        self.energies = []


class EnergyHistogram(MCObserver):
    """Histogram of the sampled energies (reference: mc_observers.py:709-761): the first
    ``buffer_size`` samples fix the range [Emin, Emax], later samples are binned directly.
    The reference sizes its histogram with ``len(self.n_bins)`` (a TypeError for the int it
    documents, :742); here the histogram has ``n_bins`` bins, and samples outside the range
    fixed by the buffer are clamped to the edge bins."""

    def __init__(self, mc_obj, buffer_size=100000, n_bins=100):
        super(EnergyHistogram, self).__init__()
        self.name = "EnergyHistogram"
        self.mc = mc_obj
        self.n_bins = int(n_bins)
        self.buffer = np.zeros(int(buffer_size))
        self.reset()

    def reset(self):
        self._next = 0
        self._histogram = None
        self.Emin = None
        self.Emax = None
        self.sample_in_buffer = True

    def _get_indx(self, E):
        if self.Emin is None or self.Emax is None:
            raise RuntimeError("the histogram range is not fixed yet")
        if self.Emax == self.Emin:
            return 0
        i = int((E - self.Emin) * (self.n_bins - 1) / (self.Emax - self.Emin))
        return min(max(i, 0), self.n_bins - 1)

    def _on_buffer_full(self):
        filled = self.buffer[:self._next] if self._next < len(self.buffer) else self.buffer
        self.Emin = float(np.min(filled))
        self.Emax = float(np.max(filled))
        self._histogram = np.zeros(self.n_bins)
        for e in filled:
            self._histogram[self._get_indx(e)] += 1
        self.sample_in_buffer = False

    device_flag = 4            # CEMC_OBS_ENERGY

    def load_device(self, mc, blk, n, new_energies):
        for E in new_energies:
            self._add(E)

    def __call__(self, system_changes):
        self._add(self.mc.current_energy_without_vib())

    def _add(self, E):
        if self.sample_in_buffer:
            self.buffer[self._next] = E
            self._next += 1
            if self._next >= len(self.buffer):
                self._on_buffer_full()
        else:
            self._histogram[self._get_indx(E)] += 1

    @property
    def histogram(self):
        if self._histogram is None:
            self._on_buffer_full()
        return self._histogram
