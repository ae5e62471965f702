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
