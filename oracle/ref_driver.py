"""TEST INFRASTRUCTURE -- drives the reference's OWN compiled CEUpdater.

``oracle/_ref/cemc_cpp_code*.so`` is built by oracle/build_ref.sh from the
unmodified sources under /root/reference (plus 4 NumPy-2 casts).  The
``cemc`` Python package itself cannot be imported here (needs ase,
ase.clease, h5py ...), so the sampler rules it applies around the updater are
restated below, each citing its reference line:

  * accept rule      cemc/mcmc/montecarlo.py:951-956
  * commit/rollback  cemc/mcmc/montecarlo.py:1015-1018
  * calculate        cemc/ce_calculator.py:345-364

``ase.clease.tools.equivalent_deco`` (called from cpp/src/cluster.cpp:78-81)
is provided by a stub module built from cemc_b200.synthetic.equivalent_deco.
"""
from __future__ import annotations

import contextlib
import ctypes
import glob
import math
import os
import sys
import types

import numpy as np

from cemc_b200.synthetic import Atoms, equivalent_deco

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF_DIR = os.path.join(_HERE, "_ref")


def available() -> bool:
    return bool(glob.glob(os.path.join(_REF_DIR, "cemc_cpp_code*.so")))


def _install_stub_modules():
    if "ase.clease.tools" in sys.modules:
        return
    ase = sys.modules.get("ase") or types.ModuleType("ase")
    clease = types.ModuleType("ase.clease")
    tools = types.ModuleType("ase.clease.tools")
    tools.equivalent_deco = equivalent_deco
    clease.tools = tools
    ase.clease = clease
    sys.modules.setdefault("ase", ase)
    sys.modules["ase.clease"] = clease
    sys.modules["ase.clease.tools"] = tools


@contextlib.contextmanager
def _quiet():
    """The reference prints progress from C++ (#define CE_DEBUG)."""
    sys.stdout.flush()
    sys.stderr.flush()
    devnull = os.open(os.devnull, os.O_WRONLY)
    o1, o2 = os.dup(1), os.dup(2)
    os.dup2(devnull, 1)
    os.dup2(devnull, 2)
    try:
        yield
    finally:
        os.dup2(o1, 1)
        os.dup2(o2, 2)
        os.close(devnull)
        os.close(o1)
        os.close(o2)


def load_module():
    if not available():
        raise ImportError("oracle/_ref is not built (run oracle/build_ref.sh)")
    _install_stub_modules()
    if _REF_DIR not in sys.path:
        sys.path.insert(0, _REF_DIR)
    import cemc_cpp_code  # noqa
    return cemc_cpp_code


class RefChain(object):
    """One chain on the reference's compiled ``PyCEUpdater``."""

    def __init__(self, settings, symbols, eci: dict, cf: dict, kT=1.0,
                 num_threads=1):
        mod = load_module()
        self.atoms = Atoms(list(symbols))
        self.settings = settings
        self.eci = dict(eci)
        self.cf0 = dict(cf)
        self._keep = (self.atoms, self.settings, self.eci, self.cf0)
        # the reference stores `atoms` without INCREF and DECREFs it in its
        # destructor (ce_updater.cpp:34,22): hold one extra reference forever.
        _LEAK.append(self.atoms)
        ctypes.pythonapi.Py_IncRef(ctypes.py_object(self.atoms))
        with _quiet():
            self.upd = mod.PyCEUpdater(self.atoms, settings, self.cf0, self.eci)
            self.upd.set_num_threads(int(num_threads))
        # never run the reference's destructor (it would DECREF freed objects
        # during interpreter shutdown): make the updater immortal.
        ctypes.pythonapi.Py_IncRef(ctypes.py_object(self.upd))
        _LEAK.append(self.upd)
        self.kT = float(kT)
        self.current_energy = self.upd.get_energy()  # montecarlo.py:753
        self.n_accepted = 0
        self.names = sorted(self.eci.keys())

    def symbols(self):
        return [a.symbol for a in self.atoms]

    def cf_vector(self):
        cf = self.upd.get_cf()
        return np.array([cf[n] for n in self.names], dtype=np.float64)

    def set_ecis(self, eci: dict):
        self.eci = dict(eci)
        self.upd.set_ecis(self.eci)
        self.current_energy = self.upd.get_energy()

    def trial(self, changes, u):
        """One _mc_step without constraints/bias/observers."""
        new_energy = self.upd.calculate(changes)        # ce_calculator.py:361
        if new_energy < self.current_energy:            # montecarlo.py:951
            accept, used = True, False
        else:
            energy_diff = new_energy - self.current_energy
            probability = math.exp(-energy_diff / self.kT)
            accept, used = (u <= probability), True     # montecarlo.py:956
        if accept:
            self.current_energy = new_energy
            self.n_accepted += 1
            self.upd.clear_history()                    # montecarlo.py:1016
        else:
            self.upd.undo_changes()                     # montecarlo.py:1018
        return accept, used

    def replay(self, species, sites, news, u):
        """Replay recorded proposals; returns (accepted, u_used, e_after)."""
        n = len(u)
        acc = np.zeros(n, dtype=np.uint8)
        used = np.zeros(n, dtype=np.uint8)
        e = np.zeros(n, dtype=np.float64)
        atoms = self.atoms
        for s in range(n):
            ch = []
            a = int(sites[s][0])
            b = int(sites[s][1])
            if b < 0:
                ch = [(a, atoms[a].symbol, species[int(news[s][0])])]
            else:
                # old symbol of the 2nd change is read BEFORE the first is
                # applied, exactly like Montecarlo._get_trial_move (:906-907)
                ch = [(a, atoms[a].symbol, species[int(news[s][0])]),
                      (b, atoms[b].symbol, species[int(news[s][1])])]
            ok, was_used = self.trial(ch, float(u[s]))
            acc[s], used[s], e[s] = ok, was_used, self.current_energy
        return acc, used, e


_LEAK = []
