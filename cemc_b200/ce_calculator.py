"""CE calculator -- host-side mirror of the reference's ``cemc.CE``
(/root/reference/cemc/ce_calculator.py:136-617) on top of the GPU updater.

The PUBLIC surface is the reference's (constructor, method and attribute names, argument
meaning, error types and messages): that is the drop-in contract.  The bodies are this
project's own; differences in behaviour (documented in DESIGN.md):

* ``initial_cf=None`` computes the correlation functions on the GPU from their
  definition (the reference delegates to ``ase.clease.CorrFunction``,
  ce_calculator.py:169-175).
* the class does not derive from ``ase.calculators.calculator.Calculator``
  (ASE is not a dependency); it offers the attributes the samplers use.
* the linear vibration correction is out of scope.
* bulk symbol changes upload the occupations and recompute the CFs on the device instead of
  replaying them one ``calculate`` at a time.
"""
from __future__ import annotations

import copy as _copy
import json

import numpy as np

from .tables import SelfInteractionError  # noqa: F401  (re-exported, reference name)
from .updater import PyCEUpdater

_SELF_INTERACTION = ("The simulation cell is so small that the same site is present multiple times "
                     "within one cluster. Increase the size of the simulation cell.")
# per-call limit of the trial log behind PyCEUpdater.calculate; larger symbol changes go in bulk
_MAX_INCREMENTAL = 400


def get_max_size_eci(eci):
    """Largest cluster size named in the ECIs: the digit after the leading ``c``
    (ce_calculator.py:120-133)."""
    return max((int(name[1]) for name in eci), default=0)


def _clusters_overlap(cluster_info):
    """True when some sub-cluster lists a site twice or contains its own reference site: the
    cell is too small for the cluster set (ce_calculator.py:596-611)."""
    for per_group in cluster_info:
        for cluster in per_group.values():
            ref = cluster["ref_indx"]
            if any(ref in sub or len(sub) != len(set(sub)) for sub in cluster["indices"]):
                return True
    return False


class CE(object):
    """Class for updating the CE when symbols change (ce_calculator.py:136)."""

    implemented_properties = ["energy"]

    def __init__(self, atoms, BC, eci=None, initial_cf=None, device=0):
        if eci is None:
            raise ValueError("ECIs have to be given")
        if _clusters_overlap(BC.cluster_info):
            raise SelfInteractionError(_SELF_INTERACTION)
        self.BC, self.eci, self.atoms, self.device = BC, eci, atoms, device
        self.results = {}
        self._linear_vib_correction = None
        # the empty cluster always has an ECI, and its CF is one (:161-165)
        if "c0" not in eci:
            eci["c0"] = 0.0
            if initial_cf is not None:
                initial_cf["c0"] = 1.0
        to_list = getattr(BC, "_info_entries_to_list", None)
        if to_list is not None:
            to_list()
        atoms.set_calculator(self)
        self._check_trans_mat_dimensions()
        wanted = [a.symbol for a in atoms]
        self.updater = PyCEUpdater(atoms, BC, initial_cf, eci, device=device)
        # the reference binds these three straight to its updater (:205-208)
        for name in ("clear_history", "undo_changes", "update_cf"):
            setattr(self, name, getattr(self.updater, name))
        self.cf = self.updater.get_cf()
        self.set_symbols(wanted)

    # ---- construction helpers ----------------------------------------------------
    def copy(self):
        """New calculator on a copy of the atoms (ce_calculator.py:217-230).  The settings object
        is deep-copied without the back reference to this calculator."""
        self.atoms.set_calculator(None)
        try:
            settings = _copy.deepcopy(self.BC)
        finally:
            self.atoms.set_calculator(self)
        return CE(self.atoms.copy(), settings, eci=dict(self.eci), initial_cf=self.get_cf(),
                  device=self.device)

    def _check_trans_mat_dimensions(self):
        n_rows = len(self.BC.trans_matrix)          # list of dicts or (N, K) array: one row per site
        n_atoms = len(self.atoms)
        if n_rows != n_atoms:
            raise ValueError("The number of atoms and the dimension of the translation matrix is "
                             "inconsistent\nNum atoms: {}. Num row trans mat: {}".format(n_atoms, n_rows))

    def _has_self_interaction(self, cluster_info):
        return _clusters_overlap(cluster_info)

    # ---- vibration correction: not part of this path ------------------
    @property
    def linear_vib_correction(self):
        return self._linear_vib_correction

    def include_linvib_in_ecis(self, T):
        return None          # no vibration ECIs: does nothing (:288-289)

    def vib_energy(self, T):
        return 0.0

    # ---- energies / correlation functions -------------------------------------------
    def get_energy(self):
        return self.updater.get_energy()

    def calculate(self, atoms, properties, system_changes):
        """Energy after ``system_changes`` [(indx, old_symb, new_symb), ...]
        were applied to the internal atoms (ce_calculator.py:345-364).  Unlike
        the reference, the CF dict is not rebuilt on every call;
        ``get_cf()`` fetches it on request."""
        self.results["energy"] = self.updater.calculate(system_changes)
        return self.results["energy"]

    def get_cf(self):
        self.cf = self.updater.get_cf()
        return self.cf

    def get_singlets(self):
        return self.updater.get_singlets()

    def update_ecis(self, new_ecis):
        self.eci = new_ecis
        self.updater.set_ecis(new_ecis)

    def set_num_threads(self, num_threads):
        self.updater.set_num_threads(num_threads)

    # ---- changing the configuration -----------------------------------------------
    def set_composition(self, comp):
        """Change the composition, e.g. {"Mg": 0.2, "Al": 0.8} (ce_calculator.py:397-434): the
        most abundant element fills the cell, the others take consecutive blocks of
        round(conc N) sites from the start, in the order of ``comp``."""
        if abs(sum(comp.values()) - 1.0) > 1e-6:
            raise ValueError("The specified concentration does not sum to 1!")
        n = len(self.atoms)
        host = max(comp, key=comp.get)              # first of equals, like a strict '>' scan
        symbols = np.full(n, host, dtype=object)
        first = 0
        for element, conc in comp.items():
            if element == host:
                continue
            count = int(round(conc * n))
            symbols[first:first + count] = element
            first += count
        self._set_symbols_bulk(list(symbols))

    def set_symbols(self, symbs):
        """Change the symbols of the entire atoms object (:436-448)."""
        if len(symbs) != len(self.atoms):
            raise ValueError("Length of the symbols array has to match"
                             "the length of the atoms object.!")
        changes = [(i, atom.symbol, new) for i, (atom, new) in enumerate(zip(self.atoms, symbs))
                   if atom.symbol != new]
        if len(changes) > _MAX_INCREMENTAL:
            self._set_symbols_bulk(symbs)
            return
        if changes:
            self.updater.calculate(changes)
        self.clear_history()

    def _set_symbols_bulk(self, symbols):
        """Upload the whole configuration and recompute the CFs from their definition."""
        upd = self.updater
        upd.batch.set_occupancy(upd.tables.occupancy(symbols)[None, :])
        upd.batch.recompute_cf()
        upd._log = []
        for atom, symbol in zip(self.atoms, symbols):
            atom.symbol = symbol

    def singlet2comp(self, singlets):
        """Convert singlets ``{"c1_<d>": value}`` to concentrations (ce_calculator.py:450-518)."""
        from .mcmc.stats import concentrations_from_named_singlets
        return concentrations_from_named_singlets(self.BC.basis_functions, singlets)

    def set_singlets(self, singlets):
        self.set_composition(self.singlet2comp(singlets))

    # ---- checkpoint: the reference's dictionary keys (ce_calculator.py:533-594) -------------
    def backup_dict(self):
        settings = dict(getattr(self.BC, "kwargs", {}), classtype=type(self.BC).__name__)
        return {"cf": self.get_cf(), "symbols": [a.symbol for a in self.atoms],
                "setting_kwargs": settings, "eci": self.eci}

    def save(self, fname):
        with open(fname, "w") as out:
            json.dump(self.backup_dict(), out, indent=2, separators=(",", ": "))

    @staticmethod
    def load(fname):
        with open(fname) as f:
            return CE.load_from_dict(json.load(f))

    @staticmethod
    def load_from_dict(backup_data):
        from . import synthetic as syn
        kw = dict(backup_data["setting_kwargs"])
        classtype = kw.pop("classtype")
        if classtype != "SyntheticSettings":
            raise ValueError("Unknown setting classtype: {}".format(classtype))
        bc = syn.fcc_settings(kw["size"][0], kw["species"], kw["families"])
        atoms = syn.Atoms(backup_data["symbols"])
        return CE(atoms, bc, eci=backup_data["eci"], initial_cf=backup_data["cf"])

    def __reduce__(self):
        return (CE.load_from_dict, (self.backup_dict(),))


def get_atoms_with_ce_calc(small_bc, bc_kwargs, eci=None, size=[1, 1, 1],
                           db_name="temp_db.db", device=0):
    """CE calculator for a supercell (ce_calculator.py:48-102): correlation
    functions are intensive, so they are evaluated on the small cell and
    handed to the large one.  ``bc_kwargs`` are the arguments of
    ``synthetic.fcc_settings``; ``size`` multiplies the small cell."""
    from . import synthetic as syn
    limit = bc_kwargs.get("max_cluster_size")
    if limit is not None and get_max_size_eci(eci) > limit:
        raise ValueError("ECI specifies a cluster size larger than "
                         "ClusterExpansionSetting tracks!")
    # a uniform small cell tiles to a uniform large cell with the same CFs
    element = small_bc.atoms[0].symbol
    if any(a.symbol != element for a in small_bc.atoms):
        raise ValueError("the small cell must be uniformly occupied")
    small_cf = CE(small_bc.atoms.copy(), small_bc, dict(eci), device=device).get_cf()
    edge = int(bc_kwargs["size"][0]) * int(size[0])
    large_bc = syn.fcc_settings(edge, bc_kwargs["species"], bc_kwargs["families"])
    atoms = large_bc.atoms.copy()
    for a in atoms:
        a.symbol = element
    CE(atoms, large_bc, eci, initial_cf=small_cf, device=device)
    return atoms
