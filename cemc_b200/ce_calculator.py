"""CE calculator -- host-side mirror of the reference's ``cemc.CE``
(/root/reference/cemc/ce_calculator.py:136-617) on top of the GPU updater.

Same constructor and method names; differences (documented in DESIGN.md):

* ``initial_cf=None`` computes the correlation functions on the GPU from their
  definition (the reference delegates to ``ase.clease.CorrFunction``,
  ce_calculator.py:169-175).
* the class does not derive from ``ase.calculators.calculator.Calculator``
  (ASE is not a dependency); it offers the attributes the samplers use.
* the linear vibration correction is out of scope.
"""
from __future__ import annotations

import json

import numpy as np

from .tables import SelfInteractionError  # noqa: F401  (re-exported, reference name)
from .updater import PyCEUpdater


def get_max_size_eci(eci):
    """Maximum cluster size named in the ECIs (ce_calculator.py:120-133)."""
    max_size = 0
    for key in eci.keys():
        size = int(key[1])
        if size > max_size:
            max_size = size
    return max_size


class CE(object):
    """Class for updating the CE when symbols change (ce_calculator.py:136)."""

    implemented_properties = ["energy"]

    def __init__(self, atoms, BC, eci=None, initial_cf=None, device=0):
        self.BC = BC
        self.results = {}
        if eci is None:
            raise ValueError("ECIs have to be given")
        if self._has_self_interaction(BC.cluster_info):
            raise SelfInteractionError(
                'The simulation cell is so small that the same site '
                'is present multiple times within one cluster. '
                'Increase the size of the simulation cell.')
        # make sure there is an ECI for the empty cluster (:161-165)
        if 'c0' not in eci.keys():
            eci['c0'] = 0.0
            if initial_cf is not None:
                initial_cf['c0'] = 1.0
        if hasattr(self.BC, "_info_entries_to_list"):
            self.BC._info_entries_to_list()
        self.eci = eci
        self.atoms = atoms
        self.atoms.set_calculator(self)
        symbols = [atom.symbol for atom in self.atoms]
        self._check_trans_mat_dimensions()
        self.device = device
        self.updater = PyCEUpdater(self.atoms, self.BC, initial_cf, self.eci,
                                   device=device)
        self.cf = self.updater.get_cf()
        # bound straight to the updater like the reference (:205-208)
        self.clear_history = self.updater.clear_history
        self.undo_changes = self.updater.undo_changes
        self.update_cf = self.updater.update_cf
        self.set_symbols(symbols)
        self._linear_vib_correction = None

    # ------------------------------------------------------------------
    def copy(self):
        """New calculator on a copy of the atoms (ce_calculator.py:217-230)."""
        from copy import deepcopy
        self.atoms.set_calculator(None)
        new_bc = deepcopy(self.BC)
        self.atoms.set_calculator(self)
        atoms = self.atoms.copy()
        return CE(atoms, new_bc, eci=dict(self.eci), initial_cf=self.get_cf(),
                  device=self.device)

    def _check_trans_mat_dimensions(self):
        tm = self.BC.trans_matrix
        n_sites = len(tm) if isinstance(tm, list) else tm.shape[0]
        if len(self.atoms) != n_sites:
            msg = "The number of atoms and the dimension of the translation "
            msg += "matrix is inconsistent\n"
            msg += "Num atoms: {}. ".format(len(self.atoms))
            msg += "Num row trans mat: {}".format(n_sites)
            raise ValueError(msg)

    @property
    def linear_vib_correction(self):
        return self._linear_vib_correction

    def include_linvib_in_ecis(self, T):
        return None          # no vibration ECIs: does nothing (:288-289)

    def vib_energy(self, T):
        return 0.0

    def get_energy(self):
        return self.updater.get_energy()

    def calculate(self, atoms, properties, system_changes):
        """Energy after ``system_changes`` [(indx, old_symb, new_symb), ...]
        were applied to the internal atoms (ce_calculator.py:345-364).  Unlike
        the reference, the CF dict is not rebuilt on every call;
        ``get_cf()`` fetches it on request."""
        energy = self.updater.calculate(system_changes)
        self.results["energy"] = energy
        return self.results["energy"]

    def get_cf(self):
        self.cf = self.updater.get_cf()
        return self.cf

    def update_ecis(self, new_ecis):
        self.eci = new_ecis
        self.updater.set_ecis(self.eci)

    def get_singlets(self):
        return self.updater.get_singlets()

    def set_composition(self, comp):
        """Change the composition, e.g. {"Mg": 0.2, "Al": 0.8}
        (ce_calculator.py:397-434)."""
        tot_conc = 0.0
        max_element = None
        max_conc = 0.0
        for key, conc in comp.items():
            tot_conc += conc
            if conc > max_conc:
                max_element = key
                max_conc = conc
        if np.abs(tot_conc - 1.0) > 1E-6:
            raise ValueError("The specified concentration does not sum to 1!")
        init_elm = max_element
        symbols = [init_elm] * len(self.atoms)
        start = 0
        for elm, conc in comp.items():
            if elm == init_elm:
                continue
            n_at = int(round(conc * len(self.atoms)))
            for i in range(start, start + n_at):
                symbols[i] = elm
            start += n_at
        self._set_symbols_bulk(symbols)

    def set_symbols(self, symbs):
        """Change the symbols of the entire atoms object (:436-448)."""
        if len(symbs) != len(self.atoms):
            raise ValueError(
                "Length of the symbols array has to match"
                "the length of the atoms object.!")
        changes = [(i, self.atoms[i].symbol, s) for i, s in enumerate(symbs)
                   if self.atoms[i].symbol != s]
        if len(changes) > 400:
            # bulk path: upload + recompute from the definition
            self._set_symbols_bulk(symbs)
            return
        for i in range(0, len(changes), 400):
            self.updater.calculate(changes[i:i + 400])
        self.clear_history()

    def _set_symbols_bulk(self, symbols):
        upd = self.updater
        upd.batch.set_occupancy(upd.tables.occupancy(symbols)[None, :])
        upd.batch.recompute_cf()
        upd._log = []
        for atom, s in zip(self.atoms, symbols):
            atom.symbol = s

    def singlet2comp(self, singlets):
        """Convert singlets ``{"c1_<d>": value}`` to concentrations (ce_calculator.py:450-518)."""
        from .mcmc.stats import concentrations_from_named_singlets
        return concentrations_from_named_singlets(self.BC.basis_functions, singlets)

    def set_singlets(self, singlets):
        self.set_composition(self.singlet2comp(singlets))

    # ---- checkpoint (ce_calculator.py:533-594) -----------------------------
    def backup_dict(self):
        backup_data = {}
        backup_data["cf"] = self.get_cf()
        backup_data["symbols"] = [atom.symbol for atom in self.atoms]
        backup_data["setting_kwargs"] = dict(getattr(self.BC, "kwargs", {}))
        backup_data["setting_kwargs"]["classtype"] = type(self.BC).__name__
        backup_data["eci"] = self.eci
        return backup_data

    def save(self, fname):
        with open(fname, 'w') as outfile:
            json.dump(self.backup_dict(), outfile, indent=2,
                      separators=(",", ": "))

    @staticmethod
    def load(fname):
        with open(fname, 'r') as infile:
            backup_data = json.load(infile)
        return CE.load_from_dict(backup_data)

    @staticmethod
    def load_from_dict(backup_data):
        from . import synthetic as syn
        kw = dict(backup_data["setting_kwargs"])
        classtype = kw.pop("classtype")
        if classtype != "SyntheticSettings":
            raise ValueError("Unknown setting classtype: {}".format(classtype))
        bc = syn.fcc_settings(kw["size"][0], kw["species"], kw["families"])
        atoms = syn.Atoms(backup_data["symbols"])
        return CE(atoms, bc, eci=backup_data["eci"],
                  initial_cf=backup_data["cf"])

    def __reduce__(self):
        return (CE.load_from_dict, (self.backup_dict(),))

    def _has_self_interaction(self, cluster_info):
        for info in cluster_info:
            for k, cluster in info.items():
                for sub in cluster['indices']:
                    if cluster['ref_indx'] in sub:
                        return True
                    if len(set(sub)) != len(sub):
                        return True
        return False

    def set_num_threads(self, num_threads):
        self.updater.set_num_threads(num_threads)


def get_atoms_with_ce_calc(small_bc, bc_kwargs, eci=None, size=[1, 1, 1],
                           db_name="temp_db.db", device=0):
    """CE calculator for a supercell (ce_calculator.py:48-102): correlation
    functions are intensive, so they are evaluated on the small cell and
    handed to the large one.  ``bc_kwargs`` are the arguments of
    ``synthetic.fcc_settings``; ``size`` multiplies the small cell."""
    from . import synthetic as syn
    max_size_eci = get_max_size_eci(eci)
    if "max_cluster_size" in bc_kwargs and max_size_eci > bc_kwargs["max_cluster_size"]:
        raise ValueError("ECI specifies a cluster size larger than "
                         "ClusterExpansionSetting tracks!")
    atoms = small_bc.atoms.copy()
    calc1 = CE(atoms, small_bc, dict(eci), device=device)
    init_cf = calc1.get_cf()
    L = int(bc_kwargs["size"][0]) * int(size[0])
    large_bc = syn.fcc_settings(L, bc_kwargs["species"], bc_kwargs["families"])
    atoms = large_bc.atoms.copy()
    # a uniform small cell tiles to a uniform large cell with the same CFs
    first = small_bc.atoms[0].symbol
    if any(a.symbol != first for a in small_bc.atoms):
        raise ValueError("the small cell must be uniformly occupied")
    for a in atoms:
        a.symbol = first
    CE(atoms, large_bc, eci, initial_cf=init_cf, device=device)
    return atoms
