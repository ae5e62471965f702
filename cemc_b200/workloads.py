"""The BASELINE.json configurations as synthetic workloads.

All inputs are synthetic and seeded: fcc primitive L^3 lattices, the standard
family set {NN pair, 2NN pair, NN triangle, NN tetrahedron} (K = 18 translation
columns, G = 90 gathered sites per changed site), i.i.d. random occupations.
"""
from __future__ import annotations

import numpy as np

from . import synthetic as syn
from .tables import FlatTables

KB = 8.617330337217213e-05   # eV/K, the value of ase.units.kB (CODATA 2014)


class Workload(object):
    def __init__(self, name, settings, eci, tables, occ, kT, eci_matrix, mode,
                 sites_changed, description):
        self.name = name
        self.settings = settings
        self.eci = eci
        self.tables = tables
        self.occ = occ                  # [R, N] int8
        self.kT = kT                    # [R]
        self.eci_matrix = eci_matrix    # [R, n_eci] or None (same ECIs everywhere)
        self.mode = mode                # "sgc" | "canonical"
        self.sites_changed = sites_changed
        self.description = description

    @property
    def R(self):
        return self.occ.shape[0]


def _occ(st, ft, conc, R, seed0, exact):
    return np.stack([ft.occupancy(syn.random_symbols(st, conc, seed=seed0 + r,
                                                     exact=exact))
                     for r in range(R)])


def c1_almg_canonical(R=1, replica_offset=0):
    """Config 1: Al-Mg fcc binary canonical MC, 4x4x4, pair+triplet+quad ECIs."""
    st = syn.fcc_settings(4, ["Al", "Mg"], ["nn", "2nn", "3nn", "tri", "iso", "tet"])
    eci = syn.almg_ecis(st)
    conc = {"Al": 0.5, "Mg": 0.5}
    ft = FlatTables(st, eci, syn.random_symbols(st, conc, seed=0))
    occ = _occ(st, ft, conc, R, 1000 + replica_offset, True)
    return Workload("C1", st, eci, ft, occ, np.full(R, 500.0 * KB), None,
                    "canonical", 2, "Al-Mg fcc 4x4x4 canonical, T=500 K")


def c2_almg_sgc_sweep(R=256, replica_offset=0, L=10):
    """Config 2: Al-Mg SGC chemical-potential x temperature sweep, 10x10x10,
    256 replicas: 16 T in [200,1000] K x 16 mu in [-1.1,-0.9] eV.  The
    chemical potential is folded into the singlet ECI per replica
    (sgc_montecarlo.py:239-261)."""
    st = syn.fcc_settings(L, ["Al", "Mg"], syn.STANDARD_FAMILIES)
    eci = syn.almg_ecis(st)
    conc = {"Al": 0.5, "Mg": 0.5}
    ft = FlatTables(st, eci, syn.random_symbols(st, conc, seed=0))
    T = np.linspace(200.0, 1000.0, 16)
    mu = np.linspace(-1.1, -0.9, 16)
    g = (np.arange(R) + replica_offset) % 256
    kT = T[g // 16] * KB
    mus = mu[g % 16]
    em = np.tile(ft.eci, (R, 1))
    em[:, ft.eci_index["c1_0"]] -= mus
    occ = _occ(st, ft, conc, R, 2000 + replica_offset, False)
    return Workload("C2", st, eci, ft, occ, kT, em, "sgc", 1,
                    "Al-Mg SGC mu x T sweep, fcc %dx%dx%d, %d replicas" % (L, L, L, R))


def c3_almgsi_canonical(R=64, replica_offset=0, L=20):
    """Config 3: Al-Mg-Si ternary canonical MC, 20x20x20 (8000 sites), up to
    4-body clusters, 64 replicas with T in [300,900] K, composition 80/10/10."""
    st = syn.fcc_settings(L, ["Al", "Mg", "Si"], syn.STANDARD_FAMILIES)
    eci = syn.synthetic_ecis(st, seed=1234)
    conc = {"Al": 0.8, "Mg": 0.1, "Si": 0.1}
    ft = FlatTables(st, eci, syn.random_symbols(st, conc, seed=0))
    g = (np.arange(R) + replica_offset) % 64
    kT = np.linspace(300.0, 900.0, 64)[g] * KB
    occ = _occ(st, ft, conc, R, 3000 + replica_offset, True)
    return Workload("C3", st, eci, ft, occ, kT, None, "canonical", 2,
                    "Al-Mg-Si canonical, fcc %dx%dx%d, %d replicas" % (L, L, L, R))


def c3s_almgsi_sgc(R=64, replica_offset=0, L=20):
    """The north-star target line: 64-replica Al-Mg-Si SGC sweep (same lattice
    and ECIs as config 3, one-site flips, mu folded into both singlets)."""
    w = c3_almgsi_canonical(R, replica_offset, L)
    ft = w.tables
    g = (np.arange(R) + replica_offset) % 64
    em = np.tile(ft.eci, (R, 1))
    em[:, ft.eci_index["c1_0"]] -= np.linspace(-0.05, 0.05, 8)[g % 8]
    em[:, ft.eci_index["c1_1"]] -= np.linspace(-0.05, 0.05, 8)[g // 8]
    return Workload("C3S", w.settings, w.eci, ft, w.occ, w.kT, em, "sgc", 1,
                    "Al-Mg-Si SGC sweep, fcc %dx%dx%d, %d replicas" % (L, L, L, R))


def c4_parallel_tempering(R=64, replica_offset=0, n_total=512, L=12):
    """Config 4: parallel tempering, 512 temperatures (geometric in
    [100,1500] K) x 12x12x12 ternary cell, replicas sharded over ranks."""
    st = syn.fcc_settings(L, ["Al", "Mg", "Si"], syn.STANDARD_FAMILIES)
    eci = syn.synthetic_ecis(st, seed=1234)
    conc = {"Al": 0.8, "Mg": 0.1, "Si": 0.1}
    ft = FlatTables(st, eci, syn.random_symbols(st, conc, seed=0))
    kT_all = np.geomspace(1500.0, 100.0, n_total) * KB     # slot 0 = Tmax
    occ = _occ(st, ft, conc, R, 4000 + replica_offset, True)
    w = Workload("C4", st, eci, ft, occ, kT_all[replica_offset:replica_offset + R],
                 None, "canonical", 2,
                 "parallel tempering, %d temperatures, fcc %d^3 ternary" % (n_total, L))
    w.kT_of_slot = kT_all
    return w


def c5_large_supercell(R=1, replica_offset=0, L=64):
    """Config 5: fcc 64x64x64 binary (262 144 sites, 90/10), single-chain exact
    canonical Metropolis (CTA cluster on one chain); R > 1 = seed ensemble."""
    st = syn.fcc_settings(L, ["Al", "Mg"], syn.STANDARD_FAMILIES)
    eci = syn.almg_ecis(st)
    conc = {"Al": 0.9, "Mg": 0.1}
    ft = FlatTables(st, eci, syn.random_symbols(st, conc, seed=0))
    occ = _occ(st, ft, conc, R, 5000 + replica_offset, True)
    return Workload("C5", st, eci, ft, occ, np.full(R, 600.0 * KB), None, "canonical", 2,
                    "fcc %dx%dx%d binary canonical, %d chain(s), T=600 K" % (L, L, L, R))


def make_updater(w: Workload, device=0, replica_offset=0, stream=None, seed=1234):
    """Upload a workload: returns a ready BatchedCEUpdater."""
    from .updater import BatchedCEUpdater
    gpu = BatchedCEUpdater(w.tables, w.R, device=device,
                           replica_offset=replica_offset, stream=stream)
    gpu.set_occupancy(w.occ)
    gpu.recompute_cf()
    gpu.set_kT(w.kT)
    if w.eci_matrix is not None:
        gpu.set_ecis(w.eci_matrix)
    gpu.seed(seed)
    return gpu


WORKLOADS = {"C1": c1_almg_canonical, "C2": c2_almg_sgc_sweep,
             "C3": c3_almgsi_canonical, "C3S": c3s_almgsi_sgc,
             "C4": c4_parallel_tempering, "C5": c5_large_supercell}
