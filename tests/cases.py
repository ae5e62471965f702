"""Shared problem builders for the tests (synthetic fcc cluster expansions)."""
import json
import os

import numpy as np

from cemc_b200 import synthetic as syn
from cemc_b200.tables import FlatTables

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN = ["almg_fcc4_canonical", "almg_fcc4_sgc", "almgsi_fcc4_canonical",
          "almgsi_fcc4_sgc", "almgsi_fcc5_canonical_cold",
          "almgsi_layered4_sgc", "almgsi_layered6_canonical"]     # two symmetry groups

# BASELINE-size fixtures: replicas of the bench workloads (configs[1], configs[2] and the
# north-star 64-replica Al-Mg-Si SGC sweep) recorded from the compiled reference
GOLDEN_WORKLOADS = ["c2_fcc10_sgc", "c3s_fcc20_sgc", "c3_fcc20_canonical"]

KB = 8.617330337217213e-05   # eV/K (ase.units.kB, CODATA 2014)


def build(L, species, families, conc, eci_kind="synthetic", seed=3,
          trans_matrix_format="auto"):
    if families == "layered":
        st = syn.layered_settings(L, species)
    else:
        st = syn.fcc_settings(L, species, families,
                              trans_matrix_format=trans_matrix_format)
    eci = syn.almg_ecis(st) if eci_kind == "almg" else \
        syn.synthetic_ecis(st, seed=1234)
    symbols = syn.random_symbols(st, conc, seed=seed)
    ft = FlatTables(st, eci, symbols)
    return st, eci, symbols, ft


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    meta = json.loads(str(z["meta"]))
    st = syn.layered_settings(meta["L"], meta["species"]) if meta["families"] == "layered" else \
        syn.fcc_settings(meta["L"], meta["species"], meta["families"])
    ft = FlatTables(st, meta["eci"], meta["symbols0"])
    assert ft.eci_names == meta["eci_names"]
    assert ft.species == meta["species_sorted"]
    return meta, st, ft, z


_WL_CACHE = {}


def load_golden_workload(name):
    """(meta, tables, arrays) of a multi-replica BASELINE-size fixture; the tables are
    rebuilt from cemc_b200.workloads and checked against what the fixture recorded."""
    if name in _WL_CACHE:
        return _WL_CACHE[name]
    from cemc_b200 import workloads as wl
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    meta = json.loads(str(z["meta"]))
    ws = [wl.WORKLOADS[meta["workload"]](R=1, replica_offset=g) for g in meta["replicas"]]
    ft = ws[0].tables
    assert ft.eci_names == meta["eci_names"] and ft.species == meta["species_sorted"]
    for r, w in enumerate(ws):      # the generator still produces the recorded inputs
        assert np.array_equal(w.occ[0], z["occ0"][r])
        assert float(w.kT[0]) == float(z["kT"][r])
        assert np.array_equal(w.eci_matrix[0] if w.eci_matrix is not None else ft.eci, z["eci"][r])
    meta["mode"] = ws[0].mode
    _WL_CACHE[name] = (meta, ft, z)
    return _WL_CACHE[name]


BINARY = dict(L=4, species=["Al", "Mg"], families=["nn", "2nn", "tri", "tet"],
              conc={"Al": 0.5, "Mg": 0.5})
# two translational symmetry groups with different cluster families (synthetic.layered_settings)
LAYERED = dict(L=4, species=["Al", "Mg", "Si"], families="layered",
               conc={"Al": 0.4, "Mg": 0.3, "Si": 0.3})
LAYERED_BINARY = dict(L=4, species=["Al", "Mg"], families="layered", conc={"Al": 0.5, "Mg": 0.5})
TERNARY = dict(L=4, species=["Al", "Mg", "Si"],
               families=["nn", "2nn", "tri", "iso", "tet"],
               conc={"Al": 0.5, "Mg": 0.25, "Si": 0.25})
