"""Shared problem builders for the tests (synthetic fcc cluster expansions)."""
import json
import os

import numpy as np

from cemc_b200 import synthetic as syn
from cemc_b200.tables import FlatTables

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN = ["almg_fcc4_canonical", "almg_fcc4_sgc", "almgsi_fcc4_canonical",
          "almgsi_fcc4_sgc", "almgsi_fcc5_canonical_cold"]

KB = 8.617330337217213e-05   # eV/K (ase.units.kB, CODATA 2014)


def build(L, species, families, conc, eci_kind="synthetic", seed=3,
          trans_matrix_format="auto"):
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
    st = syn.fcc_settings(meta["L"], meta["species"], meta["families"])
    ft = FlatTables(st, meta["eci"], meta["symbols0"])
    assert ft.eci_names == meta["eci_names"]
    assert ft.species == meta["species_sorted"]
    return meta, st, ft, z


BINARY = dict(L=4, species=["Al", "Mg"], families=["nn", "2nn", "tri", "tet"],
              conc={"Al": 0.5, "Mg": 0.5})
TERNARY = dict(L=4, species=["Al", "Mg", "Si"],
               families=["nn", "2nn", "tri", "iso", "tet"],
               conc={"Al": 0.5, "Mg": 0.25, "Si": 0.25})
