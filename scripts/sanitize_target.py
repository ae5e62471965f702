"""Small runs of every batch-kernel flavour for compute-sanitizer (memcheck / racecheck / synccheck).
    compute-sanitizer --tool racecheck python scripts/sanitize_target.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cemc_b200 import synthetic as syn
from cemc_b200.tables import FlatTables
from cemc_b200.updater import BatchedCEUpdater

variants = [int(a) for a in sys.argv[1:]] or [1, 2, 3, 6]
for species, conc in ((["Al", "Mg"], {"Al": 0.5, "Mg": 0.5}), (["Al", "Mg", "Si"], {"Al": 0.5, "Mg": 0.25, "Si": 0.25})):
    st = syn.fcc_settings(4, species, ["nn", "2nn", "tri", "tet"])
    eci = syn.synthetic_ecis(st)
    syms = syn.random_symbols(st, conc, seed=1)
    ft = FlatTables(st, eci, syms)
    for prec in (64, 32):
        for v in variants:
            gpu = BatchedCEUpdater(ft, 2)
            gpu.set_occupancy(np.stack([ft.occupancy(syms)] * 2))
            gpu.recompute_cf()
            gpu.set_kT([0.03, 0.1])
            gpu.seed(3)
            gpu.set_precision(prec)
            gpu.set_variant(v, v)
            gpu.run_sgc(200)
            gpu.run_canonical(200)
            gpu.synchronize()
            print(len(species), "species, precision", prec, "variant", v, "eval", gpu.get_batch_eval(), "ok", flush=True)
            gpu.close()
