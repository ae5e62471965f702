python -m pytest tests -m gpu -q -k "two_ecis" 2>&1 | tail -8
python - <<'PY'
import sys, numpy as np
sys.path.insert(0, '.')
from cemc_b200 import synthetic as syn, workloads as wl
from cemc_b200.tables import FlatTables
from cemc_b200.updater import BatchedCEUpdater
for species, fams in ((["Al", "Cu", "Mg", "Si"], ["nn", "2nn", "tri", "tet"]), (["Al", "Cu", "Mg", "Si", "Zn"], ["nn", "2nn", "tri"])):
    st = syn.fcc_settings(12, species, fams)
    eci = syn.synthetic_ecis(st, seed=1234)
    conc = {s: 1.0 / len(species) for s in species}
    ft = FlatTables(st, eci, syn.random_symbols(st, conc, seed=0))
    R = 64
    occ = np.stack([ft.occupancy(syn.random_symbols(st, conc, seed=10 + r)) for r in range(R)])
    for mode in ("sgc", "canonical"):
        for v in (-1, 5):
            gpu = BatchedCEUpdater(ft, R); gpu.set_occupancy(occ); gpu.recompute_cf(); gpu.set_kT(np.linspace(300, 900, R) * wl.KB); gpu.seed(3)
            if v >= 0: gpu.set_variant(v, v)
            run = gpu.run_sgc if mode == "sgc" else gpu.run_canonical
            n = 20000
            run(n); gpu.synchronize(); gpu.timer_start(); run(n); ms = gpu.timer_stop()
            print("%d species %s n_eci=%d %s: eval %d variant %s (last %d): %.0f ns/move/chain" % (len(species), fams, ft.n_eci, mode, gpu.get_batch_eval(), gpu.get_variant(), gpu.last_variant(), ms * 1e6 / n))
            gpu.close()
PY
