"""Quick throughput probe of the MC kernels on one GPU (not the bench)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cemc_b200 import synthetic as syn
from cemc_b200.tables import FlatTables
from cemc_b200.updater import BatchedCEUpdater

KB = 8.617330337217213e-05

def setup(L, species, conc, R, kTs, mus=None):
    st = syn.fcc_settings(L, species, ["nn", "2nn", "tri", "tet"])
    eci = syn.synthetic_ecis(st)
    syms = syn.random_symbols(st, conc, seed=1)
    ft = FlatTables(st, eci, syms)
    gpu = BatchedCEUpdater(ft, R)
    occ = np.stack([ft.occupancy(syn.random_symbols(st, conc, seed=r)) for r in range(R)])
    gpu.set_occupancy(occ)
    gpu.recompute_cf()
    gpu.set_kT(kTs)
    if mus is not None:
        e = np.tile(ft.eci, (R, 1))
        e[:, ft.eci_index["c1_0"]] -= mus
        gpu.set_ecis(e)
    gpu.seed(1234)
    return ft, gpu

def timeit(gpu, fn, n, reps=3):
    fn(n); gpu.synchronize()
    best = 1e30
    for _ in range(reps):
        gpu.timer_start(); fn(n); ms = gpu.timer_stop(); gpu.synchronize()
        best = min(best, ms)
    return best

if __name__ == "__main__":
    which = [a for a in sys.argv[1:] if not a.startswith("t=")] or ["c2", "c3"]
    threads = [int(a[2:]) for a in sys.argv[1:] if a.startswith("t=")] or [0]
    batches = [int(a[2:]) for a in sys.argv[1:] if a.startswith("b=")] or [0]
    which = [a for a in which if not a.startswith("b=")]
    if "c2" in which:
        T = np.linspace(200, 1000, 16); mu = np.linspace(-1.1, -0.9, 16)
        kTs = np.repeat(T * KB, 16); mus = np.tile(mu, 16)
        ft, gpu = setup(10, ["Al", "Mg"], {"Al": 0.5, "Mg": 0.5}, 256, kTs, mus * 0.0)
        gpu.set_autotune(False)
        for spin in (True, False):
          for bt in (batches if not spin else [0]):
            gpu.set_batch(bt)
            n = 50000
            ms = timeit(gpu, gpu.run_sgc, n)
            print("C2 sgc binary L=10 R=256 spin=%s batch=%d n=%d: %.2f ms -> %.1f M moves/s (%.0f ns/move/chain)" % (
                spin, bt, n, ms, 256 * n / ms / 1e3, ms * 1e6 / n))
        st, acc = gpu.get_counters(); print("  accept rate", acc.sum() / st.sum())
    if "c3" in which:
        kTs = np.linspace(300, 900, 64) * KB
        ft, gpu = setup(20, ["Al", "Mg", "Si"], {"Al": 0.8, "Mg": 0.1, "Si": 0.1}, 64, kTs)
        for bt in batches:
          for th in (1, 2):
            gpu.set_cluster(th)
            gpu.set_batch(bt)
            n = 50000
            ms = timeit(gpu, gpu.run_canonical, n)
            print("C3 canonical ternary L=20 R=64 batch=%d cluster=%d n=%d: %.2f ms -> %.1f M moves/s (%.0f ns/move/chain)" % (
                bt, th, n, ms, 64 * n / ms / 1e3, ms * 1e6 / n))
            ms = timeit(gpu, gpu.run_sgc, n)
            print("C3-lattice sgc ternary batch=%d cluster=%d: %.1f M moves/s (%.0f ns/move/chain)" % (bt, th, 64 * n / ms / 1e3, ms * 1e6 / n))
        st, acc = gpu.get_counters(); print("  accept rate", acc.sum() / st.sum())
    if "auto" in which:
        from cemc_b200 import workloads as wl
        for name in ("C2", "C3S", "C3", "C1"):
            w = wl.WORKLOADS[name](R=64) if name == "C1" else wl.WORKLOADS[name]()
            gpu = wl.make_updater(w)
            run = gpu.run_sgc if w.mode == "sgc" else gpu.run_canonical
            n = 100000
            run(n); gpu.synchronize()
            ms = timeit(gpu, run, n, reps=2)
            print("%s autotuned variant %s: %.1f M moves/s (%.0f ns/move/chain)" % (name, gpu.get_variant(), w.R * n / ms / 1e3, ms * 1e6 / n))
    if "c5" in which:
        from cemc_b200 import workloads as wl
        st = syn.fcc_settings(64, ["Al", "Mg"], ["nn", "2nn", "tri", "tet"])
        eci = syn.synthetic_ecis(st)
        syms = syn.random_symbols(st, {"Al": 0.9, "Mg": 0.1}, seed=1)
        ft = FlatTables(st, eci, syms)
        for R in (1, 8):
            gpu = BatchedCEUpdater(ft, R)
            gpu.set_occupancy(np.stack([ft.occupancy(syms)] * R)); gpu.recompute_cf(); gpu.set_kT(np.full(R, 0.03)); gpu.seed(5)
            for bt in batches:
                gpu.set_batch(bt)
                for cl in (1, 2):
                    gpu.set_cluster(cl)
                    n = 20000
                    ms = timeit(gpu, gpu.run_canonical, n)
                    print("C5 64^3 binary canonical R=%d batch=%d cluster=%d: %.2f M moves/s (%.0f ns/move/chain)" % (R, bt, cl, R * n / ms / 1e3, ms * 1e6 / n))
