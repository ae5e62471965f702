"""Where does the end-to-end step (host buffers in, results out) spend its time?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cemc_b200 import workloads as wl
for name in sys.argv[1:] or ["C2", "C3S"]:
    w = wl.WORKLOADS[name]()
    gpu = wl.make_updater(w)
    run = gpu.run_sgc if w.mode == "sgc" else gpu.run_canonical
    n = 20000
    for _ in range(3): run(n)
    gpu.synchronize()
    steps = [("set_occupancy", lambda: gpu.set_occupancy(w.occ)), ("recompute_cf", gpu.recompute_cf),
             ("set_ecis", lambda: gpu.set_ecis(w.eci_matrix)), ("set_kT", lambda: gpu.set_kT(w.kT)),
             ("reset_accumulators", gpu.reset_accumulators), ("run", lambda: run(n)),
             ("get_accumulators", gpu.get_accumulators), ("get_energy", gpu.get_energy), ("get_occupancy", gpu.get_occupancy)]
    tot = {k: 0.0 for k, _ in steps}
    reps = 5
    for _ in range(reps):
        for k, f in steps:
            gpu.synchronize(); t0 = time.perf_counter(); f(); gpu.synchronize(); tot[k] += time.perf_counter() - t0
    print(name, " ".join("%s %.0f us" % (k, v / reps * 1e6) for k, v in tot.items()), "| total %.0f us" % (sum(tot.values()) / reps * 1e6))
