"""A/B: translation by index arithmetic vs gather from the table (same trajectories)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cemc_b200 import workloads as wl
for name in sys.argv[1:] or ["C2", "C3S", "C3", "C4", "C5"]:
    w = wl.c4_parallel_tempering(R=64, n_total=64) if name == "C4" else wl.WORKLOADS[name]()
    gpu = wl.make_updater(w)
    run = gpu.run_sgc if w.mode == "sgc" else gpu.run_canonical
    n = 40000
    run(n); gpu.synchronize()       # autotune with the arithmetic on
    v = gpu.get_variant()
    for on in (True, False, True, False):
        gpu.set_lattice_arithmetic(on)
        gpu.set_variant(*[x if x >= 0 else -1 for x in v])
        run(n); gpu.synchronize()
        gpu.timer_start(); run(n); ms = gpu.timer_stop()
        print("%s variant %s lattice arithmetic=%s: %.1f ns/move/chain" % (name, v, gpu.get_lattice_arithmetic(), ms * 1e6 / n))
