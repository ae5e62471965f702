"""One pinned-variant launch of a BASELINE workload for ncu (never a bench number).
    python scripts/prof_wl.py C3S 1 4000     # workload, kernel variant, moves per replica"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cemc_b200 import workloads as wl

which, variant = sys.argv[1].upper(), int(sys.argv[2])
n = int(sys.argv[3]) if len(sys.argv) > 3 else 4000
w = wl.c4_parallel_tempering(R=64, n_total=64) if which == "C4" else wl.WORKLOADS[which]()
gpu = wl.make_updater(w)
gpu.set_variant(variant, variant)
run = gpu.run_sgc if w.mode == "sgc" else gpu.run_canonical
run(n); gpu.synchronize()
gpu.timer_start(); run(n); ms = gpu.timer_stop()
print("%s variant %d: %.0f ns/move/chain" % (which, variant, ms * 1e6 / n))
