"""How much of a launch is the bookkeeper's observer arithmetic?  run_* with the sums on / off."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cemc_b200 import workloads as wl
for name in sys.argv[1:] or ["C2", "C3S", "C3"]:
    w = wl.c4_parallel_tempering(R=64, n_total=64) if name == "C4" else wl.WORKLOADS[name]()
    gpu = wl.make_updater(w)
    run = gpu.run_sgc if w.mode == "sgc" else gpu.run_canonical
    n = 40000
    run(n); gpu.synchronize()
    for on in (True, False, True, False):
        gpu.set_observe(on)
        run(n); gpu.synchronize()
        gpu.timer_start(); run(n); ms = gpu.timer_stop()
        print("%s variant %s observe=%s: %.1f ns/move/chain" % (name, gpu.get_variant(), on, ms * 1e6 / n))
