"""ns/move/chain of the bench workloads on the library named by CEMC_B200_LIB (A/B builds)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cemc_b200 import workloads as wl
for name in sys.argv[1:] or ["C2", "C3S", "C3"]:
    w = wl.c4_parallel_tempering(R=64, n_total=64) if name == "C4" else wl.WORKLOADS[name]()
    gpu = wl.make_updater(w)
    run = gpu.run_sgc if w.mode == "sgc" else gpu.run_canonical
    n = 20000
    for _ in range(4): run(n)
    gpu.synchronize()
    best, tot = 1e9, 0.0
    for _ in range(6):
        gpu.timer_start(); run(n); ms = gpu.timer_stop(); best = min(best, ms); tot += ms
    print("%s %s variant %s: mean %.1f best %.1f ns/move/chain" % (os.environ.get("CEMC_B200_LIB", "default"), name, gpu.get_variant(), tot / 6 * 1e6 / n, best * 1e6 / n))
