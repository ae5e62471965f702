"""Where does a parallel-tempering leg's time go?  (config 4 shard: 64 chains, fcc 12^3 ternary canonical)
  * per kernel variant: 20 legs of 1728 moves (20 launches) vs ONE launch of 34 560 moves -> relaunch + restage cost
  * the whole ladder vs all chains hot (1500 K) vs all chains cold (100 K) -> chain-speed spread a leg waits for
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cemc_b200 import workloads as wl

KB = wl.KB
w = wl.c4_parallel_tempering(R=64, n_total=64)
N = w.tables.N
for label, kT in (("ladder 100-1500 K", w.kT), ("all 1500 K", np.full(64, 1500 * KB)), ("all 600 K", np.full(64, 600 * KB)),
                  ("all 100 K", np.full(64, 100 * KB))):
    for v in (3, 8, 9):
        gpu = wl.make_updater(w)
        gpu.set_kT(kT)
        gpu.set_variant(v, v)
        gpu.run_canonical(4 * N); gpu.synchronize()
        gpu.timer_start()
        for _ in range(20):
            gpu.run_canonical(N)
        ms_legs = gpu.timer_stop()
        gpu.timer_start()
        gpu.run_canonical(20 * N)
        ms_one = gpu.timer_stop()
        st, acc = gpu.get_counters()
        print("%-18s variant %d: 20 legs %.3f ms (%.0f ns/move/chain)   one launch %.3f ms (%.0f ns/move/chain)   "
              "per-leg overhead %.1f us   accept rate %.3f" % (
                  label, v, ms_legs, ms_legs * 1e6 / (20 * N), ms_one, ms_one * 1e6 / (20 * N),
                  (ms_legs - ms_one) * 1e3 / 20, acc.sum() / st.sum()))
        gpu.close()
