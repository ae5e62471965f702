"""Does the CTA -> replica order matter for config 2 (256 chains on 148 SMs)?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cemc_b200 import workloads as wl
w = wl.WORKLOADS["C2"]()
gpu = wl.make_updater(w)
gpu.set_variant(3, 3)
gpu.set_replica_order(np.arange(w.R))      # "identity" = the automatic hottest-first order switched off
R = w.R
hot = np.argsort(-w.kT, kind="stable")           # hottest first
def timeit(tag):
    gpu.run_sgc(4000); gpu.synchronize()
    best = 1e30
    for _ in range(3):
        gpu.timer_start(); gpu.run_sgc(20000); best = min(best, gpu.timer_stop())
    print("%-28s %.1f ns/move/chain" % (tag, best * 1e6 / 20000), flush=True)
timeit("identity")
# hottest 40 alone (CTAs 108..147 if the block scheduler is breadth-first), the rest hot+cold pairs
order = np.empty(R, dtype=np.int32)
alone = hot[:40]; rest = hot[40:]                 # 216 chains, hottest first
order[108:148] = alone
order[:108] = rest[:108]                          # hotter half on first-wave CTAs 0..107
order[148:] = rest[108:][::-1]                    # partner of CTA i is CTA 148+i: coldest with hottest
order2 = np.empty(R, dtype=np.int32)
order2[:148] = hot[:148]; order2[148:] = hot[148:][::-1]
gpu.set_replica_order(order); timeit("hot alone + hot/cold pairs")
gpu.set_replica_order(order2); timeit("hot first wave, cold partners")
gpu.set_replica_order(hot.astype(np.int32)); timeit("sorted hot -> cold")
gpu.set_replica_order(hot[::-1].astype(np.int32).copy()); timeit("sorted cold -> hot")
gpu.set_replica_order(None); timeit("identity again")
