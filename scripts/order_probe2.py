"""C2 (256 chains on 148 SMs, two 8-warp CTAs per SM): which chains should get an SM of their own?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cemc_b200 import workloads as wl
w = wl.WORKLOADS["C2"]()
gpu = wl.make_updater(w)
gpu.set_variant(3, 3)
n = 40000
R = w.R
hot = np.argsort(-w.kT, kind="stable")          # hottest first
def t(order, label):
    gpu.set_replica_order(order)
    gpu.run_sgc(n); gpu.synchronize()
    best = 1e9
    for _ in range(3):
        gpu.timer_start(); gpu.run_sgc(n); best = min(best, gpu.timer_stop())
    print("%-60s %.1f ns/move/chain" % (label, best * 1e6 / n))
t(None, "default (hottest first: CTA i = i-th hottest)")
t(np.arange(R, dtype=np.int32), "identity")
for lone in (40,):
    o = np.empty(R, dtype=np.int32)
    o[108:148] = hot[:lone]                      # SMs 108..147 hold one CTA each
    rest = hot[lone:]
    o[:108] = rest[:108]                         # first CTA of SM i: next hottest
    o[148:] = rest[108:][::-1]                   # second CTA of SM i: coldest first
    t(o, "40 hottest alone, rest paired hot-with-cold")
    o2 = np.empty(R, dtype=np.int32)
    o2[108:148] = hot[-lone:]                    # control: 40 COLDEST alone
    rest = hot[:-lone]
    o2[:108] = rest[:108]; o2[148:] = rest[108:][::-1]
    t(o2, "control: 40 coldest alone")
