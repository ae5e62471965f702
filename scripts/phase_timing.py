"""Per-phase cycle accounting of the MC kernel (debug build with -DCEMC_PHASE_TIMING).
Build:  nvcc ... -DCEMC_PHASE_TIMING -o cemc_b200/_cemc_b200_timing.so cemc_b200/csrc/cemc_b200.cu
Run:    CEMC_B200_LIB=cemc_b200/_cemc_b200_timing.so python scripts/phase_timing.py c2"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cemc_b200 import workloads as wl, _lib

names = ["refill", "P0", "P1", "P2a", "P2b", "P3", "endbar", "-"]
for which in sys.argv[1:] or ["C2", "C3", "C3S"]:
    w = wl.WORKLOADS[which.upper()]()
    gpu = wl.make_updater(w)
    run = gpu.run_sgc if w.mode == "sgc" else gpu.run_canonical
    n = 20000
    run(n); gpu.synchronize()
    gpu.timer_start(); run(n); ms = gpu.timer_stop()
    out = (C.c_uint64 * 8)()
    _lib.check(gpu.lib.cemc_debug_phase_cycles(gpu._h, out))
    cyc = np.array(list(out), dtype=float) / n
    print("%s: %.0f ns/move/chain; cycles/move by phase: %s  total %.0f" % (
        which, ms * 1e6 / n, ", ".join("%s %.0f" % (a, b) for a, b in zip(names, cyc) if b > 0), cyc.sum()))
