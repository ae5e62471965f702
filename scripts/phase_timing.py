"""Per-phase cycle accounting of the MC kernel (debug build with -DCEMC_PHASE_TIMING).
Build:  nvcc ... -DCEMC_PHASE_TIMING -o cemc_b200/_cemc_b200_timing.so cemc_b200/csrc/cemc_b200.cu
Run:    CEMC_B200_LIB=cemc_b200/_cemc_b200_timing.so python scripts/phase_timing.py c2"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cemc_b200 import workloads as wl, _lib

names = ["s0", "s1", "s2", "s3", "s4", "s5", "s6", "s7", "s8", "s9", "s10", "s11", "s12", "s13", "s14", "s15"]
# mc_kernel: s0 refill s1 P0 s2 P1 s3 P2a s4 P2b s5 P3 s6 end barrier
# batch_kernel: s0 refill s1 evaluation s2 hoisted decisions s3 sequential decisions s4 end sync; s8 batches s9 moves (counts/n)
batch = None
args = []
for a_ in sys.argv[1:]:
    if a_.startswith("b="): batch = int(a_[2:])
    else: args.append(a_)
for which in args or ["C2", "C3", "C3S"]:
    w = wl.WORKLOADS[which.upper()]()
    gpu = wl.make_updater(w)
    if batch is not None: gpu.set_batch(batch)
    run = gpu.run_sgc if w.mode == "sgc" else gpu.run_canonical
    n = 20000
    run(n); gpu.synchronize()
    gpu.timer_start(); run(n); ms = gpu.timer_stop()
    out = (C.c_uint64 * 16)()
    _lib.check(gpu.lib.cemc_debug_phase_cycles(gpu._h, out))
    cyc = np.array(list(out), dtype=float) / n
    print("%s: %.0f ns/move/chain; cycles/move by phase: %s  total %.0f" % (
        which, ms * 1e6 / n, ", ".join("%s %.2f" % (a, b) for a, b in zip(names, cyc) if b > 0), cyc.sum()))
