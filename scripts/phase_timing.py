"""Per-phase cycle accounting of the MC kernel (debug build with -DCEMC_PHASE_TIMING).
Build:  nvcc ... -DCEMC_PHASE_TIMING -o cemc_b200/_cemc_b200_timing.so cemc_b200/csrc/cemc_b200.cu
Run:    CEMC_B200_LIB=cemc_b200/_cemc_b200_timing.so python scripts/phase_timing.py c2"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cemc_b200 import workloads as wl, _lib

names = ["s%d" % i for i in range(24)]
# mc_kernel: s0 refill s1 P0 s2 P1 s3 P2a s4 P2b s5 P3 s6 end barrier
# batch_kernel (warp 0): s0 loop top, s5 proposal decode, s6 gather+codes (spin/product: whole evaluation in s12), s7 sums,
# s14 observer-warp work (s19 proposals, s16 CF chain, s17 energies, s18 observer sums), s15 its barrier wait,
# s12 quotients, s13 conflict mask, s1 barrier wait, s3 decisions, s4 barrier wait; per move: s8 batches s9 moves s10 accepted s11 batches cut short
batch = None
args = []
for a_ in sys.argv[1:]:
    if a_.startswith("b="): batch = int(a_[2:])
    else: args.append(a_)
for which in args or ["C2", "C3", "C3S"]:
    w = wl.c4_parallel_tempering(R=64, n_total=64) if which.upper() == "C4" else wl.WORKLOADS[which.upper()]()
    gpu = wl.make_updater(w)
    if batch is not None: gpu.set_batch(batch)
    if os.environ.get("VARIANT"): gpu.set_variant(int(os.environ["VARIANT"]), int(os.environ["VARIANT"]))
    run = gpu.run_sgc if w.mode == "sgc" else gpu.run_canonical
    n = 20000
    run(n); gpu.synchronize()
    gpu.timer_start(); run(n); ms = gpu.timer_stop()
    out = (C.c_uint64 * (24 * w.R))()
    _lib.check(gpu.lib.cemc_debug_phase_cycles(gpu._h, out))
    allc = np.array(list(out), dtype=float).reshape(w.R, 24) / n
    tot = allc[:, :8].sum(axis=1) + allc[:, 12:14].sum(axis=1)
    print("%s: %.0f ns/move/chain (launch); cycles/move/chain min %.0f median %.0f max %.0f" % (
        which, ms * 1e6 / n, tot.min(), np.median(tot), tot.max()))
    for label, r in (("fastest", int(tot.argmin())), ("slowest", int(tot.argmax()))):
        cyc = allc[r]
        print("   %s replica %d: %s  total %.0f" % (
            label, r, ", ".join("%s %.2f" % (a, b) for a, b in zip(names, cyc) if b > 0), cyc[:8].sum() + cyc[12:14].sum()))
