"""Per-warp cycle accounting of the batch kernel (debug build with -DCEMC_WARP_TIMING): for each of the first
8 warps of CTA 0, cycles per batch spent on (W) its work before the barrier -- evaluation, or proposals +
bookkeeping in the observer warp --, (B) the barrier wait, (D) the decision after it.
Build:  python -c "from cemc_b200 import _lib; _lib.build_ext(force=True, defines=['CEMC_WARP_TIMING'], out='cemc_b200/_cemc_b200_wt.so')"
Run:    CEMC_B200_LIB=cemc_b200/_cemc_b200_wt.so python scripts/warp_timing.py C2 [variant]"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from cemc_b200 import workloads as wl, _lib

which = sys.argv[1].upper()
w = wl.c4_parallel_tempering(R=64, n_total=64) if which == "C4" else wl.WORKLOADS[which]()
gpu = wl.make_updater(w)
if len(sys.argv) > 2:
    gpu.set_variant(int(sys.argv[2]), int(sys.argv[2]))
run = gpu.run_sgc if w.mode == "sgc" else gpu.run_canonical
n = 20000
for _ in range(3): run(n)
gpu.synchronize()
gpu.timer_start(); run(n); ms = gpu.timer_stop()
out = (C.c_uint64 * (24 * w.R))()
_lib.check(gpu.lib.cemc_debug_phase_cycles(gpu._h, out))
c = np.array(list(out), dtype=float).reshape(w.R, 8, 3) / n          # cycles per move
tot = c[:, 0, :].sum(axis=1)
print("%s variant %s: %.1f ns/move/chain (launch); warp 0 cycles/move min %.0f median %.0f max %.0f" % (
    which, gpu.get_variant(), ms * 1e6 / n, tot.min(), np.median(tot), tot.max()))
for label, r in (("fastest", int(tot.argmin())), ("median", int(np.argsort(tot)[len(tot) // 2])), ("slowest", int(tot.argmax()))):
    print("  %s replica %d  (cycles per move: work / barrier wait / decision)" % (label, r))
    for wp in range(8):
        if c[r, wp].sum() > 0:
            print("    warp %d: %7.1f %7.1f %7.1f" % (wp, c[r, wp, 0], c[r, wp, 1], c[r, wp, 2]))
