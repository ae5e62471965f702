"""ns/move/chain of the several-moves-per-warp spin variant (6) for CEMC_M = 2, 3, 4 against variant 3."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cemc_b200 import workloads as wl
which = sys.argv[1] if len(sys.argv) > 1 else "C2"
w = wl.WORKLOADS[which]()
gpu = wl.make_updater(w)
run = gpu.run_sgc if w.mode == "sgc" else gpu.run_canonical
n = 20000
for v, m in ((3, None), (6, 2), (6, 3), (6, 4)):
    if m: os.environ["CEMC_M"] = str(m)
    gpu.set_variant(v, v)
    for _ in range(3): run(n)
    gpu.synchronize()
    tot = 0.0
    for _ in range(5):
        gpu.timer_start(); run(n); tot += gpu.timer_stop()
    print("%s variant %d M=%s (ran %d): %.1f ns/move/chain" % (which, v, m, gpu.last_variant(), tot / 5 * 1e6 / n), flush=True)
