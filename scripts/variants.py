"""ns/move/chain of every kernel variant on the BASELINE workloads (device-timed, pinned variants).
    python scripts/variants.py [C2 C3S C3 C1] [tab=0]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cemc_b200 import workloads as wl

names = [a.upper() for a in sys.argv[1:] if "=" not in a] or ["C2", "C3S", "C3", "C1"]
opts = dict(a.split("=") for a in sys.argv[1:] if "=" in a)
n = int(opts.get("n", 20000))
for which in names:
    w = wl.WORKLOADS[which](R=64) if which == "C1" else wl.WORKLOADS[which]()
    gpu = wl.make_updater(w)
    if "tab" in opts:
        gpu.set_table_eval(int(opts["tab"]) != 0)
    run = gpu.run_sgc if w.mode == "sgc" else gpu.run_canonical
    res = []
    for v in range(9):
        gpu.set_variant(v, v)
        run(2000); gpu.synchronize()
        best = 1e30
        for _ in range(2):
            gpu.timer_start(); run(n); best = min(best, gpu.timer_stop())
        res.append("v%d %.0f" % (v, best * 1e6 / n))
    print("%s (eval %d): ns/move/chain  %s" % (which, gpu.get_batch_eval(), "  ".join(res)), flush=True)
    gpu.close()
