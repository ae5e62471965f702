"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, total time, share.
    python scripts/launch_shares.py gpurun_out/launches.csv > profiles/rNN_launch_shares.txt"""
import csv, sys, collections
rows = [l for l in open(sys.argv[1]) if l.startswith('"')]
tot = collections.OrderedDict()
for r in csv.DictReader(rows):
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"].split("(")[0]
    t = float(r["Metric Value"]) * (1e-6 if r["Metric Unit"] in ("ns", "nsecond") else 1.0)
    n, s = tot.get(name, (0, 0.0))
    tot[name] = (n + 1, s + t)
allt = sum(s for _, s in tot.values())
for name, (n, s) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("%-72s launches %3d  total %9.3f ms  share %5.1f%%" % (name[:72], n, s, 100 * s / allt))
