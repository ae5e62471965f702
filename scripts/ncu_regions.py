"""Instructions per move by kernel region of cemc_batch_kernel.cuh, from an ncu
`--page source --print-source cuda,sass --csv` dump.  usage: ncu_regions.py dump.csv n_moves [lo hi]"""
import csv, sys, os
def _f(x):
    try: return float(x)
    except ValueError: return 0.0
rows = list(csv.reader(open(sys.argv[1]))); moves = float(sys.argv[2])
for i, r in enumerate(rows):
    if r and r[0] == "Line No": H = r; start = i + 1; break
ie, smp = H.index("Instructions Executed"), H.index("# Samples")
end = len(rows)
for i in range(start, len(rows)):
    if rows[i] and rows[i][0] == "Line No": end = i; break
lines = [(int(r[0]), _f(r[ie]), _f(r[smp]), r[1]) for r in rows[start:end] if r and r[0].isdigit()]
src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "cemc_b200", "csrc", "cemc_batch_kernel.cuh")).read().split("\n")
def find(txt):
    for i, l in enumerate(src):
        if txt in l: return i + 1
marks = [("setup", 1), ("loop top / observer", find("while (sdone < a.n_steps)")), ("E1 head / P0", find("---- E1: warp b evaluates")),
         ("spin eval", find("// ---- spin evaluation")), ("tab codes", find("// ---- table evaluation")),
         ("tab sums", find("// sums over the sub-clusters, reference order")), ("tab quotients", find("// per-ECI quotients (:393-402): lane i = ECI i, both changed sites")),
         ("product eval", find("// P1: gather")), ("conflict mask", find("---- which earlier moves of the batch would invalidate")),
         ("D decide", find("---- D: warp 0 decides")), ("loop end", find("sdone += s.ctl[0];"))]
marks = sorted([(n, l) for n, l in marks if l], key=lambda x: x[1])
tot = sum(l[1] for l in lines)
print("instr/move total (this file) %.0f" % (tot / moves))
for k, (n, l0) in enumerate(marks):
    l1 = marks[k + 1][1] if k + 1 < len(marks) else 10 ** 9
    ii = sum(l[1] for l in lines if l0 <= l[0] < l1); ss = sum(l[2] for l in lines if l0 <= l[0] < l1)
    print("  %-20s lines %4d-%4d  instr/move %7.1f  samples %6d" % (n, l0, min(l1, 9999), ii / moves, ss))
if len(sys.argv) > 4:
    lo, hi = int(sys.argv[3]), int(sys.argv[4])
    for ln, i, s_, txt in lines:
        if lo <= ln <= hi and i / moves > 0.3: print("%4d inst/move %6.2f smp %5d  %s" % (ln, i / moves, s_, txt.strip()[:100]))
