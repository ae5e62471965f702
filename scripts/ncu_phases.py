"""Aggregate ncu source samples/instructions into kernel phases by marker comments."""
import csv, sys, re
src_file = sys.argv[2] if len(sys.argv) > 2 else "/root/repo/cemc_b200/csrc/cemc_kernels.cuh"
def _f(x):
    try:
        return float(x)
    except ValueError:
        return 0.0
rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if r and r[0] == "Line No":
        H = r; start = i + 1; break
ie = H.index("Instructions Executed"); smp = H.index("# Samples")
# phase boundaries from the source markers "// ---- Pxx"
marks = []
for ln, line in enumerate(open(src_file), 1):
    m = re.search(r"// ---- (P\w+|write back|stage)", line)
    if m: marks.append((ln, m.group(1)))
def phase(ln):
    p = "pre"
    for l, name in marks:
        if ln >= l: p = name
    return p
agg = {}
for r in rows[start:]:
    if len(r) <= ie or not r[0].isdigit(): continue
    p = phase(int(r[0]))
    a = agg.setdefault(p, [0.0, 0.0])
    a[0] += _f(r[ie]); a[1] += _f(r[smp])
ti = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values())
for p, a in agg.items():
    print("%-12s inst %5.1f%%  samples %5.1f%%" % (p, a[0] / ti * 100, a[1] / ts * 100))
