"""Summarise an ncu `--page source --print-source cuda,sass --csv` dump per CUDA line."""
import csv, sys
def _f(x):
    try:
        return float(x)
    except ValueError:
        return 0.0
rows = list(csv.reader(open(sys.argv[1])))
thresh = float(sys.argv[2]) if len(sys.argv) > 2 else 0.7
H = None
for i, r in enumerate(rows):
    if r and r[0] == "Line No":
        H = r; start = i + 1; break
ie = H.index("Instructions Executed"); smp = H.index("# Samples")
lines = []
for r in rows[start:]:
    if len(r) <= ie or r[0] in ("", "Line No", "File Path", "Function Name"):
        continue
    try:
        lines.append((int(r[0]), _f(r[ie]), _f(r[smp]), r[1]))
    except ValueError:
        pass
tot_i = sum(l[1] for l in lines); tot_s = sum(l[2] for l in lines)
print("total inst %.3e  samples %d" % (tot_i, tot_s))
for ln, i, s, src in lines:
    if i / tot_i * 100 > thresh or s / tot_s * 100 > thresh:
        print("%4d inst %5.1f%% smp %5.1f%%  %s" % (ln, i / tot_i * 100, s / tot_s * 100, src.strip()[:100]))
