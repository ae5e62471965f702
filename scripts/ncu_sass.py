"""Summarise an ncu `--page source --print-source sass --csv` dump per opcode and list the hottest instructions."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
H = rows[1]
isrc, iex, ismp = H.index("Source"), H.index("Instructions Executed"), H.index("# Samples")
iexc = H.index("L1 Wavefronts Shared Excessive"); iwf = H.index("L1 Wavefronts Shared")
op_i, op_s = collections.Counter(), collections.Counter()
ins = []
for r in rows[2:]:
    if len(r) <= ismp: continue
    s = r[isrc].strip()
    t = s.split()
    op = t[1] if t and t[0].startswith("@") else (t[0] if t else "")
    op = op.split(".")[0] + ("." + ".".join(op.split(".")[1:2]) if op.startswith(("LD", "ST", "MUFU", "BAR", "SHFL")) else "")
    n, sm = float(r[iex] or 0), float(r[ismp] or 0)
    op_i[op] += n; op_s[op] += sm
    ins.append((sm, n, s, float(r[iwf] or 0), float(r[iexc] or 0)))
ti, ts = sum(op_i.values()), sum(op_s.values())
print("total warp-inst %.3e samples %d" % (ti, ts))
for op, n in op_i.most_common(28):
    print("%-14s inst %5.1f%%  smp %5.1f%%" % (op, 100 * n / ti, 100 * op_s[op] / ts))
print("-- hottest instructions by samples")
for sm, n, s, wf, ex in sorted(ins, reverse=True)[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print("smp %5.2f%% inst %5.2f%% smem-wf %.2e (excess %.2e)  %s" % (100 * sm / ts, 100 * n / ti, wf, ex, s[:90]))
