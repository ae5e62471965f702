"""Key metrics of one `ncu --set full` capture (from `ncu -i X.ncu-rep --page raw --csv`).
    python scripts/ncu_summary.py raw.csv n_moves_total "command line" > profiles/rNN_<name>_summary.txt"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
H, U, V = rows[0], rows[1], rows[2]
d = dict(zip(H, V)); u = dict(zip(H, U))
moves = float(sys.argv[2])
print(sys.argv[3] if len(sys.argv) > 3 else "")
print("kernel:", d["Kernel Name"])
print("grid %s x block %s, %s registers/thread, %s KB dynamic smem/block, cluster %s" % (
    d["launch__grid_size"], d["launch__block_size"], d["launch__registers_per_thread"],
    d["launch__shared_mem_per_block_dynamic"], d.get("launch__cluster_size", "1")))
print()
keys = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__average_warp_latency_per_inst_issued.ratio",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_issued.avg.per_cycle_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_alu.sum",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor"]
for k in keys:
    if k in d:
        print("%-66s %s %s" % (k, d[k], u.get(k, "")))
def num(k):
    try: return float(d[k].replace(",", ""))
    except (KeyError, ValueError): return 0.0
st = {k.replace("smsp__pcsamp_warps_issue_stalled_", ""): num(k) for k in d
      if k.startswith("smsp__pcsamp_warps_issue_stalled_") and not k.endswith("_not_issued")}
t = sum(st.values()) or 1.0
print()
print("stall picture (pc samples): " + ", ".join("%s %.0f%%" % (k, 100 * v / t) for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:10]))
unit = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}
def nbytes(k): return num(k) * unit.get(u.get(k, "byte"), 1.0)
dur = num("gpu__time_duration.sum") * {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9, "s": 1.0, "second": 1.0}.get(u.get("gpu__time_duration.sum"), 1.0)
dram = nbytes("dram__bytes_read.sum") + nbytes("dram__bytes_write.sum")
l2 = num("lts__t_sectors.sum") * 32.0
l1 = num("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum") * 32.0
print("derived: %.0f warp-instructions / move; %.0f M moves/s under the profiler; dram %.3f B/move (%.2f GB/s); "
      "L2 %.1f B/move (%.1f GB/s); L1 global loads %.1f B/move (%.1f GB/s); issue slots busy %.1f%%" % (
          num("smsp__inst_executed.sum") / moves, moves / dur / 1e6, dram / moves, dram / dur / 1e9,
          l2 / moves, l2 / dur / 1e9, l1 / moves, l1 / dur / 1e9,
          num("smsp__issue_active.avg.pct_of_peak_sustained_active")))
