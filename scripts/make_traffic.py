"""profiles/traffic.json from the `ncu --set full` raw pages of one build.

    python scripts/make_traffic.py gpurun_out/r02_source_hash.txt C2=gpurun_out/r02_raw_C2.csv:1024000 C3S=...:256000

Each argument is <workload key>=<raw csv>:<moves in the captured launch>.  The kernel source hash
(`cemc_b200._lib.source_hash()` printed on the GPU box by the same gpurun call that made the
captures) is stored with the figures; bench.py reports them only when its own build has that hash."""
import csv, json, sys

UNIT = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}


def figures(path, moves):
    rows = list(csv.reader(open(path)))
    d, u = dict(zip(rows[0], rows[2])), dict(zip(rows[0], rows[1]))
    num = lambda k: float(d[k].replace(",", ""))
    nbytes = lambda k: num(k) * UNIT.get(u.get(k, "byte"), 1.0)
    return {
        "kernel": d["Kernel Name"],
        "moves_in_captured_launch": moves,
        "dram_bytes_per_launch": int(nbytes("dram__bytes_read.sum") + nbytes("dram__bytes_write.sum")),
        "warp_instructions_per_move": round(num("smsp__inst_executed.sum") / moves, 1),
        "issue_slots_busy_pct": round(num("smsp__issue_active.avg.pct_of_peak_sustained_active"), 1),
        "l2_bytes_per_move": round(num("lts__t_sectors.sum") * 32.0 / moves, 1),
        "source": path.replace("gpurun_out/", "profiles/").replace("_raw_", "_full_").replace(".csv", "_summary.txt"),
    }


if __name__ == "__main__":
    out = {"kernel_source_sha": open(sys.argv[1]).read().strip(),
           "note": "DRAM traffic of a launch is the one-time staging of tables / occupations (chain state stays in "
                   "shared memory), so it does not grow with the moves per launch"}
    for a in sys.argv[2:]:
        key, rest = a.split("=")
        path, moves = rest.rsplit(":", 1)
        out[key] = figures(path, int(moves))
    json.dump(out, open("profiles/traffic.json", "w"), indent=1)
    print(json.dumps(out, indent=1))
