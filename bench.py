#!/usr/bin/env python
"""bench.py -- MC trial moves/sec of the cluster-expansion Metropolis hot path.

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload (config.workload): BASELINE.json configs[1] -- the Al-Mg SGC
chemical-potential x temperature sweep on an fcc 10x10x10 cell, 256
independent replicas PER GPU (weak scaling: replicas shard over ranks with no
data-path collective).  One "step" = MOVES_PER_STEP trial moves on every
replica (one kernel launch per rank).

  value  : trial moves/s over all replicas and ranks, state resident in HBM,
           CUDA-event timed on the launching stream, max over ranks
  e2e    : the same metric through the public host API with HOST buffers:
           occupations / ECIs(mu) / kT copied host->device and the observer
           sums, energies and occupations copied back inside the timed region
  roofline, cpu_baseline: see DESIGN.md "Measurement"

--impl reference times the reference's own compiled C++ CEUpdater
(oracle/_ref, else the C oracle port) on all host cores, on the same workload.
"""
import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

MOVES_PER_STEP = 20000          # per replica, per step (GPU arm)
REPLICAS_PER_GPU = 256
METRIC = "mc_trial_moves_per_sec"
UNIT = "moves/s"


# ----------------------------------------------------------------------------
# CPU arm: the reference's compiled CEUpdater (or the oracle port) on host cores
def _cpu_worker(args):
    """One chain on one core; returns (moves, seconds, kind)."""
    replica, n_moves, warm, which = args
    sys.path.insert(0, ROOT)
    from cemc_b200 import workloads as wl
    from oracle import ref_driver
    from oracle.ce_oracle import OracleChain
    w = wl.WORKLOADS[which](R=1, replica_offset=replica)
    ft = w.tables
    eci_vec = w.eci_matrix[0]
    oc = OracleChain(ft, w.occ[0], kT=w.kT[0], seed=1234, replica=replica, eci=eci_vec)
    if ref_driver.available():
        # proposals/uniforms from the Philox chain; the reference's own C++
        # updater does every energy evaluation (its Python-side accept rule)
        cf0 = {k: float(v) for k, v in zip(ft.eci_names, oc.cf)}
        eci = {k: float(v) for k, v in zip(ft.eci_names, eci_vec)}
        st = w.settings
        if st.trans_matrix_columns is not None:
            # the reference reads a dense ndarray or a list of dicts, not our compact form
            from cemc_b200 import synthetic as syn
            kw = st.kwargs
            st = syn.fcc_settings(kw["size"][0], kw["species"], kw["families"], trans_matrix_format="list")
        rc = ref_driver.RefChain(st, ft.symbols_of(w.occ[0]), eci, cf0, kT=w.kT[0])
        tr = oc.run_sgc(warm + n_moves, trace=True)
        rc.replay(ft.species, tr[0][:warm], tr[1][:warm], tr[2][:warm])
        t0 = time.perf_counter()
        acc, _, _ = rc.replay(ft.species, tr[0][warm:], tr[1][warm:], tr[2][warm:])
        dt = time.perf_counter() - t0
        assert np.array_equal(acc, tr[3][warm:])    # the two CPU engines agree
        return n_moves, dt, "reference"
    oc.run_sgc(warm)
    t0 = time.perf_counter()
    oc.run_sgc(n_moves * 10)
    return n_moves * 10, time.perf_counter() - t0, "port"


def cpu_measure(n_moves_per_chain, n_procs=None, warm=500, which="C2"):
    n_procs = n_procs or os.cpu_count() or 1
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(n_procs) as pool:
        res = pool.map(_cpu_worker, [(r, n_moves_per_chain, warm, which) for r in range(n_procs)])
    wall = time.perf_counter() - t0
    # chains run concurrently, one per core: aggregate = sum of per-chain rates
    rate = sum(m / dt for m, dt, _ in res)
    kind = res[0][2]
    return dict(value=rate, unit=UNIT, cores=n_procs, kind=kind,
                sample="%d SGC trial moves on each of %d concurrent chains (one per host core) "
                       "of the %s workload, %s; wall %.1f s incl. setup" % (
                           res[0][0], n_procs, which,
                           "reference C++ CEUpdater driven through its Cython PyCEUpdater"
                           if kind == "reference" else "C oracle port", wall))


# ----------------------------------------------------------------------------
class ClockSampler(object):
    """nvidia-smi clocks/throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t_begin, t_end):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for t, r in self.rows if t_begin <= t <= t_end] or [r for _, r in self.rows]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except (ValueError, IndexError):
                continue
            for nme, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        if not sm:
            return None
        return dict(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)),
                    reasons=sorted(reasons), samples=len(sm))


def gpu_arm(args):
    import torch
    import torch.distributed as dist
    from cemc_b200 import workloads as wl

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # keep stdout to the one JSON line: NCCL prints its version banner (and any NCCL_DEBUG
        # output) on stdout when the first communicator is created, so stdout points at stderr
        # until that has happened
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    assert world == args.gpus, "launch with torchrun --nproc-per-node %d" % args.gpus

    # CPU baseline first (rank 0, N=1 only), before this process touches CUDA
    cpu = None
    cpu_c3s = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_measure(args.cpu_moves)
        if not args.no_extra:
            cpu_c3s = cpu_measure(args.cpu_moves // 4, which="C3S")

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    R = REPLICAS_PER_GPU
    offset = rank * R
    w = wl.c2_almg_sgc_sweep(R=R, replica_offset=offset)
    ft = w.tables
    stream = torch.cuda.Stream(dev)         # a real (non-null) stream shared with the C ABI
    torch.cuda.set_stream(stream)
    gpu = wl.make_updater(w, device=local_rank, replica_offset=offset,
                          stream=stream.cuda_stream, seed=1234)
    if args.variant >= 0:
        gpu.set_variant(args.variant, args.variant)      # profiling runs: no autotuning
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)    # > 126 MB L2

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def step_resident():
        gpu.run_sgc(MOVES_PER_STEP)

    # host-side (pinned) copies of one step's inputs / outputs for the e2e leg
    occ_h = torch.from_numpy(w.occ.copy()).pin_memory()
    eci_h = torch.from_numpy(w.eci_matrix.copy()).pin_memory()
    kT_h = torch.from_numpy(w.kT.copy()).pin_memory()

    def step_e2e():
        # the call a user makes: upload configurations + (mu, T) grid, run, read back
        gpu.set_occupancy(occ_h.numpy())
        gpu.recompute_cf()
        gpu.set_ecis(eci_h.numpy())
        gpu.set_kT(kT_h.numpy())
        gpu.reset_accumulators()
        gpu.run_sgc(MOVES_PER_STEP)
        acc = gpu.get_accumulators()
        e = gpu.get_energy()
        occ = gpu.get_occupancy()
        return acc, e, occ

    for _ in range(max(args.warmup, 3)):
        step_resident()
    gpu.synchronize()

    # ---- timed: resident --------------------------------------------------
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = gpu.launch_count()
    barrier()
    t_begin = time.perf_counter()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(args.steps)]
    for k in range(args.steps):
        flush.zero_()                       # flush L2 between timed iterations
        ev[k][0].record(stream)
        step_resident()
        ev[k][1].record(stream)
    barrier()
    t_end = time.perf_counter()
    gpu.synchronize()
    launches = gpu.launch_count() - launches0
    ms_steps = [a.elapsed_time(b) for a, b in ev]
    ms_total = float(sum(ms_steps))
    clocks = sampler.stop(t_begin, t_end) if sampler else None

    # ---- timed: end to end through the host API ------------------------------
    step_e2e()
    barrier()
    e2e_ms = 0.0
    for k in range(args.steps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev)
        a.record(stream)
        step_e2e()
        b.record(stream)
        b.synchronize()
        e2e_ms += a.elapsed_time(b)
    barrier()

    t = torch.tensor([ms_total, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms = float(t[0]), float(t[1])

    if rank == 0:
        moves = float(world) * R * MOVES_PER_STEP * args.steps
        value = moves / (ms_total * 1e-3)
        e2e_value = moves / (e2e_ms * 1e-3)
        B = ft.algorithmic_bytes_per_move(1)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except (OSError, ValueError):
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        # per launch: one rank's launch processes R * MOVES_PER_STEP moves
        launch_s = (ms_total * 1e-3) / args.steps
        achieved = B * R * MOVES_PER_STEP / launch_s / 1e9
        traffic, ncu_extra = None, {}
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            traffic = tj.get("batch_kernel_sgc_dram_bytes_per_launch", None)
            ncu_extra = {k[4:]: tj[k] for k in tj if k.startswith("ncu_")}
        except (OSError, ValueError):
            pass
        h2d = w.occ.nbytes + w.eci_matrix.nbytes + w.kT.nbytes
        d2h = R * gpu.acc_stride * 8 + R * 8 + w.occ.nbytes
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": "BASELINE configs[1]: " + w.description + " per GPU",
                "replicas_per_gpu": R, "sites": ft.N, "moves_per_step_per_replica": MOVES_PER_STEP,
                "n_eci": ft.n_eci, "K": ft.K, "G": ft.gathered_sites_per_change(),
                "algorithmic_bytes_per_move": B,
                "cache": "L2 flushed (256 MiB write) between timed iterations; per-replica state "
                         "is shared-memory resident by design",
                "parallelism": "replicas sharded over %d GPU(s), no data-path collective" % world,
                "kernel_variant": "%d (0 spin, 1-4 batch (16,2)/(16,1)/(8,1)/(4,1), 5 one move at a "
                                  "time, 6-7 batch (8,1)/(16,1) with two moves per warp, 8 batch (16,2) "
                                  "site split; autotuned unless --variant)" % gpu.get_variant()[0],
            },
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650",
                         "note": "latency / issue bound by design: one dependent chain per replica, state "
                                 "in shared memory; DRAM traffic is the one-time staging",
                         "ncu": ncu_extra},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if world == 1 and not args.no_extra:
            line["other_workloads"] = extra_workloads(local_rank)
            if cpu_c3s is not None:     # the north-star target line: 64-replica Al-Mg-Si SGC sweep
                line["other_workloads"]["C3S"]["cpu_baseline"] = cpu_c3s
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def extra_workloads(device):
    """Short device-timed runs of the other BASELINE configurations (informational:
    the north-star target line is the 64-replica Al-Mg-Si SGC sweep)."""
    from cemc_b200 import workloads as wl
    out = {}
    for name, n in (("C3S", 40000), ("C3", 40000), ("C1", 40000), ("C5", 40000)):
        w = wl.WORKLOADS[name](R=1) if name == "C5" else wl.WORKLOADS[name](R=64)
        gpu = wl.make_updater(w, device=device)
        run = gpu.run_sgc if w.mode == "sgc" else gpu.run_canonical
        run(n)
        gpu.synchronize()
        best = 1e30
        for _ in range(3):
            gpu.timer_start()
            run(n)
            best = min(best, gpu.timer_stop())
        gpu.synchronize()
        out[name] = {"workload": w.description, "moves_per_s": w.R * n / (best * 1e-3),
                     "ns_per_move_per_chain": best * 1e6 / n,
                     "algorithmic_bytes_per_move": w.tables.algorithmic_bytes_per_move(w.sites_changed)}
        gpu.close()
    out["C4"] = pt_workload(device)
    return out


def pt_workload(device, rounds=40):
    """BASELINE config 4 on one GPU's shard: 64 temperatures of the ladder, fcc 12^3 ternary,
    one exchange sweep (pt_exchange_kernel, on the device) every 1728 canonical moves."""
    import torch
    from cemc_b200 import workloads as wl
    R = 64
    w = wl.c4_parallel_tempering(R=R, n_total=R)
    gpu = wl.make_updater(w, device=device)
    dev = torch.device("cuda", device)
    slots = torch.arange(R, dtype=torch.int32, device=dev)
    kts = torch.from_numpy(np.ascontiguousarray(w.kT_of_slot)).to(dev)
    n_acc = torch.zeros(1, dtype=torch.int32, device=dev)
    torch.cuda.synchronize(dev)
    sweep = w.tables.N

    def cycle(k):
        gpu.run_canonical(sweep)
        gpu.pt_exchange(R, gpu.energy_dev_ptr(), slots.data_ptr(), kts.data_ptr(), k & 1, k,
                        n_acc.data_ptr())

    for k in range(12):         # includes the autotuner's segments
        cycle(k)
    gpu.synchronize()
    gpu.timer_start()
    for k in range(rounds):
        cycle(100 + k)
    ms = gpu.timer_stop()
    gpu.synchronize()
    out = {"workload": w.description.replace("512", str(R)) + ", %d replicas on this GPU, exchange every %d moves" % (R, sweep),
           "moves_per_s": R * sweep * rounds / (ms * 1e-3),
           "ns_per_move_per_chain": ms * 1e6 / (sweep * rounds),
           "exchange_rounds_per_s": rounds / (ms * 1e-3),
           "algorithmic_bytes_per_move": w.tables.algorithmic_bytes_per_move(2)}
    gpu.close()
    return out


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # each step = a bounded sample: every host core runs `ref_moves` moves of one chain
    for _ in range(min(args.warmup, 1)):
        cpu_measure(max(args.ref_moves // 10, 200))
    rates, last = [], None
    t0 = time.perf_counter()
    for _ in range(args.steps):
        last = cpu_measure(args.ref_moves)
        rates.append(last["value"])
    wall = time.perf_counter() - t0
    value = float(np.mean(rates))
    from cemc_b200 import workloads as wl
    w = wl.c2_almg_sgc_sweep(R=1)
    last["value"] = value
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": min(args.warmup, 1),
        "ms_per_step": wall / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "BASELINE configs[1]: " + w.description.replace("1 replicas", "one chain per host core"),
                   "sites": w.tables.N, "n_eci": w.tables.n_eci,
                   "moves_per_step_per_chain": args.ref_moves},
        "cpu_baseline": last,
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-moves", type=int, default=800000,
                    help="moves per chain of the cpu_baseline sample")
    ap.add_argument("--ref-moves", type=int, default=100000,
                    help="moves per chain per step of the --impl reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the other_workloads block")
    ap.add_argument("--variant", type=int, default=-1,
                    help="pin the kernel variant (cemc_set_variant) instead of autotuning; used "
                         "for ncu runs, whose per-launch overhead defeats the autotuner's timing")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        gpu_arm(args)


if __name__ == "__main__":
    main()
