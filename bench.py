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
  roofline, roofline_issue, cpu_baseline: see DESIGN.md "Measurement"

Further blocks of the same JSON line (all measured by this run, same rules: warm-up,
L2 flush between timed iterations, CUDA events, mean over the timed steps):

  north_star : the north-star target workload -- 64-replica Al-Mg-Si SGC sweep (fcc 20^3) --
               with its own value / e2e / roofline / cpu_baseline (N = 1 only)
  pt         : BASELINE configs[3], the one workload with a collective: parallel tempering,
               64 temperatures per GPU (512 on 8 GPUs), fcc 12^3 ternary, exchange every 1728
               moves, through cemc_b200.mcmc.ParallelTempering's sync-free round loop (NCCL
               all-gather of the energies + device-side exchange sweep), at every N
  c5_ensemble: BASELINE configs[4], fcc 64^3 single-chain exact Metropolis (CTA cluster), one
               chain (seed) per GPU, at every N
  other_workloads: configs[0], configs[2] (N = 1 only)

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
PT_REPLICAS_PER_GPU = 64
PT_ROUNDS = 100


# ----------------------------------------------------------------------------
# CPU arm: the reference's compiled CEUpdater (or the oracle port) on host cores
def _cpu_worker(args):
    """One chain on one core; returns (moves, seconds, kind)."""
    replica, n_moves, warm, which, threads = args
    sys.path.insert(0, ROOT)
    from cemc_b200 import workloads as wl
    from oracle import ref_driver
    from oracle.ce_oracle import OracleChain
    w = wl.WORKLOADS[which](R=1, replica_offset=replica)
    ft = w.tables
    eci_vec = w.eci_matrix[0] if w.eci_matrix is not None else ft.eci
    oc = OracleChain(ft, w.occ[0], kT=w.kT[0], seed=1234, replica=replica, eci=eci_vec)
    run = oc.run_sgc if w.mode == "sgc" else oc.run_canonical
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
        rc = ref_driver.RefChain(st, ft.symbols_of(w.occ[0]), eci, cf0, kT=w.kT[0], num_threads=threads)
        tr = run(warm + n_moves, trace=True)
        rc.replay(ft.species, tr[0][:warm], tr[1][:warm], tr[2][:warm])
        t0 = time.perf_counter()
        acc, _, _ = rc.replay(ft.species, tr[0][warm:], tr[1][warm:], tr[2][warm:])
        dt = time.perf_counter() - t0
        assert np.array_equal(acc, tr[3][warm:])    # the two CPU engines agree
        return n_moves, dt, "reference"
    run(warm)
    t0 = time.perf_counter()
    run(n_moves * 10)
    return n_moves * 10, time.perf_counter() - t0, "port"


def cpu_measure(n_moves_per_chain, n_procs=None, warm=500, which="C2", threads=1):
    n_procs = n_procs or os.cpu_count() or 1
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(n_procs) as pool:
        res = pool.map(_cpu_worker, [(r, n_moves_per_chain, warm, which, threads) for r in range(n_procs)])
    wall = time.perf_counter() - t0
    # chains run concurrently, one per core: aggregate = sum of per-chain rates
    rate = sum(m / dt for m, dt, _ in res)
    kind = res[0][2]
    return dict(value=rate, unit=UNIT, cores=n_procs * threads, kind=kind,
                sample="%d trial moves on each of %d concurrent chains (one per host core) "
                       "of the %s workload, %s; wall %.1f s incl. setup" % (
                           res[0][0], n_procs, which,
                           "reference C++ CEUpdater driven through its Cython PyCEUpdater"
                           if kind == "reference" else "C oracle port", wall))


def cpu_modes(which="C2", n_moves=60000):
    """SURVEY.md 8(d) CPU modes (i) and (ii) on this box: ONE chain, the reference's own OpenMP
    loop over the ECIs (ce_updater.cpp:352) with set_num_threads(t).  Mode (iii), one chain per
    core on all cores, is the cpu_baseline value itself."""
    out = {}
    ncpu = os.cpu_count() or 1
    for t in (1, 2, 4, 8):
        if t > ncpu:
            break
        r = cpu_measure(n_moves if t == 1 else n_moves // 2, n_procs=1, which=which, threads=t)
        out["one_chain_%d_thread%s" % (t, "" if t == 1 else "s")] = r["value"]
        if r["kind"] != "reference":      # the oracle port has no OpenMP mode
            break
    return out


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


def _peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except (OSError, ValueError):
        return {}


def _profile_figures(key):
    """Profiler-derived figures of workload `key` from profiles/traffic.json -- only when they
    were captured on THIS build (kernel source hash recorded with them); otherwise null."""
    from cemc_b200 import _lib
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except (OSError, ValueError):
        return None, "profiles/traffic.json missing"
    have = _lib.source_hash()
    if tj.get("kernel_source_sha") != have:
        return None, "profiles/traffic.json was captured on another build (%s, this build %s): not reported" % (
            tj.get("kernel_source_sha"), have)
    return tj.get(key), "profiles/traffic.json (ncu --set full on this build, %s)" % tj.get(key, {}).get("source", "")


def rooflines(key, bytes_per_move, moves_per_launch, launch_s, moves_per_s, clocks, n_sms=148):
    """(a) HBM roofline of SURVEY.md 8(d): algorithmic bytes / launch time vs the measured copy
    bandwidth.  (b) the bound that means something for this path (chain state lives in shared
    memory): issue slots -- warp instructions per move (ncu, this build) x moves/s vs the
    SMs' issue rate 148 SMs x 4 schedulers x SM clock."""
    peaks = _peaks()
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = bytes_per_move * moves_per_launch / launch_s / 1e9
    fig, src = _profile_figures(key)
    roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": fig.get("dram_bytes_per_launch") if fig else None,
            "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 (B200_PROFILING.md)",
            "traffic_source": src,
            "note": "latency / issue bound by design: one dependent chain per replica, state in shared "
                    "memory; DRAM traffic is the one-time staging per launch"}
    issue = None
    if fig and fig.get("warp_instructions_per_move") and clocks:
        wi = float(fig["warp_instructions_per_move"])
        peak_i = n_sms * 4 * clocks["sm_mhz"] * 1e6
        issue = {"bound": "issue", "warp_instructions_per_move": wi, "achieved": wi * moves_per_s,
                 "peak": peak_i, "unit": "warp-instr/s", "frac": wi * moves_per_s / peak_i,
                 "peak_source": "%d SMs x 4 schedulers x %.0f MHz (median SM clock sampled during the timed region)" % (
                     n_sms, clocks["sm_mhz"]),
                 "ncu_issue_slots_busy_pct": fig.get("issue_slots_busy_pct")}
    return roof, issue


class Timed(object):
    """W warm-up + K timed steps of one workload on one GPU: resident and end-to-end."""

    def __init__(self, torch, dev, stream, flush, barrier):
        self.torch, self.dev, self.stream, self.flush, self.barrier = torch, dev, stream, flush, barrier

    def run(self, step, steps, warmup):
        torch = self.torch
        for _ in range(max(warmup, 3)):
            step()
        self.barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
              for _ in range(steps)]
        t0 = time.perf_counter()
        for k in range(steps):
            self.flush.zero_()                       # flush L2 between timed iterations
            ev[k][0].record(self.stream)
            step()
            ev[k][1].record(self.stream)
        self.barrier()
        t1 = time.perf_counter()
        return float(sum(a.elapsed_time(b) for a, b in ev)), t0, t1

    def run_e2e(self, step, steps):
        torch = self.torch
        step()
        self.barrier()
        ms = 0.0
        for _ in range(steps):
            self.flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(self.dev)
            a.record(self.stream)
            step()
            b.record(self.stream)
            b.synchronize()
            ms += a.elapsed_time(b)
        self.barrier()
        return ms


def make_steps(gpu, w, moves):
    """(resident step, end-to-end step, h2d bytes, d2h bytes) of a workload."""
    import torch
    run = gpu.run_sgc if w.mode == "sgc" else gpu.run_canonical
    occ_h = torch.from_numpy(w.occ.copy()).pin_memory()
    kT_h = torch.from_numpy(w.kT.copy()).pin_memory()
    eci_h = torch.from_numpy(w.eci_matrix.copy()).pin_memory() if w.eci_matrix is not None else None

    def resident():
        run(moves)

    def e2e():
        # the call a user makes: upload configurations + (mu, T) grid, run, read back
        gpu.set_occupancy(occ_h.numpy())
        gpu.recompute_cf()
        if eci_h is not None:
            gpu.set_ecis(eci_h.numpy())
        gpu.set_kT(kT_h.numpy())
        gpu.reset_accumulators()
        run(moves)
        return gpu.get_accumulators(), gpu.get_energy(), gpu.get_occupancy()

    h2d = w.occ.nbytes + (w.eci_matrix.nbytes if w.eci_matrix is not None else 0) + w.kT.nbytes
    d2h = w.R * gpu.acc_stride * 8 + w.R * 8 + w.occ.nbytes
    return resident, e2e, int(h2d), int(d2h)


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

    # CPU baselines first (rank 0, N=1 only), before this process touches CUDA
    cpu = cpu_ns = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_measure(args.cpu_moves)
        cpu["modes"] = cpu_modes("C2")
        cpu["modes"]["one_chain_per_core_all_%d_cores" % cpu["cores"]] = cpu["value"]
        if not args.no_extra:
            cpu_ns = cpu_measure(args.cpu_moves // 4, which="C3S")

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

    timed = Timed(torch, dev, stream, flush, barrier)
    resident, e2e_step, h2d, d2h = make_steps(gpu, w, MOVES_PER_STEP)

    # ---- timed: resident, then end to end through the host API -------------------------
    for _ in range(max(args.warmup, 3)):
        resident()                      # includes the autotuner's segments
    gpu.synchronize()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    launches0 = gpu.launch_count()
    ms_total, t_begin, t_end = timed.run(resident, args.steps, 0 if args.warmup >= 3 else 3)
    gpu.synchronize()
    launches = gpu.launch_count() - launches0
    clocks = sampler.stop(t_begin, t_end) if sampler else None
    e2e_ms = timed.run_e2e(e2e_step, args.steps)

    t = torch.tensor([ms_total, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms = float(t[0]), float(t[1])
    variant = gpu.get_variant()[0]
    gpu.close()

    # ---- the other blocks (every rank takes part in pt / c5_ensemble) --------------------
    pt = c5 = north = others = None
    if not args.no_extra:
        pt = pt_block(torch, dist, dev, local_rank, rank, world, barrier)
        c5 = c5_block(torch, dist, dev, local_rank, rank, world, timed)
        if world == 1:
            north = north_star_block(torch, local_rank, timed, args, cpu_ns)
            others = extra_workloads(local_rank, timed, args)

    if rank == 0:
        moves = float(world) * R * MOVES_PER_STEP * args.steps
        value = moves / (ms_total * 1e-3)
        e2e_value = moves / (e2e_ms * 1e-3)
        B = ft.algorithmic_bytes_per_move(1)
        # per launch: one rank's launch processes R * MOVES_PER_STEP moves
        launch_s = (ms_total * 1e-3) / args.steps
        roof, issue = rooflines("C2", B, R * MOVES_PER_STEP, launch_s, value / world, clocks)
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
                "parallelism": "replicas sharded over %d GPU(s), no data-path collective (the pt block "
                               "is the workload with one)" % world,
                "kernel_variant": "%d (0 spin, 1-4 batch (16,2)/(16,1)/(8,1)/(4,1), 5 one move at a "
                                  "time, 6 batch (8,1) two moves per warp, 8/9 batch (16,2)/(8,2) site split; "
                                  "autotuned unless --variant)" % variant,
            },
            "roofline": roof,
            "roofline_issue": issue,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if north is not None:
            line["north_star"] = north
        if pt is not None:
            line["pt"] = pt
        if c5 is not None:
            line["c5_ensemble"] = c5
        if others is not None:
            line["other_workloads"] = others
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def north_star_block(torch, device, timed, args, cpu_ns):
    """The north-star target line as a full record: 64-replica Al-Mg-Si SGC sweep, fcc 20^3."""
    from cemc_b200 import workloads as wl
    w = wl.c3s_almgsi_sgc(R=64)
    gpu = wl.make_updater(w, device=device, stream=timed.stream.cuda_stream)
    resident, e2e_step, h2d, d2h = make_steps(gpu, w, MOVES_PER_STEP)
    for _ in range(3):
        resident()
    gpu.synchronize()
    sampler = ClockSampler(device)
    l0 = gpu.launch_count()
    ms, t0, t1 = timed.run(resident, args.steps, 0)
    gpu.synchronize()
    launches = gpu.launch_count() - l0
    clocks = sampler.stop(t0, t1)
    e2e_ms = timed.run_e2e(e2e_step, args.steps)
    moves = w.R * MOVES_PER_STEP * args.steps
    value = moves / (ms * 1e-3)
    B = w.tables.algorithmic_bytes_per_move(1)
    roof, issue = rooflines("C3S", B, w.R * MOVES_PER_STEP, ms * 1e-3 / args.steps, value, clocks)
    out = {"workload": "north-star target: " + w.description + " on one B200",
           "metric": METRIC, "value": value, "unit": UNIT, "steps": args.steps,
           "ms_per_step": ms / args.steps, "ns_per_move_per_chain": ms * 1e6 / (MOVES_PER_STEP * args.steps),
           "moves_per_step_per_replica": MOVES_PER_STEP, "timing": "mean over the timed steps, L2 flushed between them",
           "algorithmic_bytes_per_move": B, "kernel_variant": gpu.get_variant()[0],
           "e2e": {"value": moves / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                   "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / args.steps},
           "roofline": roof, "roofline_issue": issue, "gpu_launches": int(launches), "clocks": clocks}
    if cpu_ns is not None:
        out["cpu_baseline"] = cpu_ns
        out["e2e_vs_cpu_baseline"] = out["e2e"]["value"] / cpu_ns["value"]
    gpu.close()
    return out


def extra_workloads(device, timed, args):
    """Device-timed runs of the remaining BASELINE configurations (mean over the timed steps)."""
    from cemc_b200 import workloads as wl
    out = {}
    for name, n in (("C3", 40000), ("C1", 40000)):
        w = wl.WORKLOADS[name](R=64)
        gpu = wl.make_updater(w, device=device, stream=timed.stream.cuda_stream)
        run = gpu.run_sgc if w.mode == "sgc" else gpu.run_canonical
        ms, _, _ = timed.run(lambda: run(n), 5, 3)
        gpu.synchronize()
        out[name] = {"workload": w.description, "moves_per_s": w.R * n * 5 / (ms * 1e-3),
                     "ns_per_move_per_chain": ms * 1e6 / (n * 5), "kernel_variant": gpu.get_variant(),
                     "algorithmic_bytes_per_move": w.tables.algorithmic_bytes_per_move(w.sites_changed)}
        gpu.close()
    return out


def c5_block(torch, dist, dev, device, rank, world, timed, n=40000, steps=5):
    """BASELINE configs[4]: fcc 64^3 single-chain exact canonical Metropolis (CTA cluster of 2 on
    one chain); at N > 1 an ensemble of N seeds, one chain per GPU (replicas only, no collective)."""
    from cemc_b200 import workloads as wl
    w = wl.c5_large_supercell(R=1, replica_offset=rank)
    gpu = wl.make_updater(w, device=device, replica_offset=rank, stream=timed.stream.cuda_stream)
    ms, _, _ = timed.run(lambda: gpu.run_canonical(n), steps, 3)
    gpu.synchronize()
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0])
    out = {"workload": "BASELINE configs[4]: " + w.description + " per GPU, %d-seed ensemble" % world,
           "moves_per_s": world * n * steps / (ms * 1e-3), "ns_per_move_per_chain": ms * 1e6 / (n * steps),
           "chains": world, "kernel_variant": gpu.get_variant()[1], "scaling": "weak (one chain per GPU)"}
    gpu.close()
    return out


def pt_block(torch, dist, dev, device, rank, world, barrier, rounds=PT_ROUNDS):
    """BASELINE configs[3] through the public API (cemc_b200.mcmc.ParallelTempering): 64
    temperatures per GPU of a geometric ladder in [100, 1500] K (512 on 8 GPUs), fcc 12^3
    ternary, one exchange sweep every 1728 canonical moves.  Replicas are sharded round-robin
    (replica g on GPU g mod N); per cycle: leg kernel -> NCCL all-gather of the energies ->
    device-side exchange kernel, all on one stream without host synchronisation."""
    from cemc_b200 import synthetic as syn
    from cemc_b200.ce_calculator import CE
    from cemc_b200.mcmc import Montecarlo, ParallelTempering
    L = 12
    st = syn.fcc_settings(L, ["Al", "Mg", "Si"], syn.STANDARD_FAMILIES)
    eci = syn.synthetic_ecis(st, seed=1234)
    symbols = syn.random_symbols(st, {"Al": 0.8, "Mg": 0.1, "Si": 0.1}, seed=4000)
    atoms = syn.Atoms(symbols)
    calc = CE(atoms, st, dict(eci), device=device)
    n_total = PT_REPLICAS_PER_GPU * world
    temps = list(np.geomspace(1500.0, 100.0, n_total))
    mc = Montecarlo(atoms, temps[0], seed=7)
    pt = ParallelTempering(mc, Tmax=1500.0, Tmin=100.0, temperatures=temps,
                           temp_scheme_file="/tmp/cemc_b200_no_scheme.csv", device=device)
    sweep = len(atoms)
    pt.run(mc_args={"steps": sweep}, num_exchange_cycles=14)     # autotuner: one variant per leg
    barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = pt.gpu.launch_count()
    t0.record(pt._stream)
    pt.run(mc_args={"steps": sweep}, num_exchange_cycles=rounds, timing=True)
    t1.record(pt._stream)
    barrier()
    launches = pt.gpu.launch_count() - l0
    tm = pt.last_timing
    t = torch.tensor([t0.elapsed_time(t1), tm["leg_ms"], tm["exchange_ms"]], dtype=torch.float64, device=dev)
    tmin = t.clone()
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
    ms, leg_max, exch_max = (float(x) for x in t)
    leg_min, exch_min = float(tmin[1]), float(tmin[2])
    out = {"workload": "BASELINE configs[3]: parallel tempering, %d temperatures (geometric, 100-1500 K), fcc %d^3 "
                       "ternary canonical, %d replicas per GPU sharded round-robin over %d GPU(s), exchange every %d moves"
                       % (n_total, L, PT_REPLICAS_PER_GPU, world, sweep),
           "api": "cemc_b200.mcmc.ParallelTempering.run (sync-free round loop)",
           "rounds": rounds, "moves_per_s": n_total * sweep * rounds / (ms * 1e-3),
           "exchange_rounds_per_s": rounds / (ms * 1e-3),
           "ns_per_move_per_chain": ms * 1e6 / (sweep * rounds),
           "us_per_round": ms * 1e3 / rounds,
           "us_leg_kernel_per_round": {"max_over_ranks": leg_max * 1e3 / rounds, "min_over_ranks": leg_min * 1e3 / rounds},
           "us_allgather_plus_exchange_per_round": {"max_over_ranks": exch_max * 1e3 / rounds,
                                                    "min_over_ranks": exch_min * 1e3 / rounds},
           "fraction_of_round_outside_leg_kernel": exch_min / (leg_max + exch_min) if (leg_max + exch_min) > 0 else None,
           "note": "the all-gather waits for the slowest rank's leg: max-over-ranks of the exchange part contains "
                   "that wait, min-over-ranks is the collective + exchange kernel themselves",
           "collective": "NCCL all_gather_into_tensor of %d fp64 energies per rank" % PT_REPLICAS_PER_GPU if world > 1
                         else "none (single GPU: the exchange kernel reads the local energies)",
           "accepted_exchanges": pt.num_accepted_exchanges, "kernel_variant": pt.gpu.get_variant()[1],
           "gpu_launches": int(launches), "scaling": "weak (64 temperatures per GPU)", "n_gpus": world}
    pt.gpu.close()
    return out


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # each step = a bounded sample: every host core runs `ref_moves` moves of one chain
    for _ in range(args.warmup):
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
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
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
    ap.add_argument("--no-extra", action="store_true",
                    help="skip the north_star / pt / c5_ensemble / other_workloads blocks")
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
