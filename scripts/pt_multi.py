"""Sharded parallel tempering check: torchrun --nproc-per-node N scripts/pt_multi.py
Every rank builds the same problem; the sharded run (round-robin replicas, NCCL all-gather of the
energies, device-side exchange sweeps, sync-free round loop) must reproduce the oracle's
single-process trajectory -- slots AND energies of every replica -- exactly.
Environment: PT_L (cell, default 6), PT_N (temperatures, 16), PT_STEPS (moves per leg, 500),
PT_ROUNDS (8).  BASELINE configs[3]: PT_L=12 PT_N=512 PT_STEPS=1728."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch, torch.distributed as dist
from cemc_b200 import synthetic as syn, parallel
from cemc_b200.ce_calculator import CE
from cemc_b200.mcmc import Montecarlo, ParallelTempering
from cemc_b200.mcmc.montecarlo import KB

L = int(os.environ.get("PT_L", "6")); n_total = int(os.environ.get("PT_N", "16"))
steps = int(os.environ.get("PT_STEPS", "500")); rounds = int(os.environ.get("PT_ROUNDS", "8"))
rank, world, local = parallel.dist_info()
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
st = syn.fcc_settings(L, ["Al", "Mg", "Si"])
eci = syn.synthetic_ecis(st)
symbols = syn.random_symbols(st, {"Al": 0.8, "Mg": 0.1, "Si": 0.1}, seed=1)
atoms = syn.Atoms(symbols)
calc = CE(atoms, st, dict(eci), device=local)
temps = list(np.geomspace(1500.0, 100.0, n_total))
mc = Montecarlo(atoms, temps[0], seed=7)
pt = ParallelTempering(mc, Tmax=1500.0, Tmin=100.0, temperatures=temps, temp_scheme_file="/tmp/none.csv")
t0 = time.perf_counter()
pt.run(mc_args={"steps": steps}, num_exchange_cycles=rounds)
t_gpu = time.perf_counter() - t0
e_all = pt.gather_energies()
if rank == 0:
    from concurrent.futures import ThreadPoolExecutor
    from oracle import ce_oracle
    from oracle.ce_oracle import OracleChain
    ft = calc.updater.tables
    cf0 = calc.updater.batch.get_cf()[0]
    chains = [OracleChain(ft, ft.occupancy(symbols), cf=cf0, kT=temps[r] * KB, seed=7, replica=r) for r in range(n_total)]
    slots = np.arange(n_total, dtype=np.int32); kts = np.array(temps) * KB
    t0 = time.perf_counter()
    total = 0
    with ThreadPoolExecutor(os.cpu_count() or 4) as pool:       # the C oracle releases the GIL
        for rnd in range(rounds):
            list(pool.map(lambda c: c.run_canonical(steps), chains))
            slots, n_acc = ce_oracle.pt_exchange([c.e for c in chains], slots, kts, ce_oracle.pt_direction(7, rnd), 7, rnd)
            total += n_acc
            for r, c in enumerate(chains): c.kT = float(kts[slots[r]])
    t_cpu = time.perf_counter() - t0
    ok = np.array_equal(slots, pt.slot_of_replica) and np.array_equal(e_all, [c.e for c in chains]) \
        and total == pt.num_accepted_exchanges
    print("PT sharded over %d GPU(s): %d temperatures x fcc %d^3 ternary, %d rounds of %d moves: slots + energies of "
          "all replicas identical to the single-process oracle: %s; accepted exchanges %d; wall %.2f s (GPUs, incl. "
          "tuning) vs %.1f s (oracle on %d host threads)" % (world, n_total, L, rounds, steps, ok,
                                                             pt.num_accepted_exchanges, t_gpu, t_cpu, os.cpu_count() or 4))
    assert ok
if world > 1:
    dist.barrier(); dist.destroy_process_group()
