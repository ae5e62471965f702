"""Sharded parallel tempering check: torchrun --nproc-per-node N scripts/pt_multi.py
Every rank builds the same problem; the sharded run must reproduce the oracle's
single-process trajectory (slots, energies) exactly."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch, torch.distributed as dist
from cemc_b200 import synthetic as syn, parallel
from cemc_b200.ce_calculator import CE
from cemc_b200.mcmc import Montecarlo, ParallelTempering
from cemc_b200.mcmc.montecarlo import KB

rank, world, local = parallel.dist_info()
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
st = syn.fcc_settings(6, ["Al", "Mg", "Si"])
eci = syn.synthetic_ecis(st)
symbols = syn.random_symbols(st, {"Al": 0.8, "Mg": 0.1, "Si": 0.1}, seed=1)
atoms = syn.Atoms(symbols)
calc = CE(atoms, st, dict(eci), device=local)
n_total = 16
temps = list(np.geomspace(1500.0, 100.0, n_total))
mc = Montecarlo(atoms, temps[0], seed=7)
pt = ParallelTempering(mc, Tmax=1500.0, Tmin=100.0, temperatures=temps, temp_scheme_file="/tmp/none.csv")
pt.run(mc_args={"steps": 500}, num_exchange_cycles=8)
e_all = pt.gather_energies()
if rank == 0:
    from oracle import ce_oracle
    from oracle.ce_oracle import OracleChain
    ft = calc.updater.tables
    cf0 = calc.updater.batch.get_cf()[0]
    chains = [OracleChain(ft, ft.occupancy(symbols), cf=cf0, kT=temps[r] * KB, seed=7, replica=r) for r in range(n_total)]
    slots = np.arange(n_total, dtype=np.int32); kts = np.array(temps) * KB
    for rnd in range(8):
        for c in chains: c.run_canonical(500)
        slots, _ = ce_oracle.pt_exchange([c.e for c in chains], slots, kts, ce_oracle.pt_direction(7, rnd), 7, rnd)
        for r, c in enumerate(chains): c.kT = float(kts[slots[r]])
    ok = np.array_equal(slots, pt.slot_of_replica) and np.array_equal(e_all, [c.e for c in chains])
    print("PT sharded over %d GPU(s): slots+energies identical to oracle: %s; accepted exchanges %d" % (world, ok, pt.num_accepted_exchanges))
    assert ok
if world > 1:
    dist.barrier(); dist.destroy_process_group()
