"""World-size-2 gloo tests of the host-side sharding / parallel-tempering
plumbing (cemc_b200/parallel.py) with oracle chains as the compute engine."""
import os
import socket

import numpy as np
import pytest

from cases import BINARY, build
from cemc_b200 import parallel
from oracle import ce_oracle
from oracle.ce_oracle import OracleChain

N_TOTAL = 6
SEED = 4242


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _uniform(seed, rnd, slot):
    w = ce_oracle.philox(seed, rnd, slot, 2)
    return ((w[0] >> 5) * 67108864.0 + (w[1] >> 6)) / 9007199254740992.0


def _chains(ft, symbols, kts, ids):
    return [OracleChain(ft, ft.occupancy(symbols), kT=kts[g], seed=SEED, replica=g) for g in ids]


def _worker(rank, world, port, q, layout="contiguous"):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    st, eci, symbols, ft = build(**BINARY)
    kts = np.geomspace(0.08, 0.01, N_TOTAL)
    r, w, _ = parallel.dist_info()
    if layout == "contiguous":
        off, n_loc = parallel.shard_range(N_TOTAL, r, w)
        stride = 1
    else:       # replica g on rank g mod world (what ParallelTempering uses)
        off, stride, n_loc = parallel.shard_round_robin(N_TOTAL, r, w)
    ids = [off + i * stride for i in range(n_loc)]
    order = parallel.gather_index(N_TOTAL, w, stride)      # all-gather (rank-major) -> global order
    chains = _chains(ft, symbols, kts, ids)
    slots = np.arange(N_TOTAL, dtype=np.int32)
    for rnd in range(5):
        for c in chains:
            c.run_canonical(150)
        e_all = parallel.all_gather_array(np.array([c.e for c in chains]), w)[order]
        d = rnd % 2 if layout == "contiguous" else ce_oracle.pt_direction(SEED, rnd)
        slots, _ = parallel.exchange_sweep(e_all, slots, kts, d, SEED, rnd, _uniform)
        for g, c in zip(ids, chains):
            c.kT = float(kts[slots[g]])
    acc_all = parallel.all_gather_array(np.stack([c.acc for c in chains]), w)[order]
    e_all = parallel.all_gather_array(np.array([c.e for c in chains]), w)[order]
    if rank == 0:
        q.put((slots, e_all, acc_all))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("layout", ["contiguous", "round_robin"])
def test_sharded_parallel_tempering_gloo(layout):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, layout)) for r in range(2)]
    for p in procs:
        p.start()
    import queue
    res = None
    for _ in range(240):
        try:
            res = q.get(timeout=0.5)
            break
        except queue.Empty:
            if any(p.exitcode not in (None, 0) for p in procs):
                break
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res is not None
    slots, e_all, acc_all = res
    # single-process reference: all six chains + the oracle's exchange sweep
    st, eci, symbols, ft = build(**BINARY)
    kts = np.geomspace(0.08, 0.01, N_TOTAL)
    chains = _chains(ft, symbols, kts, range(N_TOTAL))
    ref_slots = np.arange(N_TOTAL, dtype=np.int32)
    for rnd in range(5):
        for c in chains:
            c.run_canonical(150)
        d = rnd % 2 if layout == "contiguous" else ce_oracle.pt_direction(SEED, rnd)
        ref_slots, _ = ce_oracle.pt_exchange([c.e for c in chains], ref_slots, kts, d, SEED, rnd)
        for g, c in enumerate(chains):
            c.kT = float(kts[ref_slots[g]])
    assert np.array_equal(slots, ref_slots)
    assert np.array_equal(e_all, [c.e for c in chains])
    assert np.array_equal(acc_all, np.stack([c.acc for c in chains]))


def test_shard_range_and_exchange_sweep_host():
    assert parallel.shard_range(512, 3, 8) == (192, 64)
    assert parallel.shard_round_robin(512, 3, 8) == (3, 8, 64)
    # replica g = 8 l + k is element l of rank k's block in an all-gather
    gi = parallel.gather_index(512, 8, 8)
    assert gi[3] == 3 * 64 and gi[8 + 3] == 3 * 64 + 1 and sorted(gi) == list(range(512))
    assert np.array_equal(parallel.gather_index(6, 2, 1), np.arange(6))
    with pytest.raises(ValueError):
        parallel.shard_range(10, 0, 4)
    rng = np.random.default_rng(1)
    kts = np.geomspace(0.1, 0.01, 7)
    e = rng.normal(size=7)
    slots = rng.permutation(7).astype(np.int32)
    for d in (0, 1):
        a, na = parallel.exchange_sweep(e, slots, kts, d, 9, 3, _uniform)
        b, nb = ce_oracle.pt_exchange(e, slots, kts, d, 9, 3)
        assert np.array_equal(a, b) and na == nb
