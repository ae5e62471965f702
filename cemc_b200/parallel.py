"""Replica sharding over the GPUs of one box.

Replicas are independent Markov chains: global replica g lives on rank
``g // R_local`` as local replica ``g % R_local``; the Philox stream of a
chain is keyed by its GLOBAL index, so sharding never changes a trajectory.
There is no data-path collective.  The only collectives are

* parallel tempering: one all-gather of the R_local energies per exchange
  round (NCCL over NVLink on the GPU, gloo in the CPU tests); every rank then
  evaluates the identical exchange sweep and permutes the slot<->replica map,
* end of run: all-gather of the per-replica observer sums.

The helpers take an ``engine`` (anything with get_energy / set_kT ...), so the
CPU tests can drive them with oracle chains over gloo.
"""
from __future__ import annotations

import os

import numpy as np


def dist_info():
    """(rank, world, local_rank) from the torchrun environment."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def shard_range(n_total: int, rank: int, world: int):
    """Contiguous block of global replica ids owned by ``rank``."""
    if n_total % world:
        raise ValueError("the number of replicas must be divisible by the number of ranks")
    per = n_total // world
    return rank * per, per


def shard_round_robin(n_total: int, rank: int, world: int):
    """Round-robin sharding (replica g on GPU ``g mod world``): returns
    ``(offset, stride, n_local)``; local replica r is global replica ``offset + r * stride``.
    Used by parallel tempering so that every GPU holds every ``world``-th temperature
    of the ladder (hot chains accept more and run slower: a contiguous block would leave
    one GPU with all of them)."""
    if n_total % world:
        raise ValueError("the number of replicas must be divisible by the number of ranks")
    return rank, world, n_total // world


def gather_index(n_total: int, world: int, stride: int) -> np.ndarray:
    """Position of global replica g in an all-gather (rank-major) of per-rank arrays."""
    g = np.arange(n_total)
    if stride <= 1:
        return g
    return (g % stride) * (n_total // stride) + g // stride


def all_gather_array(local: np.ndarray, world: int, device=None):
    """All-gather equally sized arrays; returns [world * n, ...]."""
    if world == 1:
        return np.ascontiguousarray(local)
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(local))
    if device is not None:
        t = t.to(device)
    out = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype,
                      device=t.device)
    dist.all_gather_into_tensor(out, t)        # concatenation along dim 0
    return out.cpu().numpy()


def exchange_sweep(energies, slot_of_replica, kT_of_slot, direction, seed, rnd,
                   uniform_fn):
    """Host restatement of the exchange sweep the device runs
    (cemc_pt_exchange; parallel_tempering.py:138-175).  ``uniform_fn(seed,
    rnd, slot)`` must return the Philox uniform of stream 2."""
    n = len(energies)
    rep_of_slot = np.empty(n, dtype=np.int64)
    rep_of_slot[np.asarray(slot_of_replica)] = np.arange(n)
    n_acc = 0
    pairs = [(i, i + 1) for i in range(0, n - 1, 2)] if direction == 0 else \
        [(i, i - 1) for i in range(n - 1, 0, -2)]
    for i, j in pairs:
        r1, r2 = rep_of_slot[i], rep_of_slot[j]
        dE = energies[r1] - energies[r2]
        db = 1.0 / kT_of_slot[i] - 1.0 / kT_of_slot[j]
        p = np.exp(db * dE)
        if uniform_fn(seed, rnd, i) < p:
            rep_of_slot[i], rep_of_slot[j] = r2, r1
            n_acc += 1
    new_slots = np.empty(n, dtype=np.int32)
    new_slots[rep_of_slot] = np.arange(n, dtype=np.int32)
    return new_slots, n_acc


class DeviceArrayF64(object):
    """A raw CUDA device pointer seen through ``__cuda_array_interface__`` so that
    ``torch.as_tensor(obj, device=...)`` aliases it (no copy): the updater's
    per-replica energies feed the NCCL all-gather of the exchange step directly."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = dict(shape=(int(n),), typestr="<f8",
                                             data=(int(ptr), False), version=2)
