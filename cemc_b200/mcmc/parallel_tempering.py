"""Parallel tempering -- mirror of ``cemc.mcmc.ParallelTempering``
(/root/reference/cemc/mcmc/parallel_tempering.py:6-191), with every replica
advanced by ONE kernel launch per cycle and the exchange sweep on the device.

The round loop is sync-free: a cycle enqueues, on ONE CUDA stream,

    leg kernel (cemc_run_canonical)  ->  all-gather of the local energies (NCCL over
    NVLink, only when sharded)  ->  exchange kernel (cemc_pt_exchange)

and the host never waits: the slot map, the per-cycle direction (drawn on the device
from the counter stream) and the accepted-exchange count stay on the device and are
read back once when ``run`` returns.  ``timing=True`` brackets the two halves of every
cycle with CUDA events (no synchronisation either) for ``last_timing``.

Differences from the reference (DESIGN.md "Parallel tempering"):

* replicas run concurrently (one CTA / cluster each) instead of sequentially;
* an accepted exchange permutes the temperature<->replica map; the reference
  copies whole configurations site by site (:146-151, up to 2N update_cf);
* the exchange uses the true energies.  The reference compares
  ``current_energy`` values from which each replica's own energy-bias probe
  was subtracted (montecarlo.py:189, parallel_tempering.py:161-162), which
  skews its criterion by the difference of the two biases;
* the temperature ladder can be given explicitly (``temperatures=``); the
  reference's bisection search (:47-136) is kept as ``_init_temperature_scheme``;
* replicas can be sharded over ranks (torch.distributed), round-robin
  (replica g on GPU ``g mod world``) so that every GPU holds
  every world-th temperature; the only collective is the all-gather of the
  energies.  Chains are keyed by their global replica id: the sharded run is
  the single-process run, bit for bit.
* only the canonical sampler is supported (the reference's class effectively
  is canonical-only as well: it never forwards chemical potentials).
"""
from __future__ import annotations

import numpy as np

from .. import parallel
from ..updater import BatchedCEUpdater
from .montecarlo import KB, Montecarlo


class ParallelTempering(object):
    def __init__(self, mc_obj=None, Tmax=1500.0, Tmin=100.0,
                 temp_scheme_file="temp_scheme.csv", temperatures=None,
                 target_accept=0.2, seed=None, device=None):
        if not isinstance(mc_obj, Montecarlo):
            raise TypeError("mc_obj has to be of type Montecarlo!")
        if getattr(mc_obj, "name", "MonteCarlo") != "MonteCarlo":
            # chemical potentials / restricted symbol lists of an SGC sampler would be
            # silently dropped (the replicas are built from the calculator's ECIs)
            raise TypeError("ParallelTempering supports the canonical Montecarlo "
                            "sampler only (got {})".format(mc_obj.name))
        self.mc = mc_obj
        mc_obj.T = Tmax
        self.natoms = len(mc_obj.atoms)
        self.Tmax, self.Tmin = Tmax, Tmin
        self.temperature_schedule_fname = temp_scheme_file
        self.seed = mc_obj.seed if seed is None else int(seed)
        self.rank, self.world, local_rank = parallel.dist_info()
        self.device = local_rank if device is None else device
        if temperatures is None:
            temperatures = self._init_temperature_scheme(target_accept)
        self._temps = [float(T) for T in temperatures]
        self.n_total = len(self._temps)
        self.offset, self.stride, self.R = parallel.shard_round_robin(
            self.n_total, self.rank, self.world)
        calc = mc_obj.atoms.get_calculator()
        self.tables = calc.updater.tables
        import torch
        self._torch = torch
        self._dev = torch.device("cuda", self.device)
        # one stream for the leg kernels, the NCCL all-gather and the exchange kernel
        self._stream = torch.cuda.Stream(self._dev)
        self.gpu = BatchedCEUpdater(self.tables, self.R, device=self.device,
                                    replica_offset=self.offset, replica_stride=self.stride,
                                    stream=self._stream.cuda_stream)
        occ = calc.updater.batch.get_occupancy()
        cf = calc.updater.batch.get_cf()
        self.gpu.set_occupancy(np.repeat(occ, self.R, axis=0))
        self.gpu.set_cf(np.repeat(cf, self.R, axis=0))
        self.gpu.set_ecis(self.tables.eci_vector(calc.eci))
        self.gpu.seed(self.seed)
        self.kT_of_slot = np.array(self._temps) * KB        # slot 0 = Tmax
        self.slot_of_replica = np.arange(self.n_total, dtype=np.int32)
        self.gpu.set_kT(self.kT_of_slot[self.global_ids()])
        self.round = 0                    # exchange cycles done: counter of the direction /
        self.num_accepted_exchanges = 0   # uniform streams, advances across run() calls
        self.last_timing = None
        self._bufs = None

    def global_ids(self):
        """Global replica ids of the local replicas."""
        return self.offset + self.stride * np.arange(self.R)

    def _log(self, msg):
        if self.rank == 0:
            print(msg)

    @property
    def temperature_scheme(self):
        return list(self._temps)

    # ---- ladder construction (parallel_tempering.py:47-136) -----------------------
    def _init_temperature_scheme_from_file(self):
        try:
            data = np.loadtxt(self.temperature_schedule_fname, delimiter=',')
            return [self.Tmax] + np.atleast_2d(data)[:, 0].tolist()
        except (IOError, OSError):
            return None

    def _accept_probability(self, E1, E2, T1, T2):
        dE = E1 - E2
        return np.exp((1.0 / (KB * T1) - 1.0 / (KB * T2)) * dE)

    def _mean_energy_at(self, T, nsteps):
        self.mc.T = T
        self.mc.runMC(steps=nsteps, equil=False)
        return self.mc.get_thermodynamic()["energy"]

    def _init_temperature_scheme(self, target_accept=0.2):
        from_file = self._init_temperature_scheme_from_file()
        if from_file is not None:
            return from_file
        nsteps = 10 * self.natoms
        temps, accs = [self.Tmax], [0.0]
        cur_E = self._mean_energy_at(self.Tmax, nsteps)
        while temps[-1] > self.Tmin:
            cur_T = temps[-1]
            trial, acc, found = cur_T / 2.0, 1.0, False
            while acc > target_accept and trial > self.Tmin:
                E = self._mean_energy_at(trial, nsteps)
                acc = self._accept_probability(cur_E, E, cur_T, trial)
                found = acc <= target_accept
                trial /= 2.0
            if not found:
                break
            upper, lower = cur_T, trial
            while True:                                   # bisection (:113-135)
                new_T = 0.5 * (upper + lower)
                new_E = self._mean_energy_at(new_T, nsteps)
                new_acc = self._accept_probability(cur_E, new_E, cur_T, new_T)
                if new_acc > target_accept:
                    upper = new_T
                else:
                    lower = new_T
                if abs(new_acc - target_accept) < 0.01 or upper - lower < 1E-4:
                    break
            temps.append(new_T)
            accs.append(float(new_acc))
            cur_E = new_E
        if self.rank == 0:
            np.savetxt(self.temperature_schedule_fname, np.vstack((temps[1:], accs[1:])).T,
                       delimiter=",", header="Temperature (K), Acceptance probabability")
        return temps

    # ---- exchange ------------------------------------------------------------------
    def _device_buffers(self):
        if self._bufs is None:
            torch, dev = self._torch, self._dev
            self._bufs = dict(
                slots=torch.from_numpy(self.slot_of_replica.copy()).to(dev),
                kts=torch.from_numpy(self.kT_of_slot.copy()).to(dev),
                e_all=torch.empty(self.n_total, dtype=torch.float64, device=dev),
                # the local energies never leave the device: alias the updater's buffer
                e_loc=torch.as_tensor(parallel.DeviceArrayF64(self.gpu.energy_dev_ptr(), self.R),
                                      device=dev),
                n_acc=torch.zeros(2, dtype=torch.int32, device=dev))
        return self._bufs

    def _enqueue_exchange(self, direction):
        """All-gather + exchange sweep of one cycle, enqueued on the stream (no host sync).
        direction: 0 "up", 1 "down", -1 drawn on the device from (seed, round)."""
        b = self._device_buffers()
        if self.world > 1:
            import torch.distributed as dist
            dist.all_gather_into_tensor(b["e_all"], b["e_loc"])      # NCCL over NVLink, stream-ordered
            e_ptr = b["e_all"].data_ptr()
        else:
            e_ptr = b["e_loc"].data_ptr()
        self.gpu.pt_exchange(self.n_total, e_ptr, b["slots"].data_ptr(), b["kts"].data_ptr(),
                             direction, self.round, b["n_acc"].data_ptr())
        self.round += 1

    def _read_back(self):
        b = self._device_buffers()
        self._stream.synchronize()
        self.slot_of_replica = b["slots"].cpu().numpy()
        self.num_accepted_exchanges = int(b["n_acc"][1].item())

    def _perform_exchange_move(self, direction="up"):
        """One exchange sweep (parallel_tempering.py:153-175), on the device; returns the
        number of accepted exchanges (this entry point reads the result back)."""
        torch = self._torch
        with torch.cuda.stream(self._stream):
            self._enqueue_exchange(0 if direction == "up" else 1)
        self._read_back()
        return int(self._bufs["n_acc"][0].item())

    def run(self, mc_args={}, num_exchange_cycles=10, timing=False):
        """``num_exchange_cycles`` x (``steps`` moves on every replica, then one exchange
        sweep) (parallel_tempering.py:177-191).  Only ``mc_args["steps"]`` is used: the
        replicas share the sampler's seed / ECIs, and equilibration or observer
        arguments of ``runMC`` do not apply to a leg."""
        unknown = set(mc_args) - {"steps", "equil", "mode"}
        if unknown:
            raise ValueError("mc_args not supported by the GPU parallel tempering: "
                             + ", ".join(sorted(unknown)))
        steps = int(mc_args.get("steps", 10 * self.natoms))
        torch = self._torch
        ev = []
        with torch.cuda.stream(self._stream):
            for _ in range(num_exchange_cycles):
                if timing:
                    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
                    e0.record(self._stream)
                self.gpu.run_canonical(steps)
                if timing:
                    e1.record(self._stream)
                # the direction of cycle k is a pure function of (seed, k): identical on every
                # rank, and a sequence of run(num_exchange_cycles=1) calls walks the same ladder
                self._enqueue_exchange(-1)
                if timing:
                    e2.record(self._stream)
                    ev.append((e0, e1, e2))
        self._read_back()
        if timing:
            leg = sum(a.elapsed_time(b) for a, b, _ in ev)
            exch = sum(b.elapsed_time(c) for _, b, c in ev)
            self.last_timing = dict(cycles=num_exchange_cycles, steps_per_leg=steps,
                                    leg_ms=leg, exchange_ms=exch)

    # ---- results -------------------------------------------------------------------
    def temperature_of_replica(self):
        return np.array(self._temps)[self.slot_of_replica]

    def gather_energies(self):
        """Energies of all replicas in global replica order."""
        e = parallel.all_gather_array(self.gpu.get_energy(), self.world,
                                      None if self.world == 1 else "cuda:%d" % self.device)
        return e[parallel.gather_index(self.n_total, self.world, self.stride)]
