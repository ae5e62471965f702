"""Parallel tempering -- mirror of ``cemc.mcmc.ParallelTempering``
(/root/reference/cemc/mcmc/parallel_tempering.py:6-191), with every replica
advanced by ONE kernel launch per cycle and the exchange sweep on the device.

Differences from the reference (DESIGN.md "Parallel tempering"):

* replicas run concurrently (one CTA / warp each) instead of sequentially;
* an accepted exchange permutes the temperature<->replica map; the reference
  copies whole configurations site by site (:146-151, up to 2N update_cf);
* the exchange uses the true energies.  The reference compares
  ``current_energy`` values from which each replica's own energy-bias probe
  was subtracted (montecarlo.py:189, parallel_tempering.py:161-162), which
  skews its criterion by the difference of the two biases;
* the temperature ladder can be given explicitly (``temperatures=``); the
  reference's bisection search (:47-136) is kept as ``_init_temperature_scheme``;
* replicas can be sharded over ranks (torch.distributed); the only collective
  is the all-gather of the energies.
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
        self.offset, self.R = parallel.shard_range(self.n_total, self.rank, self.world)
        calc = mc_obj.atoms.get_calculator()
        self.tables = calc.updater.tables
        self.gpu = BatchedCEUpdater(self.tables, self.R, device=self.device,
                                    replica_offset=self.offset)
        occ = calc.updater.batch.get_occupancy()
        cf = calc.updater.batch.get_cf()
        self.gpu.set_occupancy(np.repeat(occ, self.R, axis=0))
        self.gpu.set_cf(np.repeat(cf, self.R, axis=0))
        self.gpu.set_ecis(self.tables.eci_vector(calc.eci))
        self.gpu.seed(self.seed)
        self.kT_of_slot = np.array(self._temps) * KB        # slot 0 = Tmax
        self.slot_of_replica = np.arange(self.n_total, dtype=np.int32)
        self.gpu.set_kT(self.kT_of_slot[self.offset:self.offset + self.R])
        self.round = 0
        self.num_accepted_exchanges = 0
        self._dev = None

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
        if self._dev is None:
            import torch
            dev = torch.device("cuda", self.device)
            self._dev = dict(
                torch=torch, dev=dev,
                slots=torch.from_numpy(self.slot_of_replica.copy()).to(dev),
                kts=torch.from_numpy(self.kT_of_slot.copy()).to(dev),
                e_all=torch.empty(self.n_total, dtype=torch.float64, device=dev),
                n_acc=torch.zeros(1, dtype=torch.int32, device=dev))
        return self._dev

    def _perform_exchange_move(self, direction="up"):
        """One exchange sweep (parallel_tempering.py:153-175), on the device."""
        d = self._device_buffers()
        torch = d["torch"]
        self.gpu.synchronize()
        if self.world > 1:
            import torch.distributed as dist
            # the local energies never leave the device: wrap the updater's buffer
            e_loc = torch.as_tensor(parallel.DeviceArrayF64(self.gpu.energy_dev_ptr(), self.gpu.R),
                                    device=d["dev"])
            dist.all_gather_into_tensor(d["e_all"], e_loc)      # NCCL over NVLink
            e_ptr = d["e_all"].data_ptr()
        else:
            e_ptr = self.gpu.energy_dev_ptr()
        torch.cuda.synchronize(d["dev"])
        self.gpu.pt_exchange(self.n_total, e_ptr, d["slots"].data_ptr(), d["kts"].data_ptr(),
                             0 if direction == "up" else 1, self.round, d["n_acc"].data_ptr())
        self.gpu.synchronize()
        self.slot_of_replica = d["slots"].cpu().numpy()
        n_acc = int(d["n_acc"].item())
        self.num_accepted_exchanges += n_acc
        self.round += 1
        return n_acc

    def run(self, mc_args={}, num_exchange_cycles=10):
        """``num_exchange_cycles`` x (``steps`` moves on every replica, then
        one exchange sweep) (parallel_tempering.py:177-191)."""
        steps = int(mc_args.get("steps", 10 * self.natoms))
        rng = np.random.RandomState(self.seed & 0x7fffffff)   # same direction on every rank
        for _ in range(num_exchange_cycles):
            self.gpu.run_canonical(steps) if self.mc.name == "MonteCarlo" \
                else self.gpu.run_sgc(steps)
            direction = "up" if rng.randint(0, 2) == 0 else "down"
            self._perform_exchange_move(direction=direction)

    # ---- results -------------------------------------------------------------------
    def temperature_of_replica(self):
        return np.array(self._temps)[self.slot_of_replica]

    def gather_energies(self):
        return parallel.all_gather_array(self.gpu.get_energy(), self.world,
                                         None if self.world == 1 else "cuda:%d" % self.device)
