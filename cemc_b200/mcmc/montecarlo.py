"""Canonical (fixed composition) Monte Carlo -- host-side mirror of the
reference's ``cemc.mcmc.Montecarlo``
(/root/reference/cemc/mcmc/montecarlo.py:39-1074).

The Python-visible API is the reference's (constructor, ``runMC``, ``attach``,
``get_thermodynamic``, ``reset``, ``set_symbols`` ...).  The per-move loop
(``_mc_step`` :970, ``_get_trial_move`` :890, ``_accept`` :910) runs inside the
CUDA kernels (``cemc_run_canonical``); Python only sees chunk boundaries.

Deviations (DESIGN.md "Drivers"):
* the trial-move stream is the project's Philox stream, not NumPy's MT19937 +
  Python ``random`` (which are hash-order dependent in the reference);
* observers are called every ``interval`` steps with the NET changes of the
  chunk; constraints / bias potentials (arbitrary Python per trial move) and
  waste recycling are not supported on the GPU path and raise;
* ``atoms`` is synchronised with the device at chunk boundaries only.
"""
from __future__ import division

import datetime
import logging
import sys
import time

import numpy as np

from . import stats
from .averager import Averager
from .mc_observers import MCObserver  # noqa: F401

KB = 8.617330337217213e-05     # eV/K == ase.units.kB (CODATA 2014)


class DidNotReachEquillibriumError(Exception):
    pass


class TooFewElementsError(Exception):
    pass


class CanNotFindLegalMoveError(Exception):
    pass


def _norm_ppf(p):
    return stats.normal_quantile(p)


class Montecarlo(object):
    """Monte Carlo at fixed composition (montecarlo.py:39)."""

    def __init__(self, atoms, temp, indeces=None, logfile="",
                 plot_debug=False, min_acc_rate=0.0, recycle_waste=False,
                 max_constraint_attempts=10000,
                 accept_first_trial_move_after_reset=False, seed=None):
        self.name = "MonteCarlo"
        self.atoms = atoms
        self.T = temp
        self.min_acc_rate = min_acc_rate
        if recycle_waste:
            raise NotImplementedError("waste recycling (montecarlo.py:937-948) "
                                      "is outside the GPU hot path")
        self.recycle_waste = False
        self.indeces = range(len(self.atoms)) if indeces is None else indeces
        self.observers = []
        self.constraints = []
        self.max_allowed_constraint_pass_attempts = max_constraint_attempts
        if self.max_allowed_constraint_pass_attempts <= 0:
            raise ValueError("Max. constraint attempts has to be > 0!")
        self.bias_potentials = []
        self.current_step = 0
        self.num_accepted = 0
        self.status_every_sec = 30
        self.symbols = []
        self._build_atoms_list()
        calc = self.atoms.get_calculator()
        self._calc = calc
        self._gpu = calc.updater.batch            # BatchedCEUpdater, R = 1
        self._tables = calc.updater.tables
        if seed is None:
            seed = int(np.random.randint(0, 2 ** 31 - 1))
        self.seed = int(seed)
        self._gpu.seed(self.seed)
        E0 = calc.get_energy()
        self.current_energy = E0
        self.bias_energy = 0.0
        self.new_bias_energy = self.bias_energy
        self.new_energy = self.current_energy
        self.last_energies = np.zeros(2)
        self.trial_move = []
        self.mean_energy = Averager(ref_value=E0)
        self.energy_squared = Averager(ref_value=E0)
        self.energy_bias = 0.0
        self.update_energy_bias = True
        self.logfile = logfile
        self.logger = None
        self._init_loggers()
        self.corrtime_energies = []
        self.correlation_info = None
        self.plot_debug = plot_debug
        self._linear_vib_correction = None
        if accept_first_trial_move_after_reset:
            raise NotImplementedError("accept_first_trial_move_after_reset "
                                      "is not supported on the GPU path")
        self.accept_first_trial_move_after_reset = False
        self.is_first = False
        self.chunk_size = 100000       # moves per launch when no observer asks for less

    # ---- plumbing ------------------------------------------------------------
    def _init_loggers(self):
        self.logger = logging.getLogger("MonteCarlo")
        self.logger.setLevel(logging.DEBUG)
        if self.logfile == "":
            ch = logging.StreamHandler()
        else:
            ch = logging.FileHandler(self.logfile)
        ch.setLevel(logging.INFO)
        if not self.logger.handlers:
            self.logger.addHandler(ch)

    def log(self, msg, mode="info"):
        if mode not in ("info", "warning"):
            raise ValueError("Mode has to be one of ['info', 'warning']")
        (self.logger.info if mode == "info" else self.logger.warning)(msg)

    def _build_atoms_list(self):
        """Species present (SwapMoveIndexTracker.init_tracker,
        swap_move_index_tracker.py:22-36; the site lists live on the device)."""
        self.symbols = sorted(set(atom.symbol for atom in self.atoms))

    def _check_symbols(self):
        """At least two species with at least two atoms... (montecarlo.py:292-313)"""
        count = {}
        for atom in self.atoms:
            count[atom.symbol] = count.get(atom.symbol, 0) + 1
        if len(count.keys()) < 2:
            raise TooFewElementsError(
                "There is only one element in the given atoms object!")
        if sum(1 for v in count.values() if v >= 2) < 2:
            raise TooFewElementsError(
                "There is only one element that has more than one atom")

    def add_constraint(self, constraint):
        raise NotImplementedError(
            "Python constraints run once per trial move "
            "(montecarlo.py:979-982) and cannot be evaluated inside the GPU "
            "loop; see DESIGN.md 'out of scope'")

    def add_bias(self, potential):
        raise NotImplementedError(
            "Python bias potentials run once per trial move "
            "(montecarlo.py:928-930) and cannot be evaluated inside the GPU loop")

    def attach(self, obs, interval=1):
        """Observer called every ``interval`` MC steps (montecarlo.py:403-415)."""
        if callable(obs):
            self.observers.append((interval, obs))
        else:
            raise ValueError("The observer has to be a callable class!")

    def current_energy_without_vib(self):
        return self.current_energy

    def update_current_energy(self):
        self.current_energy = self._calc.get_energy()
        self.bias_energy = 0.0

    def set_symbols(self, symbs):
        self._calc.set_symbols(symbs)
        self._build_atoms_list()
        self.update_current_energy()

    def count_atoms(self):
        atom_count = {key: 0 for key in self.symbols}
        for atom in self.atoms:
            atom_count[atom.symbol] = atom_count.get(atom.symbol, 0) + 1
        return atom_count

    def reset(self):
        """Reset counters and averages (montecarlo.py:335-351)."""
        for interval, obs in self.observers:
            obs.reset()
        self._reset_device_observers()
        self.current_step = 0
        self.num_accepted = 0
        self.mean_energy.clear()
        self.energy_squared.clear()
        self.corrtime_energies = []
        self._gpu.reset_accumulators([self.mean_energy.ref_value])
        self._gpu.reset_counters()

    # ---- device stepping ---------------------------------------------------------
    def _device_run(self, n):
        self._gpu.run_canonical(n)

    def _sync_atoms(self, incremental=False):
        """Mirror device occupations into ``atoms``; returns the net changes since the
        previous mirror.  Only the sites that differ are touched (vectorised compare);
        ``incremental``: trust the mirror kept by the running ``_steps`` call instead of
        re-reading every ``atoms[i].symbol``."""
        occ = self._gpu.get_occupancy()[0]
        species = self._tables.species
        mirror = getattr(self, "_occ_mirror", None) if incremental else None
        if mirror is None or len(mirror) != len(occ):
            mirror = self._tables.occupancy([a.symbol for a in self.atoms])
        changes = []
        for i in np.nonzero(occ != mirror)[0]:
            i = int(i)
            changes.append((i, species[int(mirror[i])], species[int(occ[i])]))
            self.atoms[i].symbol = species[int(occ[i])]
        self._occ_mirror = occ.copy()
        return changes

    def _changes_for(self, obs):
        """Net changes since THIS observer's previous call (its own snapshot: with several
        observers on non-commensurate intervals each one still sees every change once)."""
        occ = self._occ_mirror
        snaps = self.__dict__.setdefault("_obs_snapshots", {})
        old = snaps.get(id(obs))
        if old is None:
            old = self.__dict__.get("_occ_at_start", occ)
        species = self._tables.species
        changes = [(int(i), species[int(old[i])], species[int(occ[i])])
                   for i in np.nonzero(occ != old)[0]]
        snaps[id(obs)] = occ.copy()
        return changes

    def _pull_counters(self):
        _, acc = self._gpu.get_counters()
        self.num_accepted = int(acc[0])

    def _pull_averages(self):
        acc = self._gpu.get_accumulators()[0]
        self.mean_energy.set_sums(acc[1], acc[0])
        self.energy_squared.set_sums(acc[2], acc[0])
        return acc

    # ---- observers folded on the device (cemc_set_device_observers) -------------------
    OBS_RING = 4096                 # energy samples kept on the device between two read-backs

    def _device_observer_plan(self):
        """(interval, flags) when EVERY attached observer is one of the fixed-semantics state
        observers and they share one interval: the kernels then fold them every `interval` steps
        and the host never stops the device loop.  None: the observers are called on the host."""
        obs = self._state_observers()
        if not obs:
            return None
        ivs = {iv for iv, _ in obs}
        if len(ivs) != 1 or any(getattr(o, "device_flag", 0) == 0 for _, o in obs):
            return None
        flags = 0
        for _, o in obs:
            flags |= o.device_flag
        return int(next(iter(ivs))), flags

    def _state_observers(self):
        """Attached observers except the per-step accumulators the kernels always keep
        (SGCObserver: ``device_backed``)."""
        return [(iv, o) for iv, o in self.observers if self._is_host_observer(o)]

    def _arm_device_observers(self):
        plan = self._device_observer_plan()
        if plan == getattr(self, "_dev_obs_plan", None):
            return plan
        if plan is None:
            self._gpu.set_device_observers(0, 0)
        else:
            self._gpu.set_device_observers(plan[0], plan[1], self.OBS_RING)
        self._dev_obs_plan = plan
        self._dev_obs_pulled = 0
        self._reset_device_observers()
        return plan

    def _reset_device_observers(self):
        if getattr(self, "_dev_obs_plan", None) is None:
            return
        ref = None
        for _, o in self._state_observers():
            if getattr(o, "orig_symbols", None) is not None:
                ref = self._tables.occupancy(list(o.orig_symbols))[None]
        self._gpu.reset_device_observers(ref)
        self._dev_obs_pulled = 0

    def _pull_device_observers(self):
        """Mirror the device-side observer block into the attached observer objects."""
        blk = self._gpu.get_device_observers()
        n = int(blk["n_samples"][0])
        cap = blk["energies"].shape[1]
        if n - self._dev_obs_pulled > cap:
            raise RuntimeError("device observer ring overflow")
        new_e = [float(blk["energies"][0, k % cap]) for k in range(self._dev_obs_pulled, n)]
        self._dev_obs_pulled = n
        for _, o in self._state_observers():
            o.load_device(self, blk, n, new_e)

    def _steps(self, n, observe=True):
        """n trial moves on the device, observers at their intervals."""
        plan = self._arm_device_observers() if observe else None
        if plan is not None:
            chunk = max(plan[0], min(self.chunk_size, plan[0] * (self.OBS_RING // 2)))
            done = 0
            while done < n:
                m = min(chunk, n - done)
                self._device_run(m)
                done += m
                self.current_step += m
                self._pull_device_observers()          # one read-back per chunk, not per interval
            self.current_energy = float(self._gpu.get_energy()[0])
            return
        if getattr(self, "_dev_obs_plan", None) is not None and not observe:
            self._gpu.set_device_observers(0, 0)       # legs nobody observes (bias probe, equilibration)
            self._dev_obs_plan = None
        done = 0
        intervals = [iv for iv, _ in self.observers if self._is_host_observer(_)]
        if observe and intervals:
            # what the host (atoms, hence every observer not called yet) has seen so far
            # (re-read from `atoms` once per call: the calculator's per-call API mutates them too)
            mirror = self._tables.occupancy([a.symbol for a in self.atoms])
            self._occ_mirror = mirror
            self._occ_at_start = mirror
        chunk = min([self.chunk_size] + intervals) if observe else self.chunk_size
        while done < n:
            m = min(chunk, n - done)
            if observe and intervals:
                # stop exactly on the next observer boundary
                nxt = min(iv - (self.current_step % iv) for iv in intervals)
                m = min(m, nxt)
            self._device_run(m)
            done += m
            self.current_step += m
            if observe and intervals:
                due = [o for iv, o in self.observers
                       if self._is_host_observer(o) and self.current_step % iv == 0]
                if due:
                    self._sync_atoms(incremental=True)  # waits for the stream
                    self.current_energy = float(self._gpu.get_energy()[0])
                    for o in due:
                        o(self._changes_for(o))
        self._gpu.synchronize()
        self.current_energy = float(self._gpu.get_energy()[0])

    def _is_host_observer(self, obs):
        return not getattr(obs, "device_backed", False)

    def _mc_step(self, verbose=False):
        """One trial move (kept for API compatibility; one launch per call)."""
        before = self.num_accepted
        self._steps(1)
        self._pull_counters()
        return self.current_energy, self.num_accepted > before

    # ---- correlation time / equilibration (montecarlo.py:461-697) --------------------
    # The statistics themselves live in ``stats.py`` and work on device-side sums: the energy
    # trace of the correlation-time window never leaves the GPU (cemc_energy_autocorrelation)
    # and a window of the equilibration loop is one launch plus one read of the Averager sums.
    def _estimate_correlation_time(self, window_length=1000, restart=False):
        self.log("*********** Estimating correlation time ***************")
        fresh = restart or not self.corrtime_energies
        if restart:
            self.corrtime_energies = []
        self._gpu.set_observe(False)
        self._gpu.set_trace(window_length)
        self._device_run(window_length)
        self.current_step += window_length
        if fresh:
            mean, var, lag, _ = self._gpu.energy_autocorrelation(window_length)[0]
            self.corrtime_energies = [mean]            # the trace itself stays on the device
        else:
            # appended windows (restart=False, montecarlo.py:466-470): the whole record is needed
            self.corrtime_energies += [float(x) for x in self._gpu.get_trace(window_length)[4][0]]
            d = np.array(self.corrtime_energies) - np.mean(self.corrtime_energies)
            var = float(np.mean(d * d))
            acf = np.correlate(d, d, mode="full")[len(d) - 1:]
            below = np.nonzero(acf < 0.5 * window_length * var)[0] if var > 0.0 else []
            lag = float(below[0]) if len(below) else -1.0
        self._gpu.set_trace(0)
        self._gpu.set_observe(True)
        self.current_energy = float(self._gpu.get_energy()[0])
        info = {"correlation_time_found": False, "correlation_time": 0.0, "msg": ""}
        if var == 0.0:
            info.update(msg="Zero variance leads to infinite correlation time",
                        correlation_time_found=True, correlation_time=window_length)
        else:
            tau, found = stats.correlation_time(lag, window_length)
            info["correlation_time"] = float(tau) if found else window_length
            info["correlation_time_found"] = bool(found)
            if not found:
                info["msg"] = "Window is too short. Add more samples"
            else:
                self.log("Estimated correlation time: {}".format(float(tau)))
        if info["msg"]:
            self.log(info["msg"])
        self.correlation_info = info
        return info

    def _known_correlation_time(self):
        info = self.correlation_info
        return info["correlation_time"] if info and info["correlation_time_found"] else None

    def _get_var_average_energy(self):
        return float(stats.variance_of_mean(self.mean_energy.mean, self.energy_squared.mean,
                                            self.current_step, self._known_correlation_time()))

    def _composition_reached_equillibrium(self, prev_composition, var_prev,
                                          confidence_level=0.05):
        return True, prev_composition, var_prev, 0.0       # fixed composition (:528-539)

    def _equillibriate(self, window_length="auto", confidence_level=0.05,
                       maxiter=1000, mode="stat_equiv"):
        """Run windows of ``window_length`` moves (one launch each) until the mean energies of
        two consecutive windows agree at ``confidence_level`` -- and, for the SGC sampler, the
        average singlets do too (montecarlo.py:541-697).  ``equil_history`` records
        (mean, variance of the mean, z) of every window for inspection / tests."""
        if mode not in ("stat_equiv", "fixed"):
            raise ValueError("Equilibration mode has to be one of ['stat_equiv', 'fixed']")
        if window_length == "auto":
            window_length = 10 * len(self.atoms)
        self.reset()
        self.equil_history = []
        if mode == "fixed":
            self._steps(window_length, observe=False)
            return
        previous = None
        composition, var_comp = [], []
        energy_ok = False
        for window in range(maxiter):
            self.reset()
            self._steps(window_length, observe=False)
            self._pull_averages()
            self._on_window_done()
            now = (self.mean_energy.mean, self._get_var_average_energy())
            comp_ok, composition, var_comp, comp_z = self._composition_reached_equillibrium(
                composition, var_comp, confidence_level=confidence_level)
            if previous is None:
                previous = now
                self.equil_history.append((now[0], now[1], None))
                continue
            z, frozen = stats.z_score(now[0], now[1], previous[0], previous[1])
            self.equil_history.append((now[0], now[1], float(z)))
            if frozen:
                self.log("Zero variance. System does not move.")
                comp_ok = True
            energy_ok = energy_ok or bool(stats.inside(z, confidence_level))
            if energy_ok and comp_ok:
                self.log("System reached equillibrium in {} mc steps".format(
                    (window + 1) * window_length))
                self.mean_energy.clear()
                self.energy_squared.clear()
                self.current_step = 0
                return
            previous = now
        raise DidNotReachEquillibriumError("Did not manage to reach equillibrium!")

    def _on_window_done(self):
        pass

    # ---- energy bias (montecarlo.py:178-217) ------------------------------------------
    def _probe_energy_bias(self, num_steps=1000):
        self._steps(num_steps, observe=False)
        self.energy_bias = self.current_energy
        self._remove_bias_from_empty_eci(self.energy_bias)

    def _remove_bias_from_empty_eci(self, bias):
        eci = self._calc.eci
        eci['c0'] = eci['c0'] - bias / len(self.atoms)
        self._calc.update_ecis(eci)
        self.current_energy = self._calc.get_energy()
        self.last_energies[0] = self.current_energy
        if abs(self.current_energy) > 1E-6:
            raise RuntimeError("Energy is not 0 after subtracting "
                               "the bias. Got {}".format(self.current_energy))

    def _undo_energy_bias_from_eci(self):
        eci = self._calc.eci
        eci['c0'] += self.energy_bias / len(self.atoms)
        self._calc.update_ecis(eci)

    def _has_converged_prec_mode(self, prec=0.01, confidence_level=0.05,
                                 log_status=False):
        """std of the mean energy below prec / z_{1-confidence} (montecarlo.py:699-730)."""
        return self._get_var_average_energy() < (prec / stats.normal_quantile(1.0 - confidence_level)) ** 2

    # ---- the run ---------------------------------------------------------------------
    def runMC(self, mode="fixed", steps=10, verbose=False, equil=True,
              equil_params={}, prec=0.01, prec_confidence=0.05):
        """Run Monte Carlo (montecarlo.py:732-848): warm-up move, optional
        equilibration, energy-bias probe, then ``steps`` sampled moves."""
        self._check_symbols()
        self.update_current_energy()
        if mode not in ("fixed", "prec"):
            raise ValueError("Mode has to be one of ['fixed', 'prec']")
        self._gpu.set_kT([self.T * KB])
        self._steps(1, observe=False)                  # :765
        totalenergies = [self.current_energy]
        self.current_step = 0
        if equil:
            res = self._estimate_correlation_time(restart=True)
            if not res["correlation_time_found"]:
                res["correlation_time"] = 1000
                res["correlation_time_found"] = True
            self._equillibriate(**equil_params)
        check_convergence_every = steps
        if mode == "prec":
            res = self._estimate_correlation_time(restart=True)
            while not res["correlation_time_found"]:
                res = self._estimate_correlation_time()
            self.reset()
            check_convergence_every = int(10 * self.correlation_info["correlation_time"]) + 1
        self.reset()
        self._probe_energy_bias()                      # :799
        self.reset()
        while self.current_step < steps:               # :802, in launches
            n = min(check_convergence_every, steps - self.current_step)
            self._steps(n)
            self._pull_averages()
            self._pull_counters()
            if mode == "prec" and self._has_converged_prec_mode(
                    prec=prec, confidence_level=prec_confidence):
                break
        self._pull_averages()
        self._pull_counters()
        self._sync_atoms()
        self._undo_energy_bias_from_eci()
        return totalenergies

    @property
    def meta_info(self):
        ts = time.time()
        st = datetime.datetime.fromtimestamp(ts).strftime('%Y-%m-%d %H:%M:%S')
        v = sys.version_info
        return {"timestamp": st,
                "python_version": "{}.{}.{}".format(v.major, v.minor, v.micro)}

    def get_thermodynamic(self):
        """Thermodynamic quantities (montecarlo.py:861-888)."""
        quantities = {}
        mean_energy = self.mean_energy.mean
        quantities["energy"] = mean_energy + self.energy_bias
        mean_sq = self.energy_squared.mean
        quantities["heat_capacity"] = (mean_sq - mean_energy ** 2) / (KB * self.T ** 2)
        quantities["energy_std"] = np.sqrt(self._get_var_average_energy())
        quantities["temperature"] = self.T
        for key, value in self.count_atoms().items():
            quantities["{}_conc".format(key)] = float(value) / len(self.atoms)
        quantities.update(self.meta_info)
        for obs in self.observers:
            quantities.update(obs[1].get_averages())
        return quantities

    # ---- checkpoint (montecarlo.py:1040-1074; JSON instead of dill) ---------------------
    def save(self, fname):
        import json
        self._sync_atoms()
        steps, _ = self._gpu.get_counters()
        data = {"calc": self._calc.backup_dict(), "T": self.T, "seed": self.seed,
                "philox_step": int(steps[0]), "name": self.name,
                "symbols": getattr(self, "sgc_symbols", None),
                "tracker": self._gpu.get_tracker()[0][0].tolist()}
        with open(fname, "w") as out:
            json.dump(data, out)

    @classmethod
    def load(cls, fname):
        import json
        from ..ce_calculator import CE
        with open(fname) as f:
            data = json.load(f)
        calc = CE.load_from_dict(data["calc"])
        kw = {"seed": data["seed"]}
        if data.get("symbols"):
            kw["symbols"] = data["symbols"]
        mc = cls(calc.atoms, data["T"], **kw)
        mc._gpu.set_step([data["philox_step"]])
        if data.get("tracker") is not None:
            mc._gpu.set_tracker([data["tracker"]])
        return mc
