"""Semi-grand-canonical Monte Carlo -- host-side mirror of the reference's
``cemc.mcmc.SGCMonteCarlo`` (/root/reference/cemc/mcmc/sgc_montecarlo.py:12-471).

Trial move (:62-76) = flip one site to another species; the chemical
potentials are folded into the singlet ECIs (:239-261); the SGCObserver sums
(:46, mc_observers.py:222-270) are accumulated inside the CUDA kernel every
step.
"""
import numpy as np

from . import montecarlo as mc
from . import stats
from .mc_observers import SGCObserver
from .montecarlo import KB


class InvalidChemicalPotentialError(Exception):
    pass


class SGCMonteCarlo(mc.Montecarlo):
    def __init__(self, atoms, temp, indeces=None, symbols=None, logfile="",
                 plot_debug=False, min_acc_rate=0.0, recycle_waste=False, seed=None):
        mc.Montecarlo.__init__(self, atoms, temp, indeces=indeces, logfile=logfile,
                               plot_debug=plot_debug, min_acc_rate=min_acc_rate,
                               recycle_waste=recycle_waste, seed=seed)
        if symbols is not None:
            self.symbols = list(symbols)          # :38-40
        if len(self.symbols) <= 1:
            raise ValueError("At least 2 symbols have to be specified")
        self.sgc_symbols = list(self.symbols)
        sid = self._tables.species_id
        for s in self.symbols:
            if s not in sid:
                raise ValueError("Unknown symbol {}".format(s))
        self._gpu.set_sgc_species([sid[s] for s in self.symbols])
        self.averager = SGCObserver(self.atoms.get_calculator(), self,
                                    len(self._tables.singlet_indices))
        self.averager.device_backed = True
        self.chem_pots = []
        self.chem_pot_names = []
        self.has_attached_avg = False
        self.name = "SGCMonteCarlo"
        self._chemical_potential = None
        self.chem_pot_in_ecis = False
        self.composition_correlation_time = np.zeros(len(self.symbols) - 1)
        self.current_singlets = None
        self.attach(self.averager)

    def _check_symbols(self):
        pass                                          # :78-82

    def _device_run(self, n):
        self._gpu.run_sgc(n)

    def _pull_averages(self):
        acc = mc.Montecarlo._pull_averages(self)
        self.averager.load_device_sums(acc)
        return acc

    def _on_window_done(self):
        pass

    def reset(self):
        super(SGCMonteCarlo, self).reset()
        self.averager.reset()
        # SGCObserver's Averagers have their own reference value
        # (mc_observers.py:207-208); the device keeps one per replica
        self._gpu.reset_accumulators([self.averager.energy.ref_value])

    def _pull_for_stats(self):
        self._pull_averages()

    # ---- variances / equilibrium of the composition (:86-221) --------------------
    def _mean_singlets(self):
        n = self.averager.counter
        return self.averager.singlets / n, self.averager.quantities["singlets_sq"] / n, n

    def _get_var_average_singlets(self):
        mean, mean_sq, n = self._mean_singlets()
        return stats.variance_of_mean(mean, mean_sq, n, self._known_correlation_time(),
                                      fold_negative=False)

    def _composition_reached_equillibrium(self, prev_composition, var_prev,
                                          confidence_level=0.05):
        """Window test on the average singlets (sgc_montecarlo.py:172-210): returns
        (converged, singlets, variances, largest z)."""
        singlets = self._mean_singlets()[0]
        var_n = np.maximum(self._get_var_average_singlets(), 0.0)
        if len(prev_composition) != len(singlets):
            return False, singlets, var_n, 0.0               # first window: nothing to compare with
        agree, z = stats.singlets_agree(singlets, var_n, prev_composition, var_prev, confidence_level)
        return agree, singlets, var_n, z

    def _has_converged_prec_mode(self, prec=0.01, confidence_level=0.05,
                                 log_status=False):
        limit = (prec / stats.normal_quantile(1.0 - confidence_level)) ** 2
        return bool(np.max(self._get_var_average_singlets()) < limit)

    # ---- chemical potential (:219-278) ----------------------------------------------
    @property
    def chemical_potential(self):
        return self._chemical_potential

    @chemical_potential.setter
    def chemical_potential(self, chem_pot):
        calc = self.atoms.get_calculator()
        untracked = [name for name in chem_pot if name not in calc.eci]
        if untracked:
            raise InvalidChemicalPotentialError(
                "A chemical potential that is currently not tracked is added. "
                "Make sure that all the following keys are in the ECI before "
                "the ECI are passed to the calculator: {} (if not add them "
                "with a zero value)".format(list(chem_pot.keys())))
        self._chemical_potential = chem_pot
        self.reset_ecis()                       # take a previous potential out first
        self._include_chemical_potential_in_ecis(chem_pot, calc.eci)

    def _shift_singlet_ecis(self, eci, sign):
        """eci[name] += sign * mu[name] for the potentials on record, then push the ECIs to the
        device and refresh the energy (the SGC energy is E - sum mu_i n_i, sgc_montecarlo.py:239-261)."""
        for name, mu in zip(self.chem_pot_names, self.chem_pots):
            eci[name] = eci.get(name, 0.0) + sign * mu
        calc = self.atoms.get_calculator()
        calc.update_ecis(eci)
        self.current_energy = calc.get_energy()
        return eci

    def _include_chemical_potential_in_ecis(self, chem_potential, eci):
        self.chem_pot_names = sorted(chem_potential)
        self.chem_pots = [chem_potential[name] for name in self.chem_pot_names]
        self.chem_pot_in_ecis = True
        return self._shift_singlet_ecis(eci, -1.0)

    def _reset_eci_to_original(self, eci_with_chem_pot):
        self.chem_pot_in_ecis = False
        return self._shift_singlet_ecis(eci_with_chem_pot, +1.0)

    def reset_ecis(self):
        if self.chem_pot_in_ecis:
            self._reset_eci_to_original(self.atoms.get_calculator().eci)

    def _equillibriate(self, *args, **kwargs):
        self._in_equil = True
        try:
            return mc.Montecarlo._equillibriate(self, *args, **kwargs)
        finally:
            self._in_equil = False

    def runMC(self, mode="fixed", steps=10, verbose=False, chem_potential=None,
              equil=True, equil_params={}, prec_confidence=0.05, prec=0.01):
        """Run SGC Monte Carlo (sgc_montecarlo.py:336-378)."""
        if chem_potential is None and self.chemical_potential is None:
            ex_chem_pot = {"c1_1": -0.1, "c1_2": 0.05}
            raise ValueError("No chemicalpotentials given. Has to be "
                             "dictionary of the form {}".format(ex_chem_pot))
        if chem_potential is not None:
            self.chemical_potential = chem_potential
        self.reset()
        self._gpu.set_kT([self.T * KB])
        if equil:
            res = self._estimate_correlation_time(restart=True)
            if not res["correlation_time_found"]:
                res["correlation_time_found"] = True
                res["correlation_time"] = 1000
            self._equillibriate(**equil_params)
        self.reset()
        mc.Montecarlo.runMC(self, steps=steps, verbose=verbose, equil=False, mode=mode,
                            prec_confidence=prec_confidence, prec=prec)
        self._pull_averages()

    def singlet2composition(self, avg_singlets):
        """Concentrations from the average singlets (sgc_montecarlo.py:380-396)."""
        bf = self.atoms.get_calculator().BC.basis_functions
        return stats.concentrations_from_singlets(bf, list(self.symbols), avg_singlets)

    def get_thermodynamic(self, reset_ecis=True):
        """Thermodynamic quantities of the run, under the reference's keys
        (sgc_montecarlo.py:398-448): the SGC energy / heat capacity come from the SGCObserver
        sums, the internal energy adds back mu * <singlet> * N for every chemical potential."""
        mean, mean_sq, n = self._mean_singlets()
        e_avg, e_sq = self.averager.energy.mean, self.averager.energy_sq.mean
        natoms = len(self.atoms)
        q = {"sgc_energy": e_avg + self.energy_bias,
             "sgc_heat_capacity": (e_sq - e_avg ** 2) / (KB * self.T ** 2),
             "energy": e_avg + self.energy_bias,
             "temperature": self.T,
             "n_mc_steps": n}
        for name, mu, m, m2 in zip(self.chem_pot_names, self.chem_pots, mean, mean_sq):
            q["energy"] += mu * m * natoms
            q["singlet_" + name] = m
            q["var_singlet_" + name] = m2 - m ** 2
            q["mu_" + name] = mu
        q.update(self.meta_info)
        try:
            q.update(self.singlet2composition(mean))
        except Exception as exc:           # the reference reports and carries on (:440-444)
            print("Could not find average singlets!")
            print(exc)
        if reset_ecis:
            self._reset_eci_to_original(self.atoms.get_calculator().eci)
        return q
