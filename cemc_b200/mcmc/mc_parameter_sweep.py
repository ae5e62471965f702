"""Parameter sweep -- mirror of ``cemc.mcmc.MCParameterSweep``
(/root/reference/cemc/mcmc/mc_parameter_sweep.py:6-103).

The reference walks the (temperature, chemical potential) grid sequentially;
here every grid point is one replica and the whole grid advances in one kernel
launch.  The result dictionaries carry the keys of
``SGCMonteCarlo.get_thermodynamic`` (sgc_montecarlo.py:398-448); ``save``
writes ``.npz`` (h5py is not a dependency).
"""
from __future__ import annotations

import numpy as np

from ..updater import BatchedCEUpdater
from .montecarlo import KB


class MCParameterSweep(object):
    known_parameters = {
        "MonteCarlo": ["temperature", "composition"],
        "SGCMonteCarlo": ["temperature", "chemical_potential"],
    }

    def __init__(self, parameters, mc_obj, nsteps=100000, data_getter=None,
                 outfile="default_output.npz", equil_params=None, equil_steps=None):
        self.parameters = parameters
        self.mc_obj = mc_obj
        self.nsteps = nsteps
        self.data_getter = None
        self.check_initialization()
        self.outfile = outfile
        self.equil_params = equil_params
        self.equil_steps = equil_steps
        self.results = []

    def check_initialization(self):
        if self.mc_obj.name not in self.known_parameters.keys():
            raise ValueError("Monte Carlo instance was not recognized. Known MonteCarlo "
                             "object: {}".format(self.known_parameters.keys()))
        required_params = self.known_parameters[self.mc_obj.name]
        for i in range(len(self.parameters)):
            for req_param in required_params:
                if req_param not in self.parameters[i].keys():
                    raise ValueError("Required parameter {} not given for entry {}."
                                     "".format(req_param, i))
            if "chemical_potential" in self.parameters[i].keys():
                if not isinstance(self.parameters[i]["chemical_potential"], dict):
                    raise ValueError("Chemical potential has to be given as a dictionary")
            if "temperature" in self.parameters[i].keys():
                try:
                    float(self.parameters[i]["temperature"])
                except Exception:
                    raise ValueError("Temperature has to be given as a float")

    def run(self):
        mc = self.mc_obj
        if mc.name != "SGCMonteCarlo":
            raise NotImplementedError("Parameter sweep for the MC object not supported yet!")
        calc = mc.atoms.get_calculator()
        ft = calc.updater.tables
        R = len(self.parameters)
        mc.reset_ecis()
        base = ft.eci_vector(calc.eci)
        eci = np.tile(base, (R, 1))
        kT = np.zeros(R)
        names = None
        for r, p in enumerate(self.parameters):
            kT[r] = float(p["temperature"]) * KB
            keys = sorted(p["chemical_potential"].keys())
            names = keys
            for k in keys:
                if k not in ft.eci_index:
                    raise ValueError("chemical potential for an untracked singlet " + k)
                eci[r, ft.eci_index[k]] -= p["chemical_potential"][k]
        gpu = BatchedCEUpdater(ft, R, device=calc.device)
        gpu.set_occupancy(np.repeat(calc.updater.batch.get_occupancy(), R, axis=0))
        gpu.set_cf(np.repeat(calc.updater.batch.get_cf(), R, axis=0))
        gpu.set_ecis(eci)
        gpu.set_kT(kT)
        gpu.seed(mc.seed)
        gpu.set_sgc_species([ft.species_id[s] for s in mc.symbols])
        equil = self.equil_steps if self.equil_steps is not None else 10 * len(mc.atoms)
        gpu.run_sgc(equil)                      # fixed-length equilibration window
        gpu.run_sgc(1000)                       # energy-bias probe (montecarlo.py:178)
        gpu.synchronize()
        bias = gpu.get_energy()
        eci[:, ft.eci_index["c0"]] -= bias / ft.N
        gpu.set_ecis(eci)
        ref = np.where(np.abs(bias) > 0, bias, 1.0)
        gpu.reset_accumulators(ref)
        gpu.run_sgc(self.nsteps)
        gpu.synchronize()
        acc = gpu.get_accumulators()
        natoms = ft.N
        self.results = []
        sidx = {n: d for d, n in enumerate(ft.singlet_names)}
        for r, p in enumerate(self.parameters):
            n = acc[r, 0]
            T = float(p["temperature"])
            e_mean = acc[r, 1] / n * ref[r]
            e2_mean = acc[r, 2] / n * ref[r]
            q = {"sgc_energy": e_mean + bias[r],
                 "sgc_heat_capacity": (e2_mean - e_mean ** 2) / (KB * T ** 2),
                 "energy": e_mean + bias[r], "temperature": T, "n_mc_steps": int(n)}
            for k in names:
                d = sidx[k]
                s = acc[r, 3 + 3 * d] / n
                q["energy"] += p["chemical_potential"][k] * s * natoms
                q["singlet_{}".format(k)] = s
                q["var_singlet_{}".format(k)] = acc[r, 4 + 3 * d] / n - s ** 2
                q["mu_{}".format(k)] = p["chemical_potential"][k]
            self.results.append(q)
        gpu.close()
        if self.outfile != "":
            self.save(self.results)
        return self.results

    def save(self, data):
        """Append the result arrays (one entry per grid point and key) to ``outfile``.

        ``*.h5`` / ``*.hdf5`` and an importable ``h5py``: the reference's HDF5 layout -- one
        resizable 1-d dataset per key, appended to on every call (mc_parameter_sweep.py:80-103),
        so cemc/tools post-processing reads it.  Otherwise (h5py is not part of this image) the
        same arrays go to a NumPy ``.npz`` archive, appended the same way."""
        skip = ("timestamp", "python_version")
        columns = {}
        for row in data:
            for key, value in row.items():
                if key not in skip:
                    columns.setdefault(key, []).append(value)
        columns = {k: np.asarray(v) for k, v in columns.items()}
        h5 = None
        if str(self.outfile).lower().endswith((".h5", ".hdf5")):
            try:
                import h5py as h5
            except ImportError:
                h5 = None
        if h5 is not None:
            with h5.File(self.outfile, "a") as hf:
                for key, value in columns.items():
                    if key in hf:
                        ds = hf[key]
                        ds.resize((ds.shape[0] + len(value),))
                        ds[-len(value):] = value
                    else:
                        hf.create_dataset(key, data=value, maxshape=(None,))
            return
        previous = {}
        try:
            with np.load(self.outfile) as z:
                previous = {k: z[k] for k in z.files}
        except (IOError, OSError):
            pass
        np.savez(self.outfile, **{k: np.concatenate([previous[k], v]) if k in previous else v
                                  for k, v in columns.items()})
