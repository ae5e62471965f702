"""Host-side observer logic (cemc_b200/mcmc/mc_observers.py) with stub sampler / calculator
objects: no GPU needed.  Semantics follow the reference's observers
(cemc/mcmc/mc_observers.py:81-183, 614-761)."""
import numpy as np
import pytest

from cemc_b200 import synthetic as syn
from cemc_b200.mcmc.mc_observers import (EnergyEvolution, EnergyHistogram, LowestEnergyStructure,
                                         PairCorrelationObserver, SiteOrderParameter)


class _Updater(object):
    def __init__(self, cf):
        self.cf = cf

    def get_cf(self):
        return dict(self.cf)


class _Calc(object):
    def __init__(self):
        self.eci = {"c0": 0.0, "c1_0": 0.1, "c2_nn_00": 0.2, "c2_2nn_00": -0.1, "c3_tri_000": 0.3}
        self.cf = {k: 0.0 for k in self.eci}
        self.updater = _Updater(self.cf)

    def get_cf(self):
        return dict(self.cf)


class _MC(object):
    def __init__(self, atoms):
        self.atoms = atoms
        self.current_energy = 0.0

    def current_energy_without_vib(self):
        return self.current_energy


def test_pair_correlation_observer_averages_only_pairs():
    calc = _Calc()
    obs = PairCorrelationObserver(calc)
    vals = [0.5, 0.25, -0.75]
    for v in vals:
        calc.cf["c2_nn_00"] = v
        calc.cf["c2_2nn_00"] = 2 * v
        calc.cf["c3_tri_000"] = 9.0
        obs([])
    avg = obs.get_averages()
    assert set(avg) == {"c2_nn_00", "c2_2nn_00"}
    assert avg["c2_nn_00"] == pytest.approx(np.mean(vals))
    assert obs.get_std()["c2_2nn_00"] == pytest.approx(np.std([2 * v for v in vals]))
    obs.reset()
    assert obs.n_entries == 0
    calc.updater = None
    with pytest.raises(RuntimeError):
        PairCorrelationObserver(calc)


def test_lowest_energy_structure_keeps_strict_minimum():
    atoms = syn.Atoms(["Al", "Mg", "Al", "Mg"])
    calc, mc = _Calc(), _MC(atoms)
    obs = LowestEnergyStructure(calc, mc)
    for e, sym in ((1.0, "Al"), (0.5, "Mg"), (0.5, "Al"), (0.7, "Mg")):
        mc.current_energy = e
        atoms[0].symbol = sym
        calc.cf["c1_0"] = e
        obs([])
    assert obs.lowest_energy == 0.5
    assert obs.atoms.get_chemical_symbols()[0] == "Mg"        # the first visit of 0.5, not the tie
    assert obs.lowest_energy_cf["c1_0"] == 0.5 and obs.lowest_energy_atoms is obs.atoms
    atoms[0].symbol = "Si"
    assert obs.atoms.get_chemical_symbols()[0] == "Mg"        # a copy, not a view


def test_site_order_parameter_counts_net_changes():
    atoms = syn.Atoms(["Al"] * 6)
    obs = SiteOrderParameter(atoms)
    atoms[1].symbol = "Mg"; atoms[4].symbol = "Mg"
    obs([(1, "Al", "Mg"), (4, "Al", "Mg")])                    # 2 sites differ
    atoms[1].symbol = "Al"; atoms[2].symbol = "Mg"
    obs([(1, "Mg", "Al"), (2, "Al", "Mg")])                    # still 2
    atoms[2].symbol = "Si"
    obs([(2, "Mg", "Si")])                                     # changed -> changed: still 2
    avg = obs.get_averages()
    assert avg["site_order_average"] == 2.0 and avg["site_order_std"] == 0.0
    obs.reset()                                                # recounts from the atoms, keeps the origin
    assert obs.current_num_changed == 2 and obs.num_calls == 0


def test_energy_evolution_and_histogram():
    mc = _MC(syn.Atoms(["Al"]))
    evo, hist = EnergyEvolution(mc), EnergyHistogram(mc, buffer_size=4, n_bins=3)
    seq = [0.0, 1.0, 2.0, 3.0, 3.0, -5.0, 9.0, 1.4]
    for e in seq:
        mc.current_energy = e
        evo([]); hist([])
    assert evo.energies == seq
    assert (hist.Emin, hist.Emax) == (0.0, 3.0) and not hist.sample_in_buffer
    # bins over [0, 3] with index int(E * 2 / 3); out-of-range samples clamp to the edge bins
    assert hist.histogram.tolist() == [4.0, 1.0, 3.0]
    h2 = EnergyHistogram(mc, buffer_size=100, n_bins=4)
    for e in (1.0, 1.0):
        mc.current_energy = e
        h2([])
    assert h2.histogram.sum() == 2 and h2.Emin == h2.Emax == 1.0   # range fixed on first access
