"""The reference-facing Python API (CE, Montecarlo, SGCMonteCarlo,
MCParameterSweep, ParallelTempering) on the GPU against the oracle driven by
the same rules the reference's samplers apply around the updater."""
import numpy as np
import pytest

from cases import BINARY, TERNARY, build
from cemc_b200 import synthetic as syn
from cemc_b200.ce_calculator import CE, get_atoms_with_ce_calc
from cemc_b200.mcmc import (MCParameterSweep, Montecarlo, ParallelTempering,
                            SGCMonteCarlo, TooFewElementsError)
from cemc_b200.mcmc.montecarlo import KB
from cemc_b200.mcmc.mc_observers import MCObserver
from oracle import ce_oracle
from oracle.ce_oracle import OracleChain

pytestmark = pytest.mark.gpu


def make_ce(case, seed=3):
    st, eci, symbols, ft = build(seed=seed, **case)
    atoms = syn.Atoms(symbols)
    calc = CE(atoms, st, dict(eci))
    return st, eci, symbols, ft, atoms, calc


def test_ce_calculator_energy_and_updates(cuda_device):
    st, eci, symbols, ft, atoms, calc = make_ce(TERNARY)
    oc = OracleChain(ft, ft.occupancy(symbols))
    # initial CFs come from the definition on the GPU: compare to the oracle's
    cf = calc.get_cf()
    np.testing.assert_allclose([cf[n] for n in ft.eci_names], oc.cf, rtol=0, atol=2e-14)
    assert abs(calc.get_energy() - oc.e) < 1e-10
    assert atoms.get_calculator() is calc
    # calculate / undo / clear with the reference's call pattern
    a, b = 0, [i for i in range(ft.N) if symbols[i] != symbols[0]][0]
    e0 = calc.get_energy()
    e1 = calc.calculate(atoms, ["energy"], [(a, symbols[a], symbols[b]), (b, symbols[b], symbols[a])])
    assert e1 != e0 and atoms[a].symbol == symbols[b]
    calc.undo_changes()
    assert calc.get_energy() == e0 and atoms[a].symbol == symbols[a]
    # set_symbols (bulk path) and set_composition
    new = symbols[::-1]
    calc.set_symbols(new)
    oc2 = OracleChain(ft, ft.occupancy(new))
    assert abs(calc.get_energy() - oc2.e) < 1e-10
    calc.set_composition({"Al": 0.5, "Mg": 0.25, "Si": 0.25})
    counts = {s: sum(1 for at in atoms if at.symbol == s) for s in ("Al", "Mg", "Si")}
    assert counts == {"Al": 32, "Mg": 16, "Si": 16}
    assert len(calc.get_singlets()) == 2
    copy = calc.copy()
    assert abs(copy.get_energy() - calc.get_energy()) < 1e-12


def test_supercell_cf_carry_over(cuda_device):
    """get_atoms_with_ce_calc: CFs of the small cell seed the supercell
    (ce_calculator.py:48-102, tests/test_CE_updater.py:117-157)."""
    small = syn.fcc_settings(4, ["Al", "Mg"])
    eci = syn.synthetic_ecis(small)
    atoms = get_atoms_with_ce_calc(small, small.kwargs, eci=eci, size=[2, 2, 2])
    calc = atoms.get_calculator()
    assert len(atoms) == 512
    ft = calc.updater.tables
    oc = OracleChain(ft, ft.occupancy([a.symbol for a in atoms]))
    np.testing.assert_allclose([calc.get_cf()[n] for n in ft.eci_names], oc.cf, atol=1e-13)
    calc.update_cf((7, "Al", "Mg"))
    oc.update_cf(7, ft.species_id["Mg"])
    np.testing.assert_allclose([calc.get_cf()[n] for n in ft.eci_names], oc.cf, atol=1e-13)


def _oracle_runmc(ft, occ, eci_vec, kT, seed, steps, sgc, cf0, e0_ref):
    """Montecarlo.runMC(equil=False) restated on the oracle
    (montecarlo.py:732-848): warm-up move, 1000-move bias probe with c0
    shifted by -bias/N, then `steps` sampled moves."""
    oc = OracleChain(ft, occ, cf=cf0, kT=kT, seed=seed, eci=eci_vec, ref=e0_ref)
    run = oc.run_sgc if sgc else oc.run_canonical
    run(1)
    run(1000)
    bias = oc.e
    e = oc.eci.copy()
    e[ft.eci_index["c0"]] -= bias / ft.N
    oc.set_ecis(e)
    assert abs(oc.e) < 1e-6
    oc.reset_acc()
    acc0 = oc.n_accepted.value
    run(steps)
    return oc, bias, oc.n_accepted.value - acc0


def test_montecarlo_runmc_matches_oracle(cuda_device):
    st, eci, symbols, ft, atoms, calc = make_ce(TERNARY, seed=5)
    cf0 = calc.updater.batch.get_cf()[0]
    e0 = calc.get_energy()
    T = 600.0
    mc = Montecarlo(atoms, T, seed=1234)
    seen = []

    class Obs(object):
        def __call__(self, system_changes):
            seen.append(len(system_changes))

        def reset(self):
            pass

        def get_averages(self):
            return {"obs_calls": len(seen)}
    mc.attach(Obs(), interval=500)
    mc.runMC(steps=2000, equil=False)
    oc, bias, n_acc = _oracle_runmc(ft, ft.occupancy(symbols), ft.eci, T * KB, 1234, 2000,
                                    False, cf0, e0)
    th = mc.get_thermodynamic()
    n = oc.acc[0]
    assert mc.energy_bias == bias
    assert th["energy"] == (oc.acc[1] / n) * e0 + bias
    assert th["heat_capacity"] == ((oc.acc[2] / n) * e0 - ((oc.acc[1] / n) * e0) ** 2) / (KB * T ** 2)
    assert mc.num_accepted == n_acc
    assert [a.symbol for a in atoms] == ft.symbols_of(oc.occ)
    assert th["obs_calls"] == 4 and th["temperature"] == T
    assert abs(th["Al_conc"] - 0.5) < 1e-12
    # c0 restored after the run (montecarlo.py:847)
    assert calc.eci["c0"] == pytest.approx(eci["c0"], abs=1e-15)


def test_montecarlo_errors(cuda_device):
    st = syn.fcc_settings(4, ["Al", "Mg"])
    atoms = syn.Atoms(["Al"] * 64)
    CE(atoms, st, syn.synthetic_ecis(st))
    mc = Montecarlo(atoms, 500.0, seed=1)
    with pytest.raises(TooFewElementsError):
        mc.runMC(steps=10, equil=False)
    with pytest.raises(NotImplementedError):
        mc.add_constraint(lambda ch: True)
    with pytest.raises(ValueError):
        mc.attach(3)


def test_sgc_runmc_matches_oracle(cuda_device):
    st, eci, symbols, ft, atoms, calc = make_ce(BINARY, seed=9)
    cf0 = calc.updater.batch.get_cf()[0]
    T, mu = 400.0, {"c1_0": 0.03}
    mc = SGCMonteCarlo(atoms, T, symbols=["Al", "Mg"], seed=77)
    e0 = mc.averager.energy.ref_value
    mc.runMC(steps=3000, chem_potential=mu, equil=False)
    ev = ft.eci.copy()
    ev[ft.eci_index["c1_0"]] -= mu["c1_0"]
    oc, bias, n_acc = _oracle_runmc(ft, ft.occupancy(symbols), ev, T * KB, 77, 3000, True, cf0, e0)
    th = mc.get_thermodynamic()
    n = oc.acc[0]
    singl = oc.acc[3] / n
    assert th["n_mc_steps"] == 3000
    assert th["singlet_c1_0"] == singl
    assert th["var_singlet_c1_0"] == oc.acc[4] / n - singl ** 2
    assert th["sgc_energy"] == (oc.acc[1] / n) * e0 + bias
    assert th["energy"] == (oc.acc[1] / n) * e0 + bias + mu["c1_0"] * singl * ft.N
    assert th["mu_c1_0"] == 0.03
    assert abs(th["Al_conc"] + th["Mg_conc"] - 1.0) < 1e-12
    assert [a.symbol for a in atoms] == ft.symbols_of(oc.occ)
    # the chemical potential was removed from the ECIs again (:446)
    assert calc.eci["c1_0"] == pytest.approx(eci["c1_0"], abs=1e-15)


def test_sgc_equilibration_runs(cuda_device):
    """runMC(equil=True): correlation time + window equilibration
    (reference smoke tests: tests/test_sgc_mc.py:41-104)."""
    st, eci, symbols, ft, atoms, calc = make_ce(BINARY, seed=2)
    mc = SGCMonteCarlo(atoms, 800.0, symbols=["Al", "Mg"], seed=5)
    mc.runMC(steps=2000, chem_potential={"c1_0": 0.0},
             equil_params={"window_length": 640, "maxiter": 200})
    th = mc.get_thermodynamic()
    assert th["n_mc_steps"] == 2000 and np.isfinite(th["sgc_heat_capacity"])
    assert mc.correlation_info["correlation_time_found"]


def test_parameter_sweep_matches_oracle(cuda_device, tmp_path):
    st, eci, symbols, ft, atoms, calc = make_ce(BINARY, seed=4)
    cf0 = calc.updater.batch.get_cf()[0]
    mc = SGCMonteCarlo(atoms, 500.0, symbols=["Al", "Mg"], seed=21)
    params = [{"temperature": T, "chemical_potential": {"c1_0": m}}
              for T in (300.0, 700.0) for m in (-0.02, 0.02)]
    out = str(tmp_path / "sweep.npz")
    res = MCParameterSweep(params, mc, nsteps=1500, outfile=out, equil_steps=640).run()
    for r, p in enumerate(params):
        ev = ft.eci.copy()
        ev[ft.eci_index["c1_0"]] -= p["chemical_potential"]["c1_0"]
        oc = OracleChain(ft, ft.occupancy(symbols), cf=cf0, kT=p["temperature"] * KB,
                         seed=21, replica=r, eci=ev)
        oc.run_sgc(640)
        oc.run_sgc(1000)
        bias = oc.e
        ev[ft.eci_index["c0"]] -= bias / ft.N
        oc.set_ecis(ev)
        oc.set_ref(bias)
        oc.reset_acc()
        oc.run_sgc(1500)
        s = oc.acc[3] / oc.acc[0]
        assert res[r]["singlet_c1_0"] == s
        assert res[r]["sgc_energy"] == oc.acc[1] / oc.acc[0] * bias + bias
    z = np.load(out)
    assert len(z["temperature"]) == 4
    assert len({round(float(x), 12) for x in z["singlet_c1_0"]}) == 4


def test_parallel_tempering_single_gpu(cuda_device, tmp_path):
    """Sync-free round loop: legs, device-drawn directions and exchange sweeps match the oracle;
    a sequence of run(num_exchange_cycles=1) calls continues the same direction / uniform
    streams (the reference's random.choice advances across calls, parallel_tempering.py:191)."""
    st, eci, symbols, ft, atoms, calc = make_ce(TERNARY, seed=6)
    cf0 = calc.updater.batch.get_cf()[0]
    temps = list(np.geomspace(1500.0, 100.0, 8))
    mc = Montecarlo(atoms, temps[0], seed=99)
    pt = ParallelTempering(mc, Tmax=1500.0, Tmin=100.0, temperatures=temps,
                           temp_scheme_file=str(tmp_path / "scheme.csv"))
    assert pt.temperature_scheme == temps
    pt.run(mc_args={"steps": 300}, num_exchange_cycles=4, timing=True)
    assert pt.last_timing["cycles"] == 4 and pt.last_timing["leg_ms"] > 0.0
    for _ in range(4):
        pt.run(mc_args={"steps": 300}, num_exchange_cycles=1)
    chains = [OracleChain(ft, ft.occupancy(symbols), cf=cf0, kT=temps[r] * KB, seed=99, replica=r)
              for r in range(8)]
    slots = np.arange(8, dtype=np.int32)
    kts = np.array(temps) * KB
    total = 0
    dirs = []
    for rnd in range(8):
        for c in chains:
            c.run_canonical(300)
        direction = ce_oracle.pt_direction(99, rnd)
        dirs.append(direction)
        slots, n_acc = ce_oracle.pt_exchange([c.e for c in chains], slots, kts, direction, 99, rnd)
        total += n_acc
        for r, c in enumerate(chains):
            c.kT = float(kts[slots[r]])
    assert len(set(dirs)) == 2          # both "up" and "down" sweeps occurred
    assert np.array_equal(pt.slot_of_replica, slots)
    assert pt.num_accepted_exchanges == total
    assert np.array_equal(pt.gpu.get_energy(), [c.e for c in chains])
    assert np.array_equal(pt.gpu.get_occupancy(), np.stack([c.occ for c in chains]))
    assert sorted(pt.temperature_of_replica()) == sorted(temps)
    with pytest.raises(ValueError):
        pt.run(mc_args={"steps": 10, "bogus": 1}, num_exchange_cycles=1)


def test_parallel_tempering_rejects_sgc(cuda_device, tmp_path):
    st, eci, symbols, ft, atoms, calc = make_ce(BINARY, seed=6)
    sgc = SGCMonteCarlo(atoms, 500.0, symbols=["Al", "Mg"], seed=1)
    with pytest.raises(TypeError):
        ParallelTempering(sgc, temperatures=[800.0, 400.0], temp_scheme_file=str(tmp_path / "s.csv"))


def test_checkpoint_roundtrip(cuda_device, tmp_path):
    """CE.save/load and Montecarlo.save/load (ce_calculator.py:533-594,
    montecarlo.py:1040-1074): resumed chains continue the same Philox stream."""
    st, eci, symbols, ft, atoms, calc = make_ce(BINARY, seed=8)
    mc = Montecarlo(atoms, 500.0, seed=31)
    mc._gpu.set_kT([500.0 * KB])
    mc._steps(500, observe=False)
    f = str(tmp_path / "mc.json")
    mc.save(f)
    mc._steps(300, observe=False)
    e_direct = mc.current_energy
    mc2 = Montecarlo.load(f)
    mc2._gpu.set_kT([500.0 * KB])
    mc2._steps(300, observe=False)
    assert abs(mc2.current_energy - e_direct) < 1e-9
    mc._sync_atoms(); mc2._sync_atoms()
    assert [a.symbol for a in mc2.atoms] == [a.symbol for a in mc.atoms]
    calc.save(str(tmp_path / "ce.json"))
    calc2 = CE.load(str(tmp_path / "ce.json"))
    assert abs(calc2.get_energy() - calc.get_energy()) < 1e-12


@pytest.mark.parametrize("where,variant", [("device", -1), ("device", 1), ("device", 3), ("device", 5), ("device", 9),
                                           ("host", -1)])
def test_state_observers_match_oracle(cuda_device, where, variant):
    """PairCorrelationObserver / LowestEnergyStructure / SiteOrderParameter / EnergyEvolution /
    EnergyHistogram see, on their interval boundaries, exactly the state
    the oracle chain has after the same number of moves -- folded on the device by the kernels'
    bookkeeping warp (no launch boundary per interval), or called on the host when a
    user-defined observer is attached as well."""
    from cemc_b200.mcmc import (EnergyEvolution, EnergyHistogram, LowestEnergyStructure,
                                PairCorrelationObserver, SiteOrderParameter)
    st, eci, symbols, ft, atoms, calc = make_ce(TERNARY, seed=9)
    cf0 = calc.updater.batch.get_cf()[0]
    e0 = calc.get_energy()
    T, steps, iv = 700.0, 3000, 250
    mc = Montecarlo(atoms, T, seed=77)
    pair, low = PairCorrelationObserver(calc), LowestEnergyStructure(calc, mc)
    order, evo, hist = SiteOrderParameter(atoms), EnergyEvolution(mc), EnergyHistogram(mc, buffer_size=8, n_bins=5)
    for o in (pair, low, order, evo, hist):
        mc.attach(o, interval=iv)
    seen = []
    if where == "host":
        class Mine(MCObserver):
            def __call__(self, system_changes):
                seen.append(len(system_changes))
        mc.attach(Mine(), interval=iv)
    if variant >= 0:
        mc._gpu.set_variant(variant, variant)
    l0 = mc._gpu.launch_count()
    mc.runMC(steps=steps, equil=False)
    launches = mc._gpu.launch_count() - l0
    if where == "device":
        assert mc._device_observer_plan() == (iv, 1 | 2 | 4 | 8)
        if variant >= 0:                         # (autotuning adds its own timed segments)
            assert launches < steps // iv        # the device loop did not stop on the boundaries
            assert mc._gpu.last_variant() == variant
    else:
        assert mc._device_observer_plan() is None and len(seen) == steps // iv

    # the same run on the oracle, stopped on the same boundaries
    oc = OracleChain(ft, ft.occupancy(symbols), cf=cf0, kT=T * KB, seed=77, ref=e0)
    oc.run_canonical(1)
    oc.run_canonical(1000)
    bias = oc.e
    e = oc.eci.copy()
    e[ft.eci_index["c0"]] -= bias / ft.N
    oc.set_ecis(e)
    oc.reset_acc()
    occ_start = oc.occ.copy()        # runMC resets the observers here (after the bias probe)
    pair_names = [n for n in ft.eci_names if n.startswith("c2_")]
    s1 = {n: 0.0 for n in pair_names}
    energies, changed, best_e, best_occ = [], [], np.inf, None
    occ_initial = ft.occupancy(symbols)
    for _ in range(steps // iv):
        oc.run_canonical(iv)
        for n in pair_names:
            s1[n] += oc.cf[ft.eci_index[n]]
        energies.append(oc.e)
        changed.append(int(np.count_nonzero(oc.occ != occ_initial)))
        if oc.e < best_e:
            best_e, best_occ = oc.e, oc.occ.copy()
    ncall = steps // iv
    assert pair.n_entries == ncall
    avg = pair.get_averages()
    for n in pair_names:
        assert avg[n] == s1[n] / ncall
        assert pair.get_std()[n] >= 0.0
    assert evo.energies == energies
    assert low.lowest_energy == best_e
    assert low.atoms.get_chemical_symbols() == ft.symbols_of(best_occ)
    assert set(low.lowest_energy_cf.keys()) == set(ft.eci_names)
    so = order.get_averages()
    assert so["site_order_average"] == pytest.approx(np.mean(changed), rel=0, abs=1e-12)
    assert so["site_order_std"] == pytest.approx(np.std(changed), abs=1e-9)
    h = hist.histogram
    assert h.sum() == ncall and len(h) == 5
    assert hist.Emin == min(energies[:8]) and hist.Emax == max(energies[:8])
    del occ_start


def test_equilibration_decisions_match_oracle(cuda_device):
    """The equilibration phase of runMC -- correlation time from the device-side energy
    autocorrelation (the trace never leaves the GPU), then windows until two consecutive mean
    energies agree -- takes the decisions the reference's logic takes on the oracle chain:
    same correlation time, same number of windows, same (mean, variance, z) per window
    (cemc/mcmc/montecarlo.py:461-511, :541-697; restated in oracle/sampler_oracle.py)."""
    from oracle import sampler_oracle
    st, eci, symbols, ft, atoms, calc = make_ce(TERNARY, seed=4)
    cf0 = calc.updater.batch.get_cf()[0]
    e0 = calc.get_energy()
    T, window = 900.0, 700
    mc = Montecarlo(atoms, T, seed=31)
    mc.runMC(steps=500, equil=True, equil_params={"window_length": window, "confidence_level": 0.3})
    hist = mc.equil_history
    assert len(hist) >= 2

    oc = OracleChain(ft, ft.occupancy(symbols), cf=cf0, kT=T * KB, seed=31, ref=e0)
    oc.run_canonical(1)                                      # runMC's first _mc_step (:765)
    tr = oc.run_canonical(1000, trace=True)                  # correlation-time window (:461-470)
    info = sampler_oracle.correlation_info(tr[4], 1000)
    assert mc.correlation_info["correlation_time_found"] == info["correlation_time_found"]
    assert mc.correlation_info["correlation_time"] == pytest.approx(info["correlation_time"], rel=1e-12)
    want = sampler_oracle.equilibrate(oc, oc.run_canonical, window, 0.3, info, e0)
    assert len(hist) == len(want)                            # same number of windows
    for (e_g, v_g, z_g), (e_o, v_o, z_o) in zip(hist, want):
        assert e_g == e_o                                    # the sums are bit-identical
        assert v_g == pytest.approx(v_o, rel=1e-12)
        assert (z_g is None and z_o is None) or z_g == pytest.approx(z_o, rel=1e-9, abs=1e-12)
    # ... and the chain itself: after equilibration runMC probes the bias (1000 moves) and samples
    oc.run_canonical(1000)
    bias = oc.e
    e = oc.eci.copy()
    e[ft.eci_index["c0"]] -= bias / ft.N
    oc.set_ecis(e)
    oc.run_canonical(500)
    assert [a.symbol for a in mc.atoms] == ft.symbols_of(oc.occ)


def test_energy_autocorrelation_kernel(cuda_device):
    """cemc_energy_autocorrelation against numpy on the same traced window, several replicas."""
    st, eci, symbols, ft = build(**BINARY)
    from cemc_b200.updater import BatchedCEUpdater
    R, n = 3, 900
    gpu = BatchedCEUpdater(ft, R)
    occ = np.stack([ft.occupancy(syn.random_symbols(st, BINARY["conc"], seed=50 + r)) for r in range(R)])
    gpu.set_occupancy(occ); gpu.recompute_cf(); gpu.set_kT([0.02, 0.05, 0.5]); gpu.seed(12)
    gpu.set_trace(n)
    gpu.run_canonical(n)
    out = gpu.energy_autocorrelation(n)
    e = gpu.get_trace(n)[4]
    for r in range(R):
        d = e[r] - e[r].mean()
        var = np.mean(d * d)
        acf = np.correlate(d, d, mode="full")[n - 1:] / (n * var)
        first = int(np.nonzero(acf < 0.5)[0][0])
        assert out[r, 0] == pytest.approx(e[r].mean(), rel=1e-13)
        assert out[r, 1] == pytest.approx(var, rel=1e-10)
        assert int(out[r, 2]) == first
