"""GPU parity tests: the CUDA path (through the C ABI) against the oracle on
the same seeded inputs, and against the committed golden fixtures recorded
from the reference's compiled C++ updater.  Bit-exact for accept sequences,
occupations, energies and CFs (reference operation order, fp64)."""
import numpy as np
import pytest

from cases import (BINARY, GOLDEN, GOLDEN_WORKLOADS, LAYERED, LAYERED_BINARY, TERNARY, build, load_golden,
                   load_golden_workload)
from cemc_b200.updater import BatchedCEUpdater, PyCEUpdater
from cemc_b200 import synthetic as syn
from oracle import ce_oracle
from oracle.ce_oracle import OracleChain

pytestmark = pytest.mark.gpu


def make_pair(ft, symbols_list, kTs, seed, offset=0, eci=None):
    """R replicas on the GPU and R oracle chains with identical inputs."""
    R = len(symbols_list)
    occ = np.stack([ft.occupancy(s) for s in symbols_list])
    chains = [OracleChain(ft, occ[r], kT=kTs[r], seed=seed, replica=offset + r,
                          eci=None if eci is None else eci[r]) for r in range(R)]
    gpu = BatchedCEUpdater(ft, R, replica_offset=offset)
    gpu.set_occupancy(occ)
    gpu.set_cf(np.stack([c.cf for c in chains]))
    if eci is not None:
        gpu.set_ecis(np.stack(eci))
    gpu.set_kT(kTs)
    gpu.seed(seed)
    return gpu, chains


def assert_state_equal(gpu, chains):
    occ = gpu.get_occupancy()
    cf = gpu.get_cf()
    e = gpu.get_energy()
    for r, c in enumerate(chains):
        assert np.array_equal(occ[r], c.occ), "occupancy differs, replica %d" % r
        assert np.array_equal(cf[r], c.cf), "CF differs, replica %d" % r
        assert e[r] == c.e, "energy differs, replica %d" % r


# kernel variants the recorded trajectories are put through: 0 spin kernel (warp per replica),
# 1 / 2 / 3 / 4 batch kernel (16 warps, cluster of 2) / (16,1) / (8,1) / (4,1), 5 generic
# one-move-at-a-time kernel, 6 batch kernel (8,1) with two moves per evaluation warp (binary +-1
# basis), 8 / 9 batch kernel (16,2) / (8,2) with site split (swaps)
REPLAY_VARIANTS = [0, 1, 2, 3, 4, 5, 6, 8, 9]


def _pin_replay_variant(gpu, variant, mode):
    """Pin the kernel variant of cemc_replay; skip when it does not apply to this system."""
    ev = gpu.get_batch_eval()
    if gpu.tables.n_symm > 1 and variant != 5 and not gpu.batch_kernel_applies():
        pytest.skip("several symmetry groups: this build runs them on the generic kernel")
    if variant == 0 and (ev != 1 or gpu.tables.n_symm > 1):
        pytest.skip("spin kernel: binary +-1 basis only")
    if variant == 6 and (ev != 1 or gpu.tables.K > 31):
        pytest.skip("two moves per warp: spin evaluation, K <= 31 only")
    if variant in (8, 9) and mode != "canonical":
        pytest.skip("site split: swaps only")
    gpu.set_variant(variant, variant)


@pytest.mark.parametrize("variant", REPLAY_VARIANTS)
@pytest.mark.parametrize("name", GOLDEN)
def test_replay_golden(cuda_device, name, variant):
    """Replay of trajectories recorded from the reference's own C++ CEUpdater, through every
    kernel family the samplers time."""
    meta, st, ft, z = load_golden(name)
    gpu = BatchedCEUpdater(ft, 1)
    gpu.set_occupancy(ft.occupancy(meta["symbols0"])[None])
    gpu.set_cf(z["cf0"][None])
    gpu.set_kT([meta["kT"]])
    assert gpu.get_energy()[0] == float(z["e0"])
    _pin_replay_variant(gpu, variant, meta["mode"])
    acc, e_after = gpu.replay(z["sites"][None], z["news"][None], z["u"][None])
    if ft.K <= 31 or variant in (0, 2, 3, 5):      # K = 42 (config 1): spin / wide batch / generic kernels
        assert gpu.last_variant() == variant
    assert np.array_equal(acc[0], z["accepted"])
    assert np.array_equal(e_after[0], z["e_after"])
    assert np.array_equal(gpu.get_cf()[0], z["cf_final"])
    assert np.array_equal(gpu.get_occupancy()[0], z["occ_final"])
    steps, n_acc = gpu.get_counters()
    assert steps[0] == len(z["u"]) and n_acc[0] == z["accepted"].sum()


@pytest.mark.parametrize("case", [BINARY, TERNARY, LAYERED, LAYERED_BINARY])
@pytest.mark.parametrize("mode", ["sgc", "canonical"])
def test_device_proposals_match_oracle(cuda_device, case, mode):
    """On-device Philox proposals + Metropolis == oracle chain, every step."""
    st, eci, symbols, ft = build(**case)
    R = 5
    syms = [syn.random_symbols(st, case["conc"], seed=10 + r) for r in range(R)]
    kTs = np.linspace(0.01, 0.09, R)
    gpu, chains = make_pair(ft, syms, kTs, seed=777, offset=3)
    n = 700
    gpu.set_trace(n)
    gpu.reset_accumulators()
    (gpu.run_sgc if mode == "sgc" else gpu.run_canonical)(n)
    gpu.synchronize()
    sites, news, u, acc, e = gpu.get_trace(n)
    for r, c in enumerate(chains):
        tr = c.run_sgc(n, trace=True) if mode == "sgc" else c.run_canonical(n, trace=True)
        assert np.array_equal(sites[r], tr[0])
        assert np.array_equal(news[r], tr[1])
        assert np.array_equal(u[r], tr[2])
        assert np.array_equal(acc[r], tr[3])
        assert np.array_equal(e[r], tr[4])
    assert_state_equal(gpu, chains)
    accs = gpu.get_accumulators()
    steps, n_acc = gpu.get_counters()
    for r, c in enumerate(chains):
        assert np.array_equal(accs[r], c.acc)
        assert steps[r] == n and n_acc[r] == c.n_accepted.value
    assert 0.02 < acc.mean() < 0.98


@pytest.mark.parametrize("mode", ["sgc", "canonical"])
def test_chunked_runs_equal_one_run(cuda_device, mode):
    st, eci, symbols, ft = build(**TERNARY)
    syms = [symbols, symbols]
    gpu1, chains = make_pair(ft, syms, [0.05, 0.02], seed=5)
    gpu2, _ = make_pair(ft, syms, [0.05, 0.02], seed=5)
    run1 = gpu1.run_sgc if mode == "sgc" else gpu1.run_canonical
    run2 = gpu2.run_sgc if mode == "sgc" else gpu2.run_canonical
    run1(600)
    for n in (1, 31, 32, 33, 203, 300):
        run2(n)
    gpu1.synchronize(); gpu2.synchronize()
    assert np.array_equal(gpu1.get_occupancy(), gpu2.get_occupancy())
    assert np.array_equal(gpu1.get_cf(), gpu2.get_cf())
    assert np.array_equal(gpu1.get_energy(), gpu2.get_energy())
    assert np.array_equal(gpu1.get_accumulators(), gpu2.get_accumulators())
    for c in chains:
        c.run_sgc(600) if mode == "sgc" else c.run_canonical(600)
    assert_state_equal(gpu1, chains)


def test_per_replica_ecis_chemical_potential(cuda_device):
    """SGC sweep: mu folded into the singlet ECIs per replica
    (sgc_montecarlo.py:239-261), Averager reference values per replica."""
    st, eci, symbols, ft = build(**BINARY)
    R = 4
    mus = [-0.2, -0.05, 0.05, 0.2]
    ecis = []
    for mu in mus:
        v = ft.eci.copy()
        v[ft.eci_index["c1_0"]] -= mu
        ecis.append(v)
    gpu, chains = make_pair(ft, [symbols] * R, [0.03] * R, seed=21, eci=ecis)
    refs = gpu.get_energy()
    assert all(refs[r] == chains[r].e for r in range(R))
    gpu.reset_accumulators(refs)
    for r, c in enumerate(chains):
        c.set_ref(refs[r])
    gpu.run_sgc(500)
    gpu.synchronize()
    for c in chains:
        c.run_sgc(500)
    assert_state_equal(gpu, chains)
    accs = gpu.get_accumulators()
    for r, c in enumerate(chains):
        assert np.array_equal(accs[r], c.acc)
    # different mu -> different compositions
    assert len({float(a[3]) for a in accs}) == R


@pytest.mark.parametrize("variant", REPLAY_VARIANTS)
@pytest.mark.parametrize("name", GOLDEN_WORKLOADS)
def test_replay_golden_baseline_sizes(cuda_device, name, variant):
    """BASELINE-size trajectories (replicas of bench configs[1], configs[2] and the north-star
    Al-Mg-Si SGC sweep, recorded from the compiled reference) through the timed kernels:
    accept sequence, energy after every step, final CFs and occupations bit-identical."""
    meta, ft, z = load_golden_workload(name)
    R = len(meta["replicas"])
    gpu = BatchedCEUpdater(ft, R)
    gpu.set_occupancy(z["occ0"])
    gpu.set_cf(z["cf0"])
    gpu.set_ecis(z["eci"])
    gpu.set_kT(z["kT"])
    assert np.array_equal(gpu.get_energy(), z["e0"])
    _pin_replay_variant(gpu, variant, meta["mode"])
    acc, e_after = gpu.replay(z["sites"], z["news"], z["u"])
    assert gpu.last_variant() == variant
    assert np.array_equal(acc, z["accepted"])
    assert np.array_equal(e_after, z["e_after"])
    assert np.array_equal(gpu.get_cf(), z["cf_final"])
    assert np.array_equal(gpu.get_occupancy(), z["occ_final"])
    steps, n_acc = gpu.get_counters()
    assert np.all(steps == z["u"].shape[1]) and np.array_equal(n_acc, z["accepted"].sum(axis=1))
    # the device's own CFs of the recorded start configuration equal the reference's
    gpu.set_occupancy(z["occ0"])
    gpu.recompute_cf()
    # (the device sums the 8000-site cells in another order: 24 x 8000 terms of magnitude <= 1.5)
    np.testing.assert_allclose(gpu.get_cf(), z["cf0"], rtol=0, atol=5e-13)


@pytest.mark.parametrize("variant", [-1, 0, 2, 3])
def test_replay_edge_cases_batch_kernels(cuda_device, variant):
    """No-op changes (ce_updater.cpp:315), neighbouring swap partners (A.2) and the same site
    twice, as uniform one-site / two-site records so that they run on the spin / batch kernels."""
    for case in (BINARY, TERNARY):
        st, eci, symbols, ft = build(**case)
        occ = ft.occupancy(symbols)
        S = ft.S
        nb = int(ft.trans[5, 0])
        other = [s for s in range(ft.N) if occ[s] != occ[5] and s != nb][0]
        two = np.array([[5, nb], [5, other], [7, 7], [nb, 5], [11, 12], [5, nb], [3, 3]], dtype=np.int32)
        two_new = np.array([[occ[nb], occ[5]], [occ[other], occ[5]], [(occ[7] + 1) % S, (occ[7] + 1) % S],
                            [occ[nb], occ[nb]], [occ[12], occ[11]], [occ[5], occ[nb]],
                            [(occ[3] + 1) % S, occ[3]]], dtype=np.int8)
        one = np.array([[5, -1], [9, -1], [5, -1], [nb, -1], [9, -1], [5, -1]], dtype=np.int32)
        one_new = np.array([[occ[5], 0], [(occ[9] + 1) % S, 0], [(occ[5] + 1) % S, 0], [occ[nb], 0],
                            [(occ[9] + 1) % S, 0], [occ[5], 0]], dtype=np.int8)
        for sites, news in ((one, one_new), (two, two_new)):
            u = np.linspace(0.05, 0.95, len(sites))
            gpu, chains = make_pair(ft, [symbols], [0.05], seed=0)
            if variant >= 0:
                if variant == 0 and gpu.get_batch_eval() != 1:
                    continue
                gpu.set_variant(variant, variant)
            acc, e = gpu.replay(sites[None], news[None], u[None])
            assert gpu.last_variant() != 5           # a sampler kernel took it
            acc_o, e_o = chains[0].replay(sites, news, u)
            assert np.array_equal(acc[0], acc_o)
            assert np.array_equal(e[0], e_o)
            assert_state_equal(gpu, chains)


def test_replay_edge_cases(cuda_device):
    """No-op changes (old == new, ce_updater.cpp:315), neighbouring swap
    partners (A.2), the same site twice, and one-site + two-site steps mixed."""
    st, eci, symbols, ft = build(**TERNARY)
    occ = ft.occupancy(symbols)
    nb = int(ft.trans[5, 0])           # a nearest neighbour of site 5
    other = [s for s in range(ft.N) if occ[s] != occ[5] and s != nb][0]
    sites = np.array([[5, -1], [5, nb], [5, other], [7, 7], [9, -1], [nb, 5], [11, 12]],
                     dtype=np.int32)
    news = np.array([[occ[5], 0], [occ[nb], occ[5]], [occ[other], occ[5]],
                     [(occ[7] + 1) % 3, (occ[7] + 2) % 3], [(occ[9] + 1) % 3, 0],
                     [occ[nb], occ[nb]], [occ[12], occ[11]]], dtype=np.int8)
    u = np.array([0.5, 0.9, 0.1, 0.3, 0.99, 0.2, 0.6])
    gpu, chains = make_pair(ft, [symbols], [0.05], seed=0)
    acc, e = gpu.replay(sites[None], news[None], u[None])
    acc_o, e_o = chains[0].replay(sites, news, u)
    assert np.array_equal(acc[0], acc_o)
    assert np.array_equal(e[0], e_o)
    assert_state_equal(gpu, chains)
    assert acc[0][0] == 1     # no-op step: E_new == E_cur, u <= exp(0)


def test_replay_empty_and_errors(cuda_device):
    from cemc_b200._lib import CemcError
    st, eci, symbols, ft = build(**BINARY)
    gpu, chains = make_pair(ft, [symbols], [0.05], seed=0)
    acc, e = gpu.replay(np.zeros((1, 0, 2), np.int32), np.zeros((1, 0, 2), np.int8),
                        np.zeros((1, 0)))
    assert acc.shape == (1, 0)
    with pytest.raises(CemcError):
        gpu.replay(np.array([[[ft.N, -1]]], np.int32), np.zeros((1, 1, 2), np.int8),
                   np.zeros((1, 1)))
    with pytest.raises(CemcError):
        gpu.set_occupancy(np.full((1, ft.N), 7, np.int8))
    assert_state_equal(gpu, chains)   # failed calls left the state untouched


def test_background_sites(cuda_device):
    """Background atoms never move (ce_updater.cpp:330, :1015-1027)."""
    from cemc_b200._lib import CemcError
    from cemc_b200.tables import FlatTables
    st = syn.fcc_settings(4, ["Al", "Mg"])
    bkg = [0, 17]
    st.background_indices = bkg
    st.index_by_trans_symm = [[s for s in range(64) if s not in bkg]]
    eci = syn.synthetic_ecis(st)
    symbols = syn.random_symbols(st, {"Al": 0.5, "Mg": 0.5}, seed=2)
    ft = FlatTables(st, eci, symbols)
    gpu, chains = make_pair(ft, [symbols], [0.05], seed=8)
    gpu.run_sgc(400)
    gpu.synchronize()
    chains[0].run_sgc(400)
    assert_state_equal(gpu, chains)
    occ0 = ft.occupancy(symbols)
    assert np.array_equal(gpu.get_occupancy()[0][bkg], occ0[bkg])
    with pytest.raises(CemcError, match="background"):
        gpu.replay(np.array([[[0, -1]]], np.int32),
                   np.array([[[1 - occ0[0], 0]]], np.int8), np.zeros((1, 1)))


@pytest.mark.parametrize("case", [TERNARY, LAYERED])
def test_recompute_cf_matches_definition(cuda_device, case):
    st, eci, symbols, ft = build(**case)
    gpu, chains = make_pair(ft, [symbols, symbols[::-1]], [0.05, 0.05], seed=1)
    chains[1] = OracleChain(ft, ft.occupancy(symbols[::-1]), kT=0.05, seed=1, replica=1)
    gpu.set_cf(np.zeros((2, ft.n_eci)))
    gpu.recompute_cf()
    cf = gpu.get_cf()
    for r in range(2):
        np.testing.assert_allclose(cf[r], chains[r].cf, rtol=0, atol=2e-14)
    np.testing.assert_allclose(gpu.get_energy(), [c.e for c in chains], rtol=1e-12)


def test_pyceupdater_drop_in(cuda_device):
    """calculate / undo_changes / clear_history / get_cf / get_singlets /
    set_ecis with the reference's call pattern (montecarlo.py:922,1015-1018)."""
    st, eci, symbols, ft = build(**TERNARY)
    oc = OracleChain(ft, ft.occupancy(symbols), kT=0.05)
    cf0 = {n: float(v) for n, v in zip(ft.eci_names, oc.cf)}
    atoms = syn.Atoms(symbols)
    upd = PyCEUpdater(atoms, st, cf0, dict(eci))
    assert upd.get_energy() == oc.e
    e0 = oc.e
    # a swap that is rejected, then one that is committed
    a, b = 3, [s for s in range(ft.N) if symbols[s] != symbols[3]][0]
    changes = [(a, symbols[a], symbols[b]), (b, symbols[b], symbols[a])]
    e1 = upd.calculate(changes)
    assert atoms[a].symbol == symbols[b] and atoms[b].symbol == symbols[a]
    upd.undo_changes()
    assert atoms[a].symbol == symbols[a] and atoms[b].symbol == symbols[b]
    assert upd.get_energy() == e0
    assert upd.get_cf() == cf0
    e2 = upd.calculate(changes)
    upd.clear_history()
    assert e2 == e1
    oc.update_cf(a, ft.species_id[symbols[b]])
    oc.update_cf(b, ft.species_id[symbols[a]])
    assert e2 == oc.e
    assert np.array_equal(np.array([upd.get_cf()[n] for n in ft.eci_names]), oc.cf)
    assert np.array_equal(upd.get_singlets(), oc.cf[ft.singlet_indices])
    assert upd.get_symbols() == ft.symbols_of(oc.occ)
    # chemical potential folded into an ECI (sgc_montecarlo.py:253-255)
    new_eci = dict(eci)
    new_eci["c1_0"] -= 0.37
    upd.set_ecis(new_eci)
    oc.set_ecis(ft.eci_vector(new_eci))
    assert upd.get_energy() == oc.e
    with pytest.raises(ValueError):
        upd.set_ecis({"c0": 0.0})
    # single flips through update_cf (CE.set_symbols path, ce_calculator.py:446)
    for s in (1, 2, 40):
        new = ft.species[(ft.species_id[atoms[s].symbol] + 1) % 3]
        upd.update_cf((s, atoms[s].symbol, new))
        oc.update_cf(s, ft.species_id[new])
    upd.clear_history()
    assert upd.get_energy() == oc.e


def test_pt_exchange_matches_oracle(cuda_device):
    import torch
    st, eci, symbols, ft = build(**BINARY)
    R = 8
    kts = np.geomspace(0.08, 0.01, R)
    syms = [syn.random_symbols(st, BINARY["conc"], seed=30 + r) for r in range(R)]
    gpu, chains = make_pair(ft, syms, kts, seed=4242)
    dev = torch.device("cuda:0")
    slots = torch.arange(R, dtype=torch.int32, device=dev)
    kt_slot = torch.tensor(kts, dtype=torch.float64, device=dev)
    n_acc = torch.zeros(2, dtype=torch.int32, device=dev)    # [this sweep, running total]
    slots_o = np.arange(R, dtype=np.int32)
    total = 0
    for rnd in range(6):
        gpu.run_canonical(200)
        gpu.synchronize()
        for c in chains:
            c.run_canonical(200)
        torch.cuda.synchronize()
        gpu.pt_exchange(R, gpu.energy_dev_ptr(), slots.data_ptr(), kt_slot.data_ptr(),
                        rnd % 2, rnd, n_acc.data_ptr())
        gpu.synchronize()
        slots_o, n_o = ce_oracle.pt_exchange([c.e for c in chains], slots_o, kts,
                                             rnd % 2, 4242, rnd)
        assert np.array_equal(slots.cpu().numpy(), slots_o)
        total += n_o
        assert n_acc.cpu().tolist() == [n_o, total]
        for r, c in enumerate(chains):
            c.kT = float(kts[slots_o[r]])
        assert np.array_equal(gpu.get_kT(), np.array([c.kT for c in chains]))
    assert_state_equal(gpu, chains)


def test_full_size_properties(cuda_device):
    """BASELINE config sizes (10^3 binary SGC, 20^3 ternary canonical):
    size-independent properties -- incremental CFs equal a from-scratch
    recompute, canonical moves conserve composition, replaying the device's
    own trace reproduces it, energies equal N * eci . cf."""
    for L, species, conc, mode in [(10, ["Al", "Mg"], {"Al": 0.5, "Mg": 0.5}, "sgc"),
                                   (20, ["Al", "Mg", "Si"],
                                    {"Al": 0.8, "Mg": 0.1, "Si": 0.1}, "canonical")]:
        st, eci, symbols, ft = build(L, species, ["nn", "2nn", "tri", "tet"], conc)
        R = 4
        occ0 = np.stack([ft.occupancy(symbols)] * R)
        gpu = BatchedCEUpdater(ft, R)
        gpu.set_occupancy(occ0)
        gpu.recompute_cf()
        cf0 = gpu.get_cf()
        gpu.set_kT(np.linspace(0.02, 0.08, R))
        gpu.seed(99)
        n = 3000
        gpu.set_trace(n)
        (gpu.run_sgc if mode == "sgc" else gpu.run_canonical)(n)
        gpu.synchronize()
        cf_inc, e_inc, occ1 = gpu.get_cf(), gpu.get_energy(), gpu.get_occupancy()
        tr = gpu.get_trace(n)
        assert 0.01 < tr[3].mean() < 0.99
        if mode == "canonical":
            for r in range(R):
                assert np.array_equal(np.bincount(occ1[r], minlength=3),
                                      np.bincount(occ0[r], minlength=3))
        np.testing.assert_array_equal(
            e_inc, [ce_oracle.OracleChain(ft, occ1[r], cf=cf_inc[r]).e for r in range(R)])
        gpu.recompute_cf()
        np.testing.assert_allclose(gpu.get_cf(), cf_inc, rtol=0, atol=1e-12)
        # replay of the recorded trace from the initial state: identical
        gpu2 = BatchedCEUpdater(ft, R)
        gpu2.set_occupancy(occ0)
        gpu2.set_cf(cf0)
        gpu2.set_kT(np.linspace(0.02, 0.08, R))
        acc, e = gpu2.replay(tr[0], tr[1], tr[2])
        assert np.array_equal(acc, tr[3]) and np.array_equal(e, tr[4])
        assert np.array_equal(gpu2.get_occupancy(), occ1)
        assert np.array_equal(gpu2.get_cf(), cf_inc)
        # and the oracle agrees on one replica end to end
        oc = OracleChain(ft, occ0[1], cf=cf0[1], kT=np.linspace(0.02, 0.08, R)[1])
        acc_o, e_o = oc.replay(tr[0][1], tr[1][1], tr[2][1])
        assert np.array_equal(acc_o, tr[3][1]) and np.array_equal(e_o, tr[4][1])
        assert np.array_equal(oc.cf, cf_inc[1])


def test_tree_order_mode_ternary(cuda_device):
    """CEMC_ORDER_TREE on a non-integer basis: decisions identical, energies
    and CFs within 1e-10 relative of the reference-order oracle."""
    from cemc_b200.updater import ORDER_TREE
    st, eci, symbols, ft = build(**TERNARY)
    gpu, chains = make_pair(ft, [symbols] * 3, [0.02, 0.05, 0.09], seed=31)
    gpu.set_order_mode(ORDER_TREE)
    n = 2000
    gpu.set_trace(n)
    gpu.run_canonical(n)
    gpu.synchronize()
    tr = gpu.get_trace(n)
    for r, c in enumerate(chains):
        o = c.run_canonical(n, trace=True)
        assert np.array_equal(tr[3][r], o[3])
        np.testing.assert_allclose(tr[4][r], o[4], rtol=1e-10, atol=1e-12)
    cf = gpu.get_cf()
    occ = gpu.get_occupancy()
    for r, c in enumerate(chains):
        assert np.array_equal(occ[r], c.occ)
        np.testing.assert_allclose(cf[r], c.cf, rtol=1e-10, atol=1e-13)


@pytest.mark.parametrize("threads", [32, 64, 128, 256])
def test_block_size_invariance(cuda_device, threads):
    st, eci, symbols, ft = build(**TERNARY)
    gpu, chains = make_pair(ft, [symbols] * 2, [0.03, 0.06], seed=77)
    gpu.set_block_threads(threads)
    gpu.run_canonical(300)
    gpu.run_sgc(300)
    gpu.synchronize()
    for c in chains:
        c.run_canonical(300)
        c.run_sgc(300)
    assert_state_equal(gpu, chains)


@pytest.mark.parametrize("case", [BINARY, TERNARY])
def test_generic_path_matches_oracle(cuda_device, case):
    """The shared-memory CF path (used for > 32 ECIs or several symmetry
    groups) is forced on a small problem and must equal the oracle too."""
    st, eci, symbols, ft = build(**case)
    gpu, chains = make_pair(ft, [symbols] * 2, [0.03, 0.07], seed=13)
    gpu.set_generic_path(True)
    gpu.reset_accumulators([2.0, 3.0])
    for c, ref in zip(chains, [2.0, 3.0]):
        c.set_ref(ref)
    gpu.run_sgc(400)
    gpu.run_canonical(400)
    gpu.synchronize()
    for c in chains:
        c.run_sgc(400)
        c.run_canonical(400)
    assert_state_equal(gpu, chains)
    accs = gpu.get_accumulators()
    for r, c in enumerate(chains):
        assert np.array_equal(accs[r], c.acc)


def test_many_ecis_generic_path(cuda_device):
    """> 32 ECIs: ternary with six families (generic kernel path by size)."""
    st, eci, symbols, ft = build(5, ["Al", "Cu", "Mg", "Si"],
                                 ["nn", "2nn", "tri", "iso", "tet"],
                                 {"Al": 0.4, "Cu": 0.2, "Mg": 0.2, "Si": 0.2})
    assert ft.n_eci > 32 and ft.D == 3
    gpu, chains = make_pair(ft, [symbols], [0.05], seed=3)
    gpu.run_canonical(300)
    gpu.run_sgc(300)
    gpu.synchronize()
    chains[0].run_canonical(300)
    chains[0].run_sgc(300)
    assert_state_equal(gpu, chains)


def test_averager_reference_value(cuda_device):
    """Averager sums E/ref and E^2/ref (averager.py:21-23) with ref != 1."""
    st, eci, symbols, ft = build(**TERNARY)
    gpu, chains = make_pair(ft, [symbols] * 2, [0.03, 0.07], seed=17)
    refs = gpu.get_energy()
    gpu.reset_accumulators(refs)
    for c, ref in zip(chains, refs):
        c.set_ref(ref)
    gpu.run_canonical(500)
    gpu.synchronize()
    accs = gpu.get_accumulators()
    for r, c in enumerate(chains):
        c.run_canonical(500)
        assert np.array_equal(accs[r], c.acc)


def test_exact_division_selftest(cuda_device):
    """The FMA-based division in the kernels is bit-identical to IEEE division
    (1.5e8 random operand pairs incl. the kernels' real denominators)."""
    st, eci, symbols, ft = build(**TERNARY)
    gpu = BatchedCEUpdater(ft, 1)
    assert gpu.selftest_division(seed=5, n_blocks=296, iters=2000) == 0


def test_sgc_restricted_species(cuda_device):
    """SGCMonteCarlo(symbols=[...]) restricts the inserted species
    (sgc_montecarlo.py:38-43,72-74)."""
    st, eci, symbols, ft = build(**TERNARY)
    gpu, chains = make_pair(ft, [symbols], [0.05], seed=23)
    gpu.set_sgc_species([0, 2])
    gpu.set_trace(500)
    gpu.run_sgc(500)
    gpu.synchronize()
    tr = chains[0].run_sgc(500, allowed=[0, 2], trace=True)
    sites, news, u, acc, e = gpu.get_trace(500)
    assert np.array_equal(news[0], tr[1]) and np.array_equal(acc[0], tr[3])
    assert set(np.unique(news[0][:, 0])) <= {0, 2}
    assert_state_equal(gpu, chains)


@pytest.mark.parametrize("table_eval", [True, False])
@pytest.mark.parametrize("batch", [-1, 4, 8, 16])
@pytest.mark.parametrize("mode", ["sgc", "canonical"])
def test_speculative_batch_kernel(cuda_device, batch, mode, table_eval):
    """Speculative batch evaluation (cemc_batch_kernel.cuh) keeps the chain
    exactly sequential: every batch size gives the oracle's trajectory, also on
    a tiny cell where moves of one batch collide all the time."""
    for case, R in ((TERNARY, 3), (dict(TERNARY, L=3), 2)):
        st, eci, symbols, ft = build(**case)
        kTs = np.linspace(0.02, 0.2, R)
        gpu, chains = make_pair(ft, [symbols] * R, kTs, seed=41)
        gpu.set_batch(batch)
        gpu.set_table_eval(table_eval)      # product tables vs fp64 products: same bits
        n = 1500
        gpu.set_trace(n)
        gpu.reset_accumulators()
        (gpu.run_sgc if mode == "sgc" else gpu.run_canonical)(n)
        (gpu.run_sgc if mode == "sgc" else gpu.run_canonical)(7)
        gpu.synchronize()
        tr = gpu.get_trace(7)
        for c in chains:
            c.run_sgc(n) if mode == "sgc" else c.run_canonical(n)
            o = c.run_sgc(7, trace=True) if mode == "sgc" else c.run_canonical(7, trace=True)
        assert_state_equal(gpu, chains)
        accs = gpu.get_accumulators()
        steps, n_acc = gpu.get_counters()
        for r, c in enumerate(chains):
            assert np.array_equal(accs[r], c.acc)
            assert steps[r] == n + 7 and n_acc[r] == c.n_accepted.value


@pytest.mark.parametrize("slack", [1e30, 1e6])
def test_batch_kernel_exact_decision_path(cuda_device, slack):
    """Force the batch kernel's Metropolis screen to defer to the exact
    expression (always / often): the trajectory must not change."""
    st, eci, symbols, ft = build(**TERNARY)
    gpu, chains = make_pair(ft, [symbols] * 2, [0.03, 0.08], seed=53)
    gpu.set_screen_slack(slack)
    gpu.reset_accumulators()
    gpu.run_canonical(800)
    gpu.run_sgc(800)
    gpu.synchronize()
    for c in chains:
        c.run_canonical(800)
        c.run_sgc(800)
    assert_state_equal(gpu, chains)
    accs = gpu.get_accumulators()
    for r, c in enumerate(chains):
        assert np.array_equal(accs[r], c.acc)


@pytest.mark.parametrize("species,conc", [(["Al", "Mg"], {"Al": 0.9, "Mg": 0.1}),
                                          (["Al", "Mg", "Si"], {"Al": 0.8, "Mg": 0.1, "Si": 0.1})])
def test_large_cell_global_state(cuda_device, species, conc):
    """BASELINE config 5 size (fcc 64^3 = 262 144 sites): the occupations do not
    fit in shared memory, the kernels keep them in global memory.  One chain,
    canonical + SGC moves, bit-compared with the oracle."""
    st, eci, symbols, ft = build(64, species, ["nn", "2nn", "tri", "tet"], conc)
    assert ft.N == 262144
    occ = ft.occupancy(symbols)
    gpu = BatchedCEUpdater(ft, 1)
    gpu.set_occupancy(occ[None])
    gpu.recompute_cf()
    cf0 = gpu.get_cf()[0]
    oc = OracleChain(ft, occ, cf=cf0, kT=0.04, seed=3)
    assert gpu.get_energy()[0] == oc.e
    gpu.set_kT([0.04])
    gpu.seed(3)
    gpu.run_canonical(1500)
    gpu.run_sgc(1500)
    gpu.synchronize()
    oc.run_canonical(1500)
    oc.run_sgc(1500)
    assert np.array_equal(gpu.get_occupancy()[0], oc.occ)
    assert np.array_equal(gpu.get_cf()[0], oc.cf)
    assert gpu.get_energy()[0] == oc.e
    # from-scratch CFs agree with the incrementally updated ones at this size
    gpu.recompute_cf()
    np.testing.assert_allclose(gpu.get_cf()[0], oc.cf, rtol=0, atol=1e-12)


@pytest.mark.parametrize("table_eval", [True, False])
@pytest.mark.parametrize("cluster", [1, 2])
@pytest.mark.parametrize("mode", ["sgc", "canonical"])
def test_cta_cluster_batch_kernel(cuda_device, cluster, mode, table_eval):
    """Two CTAs of a thread-block cluster cooperating on one chain (DSMEM exchange
    of proposals / quotients / conflict masks, commits to both copies of the
    state): same trajectory as the oracle, for state in shared memory (4^3, 3^3)
    and in global memory (64^3 is covered by test_large_cell_global_state)."""
    for case, R in ((TERNARY, 3), (dict(TERNARY, L=3), 2)):
        st, eci, symbols, ft = build(**case)
        kTs = np.linspace(0.02, 0.2, R)
        gpu, chains = make_pair(ft, [symbols] * R, kTs, seed=61)
        gpu.set_cluster(cluster)
        gpu.set_table_eval(table_eval)
        n = 1200
        gpu.set_trace(n)
        gpu.reset_accumulators()
        (gpu.run_sgc if mode == "sgc" else gpu.run_canonical)(n)
        gpu.synchronize()
        tr = gpu.get_trace(n)
        for r, c in enumerate(chains):
            o = c.run_sgc(n, trace=True) if mode == "sgc" else c.run_canonical(n, trace=True)
            assert np.array_equal(tr[0][r], o[0]) and np.array_equal(tr[3][r], o[3])
            assert np.array_equal(tr[4][r], o[4])
        assert_state_equal(gpu, chains)
        accs = gpu.get_accumulators()
        for r, c in enumerate(chains):
            assert np.array_equal(accs[r], c.acc)


def test_autotuned_long_run_is_invariant(cuda_device):
    """A long run is split into timed segments of different kernel variants by the
    autotuner: the trajectory must equal the oracle's regardless of the choice."""
    st, eci, symbols, ft = build(**TERNARY)
    gpu, chains = make_pair(ft, [symbols] * 2, [0.03, 0.07], seed=71)
    n = 70000
    gpu.reset_accumulators()
    gpu.run_canonical(n)
    gpu.run_sgc(n)
    gpu.synchronize()
    assert min(gpu.get_variant()) >= 0
    for c in chains:
        c.run_canonical(n)
        c.run_sgc(n)
    assert_state_equal(gpu, chains)
    accs = gpu.get_accumulators()
    for r, c in enumerate(chains):
        assert np.array_equal(accs[r], c.acc)


@pytest.mark.parametrize("batch,cluster", [(4, 1), (8, 1), (16, 1), (16, 2)])
@pytest.mark.parametrize("mode", ["sgc", "canonical"])
def test_batch_kernel_spin_evaluation(cuda_device, batch, cluster, mode):
    """Binary +-1 basis inside the batch kernel (XOR / ballot / popcount evaluation
    per warp, screened in-order decisions): oracle trajectory for every batch size,
    also on the 27-site cell and on config 1's six-family cluster set."""
    cases = [(BINARY, 3), (dict(BINARY, L=3), 2),
             (dict(BINARY, families=["nn", "2nn", "3nn", "tri", "iso", "tet"]), 2)]
    for case, R in cases:
        st, eci, symbols, ft = build(**case)
        kTs = np.linspace(0.02, 0.2, R)
        gpu, chains = make_pair(ft, [symbols] * R, kTs, seed=83)
        gpu.set_batch(batch)
        gpu.set_cluster(cluster)
        n = 1500
        gpu.set_trace(n)
        gpu.reset_accumulators()
        (gpu.run_sgc if mode == "sgc" else gpu.run_canonical)(n)
        gpu.synchronize()
        tr = gpu.get_trace(n)
        for r, c in enumerate(chains):
            o = c.run_sgc(n, trace=True) if mode == "sgc" else c.run_canonical(n, trace=True)
            assert np.array_equal(tr[0][r], o[0]) and np.array_equal(tr[3][r], o[3])
            assert np.array_equal(tr[4][r], o[4])
        assert_state_equal(gpu, chains)
        accs = gpu.get_accumulators()
        for r, c in enumerate(chains):
            assert np.array_equal(accs[r], c.acc)


@pytest.mark.parametrize("mode", ["sgc", "canonical"])
def test_two_moves_per_warp(cuda_device, mode):
    """Kernel variant 6: (8,1) batch kernel with TWO moves per evaluation warp (14-move batches,
    the two spin evaluations interleaved).  Oracle trajectory, trace, state and observer sums,
    also on the 27-site cell (batches collide constantly) and with forced exact decisions."""
    cases = [(BINARY, 3), (dict(BINARY, L=3), 2),
             (dict(BINARY, L=5, families=["nn", "2nn", "tri"]), 2)]
    for case, R in cases:
        st, eci, symbols, ft = build(**case)
        kTs = np.linspace(0.02, 0.2, R)
        for slack in (1.0, 1e30):
            gpu, chains = make_pair(ft, [symbols] * R, kTs, seed=91)
            gpu.set_variant(6, 6)
            gpu.set_screen_slack(slack)
            n = 1500 if slack == 1.0 else 300
            gpu.set_trace(n)
            gpu.reset_accumulators()
            (gpu.run_sgc if mode == "sgc" else gpu.run_canonical)(n)
            assert gpu.last_variant() == 6
            tr = gpu.get_trace(n)
            for r, c in enumerate(chains):
                o = c.run_sgc(n, trace=True) if mode == "sgc" else c.run_canonical(n, trace=True)
                assert np.array_equal(tr[0][r], o[0]) and np.array_equal(tr[3][r], o[3])
                assert np.array_equal(tr[4][r], o[4])
            assert_state_equal(gpu, chains)
            assert np.array_equal(gpu.get_accumulators(), np.stack([c.acc for c in chains]))
            gpu.close()


@pytest.mark.parametrize("order", ["reference", "tree"])
@pytest.mark.parametrize("batch,cluster", [(4, 1), (8, 1), (16, 1), (16, 2)])
def test_table_evaluation_quaternary(cuda_device, batch, cluster, order):
    """Product-table evaluation of the batch kernel on a four-species system (three
    basis functions, S^n-entry tables per decoration) and, with quadruplets, the
    fall-back to fp64 products when the tables would not fit: oracle trajectory
    either way, in the reference's summation order and in TREE order."""
    from cemc_b200.updater import ORDER_TREE
    species = ["Al", "Cu", "Mg", "Si"]
    conc = {"Al": 0.4, "Cu": 0.2, "Mg": 0.2, "Si": 0.2}
    for families, ev in ((["nn", "2nn", "tri"], 2), (["nn", "tet"], 0)):
        st, eci, symbols, ft = build(4, species, families, conc)
        assert ft.n_eci <= 32
        kTs = [0.03, 0.09]
        gpu, chains = make_pair(ft, [symbols] * 2, kTs, seed=97)
        assert gpu.get_batch_eval() == ev
        gpu.set_batch(batch)
        gpu.set_cluster(cluster)
        if order == "tree":
            gpu.set_order_mode(ORDER_TREE)
        n = 1000
        gpu.set_trace(n)
        gpu.reset_accumulators()
        gpu.run_sgc(n)
        gpu.run_canonical(n)
        gpu.synchronize()
        tr = gpu.get_trace(n)
        for r, c in enumerate(chains):
            c.run_sgc(n)
            o = c.run_canonical(n, trace=True)
            assert np.array_equal(tr[3][r], o[3])            # accept/reject sequence
            if order == "reference":
                assert np.array_equal(tr[4][r], o[4])
        if order == "reference":
            assert_state_equal(gpu, chains)
        else:
            assert np.array_equal(gpu.get_occupancy(), np.stack([c.occ for c in chains]))
            np.testing.assert_allclose(gpu.get_cf(), np.stack([c.cf for c in chains]), rtol=1e-10, atol=1e-13)


@pytest.mark.parametrize("variant", [8, 9])
@pytest.mark.parametrize("system", ["binary", "ternary_tab", "ternary_product"])
@pytest.mark.parametrize("mode", ["sgc", "canonical"])
def test_site_split_variants(cuda_device, variant, system, mode):
    """Kernel variants 8 / 9: site split -- the two CTAs of a cluster (16 / 8 warps each)
    evaluate the two changed sites of the same swaps (canonical only; SGC and the fp64
    product evaluation fall back to the default order).
    Same trajectory, trace, observer sums as the oracle for the three evaluation schemes,
    also on the 27-site cell (constant collisions)."""
    base = BINARY if system == "binary" else TERNARY
    for case, R in ((base, 3), (dict(base, L=3), 2)):
        st, eci, symbols, ft = build(**case)
        kTs = np.linspace(0.02, 0.2, R)
        gpu, chains = make_pair(ft, [symbols] * R, kTs, seed=131)
        if system != "binary":
            gpu.set_table_eval(system == "ternary_tab")
        gpu.set_variant(variant, variant)
        n = 1500
        gpu.set_trace(n)
        gpu.reset_accumulators()
        (gpu.run_sgc if mode == "sgc" else gpu.run_canonical)(n)
        gpu.synchronize()
        assert gpu.get_variant() == (variant, variant)
        if mode == "canonical" and system != "ternary_product":
            assert gpu.last_variant() == variant
        tr = gpu.get_trace(n)
        for r, c in enumerate(chains):
            o = c.run_sgc(n, trace=True) if mode == "sgc" else c.run_canonical(n, trace=True)
            assert np.array_equal(tr[0][r], o[0]) and np.array_equal(tr[3][r], o[3])
            assert np.array_equal(tr[4][r], o[4])
        assert_state_equal(gpu, chains)
        accs = gpu.get_accumulators()
        for r, c in enumerate(chains):
            assert np.array_equal(accs[r], c.acc)


def test_short_launch_autotuning_is_invariant(cuda_device):
    """Launches too short for the in-run autotuner (e.g. the legs between parallel-tempering
    exchanges) are tuned across calls: every call times another kernel variant, then the
    fastest is kept.  The trajectory must not depend on any of it."""
    st, eci, symbols, ft = build(**TERNARY)
    gpu, chains = make_pair(ft, [symbols] * 3, [0.03, 0.07, 0.15], seed=57)
    gpu.reset_accumulators()
    for _ in range(11):
        gpu.run_canonical(300)
        gpu.run_sgc(260)
    gpu.synchronize()
    assert min(gpu.get_variant()) >= 0           # both samplers settled on a variant
    for c in chains:
        for _ in range(11):
            c.run_canonical(300)
            c.run_sgc(260)
    assert_state_equal(gpu, chains)
    accs = gpu.get_accumulators()
    for r, c in enumerate(chains):
        assert np.array_equal(accs[r], c.acc)


@pytest.mark.parametrize("variant", [-1, 1, 2, 3])
def test_fp32_variant(cuda_device, variant):
    """The fp32 variant (cemc_set_precision(32)): product tables and sub-cluster sums in
    single precision, quotients / CF vector / energies in fp64.  North-star bar: the same
    accept/reject decisions as the fp64 reference, energies and CFs within 1e-5 relative
    (tolerance of this test; the fp64 default is bit-exact)."""
    from cemc_b200._lib import CemcError
    st, eci, symbols, ft = build(**dict(TERNARY, L=5))
    R = 4
    kTs = np.linspace(0.02, 0.15, R)
    gpu, chains = make_pair(ft, [symbols] * R, kTs, seed=211)
    gpu.set_precision(32)
    assert gpu.get_batch_eval() == 3
    if variant >= 0:
        gpu.set_variant(variant, variant)
    n = 4000
    gpu.set_trace(n)
    gpu.run_sgc(n)
    gpu.synchronize()
    acc_sgc = gpu.get_trace(n)[3].copy()
    gpu.run_canonical(n)
    gpu.synchronize()
    tr = gpu.get_trace(n)
    for r, c in enumerate(chains):
        o = c.run_sgc(n, trace=True)
        assert np.array_equal(acc_sgc[r], o[3])                # decisions identical
        o = c.run_canonical(n, trace=True)
        assert np.array_equal(tr[3][r], o[3])
        # energy after every move: 1e-5 relative to the energy scale of the trajectory
        np.testing.assert_allclose(tr[4][r], o[4], rtol=1e-5, atol=1e-5 * np.abs(o[4]).max())
    assert np.array_equal(gpu.get_occupancy(), np.stack([c.occ for c in chains]))
    cf, cf_ref = gpu.get_cf(), np.stack([c.cf for c in chains])
    np.testing.assert_allclose(cf, cf_ref, rtol=1e-5, atol=1e-5 * np.abs(cf_ref).max())
    e_ref = np.array([c.e for c in chains])
    np.testing.assert_allclose(gpu.get_energy(), e_ref, rtol=1e-5, atol=1e-5 * np.abs(e_ref).max())
    assert not np.array_equal(cf, cf_ref)                      # it really was single precision
    # back to fp64: bit-exact again from the oracle's state
    gpu.set_precision(64)
    gpu.set_cf(cf_ref)
    gpu.run_sgc(500)
    gpu.synchronize()
    for c in chains:
        c.run_sgc(500)
    assert_state_equal(gpu, chains)
    # a binary +-1 system is integer arithmetic in either setting
    st, eci, symbols, ft = build(**BINARY)
    gpu, chains = make_pair(ft, [symbols] * 2, [0.03, 0.1], seed=5)
    gpu.set_precision(32)
    gpu.run_sgc(1000)
    gpu.synchronize()
    for c in chains:
        c.run_sgc(1000)
    assert_state_equal(gpu, chains)
    # no table evaluation for this system (quaternary quadruplets): refused, loudly
    species = ["Al", "Cu", "Mg", "Si"]
    st, eci, symbols, ft = build(4, species, ["nn", "tet"], {"Al": 0.4, "Cu": 0.2, "Mg": 0.2, "Si": 0.2})
    gpu, chains = make_pair(ft, [symbols], [0.05], seed=5)
    with pytest.raises(CemcError):
        gpu.set_precision(32)


@pytest.mark.parametrize("variant", [1, 2, 8, 9])
def test_full_size_cluster_stress(cuda_device, variant):
    """BASELINE config 3 size (fcc 20^3 ternary, 8000 sites), 12 replicas over the whole
    temperature range, 30 000 SGC + 30 000 canonical moves: the CTA-cluster kernels with the
    async DSMEM protocol (variant 1), the site-split variant (8) and the single-CTA batch
    kernel (2) must give the oracle's occupations / CFs / energies / observer sums bit for bit
    (thousands of batches per chain, every barrier phase and ring wrap exercised)."""
    from cemc_b200 import workloads as wl
    R, n = 12, 30000
    w = wl.c3s_almgsi_sgc(R=R)
    ft = w.tables
    kT = np.linspace(w.kT.min(), w.kT.max() * 1.5, R)
    chains = [OracleChain(ft, w.occ[r], kT=kT[r], seed=4242, replica=r, eci=w.eci_matrix[r])
              for r in range(R)]
    gpu = BatchedCEUpdater(ft, R)
    gpu.set_occupancy(w.occ)
    gpu.set_cf(np.stack([c.cf for c in chains]))
    gpu.set_ecis(w.eci_matrix)
    gpu.set_kT(kT)
    gpu.seed(4242)
    gpu.set_variant(variant, variant)
    gpu.reset_accumulators()
    gpu.run_sgc(n)
    gpu.run_canonical(n)
    gpu.synchronize()
    for c in chains:
        c.run_sgc(n)
        c.run_canonical(n)
    assert_state_equal(gpu, chains)
    accs = gpu.get_accumulators()
    steps, n_acc = gpu.get_counters()
    for r, c in enumerate(chains):
        assert np.array_equal(accs[r], c.acc)
        assert steps[r] == 2 * n and n_acc[r] == c.n_accepted.value


@pytest.mark.parametrize("variant", [3, 6])
def test_full_size_bench_workload_stress(cuda_device, variant):
    """BASELINE config 2 (the bench workload: fcc 10^3 Al-Mg, replicas across the whole mu x T grid
    from its coldest to its hottest corner), 12 replicas x 250 000 SGC moves in two launches through
    the bench kernels -- the (8,1) batch kernel where every warp decides (variant 3) and its
    two-moves-per-warp form (6, the tuner's pick): occupations / CFs / energies / observer sums /
    counters equal the oracle's.  (Some corners of the grid end up single-element, so no swaps here.)"""
    from cemc_b200 import workloads as wl
    w = wl.c2_almg_sgc_sweep()
    R, n, n2 = 12, 200000, 50000
    pick = np.linspace(0, w.R - 1, R).astype(int)
    ft = w.tables
    chains = [OracleChain(ft, w.occ[q], kT=w.kT[q], seed=777, replica=r, eci=w.eci_matrix[q])
              for r, q in enumerate(pick)]
    gpu = BatchedCEUpdater(ft, R)
    gpu.set_occupancy(w.occ[pick])
    gpu.set_cf(np.stack([c.cf for c in chains]))
    gpu.set_ecis(w.eci_matrix[pick])
    gpu.set_kT(w.kT[pick])
    gpu.seed(777)
    gpu.set_variant(variant, variant)
    gpu.reset_accumulators()
    gpu.run_sgc(n)
    assert gpu.last_variant() == variant
    gpu.run_sgc(n2)
    gpu.synchronize()
    for c in chains:
        c.run_sgc(n)
        c.run_sgc(n2)
    assert_state_equal(gpu, chains)
    accs = gpu.get_accumulators()
    steps, n_acc = gpu.get_counters()
    for r, c in enumerate(chains):
        assert np.array_equal(accs[r], c.acc)
        assert steps[r] == n + n2 and n_acc[r] == c.n_accepted.value


def test_replica_order_is_invisible(cuda_device):
    """CTA -> replica order of the batch kernel (load balance): a user permutation and the
    automatic hottest-first order used when there are more chains than SMs (R = 160 > 148) both
    leave every chain's trajectory bit-identical to the oracle."""
    st, eci, symbols, ft = build(**BINARY)
    rng = np.random.default_rng(5)
    for R, user in ((160, False), (7, True)):
        kTs = rng.uniform(0.02, 0.2, R)
        gpu, chains = make_pair(ft, [symbols] * R, kTs, seed=19)
        gpu.set_variant(3, 3)
        if user:
            gpu.set_replica_order(rng.permutation(R))
        gpu.reset_accumulators()
        gpu.run_canonical(400)
        gpu.run_sgc(700)
        gpu.set_kT(kTs[::-1].copy())              # the automatic order follows the temperatures
        gpu.run_sgc(500)
        gpu.synchronize()
        for r, c in enumerate(chains):
            c.run_canonical(400)
            c.run_sgc(700)
            c.kT = float(kTs[::-1][r])
            c.run_sgc(500)
        assert_state_equal(gpu, chains)
        accs = gpu.get_accumulators()
        for r, c in enumerate(chains):
            assert np.array_equal(accs[r], c.acc)
    with pytest.raises(Exception):
        gpu.set_replica_order([0] * 7)            # not a permutation


@pytest.mark.parametrize("mode", ["sgc", "canonical"])
@pytest.mark.parametrize("case", [BINARY, TERNARY])
def test_lattice_arithmetic_matches_table(cuda_device, case, mode):
    """A translation-invariant lattice is detected (and verified) at create; T(site, col) by index
    arithmetic gives the trajectory the table gather gives (0 bytes of T traffic)."""
    st, eci, symbols, ft = build(**dict(case, L=5))
    syms = [syn.random_symbols(st, case["conc"], seed=70 + r) for r in range(3)]
    out = []
    for on in (True, False):
        gpu, chains = make_pair(ft, syms, [0.03, 0.05, 0.09], seed=321)
        gpu.set_lattice_arithmetic(True)             # verified at create: fcc L^3 with site = (i L + j) L + k
        assert gpu.get_lattice_arithmetic()          # (default: only for tables larger than L1)
        gpu.set_lattice_arithmetic(on)
        assert gpu.get_lattice_arithmetic() == on
        for v in (2, 3):
            gpu.set_variant(v, v)
            (gpu.run_sgc if mode == "sgc" else gpu.run_canonical)(700)
            assert gpu.last_variant() == v
        for c in chains:
            (c.run_sgc if mode == "sgc" else c.run_canonical)(1400)
        assert_state_equal(gpu, chains)
        out.append(gpu.get_accumulators())
    assert np.array_equal(out[0], out[1])


def test_lattice_arithmetic_rejects_non_lattice_tables(cuda_device):
    """A table that is not a periodic shift on the hinted grid keeps the gather."""
    st, eci, symbols, ft = build(**BINARY)
    ft.lattice_dims = np.array([4, 4, 4], dtype=np.int32)
    perm = np.arange(ft.N)
    perm[[3, 17]] = perm[[17, 3]]                # relabel two sites: same physics, no longer a grid
    ft.trans = np.ascontiguousarray(perm[ft.trans[np.argsort(perm)]].astype(np.int32))
    gpu = BatchedCEUpdater(ft, 1)
    assert not gpu.get_lattice_arithmetic()
    gpu.set_lattice_arithmetic(True)             # cannot be forced on
    assert not gpu.get_lattice_arithmetic()
    chain = OracleChain(ft, ft.occupancy(symbols), kT=0.05, seed=3)
    gpu.set_occupancy(ft.occupancy(symbols)[None]); gpu.set_cf(chain.cf[None]); gpu.set_kT([0.05]); gpu.seed(3)
    gpu.run_sgc(500); chain.run_sgc(500)
    assert_state_equal(gpu, [chain])


@pytest.mark.parametrize("mode", ["sgc", "canonical"])
@pytest.mark.parametrize("batch", [8, 16])
def test_table_evaluation_wide_columns(cuda_device, batch, mode):
    """Ternary system with all six cluster families (K = 42 translation columns, 27 ECIs): the
    table evaluation of the batch kernel takes two columns per lane (32 <= K <= 63)."""
    case = dict(L=5, species=["Al", "Mg", "Si"], families=["nn", "2nn", "3nn", "tri", "iso", "tet"],
                conc={"Al": 0.5, "Mg": 0.25, "Si": 0.25})
    st, eci, symbols, ft = build(**case)
    assert ft.K == 42 and ft.n_eci <= 32
    syms = [syn.random_symbols(st, case["conc"], seed=90 + r) for r in range(3)]
    gpu, chains = make_pair(ft, syms, [0.02, 0.05, 0.12], seed=77)
    assert gpu.batch_kernel_applies() and gpu.get_batch_eval() == 2
    gpu.set_batch(batch)
    n = 1200
    gpu.set_trace(n)
    (gpu.run_sgc if mode == "sgc" else gpu.run_canonical)(n)
    assert gpu.last_variant() == (2 if batch == 16 else 3)
    tr = gpu.get_trace(n)
    for r, c in enumerate(chains):
        o = c.run_sgc(n, trace=True) if mode == "sgc" else c.run_canonical(n, trace=True)
        assert np.array_equal(tr[3][r], o[3]) and np.array_equal(tr[4][r], o[4])
    assert_state_equal(gpu, chains)
    assert np.array_equal(gpu.get_accumulators(), np.stack([c.acc for c in chains]))


@pytest.mark.parametrize("mode", ["sgc", "canonical"])
@pytest.mark.parametrize("system,batch", [("quaternary_products", 4), ("quinary_tables", 8), ("quinary_tables", 16)])
def test_two_ecis_per_lane(cuda_device, system, batch, mode):
    """33..64 ECIs: the batch kernel keeps two ECIs per lane (CF vector, quotients, ordered
    energy dot over both slots).  Quaternary standard families = 41 ECIs (fp64 product
    evaluation: the quadruplet tables do not fit), quinary pairs + triplets = 45 ECIs (product
    tables).  Trajectory, trace, CFs, energies and observer sums equal the oracle's."""
    if system == "quaternary_products":
        species, fams, ev = ["Al", "Cu", "Mg", "Si"], ["nn", "2nn", "tri", "tet"], 0
    else:
        species, fams, ev = ["Al", "Cu", "Mg", "Si", "Zn"], ["nn", "2nn", "tri"], 2
    conc = {s: 1.0 / len(species) for s in species}
    st, eci, symbols, ft = build(4, species, fams, conc)
    assert 32 < ft.n_eci <= 64
    syms = [syn.random_symbols(st, conc, seed=40 + r) for r in range(3)]
    gpu, chains = make_pair(ft, syms, [0.03, 0.06, 0.15], seed=55)
    assert gpu.batch_kernel_applies() and gpu.get_batch_eval() == ev
    gpu.set_batch(batch)
    n = 1000
    gpu.set_trace(n)
    gpu.reset_accumulators()
    (gpu.run_sgc if mode == "sgc" else gpu.run_canonical)(n)
    if system == "quaternary_products" and mode == "canonical":
        assert gpu.last_variant() == 5       # the product scratch of two changed sites does not fit: generic kernel
    else:
        assert gpu.last_variant() == {16: 2, 8: 3, 4: 4}[batch]
    tr = gpu.get_trace(n)
    for r, c in enumerate(chains):
        o = c.run_sgc(n, trace=True) if mode == "sgc" else c.run_canonical(n, trace=True)
        assert np.array_equal(tr[3][r], o[3]) and np.array_equal(tr[4][r], o[4])
    assert_state_equal(gpu, chains)
    assert np.array_equal(gpu.get_accumulators(), np.stack([c.acc for c in chains]))
    # forced exact decisions (the screen always defers): the ordered dot over both ECI slots
    gpu2, chains2 = make_pair(ft, syms, [0.03, 0.06, 0.15], seed=56)
    gpu2.set_batch(batch)
    gpu2.set_screen_slack(1e30)
    (gpu2.run_sgc if mode == "sgc" else gpu2.run_canonical)(300)
    for c in chains2:
        (c.run_sgc if mode == "sgc" else c.run_canonical)(300)
    assert_state_equal(gpu2, chains2)


@pytest.mark.parametrize("variant", [1, 2, 3, 4, 8, 9])
@pytest.mark.parametrize("case", [LAYERED, LAYERED_BINARY])
@pytest.mark.parametrize("mode", ["sgc", "canonical"])
def test_multi_group_batch_kernel(cuda_device, mode, case, variant):
    """Two translational symmetry groups with different cluster families (crystal with a basis,
    ce_updater.cpp:379-384): the batch kernel's table evaluation picks descriptors, task list
    and per-ECI constants by the changed site's group.  Every batch flavour reproduces the
    oracle's trajectory, trace, CFs, energies and observer sums."""
    if variant in (8, 9) and mode != "canonical":
        pytest.skip("site split: swaps only")
    for L in (4, 6):
        st, eci, symbols, ft = build(**dict(case, L=L))
        assert ft.n_symm == 2
        syms = [syn.random_symbols(st, case["conc"], seed=60 + r) for r in range(3)]
        gpu, chains = make_pair(ft, syms, [0.02, 0.05, 0.2], seed=909)
        assert gpu.batch_kernel_applies() and gpu.get_batch_eval() == 2
        gpu.set_variant(variant, variant)
        n = 1200
        gpu.set_trace(n)
        gpu.reset_accumulators()
        (gpu.run_sgc if mode == "sgc" else gpu.run_canonical)(n)
        assert gpu.last_variant() == variant
        tr = gpu.get_trace(n)
        for r, c in enumerate(chains):
            o = c.run_sgc(n, trace=True) if mode == "sgc" else c.run_canonical(n, trace=True)
            assert np.array_equal(tr[0][r], o[0]) and np.array_equal(tr[3][r], o[3])
            assert np.array_equal(tr[4][r], o[4])
        assert_state_equal(gpu, chains)
        assert np.array_equal(gpu.get_accumulators(), np.stack([c.acc for c in chains]))
        # incremental CFs equal the definition (every cluster is listed in each member's table)
        cf_run = gpu.get_cf()
        gpu.recompute_cf()
        np.testing.assert_allclose(gpu.get_cf(), cf_run, rtol=0, atol=1e-13)


@pytest.mark.parametrize("case", [BINARY, TERNARY, LAYERED, LAYERED_BINARY,
                                  dict(L=5, species=["Al", "Mg", "Si"], families=["nn", "2nn", "3nn", "tri", "iso", "tet"],
                                       conc={"Al": 0.5, "Mg": 0.25, "Si": 0.25})])
def test_recompute_cf_table_kernel(cuda_device, case):
    """cemc_recompute_cf with the product tables (codes per sub-cluster + table sums) against the
    oracle's brute-force definition and against the item-by-item kernel (generic path)."""
    st, eci, symbols, ft = build(**case)
    R = 5
    syms = [syn.random_symbols(st, case["conc"], seed=20 + r) for r in range(R)]
    occ = np.stack([ft.occupancy(s) for s in syms])
    want = np.stack([OracleChain(ft, occ[r]).cf for r in range(R)])
    out = []
    for generic in (False, True):
        gpu = BatchedCEUpdater(ft, R)
        gpu.set_generic_path(generic)
        gpu.set_occupancy(occ)
        gpu.recompute_cf()
        cf = gpu.get_cf()
        np.testing.assert_allclose(cf, want, rtol=0, atol=2e-14)
        gpu.recompute_cf()
        assert np.array_equal(gpu.get_cf(), cf)          # reproducible bits
        out.append(cf)
    np.testing.assert_allclose(out[0], out[1], rtol=0, atol=2e-14)
