"""CPU tests: the oracle against the reference's golden vectors and against
the reference's own compiled CEUpdater (when oracle/_ref is present), the
host-side table logic, and the C-ABI symbol table (no compute calls)."""
import ctypes
import os
import re

import numpy as np
import pytest

from cases import (BINARY, GOLDEN, GOLDEN_WORKLOADS, LAYERED, LAYERED_BINARY, TERNARY, build, load_golden,
                   load_golden_workload)
from cemc_b200 import synthetic as syn
from cemc_b200.tables import FlatTables, SelfInteractionError
from oracle import ce_oracle, ref_driver
from oracle.ce_oracle import OracleChain

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("name", GOLDEN)
def test_oracle_matches_golden(name):
    """Fixtures were produced by the reference's compiled C++ updater
    (tests/golden/make_golden.py): the oracle must agree bit for bit."""
    meta, st, ft, z = load_golden(name)
    oc = OracleChain(ft, ft.occupancy(meta["symbols0"]), cf=z["cf0"], kT=meta["kT"])
    assert oc.e == float(z["e0"])
    acc, e_after = oc.replay(z["sites"], z["news"], z["u"])
    assert np.array_equal(acc, z["accepted"])
    assert np.array_equal(e_after, z["e_after"])
    assert np.array_equal(oc.cf, z["cf_final"])
    assert np.array_equal(oc.occ, z["occ_final"])


@pytest.mark.parametrize("name", GOLDEN_WORKLOADS)
def test_oracle_matches_golden_baseline_sizes(name):
    """BASELINE-size fixtures (replicas of the bench workloads, recorded from the compiled
    reference): the oracle agrees bit for bit, and its Philox chain regenerates the proposals."""
    meta, ft, z = load_golden_workload(name)
    for r, g in enumerate(meta["replicas"]):
        oc = OracleChain(ft, z["occ0"][r], cf=z["cf0"][r], kT=float(z["kT"][r]), eci=z["eci"][r],
                         seed=2024, replica=g)
        assert oc.e == float(z["e0"][r])
        acc, e_after = oc.replay(z["sites"][r], z["news"][r], z["u"][r])
        assert np.array_equal(acc, z["accepted"][r])
        assert np.array_equal(e_after, z["e_after"][r])
        assert np.array_equal(oc.cf, z["cf_final"][r])
        assert np.array_equal(oc.occ, z["occ_final"][r])
        oc2 = OracleChain(ft, z["occ0"][r], cf=z["cf0"][r], kT=float(z["kT"][r]), eci=z["eci"][r],
                          seed=2024, replica=g)
        n = z["u"].shape[1]
        tr = oc2.run_canonical(n, trace=True) if meta["mode"] == "canonical" else oc2.run_sgc(n, trace=True)
        assert np.array_equal(tr[0], z["sites"][r]) and np.array_equal(tr[2], z["u"][r])
        assert np.array_equal(tr[3], z["accepted"][r])


@pytest.mark.parametrize("name", GOLDEN[:2])
def test_oracle_proposals_reproduce_golden(name):
    """The golden proposals were drawn by the oracle's Philox chain (seed 2024,
    replica 5): regenerating them pins the proposal generator too."""
    meta, st, ft, z = load_golden(name)
    oc = OracleChain(ft, ft.occupancy(meta["symbols0"]), cf=z["cf0"],
                     kT=meta["kT"], seed=2024, replica=5)
    n = len(z["u"])
    tr = oc.run_canonical(n, trace=True) if meta["mode"] == "canonical" \
        else oc.run_sgc(n, trace=True)
    assert np.array_equal(tr[0], z["sites"])
    assert np.array_equal(tr[1], z["news"])
    assert np.array_equal(tr[2], z["u"])
    assert np.array_equal(tr[3], z["accepted"])


@pytest.mark.skipif(not ref_driver.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("case,mode", [(BINARY, "sgc"), (BINARY, "canonical"),
                                       (TERNARY, "sgc"), (TERNARY, "canonical"),
                                       (LAYERED, "sgc"), (LAYERED, "canonical"),
                                       (LAYERED_BINARY, "sgc"), (LAYERED_BINARY, "canonical")])
def test_oracle_vs_compiled_reference(case, mode):
    st, eci, symbols, ft = build(**case)
    oc = OracleChain(ft, ft.occupancy(symbols), kT=0.04, seed=99, replica=1)
    cf0 = {n: float(v) for n, v in zip(ft.eci_names, oc.cf)}
    rc = ref_driver.RefChain(st, symbols, eci, cf0, kT=0.04)
    assert rc.current_energy == oc.e
    tr = oc.run_sgc(400, trace=True) if mode == "sgc" else oc.run_canonical(400, trace=True)
    acc, used, e = rc.replay(ft.species, tr[0], tr[1], tr[2])
    assert np.array_equal(acc, tr[3])
    assert np.array_equal(e, tr[4])
    assert np.array_equal(rc.cf_vector(), oc.cf)
    assert rc.symbols() == ft.symbols_of(oc.occ)


@pytest.mark.skipif(not ref_driver.available(), reason="oracle/_ref not built")
def test_reference_list_of_dict_translation_matrix():
    """Both trans_matrix formats the reference accepts give the same tables."""
    st_a, eci, symbols, ft_a = build(**BINARY)
    st_l, _, _, ft_l = build(trans_matrix_format="list", **BINARY)
    assert np.array_equal(ft_a.trans, ft_l.trans)
    oc = OracleChain(ft_l, ft_l.occupancy(symbols), kT=0.05, seed=1)
    cf0 = {n: float(v) for n, v in zip(ft_l.eci_names, oc.cf)}
    rc = ref_driver.RefChain(st_l, symbols, eci, cf0, kT=0.05)
    tr = oc.run_sgc(100, trace=True)
    acc, _, e = rc.replay(ft_l.species, tr[0], tr[1], tr[2])
    assert np.array_equal(acc, tr[3]) and np.array_equal(e, tr[4])


@pytest.mark.parametrize("case", [BINARY, TERNARY])
def test_incremental_matches_brute_force(case):
    """Same check as the reference's tests/test_CE_updater.py:40-115 (its oracle
    is ase.clease CorrFunction; ours is the CF definition)."""
    st, eci, symbols, ft = build(**case)
    oc = OracleChain(ft, ft.occupancy(symbols), kT=0.05, seed=4)
    oc.run_canonical(300)
    oc.run_sgc(300)
    inc = oc.cf.copy()
    oc.recompute_cf()
    np.testing.assert_allclose(inc, oc.cf, rtol=0, atol=1e-13)
    assert oc.cf[ft.eci_index["c0"]] == 1.0


def test_equivalent_deco():
    assert syn.equivalent_deco([0, 1], []) == [[0, 1]]
    assert syn.equivalent_deco([0, 1], [[0, 1]]) == [[0, 1], [1, 0]]
    assert syn.equivalent_deco([0, 0, 1], [[0, 1, 2]]) == [[0, 0, 1], [0, 1, 0], [1, 0, 0]]
    assert syn.equivalent_deco([1, 0, 1], [[1, 2]]) == [[1, 0, 1], [1, 1, 0]]
    assert len(syn.equivalent_deco([0, 0, 1, 1], [[0, 1, 2, 3]])) == 6


def test_tables_shapes_and_names():
    st, eci, symbols, ft = build(**TERNARY)
    assert ft.eci_names == sorted(eci.keys())
    assert ft.K == 18 and ft.S == 3 and ft.D == 2
    assert ft.singlet_names == ["c1_0", "c1_1"]
    assert ft.gathered_sites_per_change() == 12 + 6 + 48 + 72 + 24
    st2, eci2, symbols2, ft2 = build(**BINARY)
    # standard binary set is 258 B/move (SGC), K=18, G=90
    assert ft2.algorithmic_bytes_per_move(1) == 258
    with pytest.raises(ValueError):
        ft.eci_vector({"c0": 1.0})


def test_self_interaction_rejected():
    with pytest.raises(ValueError):
        syn.fcc_settings(2, ["Al", "Mg"])
    st = syn.fcc_settings(4, ["Al", "Mg"])
    fam = next(iter(st.cluster_info[0].values()))
    fam["indices"][0][0] = 0
    with pytest.raises(SelfInteractionError):
        FlatTables(st, syn.synthetic_ecis(st), ["Al"] * 64)


def test_philox_known_answer():
    """Random123 known-answer vectors for Philox4x32-10."""
    lib = ce_oracle._lib()
    out = (ctypes.c_uint32 * 4)()
    # counter = 0, key = 0
    lib.oracle_philox(ctypes.c_uint64(0), ctypes.c_uint64(0), 0, 0, out)
    assert [hex(x) for x in out] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    # counter = ff.., key = ff..
    lib.oracle_philox(ctypes.c_uint64(2**64 - 1), ctypes.c_uint64(2**64 - 1),
                      0xffffffff, 0xffffffff, out)
    assert [hex(x) for x in out] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]


def test_pt_exchange_oracle_properties():
    rng = np.random.default_rng(0)
    n = 9
    kts = np.geomspace(0.1, 0.01, n)
    e = rng.normal(size=n)
    slots = np.arange(n, dtype=np.int32)
    for rnd in range(20):
        new_slots, n_acc = ce_oracle.pt_exchange(e, slots, kts, rnd % 2, 7, rnd)
        assert sorted(new_slots.tolist()) == list(range(n))
        moved = np.nonzero(new_slots != slots)[0]
        assert len(moved) == 2 * n_acc
        slots = new_slots
    # equal energies: p = exp(0) = 1 and u < 1 always -> every pair swaps
    new_slots, n_acc = ce_oracle.pt_exchange(np.zeros(n), np.arange(n), kts, 0, 7, 0)
    assert n_acc == n // 2


def test_c_abi_exports_every_declared_symbol():
    """The shared library loads and exports everything include/*.h declares."""
    from cemc_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "cemc_b200.h")).read()
    declared = set(re.findall(r"\b(cemc_[a-z_A-Z0-9]+)\s*\(", hdr))
    declared -= {"cemc_tables", "cemc_handle"}
    assert len(declared) > 30
    lib = _lib.load()
    for name in sorted(declared):
        assert hasattr(lib, name), name
    assert set(_lib.SIGNATURES) | {"cemc_last_error", "cemc_version"} == declared
    assert lib.cemc_version() >= 100


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under cemc_b200/ may use it."""
    pkg = os.path.join(ROOT, "cemc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("oracle/", "").lower() or \
                    "import oracle" not in txt and "from oracle" not in txt
                assert "from oracle" not in txt and "import oracle" not in txt
                assert "ce_oracle" not in txt


def test_ce_wrapper_pure_helpers():
    """Host helpers of the CE wrapper that need no device: largest cluster size named in the ECIs
    (ce_calculator.py:120-133) and the self-interaction check (:596-611)."""
    from cemc_b200.ce_calculator import get_max_size_eci, _clusters_overlap
    assert get_max_size_eci({"c0": 0.0, "c1_0": 0.1, "c2_d0000_0_00": 0.2, "c4_d0000_0_0000": 0.3}) == 4
    assert get_max_size_eci({}) == 0
    ok = [{"c2_a": {"ref_indx": 0, "indices": [[1], [2]]}, "c3_a": {"ref_indx": 0, "indices": [[1, 2], [3, 4]]}}]
    assert not _clusters_overlap(ok)
    own_site = [{"c2_a": {"ref_indx": 0, "indices": [[1], [0]]}}]                 # the reference site in its own sub-cluster
    twice = [{"c3_a": {"ref_indx": 0, "indices": [[1, 2], [3, 3]]}}]              # the same site twice
    assert _clusters_overlap(own_site) and _clusters_overlap(twice)
