"""Generate the golden replay fixtures with the reference's OWN compiled
CEUpdater (oracle/_ref, built by oracle/build_ref.sh from /root/reference).

Run here (the GPU box has no /root/reference):  python tests/golden/make_golden.py

Each fixture is one Appendix-D replay record set: proposals and
uniforms (drawn by the oracle's Philox chain so that acceptance rates are
realistic), and what the reference did with them: accepted flags, whether the
uniform was consumed, the energy after every step, final CFs and symbols.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from cemc_b200 import synthetic as syn          # noqa: E402
from cemc_b200.tables import FlatTables         # noqa: E402
from oracle import ref_driver                   # noqa: E402
from oracle.ce_oracle import OracleChain        # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = [
    # name, L, species, families, conc, eci kind, kT, mode, steps
    ("almg_fcc4_canonical", 4, ["Al", "Mg"], ["nn", "2nn", "3nn", "tri", "iso", "tet"],
     {"Al": 0.5, "Mg": 0.5}, "almg", 0.0430866, "canonical", 1500),   # config 1, T=500 K
    ("almg_fcc4_sgc", 4, ["Al", "Mg"], ["nn", "2nn", "tri", "tet"],
     {"Al": 0.7, "Mg": 0.3}, "synthetic", 0.03, "sgc", 1500),
    ("almgsi_fcc4_canonical", 4, ["Al", "Mg", "Si"], ["nn", "2nn", "tri", "iso", "tet"],
     {"Al": 0.5, "Mg": 0.25, "Si": 0.25}, "synthetic", 0.04, "canonical", 1200),
    ("almgsi_fcc4_sgc", 4, ["Al", "Mg", "Si"], ["nn", "2nn", "tri", "iso", "tet"],
     {"Al": 0.5, "Mg": 0.25, "Si": 0.25}, "synthetic", 0.05, "sgc", 1200),
    ("almgsi_fcc5_canonical_cold", 5, ["Al", "Mg", "Si"], ["nn", "2nn", "tri", "tet"],
     {"Al": 0.8, "Mg": 0.1, "Si": 0.1}, "synthetic", 0.008, "canonical", 1200),
    # two translational symmetry groups (cemc_b200.synthetic.layered_settings; ce_updater.cpp:379-384)
    ("almgsi_layered4_sgc", 4, ["Al", "Mg", "Si"], "layered",
     {"Al": 0.4, "Mg": 0.3, "Si": 0.3}, "synthetic", 0.05, "sgc", 1200),
    ("almgsi_layered6_canonical", 6, ["Al", "Mg", "Si"], "layered",
     {"Al": 0.4, "Mg": 0.3, "Si": 0.3}, "synthetic", 0.04, "canonical", 1200),
]


def build_case(L, species, families, conc, eci_kind, seed=3):
    st = syn.layered_settings(L, species) if families == "layered" else syn.fcc_settings(L, species, families)
    eci = syn.almg_ecis(st) if eci_kind == "almg" else syn.synthetic_ecis(st, seed=1234)
    symbols = syn.random_symbols(st, conc, seed=seed)
    ft = FlatTables(st, eci, symbols)
    return st, eci, symbols, ft


# BASELINE-size fixtures: replicas of the bench workloads themselves (cemc_b200.workloads), a few
# thousand recorded steps each.  name, workload, global replica ids, steps
WORKLOAD_CASES = [
    ("c2_fcc10_sgc", "C2", [0, 85, 170, 255], 2500),       # BASELINE configs[1]: corners + middle of the mu x T grid
    ("c3s_fcc20_sgc", "C3S", [0, 31, 63], 2500),           # north-star target line (64-replica Al-Mg-Si SGC sweep)
    ("c3_fcc20_canonical", "C3", [0, 31, 63], 2500),       # BASELINE configs[2]
]


def workload_replica(which, g):
    """Replica g of a bench workload, plus the settings object in a form the reference reads."""
    from cemc_b200 import workloads as wl
    w = wl.WORKLOADS[which](R=1, replica_offset=g)
    st = w.settings
    if st.trans_matrix_columns is not None:     # compact form: the reference needs list-of-dicts
        kw = st.kwargs
        st = syn.fcc_settings(kw["size"][0], kw["species"], kw["families"], trans_matrix_format="list")
    return w, st


def make_workload_case(name, which, replicas, steps):
    rows = dict(sites=[], news=[], u=[], accepted=[], u_used=[], e_after=[], cf0=[], e0=[],
                cf_final=[], occ0=[], occ_final=[], kT=[], eci=[])
    for g in replicas:
        w, st = workload_replica(which, g)
        ft = w.tables
        eci_vec = w.eci_matrix[0] if w.eci_matrix is not None else ft.eci
        kT = float(w.kT[0])
        oc = OracleChain(ft, w.occ[0], kT=kT, seed=2024, replica=g, eci=eci_vec)
        cf0 = oc.cf.copy()
        tr = oc.run_canonical(steps, trace=True) if w.mode == "canonical" else oc.run_sgc(steps, trace=True)
        eci = {n: float(v) for n, v in zip(ft.eci_names, eci_vec)}
        rc = ref_driver.RefChain(st, ft.symbols_of(w.occ[0]), eci,
                                 {n: float(v) for n, v in zip(ft.eci_names, cf0)}, kT=kT)
        e0 = rc.current_energy
        acc, used, e_after = rc.replay(ft.species, tr[0], tr[1], tr[2])
        for k, v in (("sites", tr[0]), ("news", tr[1]), ("u", tr[2]), ("accepted", acc), ("u_used", used),
                     ("e_after", e_after), ("cf0", cf0), ("e0", e0), ("cf_final", rc.cf_vector()),
                     ("occ0", w.occ[0]), ("occ_final", ft.occupancy(rc.symbols())), ("kT", kT),
                     ("eci", eci_vec)):
            rows[k].append(v)
    out = os.path.join(HERE, name + ".npz")
    meta = dict(workload=which, replicas=list(replicas), steps=steps, eci_names=ft.eci_names,
                species_sorted=ft.species)
    np.savez_compressed(out, meta=json.dumps(meta), **{k: np.stack(v) for k, v in rows.items()})
    print("{}: {} replicas x {} steps, accept rates {}, {} bytes".format(
        name, len(replicas), steps, [round(float(a.mean()), 3) for a in rows["accepted"]],
        os.path.getsize(out)))


def main():
    only = sys.argv[1:]
    for name, which, replicas, steps in WORKLOAD_CASES:
        if not only or name in only:
            make_workload_case(name, which, replicas, steps)
    for name, L, species, fams, conc, eci_kind, kT, mode, steps in CASES:
        if only and name not in only:
            continue
        st, eci, symbols, ft = build_case(L, species, fams, conc, eci_kind)
        oc = OracleChain(ft, ft.occupancy(symbols), kT=kT, seed=2024, replica=5)
        cf0 = oc.cf.copy()
        tr = oc.run_canonical(steps, trace=True) if mode == "canonical" \
            else oc.run_sgc(steps, trace=True)
        sites, news, u = tr[0], tr[1], tr[2]
        rc = ref_driver.RefChain(st, symbols, eci,
                                 {n: float(v) for n, v in zip(ft.eci_names, cf0)}, kT=kT)
        e0 = rc.current_energy
        acc, used, e_after = rc.replay(ft.species, sites, news, u)
        out = os.path.join(HERE, name + ".npz")
        meta = dict(L=L, species=species, families=fams, conc=conc, eci_kind=eci_kind,
                    kT=kT, mode=mode, eci=eci, eci_names=ft.eci_names,
                    symbols0=symbols, species_sorted=ft.species)
        np.savez_compressed(
            out, meta=json.dumps(meta), cf0=cf0, e0=np.float64(e0), sites=sites, news=news,
            u=u, accepted=acc, u_used=used, e_after=e_after, cf_final=rc.cf_vector(),
            occ_final=ft.occupancy(rc.symbols()))
        print("{}: {} steps, accept rate {:.3f}, uniforms used {:.3f}, {} bytes".format(
            name, steps, acc.mean(), used.mean(), os.path.getsize(out)))


if __name__ == "__main__":
    main()
