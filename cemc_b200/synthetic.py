"""Synthetic cluster-expansion problems (no ASE / ase.clease needed).

The reference takes its lattice tables from ``ase.clease`` settings objects
(``CEBulk``), which are not vendored and not installable here.  This module
emits objects with exactly the attributes ``CEUpdater::init`` reads
(/root/reference/cpp/src/ce_updater.cpp:32-234):

    unique_elements, num_unique_elements, index_by_trans_symm,
    background_indices, cluster_info, basis_functions, trans_matrix

so that the SAME object can be handed to the compiled reference
(``oracle/_ref``), to the oracle restatement and to the CUDA path.

Lattice: fcc primitive cell replicated L x L x L, site index
``(i*L + j)*L + k``.  Cluster families are enumerated geometrically from the
origin's neighbour shells; vertex order inside a figure is canonicalised by a
distance signature, which yields the ``order`` / ``equiv_sites`` tables in the
shape ``Cluster::parse_info_dict`` expects (cpp/src/cluster.cpp:185-232).
"""
from __future__ import annotations

import itertools
import math
from typing import Dict, List, Sequence

import numpy as np

# fcc primitive vectors in units of a/2
_FCC_PRIM = np.array([[0, 1, 1], [1, 0, 1], [1, 1, 0]], dtype=np.int64)


class Atom(object):
    """Minimal stand-in for ase.Atom: the updater only needs ``.symbol``."""
    __slots__ = ("symbol", "index")

    def __init__(self, symbol, index):
        self.symbol = symbol
        self.index = index

    def __repr__(self):
        return "Atom({!r}, {})".format(self.symbol, self.index)


class Atoms(object):
    """Minimal stand-in for ase.Atoms (sequence of Atom + calculator slot)."""

    def __init__(self, symbols: Sequence[str]):
        self._atoms = [Atom(s, i) for i, s in enumerate(symbols)]
        self._calc = None

    def __len__(self):
        return len(self._atoms)

    def __getitem__(self, i):
        return self._atoms[i]

    def __iter__(self):
        return iter(self._atoms)

    def get_calculator(self):
        return self._calc

    def set_calculator(self, calc):
        self._calc = calc

    def get_chemical_symbols(self):
        return [a.symbol for a in self._atoms]

    def copy(self):
        return Atoms(self.get_chemical_symbols())


def equivalent_deco(deco, equiv_sites):
    """Equivalent decoration numbers.

    Restatement of ``ase.clease.tools.equivalent_deco`` (third-party, not in
    /root/reference, no version pinned; called from
    /root/reference/cpp/src/cluster.cpp:78-118).  For every group of
    equivalent vertex positions, all permutations of the decoration numbers on
    those positions, de-duplicated in order of first appearance
    (``itertools.permutations`` x ``itertools.product`` order).
    """
    if not equiv_sites:
        return [list(deco)]
    perms = [list(itertools.permutations(grp)) for grp in equiv_sites]
    out = []
    for comb in itertools.product(*perms):
        order = []
        for item in comb:
            order += list(item)
        orig = list(range(len(deco)))
        for i, srt in enumerate(sorted(order)):
            orig[srt] = order[i]
        out.append([deco[j] for j in orig])
    unique = []
    for d in out:
        if d not in unique:
            unique.append(d)
    return unique


def basis_functions_for(species: Sequence[str]) -> List[Dict[str, float]]:
    """Orthonormal polynomial site basis (the clease 'polynomial' flavour).

    binary  : {+1, -1}
    ternary : sqrt(3/2)*s, sqrt(2)*(1 - 3/2 s^2) on s in {-1, 0, 1}
    """
    S = len(species)
    if S == 2:
        return [{species[0]: 1.0, species[1]: -1.0}]
    if S == 3:
        sig = [-1.0, 0.0, 1.0]
        b0 = {sp: math.sqrt(1.5) * s for sp, s in zip(species, sig)}
        b1 = {sp: math.sqrt(2.0) * (1.0 - 1.5 * s * s)
              for sp, s in zip(species, sig)}
        return [b0, b1]
    # generic: Gram-Schmidt on monomials of equally spaced spins
    sig = np.linspace(-1.0, 1.0, S)
    V = np.vander(sig, S, increasing=True)
    q, _ = np.linalg.qr(V)
    q = q * math.sqrt(S)
    return [{sp: float(q[i, d + 1]) for i, sp in enumerate(species)}
            for d in range(S - 1)]


class SyntheticSettings(object):
    """Duck-typed ``ClusterExpansionSetting`` (see module docstring)."""

    def __init__(self):
        self.unique_elements = []
        self.num_unique_elements = 0
        self.index_by_trans_symm = []
        self.background_indices = []
        self.cluster_info = []
        self.basis_functions = []
        self.trans_matrix = None
        self.trans_matrix_columns = None   # compact form only (see fcc_settings)
        self.atoms = None
        self.size = None
        self.max_cluster_dia = 0.0
        self.kwargs = {}

    def _info_entries_to_list(self):  # called by cemc/ce_calculator.py:168
        pass

    # ------------------------------------------------------------------
    def eci_names(self) -> List[str]:
        """All symmetry-distinct CF/ECI names for these families."""
        D = self.num_unique_elements - 1
        names = ["c0"] + ["c1_{}".format(d) for d in range(D)]
        seen = set()
        for info in self.cluster_info:
            for prefix, fam in info.items():
                if prefix in seen:
                    continue
                seen.add(prefix)
                n = fam["size"]
                done = []
                for deco in itertools.product(range(D), repeat=n):
                    eq = equivalent_deco(list(deco), fam["equiv_sites"])
                    key = min(tuple(e) for e in eq)
                    if key in done:
                        continue
                    done.append(key)
                    names.append(prefix + "_" + "".join(str(x) for x in key))
        return sorted(names)


def _sig(points, v):
    return tuple(sorted(int(((points[v] - points[w]) ** 2).sum())
                        for w in range(len(points)) if w != v))


def _canonical_order(points):
    """Return (order, equiv_sites): ``sorted[k] = raw[order[k]]``."""
    n = len(points)
    sigs = [_sig(points, v) for v in range(n)]
    order = sorted(range(n), key=lambda v: (sigs[v], v))
    equiv = []
    k = 0
    while k < n:
        j = k
        while j + 1 < n and sigs[order[j + 1]] == sigs[order[k]]:
            j += 1
        if j > k:
            equiv.append(list(range(k, j + 1)))
        k = j + 1
    return order, equiv


# named figures: (size, sorted tuple of squared pair distances in (a/2)^2)
FIGURES = {
    "nn": (2, (2,)),
    "2nn": (2, (4,)),
    "3nn": (2, (6,)),
    "tri": (3, (2, 2, 2)),          # NN equilateral triangle
    "iso": (3, (2, 2, 4)),          # right-isosceles: two NN legs + 2NN
    "tet": (4, (2, 2, 2, 2, 2, 2)),  # regular NN tetrahedron
}
STANDARD_FAMILIES = ("nn", "2nn", "tri", "tet")


def fcc_settings(L: int, species: Sequence[str] = ("Al", "Mg"),
                 families: Sequence[str] = STANDARD_FAMILIES,
                 trans_matrix_format: str = "auto") -> SyntheticSettings:
    """fcc primitive L^3 cell with the requested cluster families."""
    if L < 3:
        raise ValueError("L >= 3 required (self interaction otherwise)")
    N = L ** 3
    species = list(species)
    st = SyntheticSettings()
    st.unique_elements = sorted(species)
    st.num_unique_elements = len(species)
    st.index_by_trans_symm = [list(range(N))]
    st.background_indices = []
    st.basis_functions = basis_functions_for(st.unique_elements)
    st.size = [L, L, L]
    st.kwargs = dict(crystalstructure="fcc", size=[L, L, L],
                     species=list(species), families=list(families))

    # candidate neighbour offsets (primitive coords) within the 3NN shell
    rng = range(-2, 3)
    offs = []
    for o in itertools.product(rng, rng, rng):
        if o == (0, 0, 0):
            continue
        cart = np.array(o) @ _FCC_PRIM
        d2 = int((cart ** 2).sum())
        if d2 <= 6:
            offs.append((o, cart, d2))

    def col_of(o):
        return ((o[0] % L) * L + (o[1] % L)) * L + (o[2] % L)

    info = {}
    shell_counter = {}
    max_dia = 0.0
    for fam_name in families:
        n, dist_key = FIGURES[fam_name]
        subs = []
        for combo in itertools.combinations(range(len(offs)), n - 1):
            pts = [np.zeros(3, dtype=np.int64)] + [offs[c][1] for c in combo]
            d2s = tuple(sorted(int(((pts[a] - pts[b]) ** 2).sum())
                               for a in range(n) for b in range(a + 1, n)))
            if d2s != dist_key:
                continue
            cols = [col_of(offs[c][0]) for c in combo]
            # deterministic member order: ascending column id
            perm = sorted(range(n - 1), key=lambda q: cols[q])
            cols = [cols[q] for q in perm]
            pts = [pts[0]] + [pts[1 + q] for q in perm]
            order, equiv = _canonical_order(pts)
            subs.append((cols, order, equiv))
        if not subs:
            raise ValueError("no figures for family " + fam_name)
        subs.sort(key=lambda t: t[0])
        equiv = subs[0][2]
        for s in subs:
            assert s[2] == equiv
            if 0 in s[0] or len(set(s[0])) != len(s[0]):
                raise ValueError("self interaction: cell too small")
        idx = shell_counter.get(n, 0)
        shell_counter[n] = idx + 1
        prefix = "c{}_d{:04d}_0".format(n, idx)
        dia = math.sqrt(max(dist_key)) * 0.5
        max_dia = max(max_dia, dia)
        info[prefix] = {
            "ref_indx": 0,
            "size": n,
            "max_cluster_dia": float(dia),
            "symm_group": 0,
            "name": prefix,
            "descriptor": fam_name,
            "indices": [list(s[0]) for s in subs],
            "order": [list(s[1]) for s in subs],
            "equiv_sites": [list(g) for g in equiv],
        }
    st.cluster_info = [info]
    st.max_cluster_dia = max_dia

    # translation matrix T[s][col] = site(s + col) (periodic)
    used_cols = sorted({c for f in info.values() for sub in f["indices"]
                        for c in sub})
    ijk = np.stack(np.unravel_index(np.arange(N), (L, L, L)), axis=1)
    if trans_matrix_format == "auto":
        # a dense [N, max_col + 1] table (the reference's ndarray form) is O(N^2) memory
        trans_matrix_format = "ndarray" if N <= 4096 else "compact"
    if trans_matrix_format == "compact":
        # only the used columns: trans_matrix[:, k] belongs to column id
        # trans_matrix_columns[k] (extension understood by cemc_b200.tables.FlatTables;
        # the reference needs the dense or the list-of-dict form)
        tm = np.zeros((N, len(used_cols)), dtype=np.int32)
        for k, c in enumerate(used_cols):
            cijk = np.array(np.unravel_index(c, (L, L, L)))
            t = (ijk + cijk[None, :]) % L
            tm[:, k] = (t[:, 0] * L + t[:, 1]) * L + t[:, 2]
        st.trans_matrix = tm
        st.trans_matrix_columns = list(used_cols)
    elif trans_matrix_format == "ndarray":
        ncol = max(used_cols) + 1
        tm = np.zeros((N, ncol), dtype=np.int32)
        for c in used_cols:
            cijk = np.array(np.unravel_index(c, (L, L, L)))
            t = (ijk + cijk[None, :]) % L
            tm[:, c] = (t[:, 0] * L + t[:, 1]) * L + t[:, 2]
        st.trans_matrix = tm
    elif trans_matrix_format == "list":
        colmap = {}
        for c in used_cols:
            cijk = np.array(np.unravel_index(c, (L, L, L)))
            t = (ijk + cijk[None, :]) % L
            colmap[c] = ((t[:, 0] * L + t[:, 1]) * L + t[:, 2]).tolist()
        st.trans_matrix = [{c: colmap[c][s] for c in used_cols}
                           for s in range(N)]
    else:
        raise ValueError(trans_matrix_format)
    st.atoms = Atoms([st.unique_elements[0]] * N)
    return st


def layered_settings(L: int, species: Sequence[str] = ("Al", "Mg", "Si"),
                     trans_matrix_format: str = "ndarray") -> SyntheticSettings:
    """Two-sublattice crystal: simple cubic L^3 cell whose (001) layers alternate between two
    inequivalent site types A (k even) and B (k odd) -- TWO translational symmetry groups
    (``index_by_trans_symm``), each with its own ``cluster_info`` dict, the way ase.clease
    describes a crystal with a basis (the reference selects the family table by the changed
    site's group: cpp/src/ce_updater.cpp:379-384; 4-sublattice case in
    tests/test_CE_updater.py:206-228).  Families (every physical cluster is listed once in
    the table of EACH of its member sites, so incremental updates and the definition agree):

      c2_d0000_0  A-A in-plane nearest neighbours     group 0 only, M = 4, equivalent vertices
      c2_d0001_0  B-B in-plane nearest neighbours     group 1 only, M = 4, equivalent vertices
      c2_d0002_0  A-B inter-layer nearest neighbours  both groups, M = 2; sorted order (A, B): the
                                                      changed site is position 0 (A) or 1 (B)
      c3_d0000_0  right-angle triplet (A corner, A', B above the corner): three inequivalent
                  vertices, sorted (corner, far A, B); M = 16 seen from an A site (8 as the corner,
                  8 as the far A: ``order`` [0,1,2] / [1,0,2]) and M = 8 from a B site ([1,2,0])
    """
    if L < 4 or L % 2:
        raise ValueError("L must be even and >= 4")
    N = L ** 3
    species = list(species)
    st = SyntheticSettings()
    st.unique_elements = sorted(species)
    st.num_unique_elements = len(species)
    st.background_indices = []
    st.basis_functions = basis_functions_for(st.unique_elements)
    st.size = [L, L, L]
    st.kwargs = dict(crystalstructure="layered_sc", size=[L, L, L], species=list(species))
    ijk = np.stack(np.unravel_index(np.arange(N), (L, L, L)), axis=1)
    st.index_by_trans_symm = [[int(s) for s in range(N) if ijk[s, 2] % 2 == 0],
                              [int(s) for s in range(N) if ijk[s, 2] % 2 == 1]]

    def col(o):
        return ((o[0] % L) * L + (o[1] % L)) * L + (o[2] % L)

    inplane = [(1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0)]
    updown = [(0, 0, 1), (0, 0, -1)]

    def fam(prefix, n, g, indices, order, equiv, dia, descr):
        return {"ref_indx": 0, "size": n, "max_cluster_dia": float(dia), "symm_group": g,
                "name": prefix, "descriptor": descr, "indices": indices, "order": order,
                "equiv_sites": equiv}

    add = lambda a, b: (a[0] + b[0], a[1] + b[1], a[2] + b[2])      # noqa: E731
    info0, info1 = {}, {}
    info0["c2_d0000_0"] = fam("c2_d0000_0", 2, 0, [[col(o)] for o in inplane], [[0, 1]] * 4, [[0, 1]], 0.5, "AA")
    info1["c2_d0001_0"] = fam("c2_d0001_0", 2, 1, [[col(o)] for o in inplane], [[0, 1]] * 4, [[0, 1]], 0.5, "BB")
    info0["c2_d0002_0"] = fam("c2_d0002_0", 2, 0, [[col(o)] for o in updown], [[0, 1]] * 2, [], 0.5, "AB")
    info1["c2_d0002_0"] = fam("c2_d0002_0", 2, 1, [[col(o)] for o in updown], [[1, 0]] * 2, [], 0.5, "AB")
    # triplet seen from an A site: as the corner -> raw (ref, A', B) is already sorted; as the far A ->
    # raw (ref, corner, B above the corner), sorted (corner, ref, B) = raw[1], raw[0], raw[2]
    ind, order = [], []
    for a in inplane:
        for b in updown:
            ind.append([col(a), col(b)]); order.append([0, 1, 2])
    for a in inplane:
        for b in updown:
            ind.append([col(a), col(add(a, b))]); order.append([1, 0, 2])
    info0["c3_d0000_0"] = fam("c3_d0000_0", 3, 0, ind, order, [], 0.7071, "AAB")
    # seen from the B site: raw (ref, corner, far A), sorted (corner, far, B) = raw[1], raw[2], raw[0]
    ind, order = [], []
    for b in updown:
        for a in inplane:
            ind.append([col(b), col(add(b, a))]); order.append([1, 2, 0])
    info1["c3_d0000_0"] = fam("c3_d0000_0", 3, 1, ind, order, [], 0.7071, "AAB")
    st.cluster_info = [info0, info1]
    st.max_cluster_dia = 0.7071
    used = sorted({c for info in st.cluster_info for f in info.values() for sub in f["indices"] for c in sub})
    if any(c == 0 for c in used):
        raise ValueError("self interaction: cell too small")

    def shifted(c):
        cijk = np.array(np.unravel_index(c, (L, L, L)))
        t = (ijk + cijk[None, :]) % L
        return (t[:, 0] * L + t[:, 1]) * L + t[:, 2]
    if trans_matrix_format == "ndarray":
        tm = np.zeros((N, max(used) + 1), dtype=np.int32)
        for c in used:
            tm[:, c] = shifted(c)
        st.trans_matrix = tm
    elif trans_matrix_format == "list":
        cm = {c: shifted(c).tolist() for c in used}
        st.trans_matrix = [{c: cm[c][s] for c in used} for s in range(N)]
    else:
        raise ValueError(trans_matrix_format)
    st.atoms = Atoms([st.unique_elements[0]] * N)
    return st


def random_symbols(settings: SyntheticSettings, conc: Dict[str, float],
                   seed: int, exact: bool = True) -> List[str]:
    """Random occupation at the stated composition (default_rng(seed))."""
    N = len(settings.atoms)
    rng = np.random.default_rng(seed)
    names = list(conc.keys())
    if exact:
        counts = [int(round(conc[k] * N)) for k in names]
        counts[0] += N - sum(counts)
        symbs = []
        for k, c in zip(names, counts):
            symbs += [k] * c
        symbs = np.array(symbs)
        rng.shuffle(symbs)
        return symbs.tolist()
    p = np.array([conc[k] for k in names], dtype=float)
    return rng.choice(names, size=N, p=p / p.sum()).tolist()


def synthetic_ecis(settings: SyntheticSettings, seed: int = 1234,
                   scale: float = 0.02) -> Dict[str, float]:
    """c0 = 0, singlets 0, n-body ~ N(0, scale/n) eV."""
    rng = np.random.default_rng(seed)
    eci = {}
    for name in settings.eci_names():
        if name == "c0" or name.startswith("c1"):
            eci[name] = 0.0
        else:
            n = int(name[1])
            eci[name] = float(rng.normal(0.0, scale / n))
    return eci


# Al-Mg ECIs published with the reference
# (/root/reference/examples_depr/test_al_mg.py:15-28), mapped onto our shells.
ALMG_ECI_VALUES = {
    "c0": -2.6466290360293874,
    "c1_0": -1.0666948263880078,
    "nn": 0.01078731693580544,       # c2_707_1_1
    "2nn": -0.012304759727020153,    # c2_1000_1_1
    "3nn": -0.010814400169849577,    # c2_1225_1_1
    "tri": -0.017523737495758165,    # c3_1225_2_1
    "iso": -0.011318935831421125,    # c3_1225_3_1
    "tet": 0.0016577886586285448,    # c4_1000_1_1
}


def almg_ecis(settings: SyntheticSettings) -> Dict[str, float]:
    """Binary Al-Mg ECI set of the reference example on our family names."""
    eci = {"c0": ALMG_ECI_VALUES["c0"], "c1_0": ALMG_ECI_VALUES["c1_0"]}
    for prefix, fam in settings.cluster_info[0].items():
        n = fam["size"]
        eci[prefix + "_" + "0" * n] = ALMG_ECI_VALUES[fam["descriptor"]]
    return eci
