"""Flatten a cluster-expansion settings object into the read-only tables of
``include/cemc_b200.h`` (struct ``cemc_tables``).

This is the host-side equivalent of ``CEUpdater::init``
(/root/reference/cpp/src/ce_updater.cpp:32-234): same inputs (atoms, settings
``BC``, ECI dict), same derived quantities (species ids, ECI order,
symmetry-group counts, equivalent decorations), but the result is a handful of
dense arrays instead of string-keyed maps.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Sequence

import numpy as np

from .synthetic import equivalent_deco

POS_REF = -1
KIND_EMPTY, KIND_SINGLET, KIND_CLUSTER = 0, 1, 2


class CemcTablesStruct(C.Structure):
    """ctypes mirror of ``struct cemc_tables`` (include/cemc_b200.h)."""
    _fields_ = [
        ("n_sites", C.c_int32), ("n_species", C.c_int32), ("n_bf", C.c_int32),
        ("n_cols", C.c_int32), ("n_eci", C.c_int32), ("n_symm", C.c_int32),
        ("n_fam", C.c_int32), ("n_deco", C.c_int32),
        ("trans", C.POINTER(C.c_int32)),
        ("symm_of_site", C.POINTER(C.c_int32)),
        ("symm_count", C.POINTER(C.c_int32)),
        ("bf", C.POINTER(C.c_double)),
        ("eci", C.POINTER(C.c_double)),
        ("eci_kind", C.POINTER(C.c_int32)),
        ("eci_bf", C.POINTER(C.c_int32)),
        ("fam_size", C.POINTER(C.c_int32)),
        ("fam_nsub", C.POINTER(C.c_int32)),
        ("fam_pos_off", C.POINTER(C.c_int32)),
        ("fam_pos", C.POINTER(C.c_int32)),
        ("term_fam", C.POINTER(C.c_int32)),
        ("term_count", C.POINTER(C.c_int32)),
        ("term_deco_off", C.POINTER(C.c_int32)),
        ("deco", C.POINTER(C.c_int8)),
        ("lattice_dims", C.POINTER(C.c_int32)),
    ]


class SelfInteractionError(Exception):
    """Same site twice in one cluster (cemc/ce_calculator.py:19, :596-608)."""


def _ptr(a, ctype):
    return a.ctypes.data_as(C.POINTER(ctype))


class FlatTables(object):
    """Dense tables + the name bookkeeping the Python API needs."""

    def __init__(self, settings, eci: Dict[str, float],
                 symbols: Sequence[str]):
        self.eci_names: List[str] = sorted(eci.keys())  # std::map order,
        # cpp/include/additional_tools.tpp:37-45, ce_updater.cpp:197-203
        n_eci = len(self.eci_names)
        if n_eci == 0:
            raise ValueError("No ECIs given")
        self.eci_index = {n: i for i, n in enumerate(self.eci_names)}

        # species ids: rank in sorted(unique_elements U symbols present)
        # (ce_updater.cpp:67-78, symbols_with_numbers.cpp:11-21)
        uniq = set(settings.unique_elements)
        uniq.update(symbols)
        self.species: List[str] = sorted(uniq)
        self.species_id = {s: i for i, s in enumerate(self.species)}
        S = len(self.species)
        N = len(symbols)
        D = int(settings.num_unique_elements) - 1
        if D < 1 or D > 10:
            raise ValueError("1..10 basis functions supported "
                             "(single-digit decoration numbers)")
        if S > 127:
            raise ValueError("too many species for int8 occupancy")

        # symmetry groups (ce_updater.cpp:733-776)
        bkg = set(int(i) for i in settings.background_indices)
        symm_of_site = np.full(N, -1, dtype=np.int32)
        groups = settings.index_by_trans_symm
        for g, grp in enumerate(groups):
            for s in grp:
                if symm_of_site[s] != -1:
                    raise RuntimeError("One site appears to be present in "
                                       "more than one translation symmetry "
                                       "group!")
                symm_of_site[s] = g
        for s in range(N):
            if symm_of_site[s] == -1 and s not in bkg:
                raise RuntimeError("Site {} has not been assigned to any "
                                   "translational symmetry group!".format(s))
        n_symm = len(groups)
        symm_count = np.zeros(max(n_symm, 1), dtype=np.int32)
        for s in range(N):
            if symm_of_site[s] >= 0:
                symm_count[symm_of_site[s]] += 1

        # basis functions (ce_updater.cpp:157-181, basis_function.cpp:4-19)
        bfs = settings.basis_functions
        if len(bfs) < D:
            raise ValueError("fewer basis functions than decorations")
        bf = np.zeros((D, S), dtype=np.float64)
        for d in range(D):
            for sym, val in bfs[d].items():
                if sym in self.species_id:
                    bf[d, self.species_id[sym]] = float(val)

        # cluster families; count[prefix] summed over groups (:137-144)
        cluster_info = settings.cluster_info
        if len(cluster_info) != n_symm:
            raise ValueError("cluster_info must have one dict per "
                             "translational symmetry group")
        count = {}
        used_cols = set()
        for info in cluster_info:
            for prefix, fam in info.items():
                n = int(fam["size"])
                if n < 2:
                    continue
                if n > 4:
                    raise ValueError("Only cluster sizes 2, 3 and 4 are "
                                     "supported!")  # cluster.cpp:168
                count[prefix] = count.get(prefix, 0) + len(fam["indices"])
                for sub in fam["indices"]:
                    if len(sub) != n - 1:
                        raise ValueError("indices rows must have size-1 "
                                         "entries")
                    if int(fam["ref_indx"]) in sub or len(set(sub)) != len(sub):
                        raise SelfInteractionError(
                            "The simulation cell is so small that the same "
                            "site is present multiple times within one "
                            "cluster. Increase the size of the simulation "
                            "cell.")
                    used_cols.update(int(c) for c in sub)
        if not used_cols:
            raise RuntimeError("It looks like no clusters are present.")
        self.cols: List[int] = sorted(used_cols)
        col_lut = {c: k for k, c in enumerate(self.cols)}
        K = len(self.cols)

        # translation matrix restricted to used columns (:899-979)
        tm = settings.trans_matrix
        trans = np.zeros((N, K), dtype=np.int32)
        compact_cols = getattr(settings, "trans_matrix_columns", None)
        if isinstance(tm, np.ndarray) and compact_cols is not None:
            pos = {c: k for k, c in enumerate(compact_cols)}
            if tm.shape != (N, len(compact_cols)) or any(c not in pos for c in self.cols):
                raise ValueError("compact translation matrix does not cover "
                                 "the columns used by the clusters")
            trans[:, :] = tm[:, [pos[c] for c in self.cols]]
        elif isinstance(tm, np.ndarray):
            if tm.shape[0] != N:
                raise ValueError("The number of atoms and the dimension of "
                                 "the translation matrix is inconsistent")
            if max(self.cols) + 1 > tm.shape[1]:
                raise ValueError("Something is wrong with the translation "
                                 "matrix passed.")
            trans[:, :] = tm[:, self.cols]
        else:
            if len(tm) != N:
                raise ValueError("The number of atoms and the dimension of "
                                 "the translation matrix is inconsistent")
            for s in range(N):
                if s in bkg:
                    continue
                row = tm[s]
                for k, c in enumerate(self.cols):
                    if c not in row:
                        raise ValueError("Requested value {} is not a key in "
                                         "the dictionary!".format(c))
                    trans[s, k] = int(row[c])
        if trans.min() < 0 or trans.max() >= N:
            raise ValueError("translation matrix entry out of range")

        fam_size, fam_nsub, fam_pos_off, fam_pos = [], [], [0], []
        fam_id = {}
        for g, info in enumerate(cluster_info):
            for prefix, fam in info.items():
                n = int(fam["size"])
                if n < 2:
                    continue
                M = len(fam["indices"])
                order = fam["order"]
                pos = np.zeros((M, n), dtype=np.int32)
                for m in range(M):
                    if sorted(order[m]) != list(range(n)):
                        raise ValueError("order rows must be permutations")
                    for k in range(n):
                        src = int(order[m][k])
                        pos[m, k] = POS_REF if src == 0 else \
                            col_lut[int(fam["indices"][m][src - 1])]
                fam_id[(g, prefix)] = len(fam_size)
                fam_size.append(n)
                fam_nsub.append(M)
                fam_pos.append(pos.ravel())
                fam_pos_off.append(fam_pos_off[-1] + M * n)

        # ECI terms (ce_updater.cpp:353-404)
        eci_kind = np.zeros(n_eci, dtype=np.int32)
        eci_bf = np.zeros(n_eci, dtype=np.int32)
        term_fam = np.full((max(n_symm, 1), n_eci), -1, dtype=np.int32)
        term_count = np.ones((max(n_symm, 1), n_eci), dtype=np.int32)
        term_deco_off = [0]
        deco_rows: List[List[int]] = []
        self.singlet_indices: List[int] = []
        for i, name in enumerate(self.eci_names):
            if name.startswith("c0"):
                eci_kind[i] = KIND_EMPTY
            elif name.startswith("c1"):
                eci_kind[i] = KIND_SINGLET
                d = ord(name[name.rfind("_") + 1]) - ord("0")
                if d < 0 or d >= D:
                    raise ValueError("bad singlet name " + name)
                eci_bf[i] = d
            else:
                eci_kind[i] = KIND_CLUSTER
        for g in range(n_symm):
            for i, name in enumerate(self.eci_names):
                if eci_kind[i] == KIND_CLUSTER:
                    pos_ = name.rfind("_")
                    prefix, dec_str = name[:pos_], name[pos_ + 1:]
                    key = (g, prefix)
                    if key in fam_id:
                        f = fam_id[key]
                        n = fam_size[f]
                        deco = [ord(ch) - ord("0") for ch in dec_str]
                        if len(deco) != n or min(deco) < 0 or max(deco) >= D:
                            raise ValueError(
                                "decoration of {} does not fit a {}-body "
                                "cluster with {} basis functions".format(
                                    name, n, D))
                        eq = equivalent_deco(
                            deco, cluster_info[g][prefix]["equiv_sites"])
                        for e in eq:
                            deco_rows.append(list(e) + [0] * (4 - n))
                        term_fam[g, i] = f
                        term_count[g, i] = count[prefix]
                term_deco_off.append(len(deco_rows))
        # singlets in sorted-name order (ce_updater.cpp:208-221)
        self.singlet_indices = [i for i, n_ in enumerate(self.eci_names)
                                if n_.startswith("c1")]
        self.singlet_names = [self.eci_names[i] for i in self.singlet_indices]

        self.N, self.S, self.D, self.K = N, S, D, K
        self.n_eci, self.n_symm = n_eci, n_symm
        self.trans = np.ascontiguousarray(trans)
        self.symm_of_site = symm_of_site
        self.symm_count = symm_count
        self.bf = np.ascontiguousarray(bf)
        self.eci = np.array([float(eci[n]) for n in self.eci_names],
                            dtype=np.float64)
        self.eci_kind, self.eci_bf = eci_kind, eci_bf
        self.fam_size = np.array(fam_size, dtype=np.int32)
        self.fam_nsub = np.array(fam_nsub, dtype=np.int32)
        self.fam_pos_off = np.array(fam_pos_off, dtype=np.int32)
        self.fam_pos = (np.concatenate(fam_pos).astype(np.int32)
                        if fam_pos else np.zeros(1, dtype=np.int32))
        self.term_fam = np.ascontiguousarray(term_fam)
        self.term_count = np.ascontiguousarray(term_count)
        self.term_deco_off = np.array(term_deco_off, dtype=np.int32)
        self.deco = (np.array(deco_rows, dtype=np.int8).reshape(-1, 4)
                     if deco_rows else np.zeros((1, 4), dtype=np.int8))
        self.n_deco = len(deco_rows)
        self.background = sorted(bkg)
        # hint for the index-arithmetic translation (verified by cemc_create against `trans`)
        size = getattr(settings, "size", None)
        self.lattice_dims = np.zeros(3, dtype=np.int32)
        try:
            if size is not None and len(size) == 3 and int(np.prod(size)) == N and not bkg:
                self.lattice_dims = np.array([int(x) for x in size], dtype=np.int32)
        except (TypeError, ValueError):
            pass

    # ------------------------------------------------------------------
    def occupancy(self, symbols: Sequence[str]) -> np.ndarray:
        """int8 species ids for a list of chemical symbols."""
        return np.array([self.species_id[s] for s in symbols], dtype=np.int8)

    def symbols_of(self, occ: np.ndarray) -> List[str]:
        return [self.species[int(v)] for v in occ]

    def eci_vector(self, eci: Dict[str, float]) -> np.ndarray:
        """ECI dict -> vector in table order; names must match exactly
        (CEUpdater::set_ecis / all_eci_corresponds_to_cf, :615-629,:778)."""
        if set(eci.keys()) != set(self.eci_names):
            raise ValueError("All ECIs has to correspond to a correlation "
                             "function!")
        return np.array([float(eci[n]) for n in self.eci_names],
                        dtype=np.float64)

    def cf_vector(self, cf: Dict[str, float]) -> np.ndarray:
        """CF dict -> vector; names absent from the dict stay 0.0
        (cpp/src/cf_history_tracker.cpp:82-95)."""
        return np.array([float(cf.get(n, 0.0)) for n in self.eci_names],
                        dtype=np.float64)

    def gathered_sites_per_change(self) -> int:
        """G of SURVEY.md 8(d): sum over families of M*(n-1)."""
        return int(sum(int(m) * (int(n) - 1) for m, n in
                       zip(self.fam_nsub, self.fam_size)))

    def algorithmic_bytes_per_move(self, sites_changed: int) -> int:
        """B = c*(4K + G) + 16 n_eci   (SURVEY.md 8d)."""
        return sites_changed * (4 * self.K + self.gathered_sites_per_change()) \
            + 16 * self.n_eci

    def as_struct(self) -> CemcTablesStruct:
        st = CemcTablesStruct()
        st.n_sites, st.n_species, st.n_bf = self.N, self.S, self.D
        st.n_cols, st.n_eci, st.n_symm = self.K, self.n_eci, self.n_symm
        st.n_fam, st.n_deco = len(self.fam_size), self.n_deco
        st.trans = _ptr(self.trans, C.c_int32)
        st.symm_of_site = _ptr(self.symm_of_site, C.c_int32)
        st.symm_count = _ptr(self.symm_count, C.c_int32)
        st.bf = _ptr(self.bf, C.c_double)
        st.eci = _ptr(self.eci, C.c_double)
        st.eci_kind = _ptr(self.eci_kind, C.c_int32)
        st.eci_bf = _ptr(self.eci_bf, C.c_int32)
        st.fam_size = _ptr(self.fam_size, C.c_int32)
        st.fam_nsub = _ptr(self.fam_nsub, C.c_int32)
        st.fam_pos_off = _ptr(self.fam_pos_off, C.c_int32)
        st.fam_pos = _ptr(self.fam_pos, C.c_int32)
        st.term_fam = _ptr(self.term_fam, C.c_int32)
        st.term_count = _ptr(self.term_count, C.c_int32)
        st.term_deco_off = _ptr(self.term_deco_off, C.c_int32)
        st.deco = _ptr(self.deco, C.c_int8)
        st.lattice_dims = _ptr(self.lattice_dims, C.c_int32)
        st._keepalive = self  # arrays must outlive the struct
        return st
