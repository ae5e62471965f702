"""Host-side mirrors of the reference's updater interface.

``BatchedCEUpdater``  -- the new batched surface: R replicas on
                         one GPU behind one C-ABI handle.
``PyCEUpdater``       -- same Python-visible names as the reference's Cython
                         class (/root/reference/cemc/cpp_ext/pyce_updater.pyx:5-60)
                         so ``CE`` (cemc/ce_calculator.py:199-208) works
                         unchanged; one replica of a ``BatchedCEUpdater``.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Sequence

import numpy as np

from . import _lib
from .tables import FlatTables

ORDER_REFERENCE, ORDER_TREE = 0, 1


def _p(a, ct):
    return None if a is None else a.ctypes.data_as(C.POINTER(ct))


class BatchedCEUpdater(object):
    """R independent chains on one CUDA device."""

    def __init__(self, tables: FlatTables, n_replicas: int, device: int = 0,
                 replica_offset: int = 0, stream=None, replica_stride: int = 1):
        self.lib = _lib.load()
        self.tables = tables
        self.R = int(n_replicas)
        self.N = tables.N
        self.n_eci = tables.n_eci
        self.n_singlets = len(tables.singlet_indices)
        self.acc_stride = 3 + 3 * self.n_singlets
        self.device = int(device)
        self.replica_offset = int(replica_offset)
        self._struct = tables.as_struct()
        self._h = C.c_void_p()
        _lib.check(self.lib.cemc_create(C.byref(self._struct), self.R,
                                        self.replica_offset, self.device,
                                        stream, C.byref(self._h)))
        self.replica_stride = int(replica_stride)
        if self.replica_stride != 1:
            _lib.check(self.lib.cemc_set_replica_stride(self._h, self.replica_stride))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self.lib.cemc_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- state ----------------------------------------------------------
    def set_occupancy(self, occ):
        occ = np.ascontiguousarray(occ, dtype=np.int8).reshape(self.R, self.N)
        _lib.check(self.lib.cemc_set_occupancy(self._h, _p(occ, C.c_int8)))

    def get_occupancy(self):
        occ = np.zeros((self.R, self.N), dtype=np.int8)
        _lib.check(self.lib.cemc_get_occupancy(self._h, _p(occ, C.c_int8)))
        return occ

    def set_cf(self, cf):
        cf = np.ascontiguousarray(
            np.broadcast_to(np.asarray(cf, dtype=np.float64),
                            (self.R, self.n_eci)))
        _lib.check(self.lib.cemc_set_cf(self._h, _p(cf, C.c_double)))

    def get_cf(self):
        cf = np.zeros((self.R, self.n_eci), dtype=np.float64)
        _lib.check(self.lib.cemc_get_cf(self._h, _p(cf, C.c_double)))
        return cf

    def recompute_cf(self):
        _lib.check(self.lib.cemc_recompute_cf(self._h))

    def set_ecis(self, eci):
        eci = np.ascontiguousarray(eci, dtype=np.float64)
        if eci.ndim == 1:
            if eci.shape[0] != self.n_eci:
                raise ValueError("All ECIs has to correspond to a correlation "
                                 "function!")
            _lib.check(self.lib.cemc_set_ecis(self._h, _p(eci, C.c_double), 0))
        else:
            if eci.shape != (self.R, self.n_eci):
                raise ValueError("ECI array must be [R, n_eci]")
            _lib.check(self.lib.cemc_set_ecis(self._h, _p(eci, C.c_double), 1))

    def get_ecis(self):
        eci = np.zeros((self.R, self.n_eci), dtype=np.float64)
        _lib.check(self.lib.cemc_get_ecis(self._h, _p(eci, C.c_double)))
        return eci

    def get_energy(self):
        e = np.zeros(self.R, dtype=np.float64)
        _lib.check(self.lib.cemc_get_energy(self._h, _p(e, C.c_double)))
        return e

    def set_kT(self, kT):
        kT = np.ascontiguousarray(
            np.broadcast_to(np.asarray(kT, dtype=np.float64), (self.R,)))
        _lib.check(self.lib.cemc_set_kT(self._h, _p(kT, C.c_double)))

    def get_kT(self):
        kT = np.zeros(self.R, dtype=np.float64)
        _lib.check(self.lib.cemc_get_kT(self._h, _p(kT, C.c_double)))
        return kT

    def seed(self, seed: int):
        _lib.check(self.lib.cemc_seed(self._h, C.c_uint64(int(seed))))

    def set_step(self, steps):
        steps = np.ascontiguousarray(
            np.broadcast_to(np.asarray(steps, dtype=np.uint64), (self.R,)))
        _lib.check(self.lib.cemc_set_step(self._h, _p(steps, C.c_uint64)))

    def set_sgc_species(self, allowed: Sequence[int]):
        a = np.ascontiguousarray(allowed, dtype=np.int8)
        _lib.check(self.lib.cemc_set_sgc_species(self._h, len(a),
                                                 _p(a, C.c_int8)))

    def set_order_mode(self, mode: int):
        _lib.check(self.lib.cemc_set_order_mode(self._h, int(mode)))

    def set_block_threads(self, n: int):
        _lib.check(self.lib.cemc_set_block_threads(self._h, int(n)))

    def get_tracker(self):
        lst = np.zeros((self.R, self.N), dtype=np.int32)
        off = np.zeros((self.R, self.tables.S + 1), dtype=np.int32)
        _lib.check(self.lib.cemc_get_tracker(self._h, _p(lst, C.c_int32),
                                             _p(off, C.c_int32)))
        return lst, off

    def set_tracker(self, lst):
        lst = np.ascontiguousarray(lst, dtype=np.int32).reshape(self.R, self.N)
        _lib.check(self.lib.cemc_set_tracker(self._h, _p(lst, C.c_int32)))

    def set_screen_slack(self, factor: float):
        _lib.check(self.lib.cemc_set_screen_slack(self._h, C.c_double(factor)))

    def set_spin_kernel(self, on: bool):
        _lib.check(self.lib.cemc_set_spin_kernel(self._h, int(bool(on))))

    def set_autotune(self, on: bool):
        _lib.check(self.lib.cemc_set_autotune(self._h, int(bool(on))))

    def set_table_eval(self, on: bool):
        """Testing hook: product tables (default) vs fp64 products in the batch kernel."""
        _lib.check(self.lib.cemc_set_table_eval(self._h, 1 if on else 0))

    def set_precision(self, bits: int):
        """64 (default, bit-identical to the reference) or 32: the fp32 variant
        (single-precision product tables and sub-cluster sums, fp64 everything else)."""
        _lib.check(self.lib.cemc_set_precision(self._h, int(bits)))

    def set_replica_order(self, order=None):
        """CTA i of the batch kernel works on replica ``order[i]`` (load balance only)."""
        if order is None:
            _lib.check(self.lib.cemc_set_replica_order(self._h, None))
        else:
            o = np.ascontiguousarray(order, dtype=np.int32)
            _lib.check(self.lib.cemc_set_replica_order(self._h, _p(o, C.c_int32)))

    def get_batch_eval(self) -> int:
        """0 fp64 products, 1 binary spin, 2 product tables."""
        v = C.c_int32(-1)
        _lib.check(self.lib.cemc_get_batch_eval(self._h, C.byref(v)))
        return v.value

    def set_lattice_arithmetic(self, on: bool):
        _lib.check(self.lib.cemc_set_lattice_arithmetic(self._h, 1 if on else 0))

    def get_lattice_arithmetic(self) -> bool:
        v = C.c_int32(0)
        _lib.check(self.lib.cemc_get_lattice_arithmetic(self._h, C.byref(v)))
        return bool(v.value)

    def set_variant(self, sgc: int = -1, canonical: int = -1):
        _lib.check(self.lib.cemc_set_variant(self._h, int(sgc), int(canonical)))

    def get_variant(self):
        a, b = C.c_int32(-1), C.c_int32(-1)
        _lib.check(self.lib.cemc_get_variant(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def batch_kernel_applies(self) -> bool:
        """True when the speculative batch kernel takes this system (else the generic
        one-move-at-a-time kernel runs it); probes with a zero-move launch request."""
        v = C.c_int32(0)
        _lib.check(self.lib.cemc_batch_applicable(self._h, C.byref(v)))
        return bool(v.value)

    def last_variant(self) -> int:
        """Kernel variant of the most recent Metropolis launch (-1: none)."""
        v = C.c_int32(-1)
        _lib.check(self.lib.cemc_last_variant(self._h, C.byref(v)))
        return v.value

    def set_cluster(self, c: int):
        _lib.check(self.lib.cemc_set_cluster(self._h, int(c)))

    def set_batch(self, b: int):
        _lib.check(self.lib.cemc_set_batch(self._h, int(b)))

    def set_generic_path(self, on: bool):
        _lib.check(self.lib.cemc_set_generic_path(self._h, int(bool(on))))

    def selftest_division(self, seed=1, n_blocks=296, iters=2000) -> int:
        bad = C.c_uint64(0)
        _lib.check(self.lib.cemc_selftest_division(
            self._h, C.c_uint64(seed), int(n_blocks), int(iters), C.byref(bad)))
        return bad.value

    def get_counters(self):
        steps = np.zeros(self.R, dtype=np.uint64)
        acc = np.zeros(self.R, dtype=np.uint64)
        _lib.check(self.lib.cemc_get_counters(self._h, _p(steps, C.c_uint64),
                                              _p(acc, C.c_uint64)))
        return steps, acc

    def reset_counters(self):
        _lib.check(self.lib.cemc_reset_counters(self._h))

    # ---- reference per-call surface --------------------------------------
    def trial_changes(self, replica, sites, olds, news):
        n = len(sites)
        s = np.ascontiguousarray(sites, dtype=np.int32)
        o = None if olds is None else np.ascontiguousarray(olds, dtype=np.int8)
        w = np.ascontiguousarray(news, dtype=np.int8)
        e = C.c_double(0.0)
        _lib.check(self.lib.cemc_trial_changes(
            self._h, int(replica), n, _p(s, C.c_int32), _p(o, C.c_int8),
            _p(w, C.c_int8), C.byref(e)))
        return e.value

    def undo_changes(self, replica=0):
        _lib.check(self.lib.cemc_undo_changes(self._h, int(replica)))

    def clear_history(self, replica=0):
        _lib.check(self.lib.cemc_clear_history(self._h, int(replica)))

    # ---- batched Metropolis ------------------------------------------------
    def replay(self, sites, news, u):
        """sites [R,n,2] int32 (sites[...,1] < 0: one-site step), news [R,n,2]
        int8, u [R,n] -> (accepted [R,n] uint8, e_after [R,n])."""
        u = np.ascontiguousarray(u, dtype=np.float64).reshape(self.R, -1)
        n = u.shape[1]
        sites = np.ascontiguousarray(sites, dtype=np.int32).reshape(self.R, n, 2)
        news = np.ascontiguousarray(news, dtype=np.int8).reshape(self.R, n, 2)
        acc = np.zeros((self.R, n), dtype=np.uint8)
        e = np.zeros((self.R, n), dtype=np.float64)
        _lib.check(self.lib.cemc_replay(
            self._h, n, _p(sites, C.c_int32), _p(news, C.c_int8),
            _p(u, C.c_double), _p(acc, C.c_uint8), _p(e, C.c_double)))
        return acc, e

    def run_sgc(self, n_steps: int):
        _lib.check(self.lib.cemc_run_sgc(self._h, C.c_int64(int(n_steps))))

    def run_canonical(self, n_steps: int):
        _lib.check(self.lib.cemc_run_canonical(self._h, C.c_int64(int(n_steps))))

    def synchronize(self):
        _lib.check(self.lib.cemc_synchronize(self._h))

    def set_trace(self, capacity: int):
        _lib.check(self.lib.cemc_set_trace(self._h, C.c_int64(int(capacity))))

    def get_trace(self, n: int):
        sites = np.zeros((self.R, n, 2), dtype=np.int32)
        news = np.zeros((self.R, n, 2), dtype=np.int8)
        u = np.zeros((self.R, n), dtype=np.float64)
        acc = np.zeros((self.R, n), dtype=np.uint8)
        e = np.zeros((self.R, n), dtype=np.float64)
        _lib.check(self.lib.cemc_get_trace(
            self._h, C.c_int64(n), _p(sites, C.c_int32), _p(news, C.c_int8),
            _p(u, C.c_double), _p(acc, C.c_uint8), _p(e, C.c_double)))
        return sites, news, u, acc, e

    def energy_autocorrelation(self, n: int):
        """[R, 4]: mean, variance, first lag with normalised autocorrelation < 1/2 (or -1),
        smallest normalised autocorrelation seen -- of the energies traced by the last run."""
        out = np.zeros((self.R, 4), dtype=np.float64)
        _lib.check(self.lib.cemc_energy_autocorrelation(self._h, C.c_int64(int(n)), _p(out, C.c_double)))
        return out

    # ---- observers -----------------------------------------------------------
    def set_observe(self, on: bool):
        """Switch the per-step Averager / SGCObserver sums of run_* on or off."""
        _lib.check(self.lib.cemc_set_observe(self._h, 1 if on else 0))

    def reset_accumulators(self, ref=None):
        if ref is not None:
            ref = np.ascontiguousarray(
                np.broadcast_to(np.asarray(ref, dtype=np.float64), (self.R,)))
        _lib.check(self.lib.cemc_reset_accumulators(self._h,
                                                    _p(ref, C.c_double)))

    def get_accumulators(self):
        acc = np.zeros((self.R, self.acc_stride), dtype=np.float64)
        _lib.check(self.lib.cemc_get_accumulators(self._h,
                                                  _p(acc, C.c_double)))
        return acc

    # ---- device-side state observers ---------------------------------------------
    OBS_CF_SUMS, OBS_LOWEST, OBS_ENERGY, OBS_SITE_ORDER = 1, 2, 4, 8

    def set_device_observers(self, interval: int, flags: int, capacity: int = 0):
        _lib.check(self.lib.cemc_set_device_observers(self._h, C.c_int64(int(interval)),
                                                      int(flags), C.c_int64(int(capacity))))
        self._obs_capacity = int(capacity) if interval > 0 and flags else 0

    def reset_device_observers(self, occ_ref=None):
        if occ_ref is not None:
            occ_ref = np.ascontiguousarray(occ_ref, dtype=np.int8).reshape(self.R, self.N)
        _lib.check(self.lib.cemc_reset_device_observers(self._h, _p(occ_ref, C.c_int8)))

    def get_device_observers(self):
        """dict: n_samples [R], cf_sum / cf_sq [R, n_eci], best_energy [R], best_cf [R, n_eci],
        best_occ [R, N], site_order [R, 2] (sum, sum of squares), energies [R, capacity] (ring:
        the energy of boundary k is energies[:, k % capacity])."""
        n = np.zeros(self.R, dtype=np.uint64)
        _lib.check(self.lib.cemc_get_device_observers(self._h, _p(n, C.c_uint64), None, None, None, None,
                                                      None, None, None, C.c_int64(0)))
        ne = int(getattr(self, "_obs_capacity", 0))      # the whole ring; sample k sits at k % capacity
        out = dict(n_samples=n, cf_sum=np.zeros((self.R, self.n_eci)), cf_sq=np.zeros((self.R, self.n_eci)),
                   best_energy=np.zeros(self.R), best_cf=np.zeros((self.R, self.n_eci)),
                   best_occ=np.zeros((self.R, self.N), dtype=np.int8), site_order=np.zeros((self.R, 2)),
                   energies=np.zeros((self.R, max(ne, 1))))
        _lib.check(self.lib.cemc_get_device_observers(
            self._h, None, _p(out["cf_sum"], C.c_double), _p(out["cf_sq"], C.c_double),
            _p(out["best_energy"], C.c_double), _p(out["best_cf"], C.c_double),
            _p(out["best_occ"], C.c_int8), _p(out["site_order"], C.c_double),
            _p(out["energies"], C.c_double) if ne else None, C.c_int64(ne)))
        out["energies"] = out["energies"][:, :ne]
        return out

    # ---- parallel tempering -----------------------------------------------------
    def energy_dev_ptr(self) -> int:
        ptr = C.c_void_p()
        _lib.check(self.lib.cemc_energy_dev(self._h, C.byref(ptr)))
        return ptr.value

    def pt_exchange(self, n_total, energies_dev, slot_of_replica_dev,
                    kT_of_slot_dev, direction, rnd, n_accepted_dev=None):
        """All pointer arguments are raw CUDA device addresses (ints)."""
        _lib.check(self.lib.cemc_pt_exchange(
            self._h, int(n_total), C.c_void_p(energies_dev),
            C.c_void_p(slot_of_replica_dev), C.c_void_p(kT_of_slot_dev),
            int(direction), C.c_uint64(int(rnd)),
            C.c_void_p(n_accepted_dev) if n_accepted_dev else None))

    # ---- timing -----------------------------------------------------------------
    def timer_start(self):
        _lib.check(self.lib.cemc_timer_start(self._h))

    def timer_stop(self) -> float:
        ms = C.c_float(0.0)
        _lib.check(self.lib.cemc_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def launch_count(self) -> int:
        n = C.c_uint64(0)
        _lib.check(self.lib.cemc_launch_count(self._h, C.byref(n)))
        return n.value


class PyCEUpdater(object):
    """Drop-in for ``cemc_cpp_code.PyCEUpdater`` (pyce_updater.pyx:5-60).

    Same constructor and method names; the state lives on the GPU.  The
    caller's ``atoms`` are mutated on every change like the reference does
    (ce_updater.cpp:338-347, :435-446).
    """

    def __init__(self, atoms, bc, corr_func, eci, device: int = 0):
        self.atoms = atoms
        self.bc = bc
        self.corr_func = corr_func
        self.eci = eci
        symbols = [a.symbol for a in atoms]
        self.tables = FlatTables(bc, eci, symbols)
        self.batch = BatchedCEUpdater(self.tables, 1, device=device)
        self.batch.set_occupancy(self.tables.occupancy(symbols)[None, :])
        if corr_func is None:
            # from the definition, on the GPU (the reference needs
            # ase.clease.CorrFunction for this, ce_calculator.py:169-175)
            self.batch.recompute_cf()
        else:
            self.batch.set_cf(self.tables.cf_vector(corr_func)[None, :])
        self._log = []          # (index, old_symbol) since clear_history

    # -- reference names ---------------------------------------------------
    def clear_history(self):
        self.batch.clear_history(0)
        self._log = []

    def undo_changes(self):
        self.batch.undo_changes(0)
        for indx, old in reversed(self._log):
            self.atoms[indx].symbol = old
        self._log = []

    def update_cf(self, system_changes):
        self.calculate([system_changes])

    def calculate(self, system_changes):
        if len(system_changes) == 0:
            return self.get_energy()
        sid = self.tables.species_id
        sites, olds, news = [], [], []
        for indx, old, new in system_changes:
            if new not in sid or old not in sid:
                raise ValueError("unknown symbol in system change")
            sites.append(int(indx))
            olds.append(sid[old])
            news.append(sid[new])
        # like the reference, the stored occupancy (not old_symb) is the truth
        # (ce_updater.cpp:334); a mismatch is reported instead of corrupting undo
        e = self.batch.trial_changes(0, sites, None, news)
        for indx, old, new in system_changes:
            cur = self.atoms[indx].symbol
            if cur != new:
                self._log.append((indx, cur))
                self.atoms[indx].symbol = new
        return e

    def add_linear_vib_correction(self, value):
        raise NotImplementedError("linear vibration correction is outside "
                                  "the hot path (SURVEY.md N9)")

    def vib_energy(self, T):
        return 0.0

    def get_cf(self) -> Dict[str, float]:
        cf = self.batch.get_cf()[0]
        return {n: float(v) for n, v in zip(self.tables.eci_names, cf)}

    def set_ecis(self, ecis):
        self.eci = ecis
        self.batch.set_ecis(self.tables.eci_vector(ecis))

    def get_singlets(self):
        cf = self.batch.get_cf()[0]
        return cf[self.tables.singlet_indices].copy()

    def get_energy(self):
        return float(self.batch.get_energy()[0])

    def get_symbols(self):
        return self.tables.symbols_of(self.batch.get_occupancy()[0])

    def set_num_threads(self, num_threads):
        pass  # OpenMP knob of the reference (ce_updater.hpp:144); no-op here
