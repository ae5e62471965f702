"""TEST INFRASTRUCTURE -- ctypes front-end of oracle/ce_oracle.c.

Not product code: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from cemc_b200.tables import CemcTablesStruct, FlatTables

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    path = os.path.join(_HERE, "libce_oracle.so")
    src = os.path.join(_HERE, "ce_oracle.c")
    if force or not os.path.exists(path) or \
            os.path.getmtime(path) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "libce_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return path


class _Obs(C.Structure):
    _fields_ = [("n_singlets", C.c_int), ("singlet_idx", C.POINTER(C.c_int32)),
                ("ref", C.c_double), ("acc", C.POINTER(C.c_double))]


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.oracle_energy.restype = C.c_double
    return _LIB


def _p(a, ct):
    return None if a is None else a.ctypes.data_as(C.POINTER(ct))


class OracleChain(object):
    """One Markov chain evaluated by the C oracle."""

    def __init__(self, tables: FlatTables, occ, cf=None, eci=None, kT=1.0,
                 seed=0, replica=0, ref=1.0):
        self.t = tables
        self.st: CemcTablesStruct = tables.as_struct()
        self.lib = _lib()
        self.occ = np.array(occ, dtype=np.int8).copy()
        self.eci = (tables.eci if eci is None else
                    np.array(eci, dtype=np.float64)).copy()
        self.cf = np.zeros(tables.n_eci, dtype=np.float64)
        if cf is None:
            self.recompute_cf()
        else:
            self.cf[:] = cf
        self.kT = float(kT)
        self.seed, self.replica = int(seed), int(replica)
        self.step = 0
        self.n_accepted = C.c_uint64(0)
        self.e_cur = C.c_double(self.energy())
        self._sidx = np.array(tables.singlet_indices, dtype=np.int32)
        self.acc = np.zeros(3 + 3 * len(self._sidx), dtype=np.float64)
        self._obs = _Obs(len(self._sidx), _p(self._sidx, C.c_int32),
                         float(ref), _p(self.acc, C.c_double))
        self._tracker = None

    # -- CEUpdater surface ------------------------------------------------
    def recompute_cf(self):
        self.lib.oracle_full_cf(C.byref(self.st), _p(self.occ, C.c_int8),
                                _p(self.cf, C.c_double))
        return self.cf

    def energy(self, cf=None):
        cf = self.cf if cf is None else cf
        return self.lib.oracle_energy(C.byref(self.st),
                                      _p(self.eci, C.c_double),
                                      _p(cf, C.c_double))

    def update_cf(self, site, new_sp):
        """Committed single-site change (update_cf + clear_history)."""
        nxt = np.zeros_like(self.cf)
        err = self.lib.oracle_update_cf(C.byref(self.st),
                                        _p(self.occ, C.c_int8),
                                        _p(self.cf, C.c_double),
                                        _p(nxt, C.c_double), int(site),
                                        int(new_sp))
        if err:
            raise RuntimeError("Attempting to move a background atom!")
        self.cf[:] = nxt
        self.e_cur.value = self.energy()

    def set_ecis(self, eci):
        self.eci[:] = eci
        self.e_cur.value = self.energy()

    def set_ref(self, ref):
        self._obs.ref = float(ref)

    def reset_acc(self):
        self.acc[:] = 0.0

    @property
    def e(self):
        return self.e_cur.value

    # -- Metropolis -------------------------------------------------------
    def replay(self, sites, news, u):
        sites = np.ascontiguousarray(sites, dtype=np.int32).reshape(-1, 2)
        news = np.ascontiguousarray(news, dtype=np.int8).reshape(-1, 2)
        u = np.ascontiguousarray(u, dtype=np.float64)
        n = len(u)
        acc = np.zeros(n, dtype=np.uint8)
        e = np.zeros(n, dtype=np.float64)
        err = self.lib.oracle_replay(
            C.byref(self.st), _p(self.eci, C.c_double), _p(self.occ, C.c_int8),
            _p(self.cf, C.c_double), C.byref(self.e_cur), C.c_double(self.kT),
            n, _p(sites, C.c_int32), _p(news, C.c_int8), _p(u, C.c_double),
            _p(acc, C.c_uint8), _p(e, C.c_double), C.byref(self._obs),
            C.byref(self.n_accepted))
        if err:
            raise RuntimeError("oracle_replay failed")
        self.step += n
        self._tracker = None
        return acc, e

    def _trace(self, n, trace):
        if not trace:
            return None, None, None, None, None
        return (np.zeros((n, 2), dtype=np.int32), np.zeros((n, 2), dtype=np.int8),
                np.zeros(n), np.zeros(n, dtype=np.uint8), np.zeros(n))

    def run_sgc(self, n_steps, allowed=None, trace=False):
        t = self.t
        active = np.array([s for s in range(t.N) if t.symm_of_site[s] >= 0],
                          dtype=np.int32)
        allowed = np.arange(t.S, dtype=np.int8) if allowed is None else \
            np.array(allowed, dtype=np.int8)
        ts, tn, tu, ta, te = self._trace(n_steps, trace)
        err = self.lib.oracle_run_sgc(
            C.byref(self.st), _p(self.eci, C.c_double), _p(self.occ, C.c_int8),
            _p(self.cf, C.c_double), C.byref(self.e_cur), C.c_double(self.kT),
            C.c_uint64(self.seed), C.c_uint32(self.replica),
            C.c_uint64(self.step), C.c_int64(n_steps), len(active),
            _p(active, C.c_int32), len(allowed), _p(allowed, C.c_int8),
            C.byref(self._obs), C.byref(self.n_accepted), _p(ts, C.c_int32),
            _p(tn, C.c_int8), _p(tu, C.c_double), _p(ta, C.c_uint8),
            _p(te, C.c_double))
        if err:
            raise RuntimeError("oracle_run_sgc failed")
        self.step += n_steps
        self._tracker = None
        return (ts, tn, tu, ta, te) if trace else None

    def tracker(self):
        if self._tracker is None:
            t = self.t
            lst = np.zeros(t.N, dtype=np.int32)
            loc = np.zeros(t.N, dtype=np.int32)
            off = np.zeros(t.S + 1, dtype=np.int32)
            self.lib.oracle_tracker_init(C.byref(self.st),
                                         _p(self.occ, C.c_int8),
                                         _p(lst, C.c_int32),
                                         _p(loc, C.c_int32),
                                         _p(off, C.c_int32))
            self._tracker = (lst, loc, off)
        return self._tracker

    def run_canonical(self, n_steps, trace=False):
        lst, loc, off = self.tracker()
        ts, tn, tu, ta, te = self._trace(n_steps, trace)
        err = self.lib.oracle_run_canonical(
            C.byref(self.st), _p(self.eci, C.c_double), _p(self.occ, C.c_int8),
            _p(self.cf, C.c_double), C.byref(self.e_cur), C.c_double(self.kT),
            C.c_uint64(self.seed), C.c_uint32(self.replica),
            C.c_uint64(self.step), C.c_int64(n_steps), _p(lst, C.c_int32),
            _p(loc, C.c_int32), _p(off, C.c_int32), C.byref(self._obs),
            C.byref(self.n_accepted), _p(ts, C.c_int32), _p(tn, C.c_int8),
            _p(tu, C.c_double), _p(ta, C.c_uint8), _p(te, C.c_double))
        if err == 2:
            raise RuntimeError("There is only one element in the given atoms "
                               "object!")
        if err:
            raise RuntimeError("oracle_run_canonical failed")
        self.step += n_steps
        return (ts, tn, tu, ta, te) if trace else None


def philox(seed, step, replica, stream):
    out = (C.c_uint32 * 4)()
    _lib().oracle_philox(C.c_uint64(seed), C.c_uint64(step),
                         C.c_uint32(replica), C.c_uint32(stream), out)
    return [int(x) for x in out]


def pt_direction(seed, rnd):
    """0 = "up", 1 = "down": the per-cycle direction of the device-side round loop."""
    return int(_lib().oracle_pt_direction(C.c_uint64(seed), C.c_uint64(rnd)))


def pt_exchange(energies, slot_of_replica, kT_of_slot, direction, seed, rnd):
    e = np.ascontiguousarray(energies, dtype=np.float64)
    slots = np.array(slot_of_replica, dtype=np.int32).copy()
    kts = np.ascontiguousarray(kT_of_slot, dtype=np.float64)
    n_acc = _lib().oracle_pt_exchange(len(e), _p(e, C.c_double),
                                      _p(slots, C.c_int32),
                                      _p(kts, C.c_double), int(direction),
                                      C.c_uint64(seed), C.c_uint64(rnd))
    return slots, n_acc
