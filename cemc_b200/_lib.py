"""ctypes binding of the C-ABI shared library (include/cemc_b200.h).

The library is built in-tree (``__graft_entry__.build()`` / ``build_ext()``
below) as ``cemc_b200/_cemc_b200.so``.  There is NO fallback: if the library
is missing or no CUDA device is usable, the product raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

from .tables import CemcTablesStruct

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CEMC_B200_LIB") or os.path.join(_HERE, "_cemc_b200.so")
_CSRC = os.path.join(_HERE, "csrc")
# translation units (compiled in parallel, linked into one library) and their headers
UNITS = ["cemc_b200.cu", "cemc_batch_product.cu", "cemc_batch_spin.cu", "cemc_batch_tab.cu",
         "cemc_batch_tab32.cu"]
HEADERS = ["cemc_kernels.cuh", "cemc_spin_kernel.cuh", "cemc_batch_kernel.cuh",
           "cemc_batch_launch.cuh"]
SOURCES = [os.path.join(_CSRC, f) for f in UNITS + HEADERS] + \
    [os.path.join(os.path.dirname(_HERE), "include", "cemc_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3",
              "-std=c++17", "-Xcompiler", "-fPIC", "-fmad=false"]

_lib = None


def source_hash() -> str:
    """sha256 (first 16 hex digits) over the kernel sources and the C header: identifies the
    build that profiler-derived figures (profiles/traffic.json) were taken from."""
    import hashlib
    h = hashlib.sha256()
    for path in sorted(SOURCES):
        h.update(os.path.basename(path).encode())
        with open(path, "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


class CemcError(RuntimeError):
    pass


def build_ext(force: bool = False, verbose: bool = False, defines=(), out: str = None) -> str:
    """Compile the CUDA extension for sm_100a (nvcc cross-compiles w/o GPU):
    one object per translation unit, in parallel, then one shared library."""
    out = out or LIB_PATH
    newest = max(os.path.getmtime(s) for s in SOURCES)
    if not force and os.path.exists(out) and os.path.getmtime(out) >= newest:
        return out
    nvcc = os.environ.get("NVCC", "nvcc")
    tag = os.path.splitext(os.path.basename(out))[0]
    objdir = os.path.join(_CSRC, "_build", tag)
    os.makedirs(objdir, exist_ok=True)
    flags = NVCC_FLAGS + ["-D" + d for d in defines] + (["-Xptxas", "-v"] if verbose else [])
    procs, objs = [], []
    for u in UNITS:
        obj = os.path.join(objdir, u.replace(".cu", ".o"))
        objs.append(obj)
        procs.append((u, subprocess.Popen([nvcc] + flags + ["-c", "-o", obj, os.path.join(_CSRC, u)])))
    failed = [u for u, p in procs if p.wait() != 0]
    if failed:
        raise RuntimeError("nvcc failed for " + ", ".join(failed))
    subprocess.check_call([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a",
                           "-o", out] + objs)
    return out


_u64p = C.POINTER(C.c_uint64)
_i32p = C.POINTER(C.c_int32)
_i8p = C.POINTER(C.c_int8)
_u8p = C.POINTER(C.c_uint8)
_f64p = C.POINTER(C.c_double)
_H = C.c_void_p

# name -> argtypes; every function returns int except the two noted below.
SIGNATURES = {
    "cemc_create": [C.POINTER(CemcTablesStruct), C.c_int, C.c_int, C.c_int,
                    C.c_void_p, C.POINTER(_H)],
    "cemc_destroy": [_H],
    "cemc_set_replica_stride": [_H, C.c_int],
    "cemc_set_stream": [_H, C.c_void_p],
    "cemc_synchronize": [_H],
    "cemc_set_order_mode": [_H, C.c_int],
    "cemc_set_block_threads": [_H, C.c_int],
    "cemc_set_generic_path": [_H, C.c_int],
    "cemc_set_batch": [_H, C.c_int],
    "cemc_set_cluster": [_H, C.c_int],
    "cemc_set_autotune": [_H, C.c_int],
    "cemc_get_variant": [_H, _i32p, _i32p],
    "cemc_last_variant": [_H, _i32p],
    "cemc_batch_applicable": [_H, _i32p],
    "cemc_set_variant": [_H, C.c_int, C.c_int],
    "cemc_set_spin_kernel": [_H, C.c_int],
    "cemc_set_table_eval": [_H, C.c_int],
    "cemc_set_precision": [_H, C.c_int],
    "cemc_set_replica_order": [_H, _i32p],
    "cemc_get_batch_eval": [_H, _i32p],
    "cemc_set_lattice_arithmetic": [_H, C.c_int],
    "cemc_get_lattice_arithmetic": [_H, _i32p],
    "cemc_set_screen_slack": [_H, C.c_double],
    "cemc_debug_phase_cycles": [_H, _u64p],
    "cemc_selftest_division": [_H, C.c_uint64, C.c_int, C.c_int, _u64p],
    "cemc_set_occupancy": [_H, _i8p],
    "cemc_get_occupancy": [_H, _i8p],
    "cemc_set_cf": [_H, _f64p],
    "cemc_get_cf": [_H, _f64p],
    "cemc_recompute_cf": [_H],
    "cemc_set_ecis": [_H, _f64p, C.c_int],
    "cemc_get_ecis": [_H, _f64p],
    "cemc_get_energy": [_H, _f64p],
    "cemc_set_kT": [_H, _f64p],
    "cemc_get_kT": [_H, _f64p],
    "cemc_seed": [_H, C.c_uint64],
    "cemc_set_step": [_H, _u64p],
    "cemc_set_sgc_species": [_H, C.c_int, _i8p],
    "cemc_get_tracker": [_H, _i32p, _i32p],
    "cemc_set_tracker": [_H, _i32p],
    "cemc_get_counters": [_H, _u64p, _u64p],
    "cemc_reset_counters": [_H],
    "cemc_trial_changes": [_H, C.c_int, C.c_int, _i32p, _i8p, _i8p, _f64p],
    "cemc_undo_changes": [_H, C.c_int],
    "cemc_clear_history": [_H, C.c_int],
    "cemc_replay": [_H, C.c_int, _i32p, _i8p, _f64p, _u8p, _f64p],
    "cemc_run_sgc": [_H, C.c_int64],
    "cemc_run_canonical": [_H, C.c_int64],
    "cemc_set_trace": [_H, C.c_int64],
    "cemc_get_trace": [_H, C.c_int64, _i32p, _i8p, _f64p, _u8p, _f64p],
    "cemc_energy_autocorrelation": [_H, C.c_int64, _f64p],
    "cemc_set_observe": [_H, C.c_int],
    "cemc_reset_accumulators": [_H, _f64p],
    "cemc_get_accumulators": [_H, _f64p],
    "cemc_set_device_observers": [_H, C.c_int64, C.c_int, C.c_int64],
    "cemc_reset_device_observers": [_H, _i8p],
    "cemc_get_device_observers": [_H, _u64p, _f64p, _f64p, _f64p, _f64p, _i8p, _f64p, _f64p, C.c_int64],
    "cemc_pt_exchange": [_H, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                         C.c_int, C.c_uint64, C.c_void_p],
    "cemc_energy_dev": [_H, C.POINTER(C.c_void_p)],
    "cemc_timer_start": [_H],
    "cemc_timer_stop": [_H, C.POINTER(C.c_float)],
    "cemc_launch_count": [_H, _u64p],
}


def load():
    """Load the shared library; raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "cemc_b200: CUDA extension {} is missing -- run "
            "`python -c 'import __graft_entry__ as g; g.build()'`; there is "
            "no CPU fallback".format(LIB_PATH))
    lib = C.CDLL(LIB_PATH)
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = C.c_int
    lib.cemc_last_error.restype = C.c_char_p
    lib.cemc_last_error.argtypes = []
    lib.cemc_version.restype = C.c_int
    lib.cemc_version.argtypes = []
    _lib = lib
    return lib


def check(rc: int):
    if rc != 0:
        raise CemcError(load().cemc_last_error().decode("utf-8", "replace"))
