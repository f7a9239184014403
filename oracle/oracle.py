"""numpy/ctypes front-end of the C oracle (oracle/cuco_oracle.c). TEST INFRASTRUCTURE ONLY.

Importable only from tests/, __graft_entry__.smoke() and bench.py's CPU legs; the product package
(cucollections_b200) never imports it.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_DIR = Path(__file__).resolve().parent
_LIB_PATH = _DIR / "_build" / "liboracle.so"
_BASELINE_PATH = _DIR / "_build" / "libcpu_baseline.so"

XXHASH32, XXHASH64, MURMUR3_32 = 0, 1, 2
LINEAR, DOUBLE = 0, 1
PLUS, MIN, MAX = 0, 1, 2


def build() -> None:
    subprocess.run(["make", "-s", "-C", str(_DIR)], check=True)


def _load(path: Path) -> C.CDLL:
    if not path.exists():
        build()
    return C.CDLL(str(path))


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        L = _load(_LIB_PATH)
        vp, i64, u64, u32, dbl, i = C.c_void_p, C.c_int64, C.c_uint64, C.c_uint32, C.c_double, C.c_int
        L.oracle_xxhash32.restype, L.oracle_xxhash32.argtypes = u32, [vp, u64, u32]
        L.oracle_xxhash64.restype, L.oracle_xxhash64.argtypes = u64, [vp, u64, u64]
        L.oracle_murmur3_32.restype, L.oracle_murmur3_32.argtypes = u32, [vp, u64, u32]
        L.oracle_murmur3_x64_128.restype, L.oracle_murmur3_x64_128.argtypes = None, [vp, u64, u64, vp]
        L.oracle_murmur3_x86_128.restype, L.oracle_murmur3_x86_128.argtypes = None, [vp, u64, u32, vp]
        L.oracle_murmur3_fmix32.restype, L.oracle_murmur3_fmix32.argtypes = u32, [u32, u32]
        L.oracle_murmur3_fmix64.restype, L.oracle_murmur3_fmix64.argtypes = u64, [u64, u64]
        L.oracle_prime_at_least.restype, L.oracle_prime_at_least.argtypes = u64, [u64]
        L.oracle_num_windows.restype, L.oracle_num_windows.argtypes = u64, [i64, i, i]
        L.oracle_ceil_div_lf.restype, L.oracle_ceil_div_lf.argtypes = u64, [u64, dbl]
        L.oracle_create.restype = vp
        L.oracle_create.argtypes = [i, i, i, i, i, i, i64, dbl, i64, i64, i, i64]
        L.oracle_destroy.restype, L.oracle_destroy.argtypes = None, [vp]
        L.oracle_capacity.restype, L.oracle_capacity.argtypes = i64, [vp]
        L.oracle_size.restype, L.oracle_size.argtypes = i64, [vp]
        L.oracle_clear.restype, L.oracle_clear.argtypes = None, [vp]
        L.oracle_insert.restype, L.oracle_insert.argtypes = i64, [vp, vp, vp, i64]
        L.oracle_insert_if.restype, L.oracle_insert_if.argtypes = i64, [vp, vp, vp, vp, i64]
        L.oracle_find.restype, L.oracle_find.argtypes = None, [vp, vp, vp, i64]
        L.oracle_contains.restype, L.oracle_contains.argtypes = None, [vp, vp, vp, i64]
        L.oracle_contains_if.restype, L.oracle_contains_if.argtypes = None, [vp, vp, vp, vp, i64]
        L.oracle_insert_and_find.restype, L.oracle_insert_and_find.argtypes = None, [vp, vp, vp, vp, vp, i64]
        L.oracle_insert_or_assign.restype, L.oracle_insert_or_assign.argtypes = None, [vp, vp, vp, i64]
        L.oracle_insert_or_apply.restype = None
        L.oracle_insert_or_apply.argtypes = [vp, vp, vp, i64, i, i, i64]
        L.oracle_erase.restype, L.oracle_erase.argtypes = None, [vp, vp, i64]
        L.oracle_retrieve_all.restype, L.oracle_retrieve_all.argtypes = i64, [vp, vp, vp]
        L.oracle_probe_sequence.restype, L.oracle_probe_sequence.argtypes = None, [vp, i64, i, vp, i]
        L.oracle_set_allows_duplicates.restype, L.oracle_set_allows_duplicates.argtypes = None, [vp, i]
        L.oracle_count.restype, L.oracle_count.argtypes = i64, [vp, vp, i64, i]
        L.oracle_retrieve.restype, L.oracle_retrieve.argtypes = i64, [vp, vp, i64, i, vp, vp, vp]
        _lib = L
    return _lib


def _bytes_of(value, dtype) -> bytes:
    return np.asarray(value, dtype=dtype).tobytes()


def xxhash32(data: bytes, seed=0) -> int:
    return int(lib().oracle_xxhash32(data, len(data), seed))


def xxhash64(data: bytes, seed=0) -> int:
    return int(lib().oracle_xxhash64(data, len(data), seed))


def murmur3_32(data: bytes, seed=0) -> int:
    return int(lib().oracle_murmur3_32(data, len(data), seed))


def murmur3_x64_128(data: bytes, seed=0):
    out = (C.c_uint64 * 2)()
    lib().oracle_murmur3_x64_128(data, len(data), seed, out)
    return [int(out[0]), int(out[1])]


def murmur3_x86_128(data: bytes, seed=0):
    out = (C.c_uint32 * 4)()
    lib().oracle_murmur3_x86_128(data, len(data), seed, out)
    return [int(v) for v in out]


def _i64(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a).astype(np.int64, copy=False))


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


# geometry of the C-ABI kinds (include/cuco_b200.h) in oracle terms:
# kind -> (key_bytes, value_bytes, cg, w, probing, hash)
KIND_GEOMETRY = {
    0: (4, 0, 4, 1, DOUBLE, XXHASH32),
    1: (8, 8, 1, 1, LINEAR, XXHASH32),
    2: (8, 8, 8, 1, DOUBLE, XXHASH32),
    3: (4, 4, 4, 1, LINEAR, XXHASH32),
    4: (8, 8, 4, 1, LINEAR, XXHASH32),
    5: (8, 0, 4, 1, DOUBLE, XXHASH32),
    6: (8, 8, 1, 2, LINEAR, XXHASH32),
    7: (4, 4, 2, 2, DOUBLE, MURMUR3_32),
    8: (4, 8, 1, 1, LINEAR, XXHASH32),
    9: (8, 8, 8, 1, DOUBLE, XXHASH64),
    10: (4, 0, 4, 2, DOUBLE, XXHASH32),
    11: (8, 0, 1, 2, LINEAR, XXHASH32),
    12: (8, 8, 4, 1, LINEAR, XXHASH32),
    13: (8, 8, 1, 1, LINEAR, XXHASH64),
}
MULTI_KINDS = {10, 11, 12}  # static_multiset / multimap instantiations: equal keys are stored repeatedly


class Table:
    """Sequential CPU table with the reference's semantics. Arrays in and out are numpy int64."""

    def __init__(self, key_bytes, value_bytes, cg, w, probing, hash, size, load_factor=0.0,
                 empty_key=-1, empty_value=-1, erased_key=None, allows_duplicates=False):
        self.is_map = value_bytes != 0
        self._h = lib().oracle_create(key_bytes, value_bytes, cg, w, probing, hash, int(size),
                                      float(load_factor), int(empty_key), int(empty_value),
                                      0 if erased_key is None else 1,
                                      0 if erased_key is None else int(erased_key))
        if not self._h:
            raise ValueError("oracle_create rejected the arguments")
        if allows_duplicates:
            lib().oracle_set_allows_duplicates(self._h, 1)

    @classmethod
    def for_kind(cls, kind, size, load_factor=0.0, empty_key=-1, empty_value=-1, erased_key=None):
        return cls(*KIND_GEOMETRY[kind], size, load_factor, empty_key, empty_value, erased_key,
                   allows_duplicates=kind in MULTI_KINDS)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().oracle_destroy(self._h)
            self._h = None

    def capacity(self):
        return int(lib().oracle_capacity(self._h))

    def size(self):
        return int(lib().oracle_size(self._h))

    def clear(self):
        lib().oracle_clear(self._h)

    def insert(self, keys, values=None):
        k = _i64(keys)
        v = None if values is None else _i64(values)
        return int(lib().oracle_insert(self._h, _p(k), _p(v), k.size))

    def insert_if(self, keys, stencil, values=None):
        k = _i64(keys)
        v = None if values is None else _i64(values)
        s = np.ascontiguousarray(np.asarray(stencil).astype(np.uint8))
        return int(lib().oracle_insert_if(self._h, _p(k), _p(v), _p(s), k.size))

    def find(self, keys):
        k = _i64(keys)
        out = np.empty(k.size, dtype=np.int64)
        lib().oracle_find(self._h, _p(k), _p(out), k.size)
        return out

    def contains(self, keys, stencil=None):
        k = _i64(keys)
        out = np.empty(k.size, dtype=np.uint8)
        if stencil is None:
            lib().oracle_contains(self._h, _p(k), _p(out), k.size)
        else:
            s = np.ascontiguousarray(np.asarray(stencil).astype(np.uint8))
            lib().oracle_contains_if(self._h, _p(k), _p(s), _p(out), k.size)
        return out.astype(bool)

    def contains_if(self, keys, stencil):
        return self.contains(keys, stencil)

    def insert_and_find(self, keys, values=None):
        k = _i64(keys)
        v = None if values is None else _i64(values)
        found = np.empty(k.size, dtype=np.int64)
        ins = np.empty(k.size, dtype=np.uint8)
        lib().oracle_insert_and_find(self._h, _p(k), _p(v), _p(found), _p(ins), k.size)
        return found, ins.astype(bool)

    def insert_or_assign(self, keys, values):
        k, v = _i64(keys), _i64(values)
        lib().oracle_insert_or_assign(self._h, _p(k), _p(v), k.size)

    def insert_or_apply(self, keys, values, op=PLUS, init=None):
        k, v = _i64(keys), _i64(values)
        lib().oracle_insert_or_apply(self._h, _p(k), _p(v), k.size, op,
                                     0 if init is None else 1, 0 if init is None else int(init))

    def erase(self, keys):
        k = _i64(keys)
        lib().oracle_erase(self._h, _p(k), k.size)

    def retrieve_all(self):
        cap = self.capacity()
        k = np.empty(cap, dtype=np.int64)
        v = np.empty(cap, dtype=np.int64)
        n = int(lib().oracle_retrieve_all(self._h, _p(k), _p(v)))
        return (k[:n], v[:n]) if self.is_map else k[:n]

    def count(self, keys, outer=False):
        k = _i64(keys)
        return int(lib().oracle_count(self._h, _p(k), k.size, 1 if outer else 0))

    def retrieve(self, keys, outer=False):
        """(probe keys, matched keys[, matched payloads]) in input order / probe order."""
        k = _i64(keys)
        rows = self.count(k, outer)
        probe = np.empty(rows, dtype=np.int64)
        mk = np.empty(rows, dtype=np.int64)
        mv = np.empty(rows, dtype=np.int64) if self.is_map else None
        n = int(lib().oracle_retrieve(self._h, _p(k), k.size, 1 if outer else 0, _p(probe), _p(mk), _p(mv)))
        assert n == rows
        return (probe, mk, mv) if self.is_map else (probe, mk)

    def probe_sequence(self, key, rank=0, length=8):
        out = np.empty(length, dtype=np.int64)
        lib().oracle_probe_sequence(self._h, int(key), rank, _p(out), length)
        return out


# ---- host std::unordered_map baseline ------------------------------------------------------------
_baseline = None


def baseline() -> C.CDLL:
    global _baseline
    if _baseline is None:
        L = _load(_BASELINE_PATH)
        vp, i64, dbl, i = C.c_void_p, C.c_int64, C.c_double, C.c_int
        L.cpu_baseline_map_i64.restype = i
        L.cpu_baseline_map_i64.argtypes = [vp, vp, i64, vp, i64, i, dbl, C.POINTER(dbl), C.POINTER(dbl), C.POINTER(i64)]
        L.cpu_baseline_set_i32.restype = i
        L.cpu_baseline_set_i32.argtypes = [vp, i64, vp, i64, i, dbl, C.POINTER(dbl), C.POINTER(dbl), C.POINTER(i64)]
        L.cpu_baseline_hardware_threads.restype = i
        _baseline = L
    return _baseline


def baseline_map_i64(keys, values, queries, threads, load_factor):
    """Returns (insert_seconds, find_seconds, checksum) of the sharded std::unordered_map baseline."""
    k, v, q = _i64(keys), _i64(values), _i64(queries)
    ti, tf, cs = C.c_double(), C.c_double(), C.c_int64()
    baseline().cpu_baseline_map_i64(_p(k), _p(v), k.size, _p(q), q.size, threads, load_factor,
                                    C.byref(ti), C.byref(tf), C.byref(cs))
    return ti.value, tf.value, cs.value


def hardware_threads() -> int:
    return int(baseline().cpu_baseline_hardware_threads())
