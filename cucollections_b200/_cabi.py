"""ctypes binding of include/cuco_b200.h.

`Library(path)` binds one shared object; the product library is `native()`. The parity tests bind
oracle/_ref/libcuco_ref.so (cuco's own build of the same shim) through the very same class, which is
what makes the two comparable call for call. There is no fallback: if the CUDA library is missing,
using the product path raises.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_PKG = Path(__file__).resolve().parent
NATIVE_PATH = _PKG / "libcuco_b200.so"
REFERENCE_PATH = _PKG.parent / "oracle" / "_ref" / "libcuco_ref.so"

# enum cuco_b200_kind
SET_I32_DH4 = 0
MAP_I64_LP1 = 1
MAP_I64_DH8 = 2
MAP_I32_LP4 = 3
MAP_I64_LP4 = 4
SET_I64_DH4 = 5
MAP_I64_LP1_W2 = 6
MAP_I32_DH2_W2_MM = 7
MAP_I32I64_LP1 = 8
MAP_I64_DH8_X64 = 9
MULTISET_I32_DH4_W2 = 10
MULTISET_I64_LP1_W2 = 11
MULTIMAP_I64_LP4 = 12
MAP_I64_LP1_X64 = 13
NUM_KINDS = 14

PLUS, MIN, MAX = 0, 1, 2

_vp, _i64, _int, _dbl, _u64 = C.c_void_p, C.c_int64, C.c_int, C.c_double, C.c_uint64
_pi64 = C.POINTER(C.c_int64)
_u32 = C.c_uint32
_pu32 = C.POINTER(C.c_uint32)
_pvp = C.POINTER(C.c_void_p)

_PROTOTYPES = {
    "cuco_b200_build_info": (C.c_char_p, []),
    "cuco_b200_last_error": (C.c_char_p, []),
    "cuco_b200_create": (_int, [_int, _i64, _dbl, _i64, _i64, _int, _i64, _vp, C.POINTER(_vp)]),
    "cuco_b200_destroy": (_int, [_vp]),
    "cuco_b200_kind_of": (_int, [_vp]),
    "cuco_b200_key_bytes": (_int, [_vp]),
    "cuco_b200_value_bytes": (_int, [_vp]),
    "cuco_b200_capacity": (_i64, [_vp]),
    "cuco_b200_size": (_int, [_vp, _vp, _pi64]),
    "cuco_b200_clear": (_int, [_vp, _vp]),
    "cuco_b200_insert": (_int, [_vp, _vp, _vp, _i64, _vp, _pi64]),
    "cuco_b200_insert_if": (_int, [_vp, _vp, _vp, _vp, _i64, _vp, _pi64]),
    "cuco_b200_find": (_int, [_vp, _vp, _vp, _i64, _vp]),
    "cuco_b200_contains": (_int, [_vp, _vp, _vp, _i64, _vp]),
    "cuco_b200_contains_if": (_int, [_vp, _vp, _vp, _vp, _i64, _vp]),
    "cuco_b200_insert_and_find": (_int, [_vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "cuco_b200_insert_or_assign": (_int, [_vp, _vp, _vp, _i64, _vp]),
    "cuco_b200_insert_or_apply": (_int, [_vp, _vp, _vp, _i64, _int, _int, _i64, _vp]),
    "cuco_b200_erase": (_int, [_vp, _vp, _i64, _vp]),
    "cuco_b200_retrieve_all": (_int, [_vp, _vp, _vp, _pi64, _vp]),
    "cuco_b200_rehash": (_int, [_vp, _i64, _vp]),
    "cuco_b200_count": (_int, [_vp, _vp, _i64, _int, _vp, _pi64]),
    "cuco_b200_retrieve": (_int, [_vp, _vp, _i64, _int, _vp, _vp, _pi64, _vp]),
    "cuco_b200_insert_host": (_int, [_vp, _vp, _vp, _i64, _vp]),
    "cuco_b200_find_host": (_int, [_vp, _vp, _vp, _i64, _vp]),
    "cuco_b200_contains_host": (_int, [_vp, _vp, _vp, _i64, _vp]),
    "cuco_b200_insert_and_find_host": (_int, [_vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "cuco_b200_exchange_plan": (_int, [_vp, _i64, _int, _pu32, _pu32, _pu32]),
    "cuco_b200_exchange_route": (_int, [_vp, _vp, _vp, _i64, _int, _u32, _u32, _u32, _int, _int, _u64,
                                        _pvp, _pvp, _pvp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "cuco_b200_exchange_mutate": (_int, [_vp, _vp, _vp, _u32, _u32, _int, _int, _vp]),
    "cuco_b200_exchange_lookup": (_int, [_vp, _vp, _vp, _pvp, _u32, _u32, _int, _int, _int, _vp]),
    "cuco_b200_exchange_unpermute": (_int, [_vp, _vp, _vp, _i64, _vp, _int, _vp]),
    "cuco_b200_exchange_stage_plan": (_int, [_vp, _i64, _int, _int, _pu32, _pu32]),
    "cuco_b200_exchange_stage": (_int, [_vp, _vp, _vp, _i64, _int, _int, _u32, _u32, _int, _int, _u64,
                                        _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "cuco_b200_exchange_publish": (_int, [_vp, _vp, _pvp, _pvp, _int, _u32, _int, _int, _int, _vp]),
    "cuco_b200_exchange_fine_regions": (_int, [_vp, _int, _pu32]),
    "cuco_b200_exchange_probe": (_int, [_vp, _vp, _vp, _u32, _u32, _int, _u32, _u32, _int, _vp]),
    "cuco_b200_copy_async": (_int, [_vp, _vp, _i64, _vp]),
    "cuco_b200_push_async": (_int, [_pvp, _pvp, C.POINTER(C.c_int64), _int, _int, _vp]),
    "cuco_b200_exchange_apply": (_int, [_vp, _vp, _vp, _u32, _int, _int, _int, _int, _vp]),
    "cuco_b200_exchange_lookup_local": (_int, [_vp, _vp, _vp, _vp, _u32, _int, _int, _vp]),
    "cuco_b200_set_tuning": (_int, [_int, _int, _int, _int, _int, _int, _int]),
    "cuco_b200_set_blocking": (_int, [_int, _int]),
    "cuco_b200_set_blocking_variant": (_int, [_int, _int, _int]),
    "cuco_b200_set_stream_variant": (_int, [_int, _int, _int]),
    "cuco_b200_partition_count": (_int, [_vp, _int, _int, _i64, _int, _u64, _vp, _vp]),
    "cuco_b200_partition_scatter": (
        _int, [_vp, _vp, _int, _int, _int, _i64, _int, _u64, _vp, _vp, _vp, _vp, _vp]),
    "cuco_b200_scatter_by_index": (_int, [_vp, _vp, _vp, _int, _i64, _vp]),
}

EXPORTED_SYMBOLS = tuple(_PROTOTYPES)


class CucoError(RuntimeError):
    """A C-ABI call returned a non-zero status (message from cuco_b200_last_error)."""

    def __init__(self, code: int, message: str):
        super().__init__(f"[{code}] {message}")
        self.code = code


class Library:
    def __init__(self, path: os.PathLike | str):
        path = Path(path)
        if not path.exists():
            raise FileNotFoundError(
                f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`"
                " (there is no CPU fallback for the hash table path)")
        self.path = path
        self._dll = C.CDLL(str(path), mode=os.RTLD_LOCAL | os.RTLD_NOW)
        for name, (restype, argtypes) in _PROTOTYPES.items():
            fn = getattr(self._dll, name)  # AttributeError if a declared symbol is not exported
            fn.restype = restype
            fn.argtypes = argtypes
            setattr(self, name[len("cuco_b200_"):], fn)

    @property
    def flavour(self) -> str:
        return self.build_info().decode()

    def check(self, status: int) -> None:
        if status != 0:
            raise CucoError(status, self.last_error().decode())


_native: Library | None = None
_reference: Library | None = None


def native() -> Library:
    """The product library (hand-written sm_100a kernels). Raises if it has not been built."""
    global _native
    if _native is None:
        # CUCO_B200_LIB: a development build of the same library (cucollections_b200/build.py dev)
        _native = Library(os.environ.get("CUCO_B200_LIB") or NATIVE_PATH)
    return _native


def reference() -> Library:
    """cuco's own build of the same shim (test / bench baseline only)."""
    global _reference
    if _reference is None:
        _reference = Library(REFERENCE_PATH)
    return _reference
