"""b200-hashtable: Blackwell-native open-addressing hash table behind cuCollections' surface.

The product is the C++ header tree in include/cuco (drop-in for the reference's static_map /
static_set) and the C-ABI library built from it (libcuco_b200.so, declared in include/cuco_b200.h).
This package is the thin host mirror of that API for Python callers, tests and the benchmark:
torch supplies device memory, streams and torch.distributed; every table operation runs in the
hand-written sm_100a kernels. There is no CPU fallback.
"""
from . import _cabi
from ._cabi import CucoError, Library, native
from .containers import KINDS, find_kind, static_map, static_multimap, static_multiset, static_set

__all__ = ["CucoError", "Library", "native", "KINDS", "find_kind", "static_map", "static_multimap", "static_multiset", "static_set",
           "_cabi"]
