"""Host-side mirror of cuco::static_map / static_set / static_multiset over the C ABI, with torch tensors as the
device buffers and torch's current CUDA stream as the `cuda::stream_ref`.

Names and argument meaning follow the reference classes (include/cuco/static_map.cuh:88-986,
include/cuco/static_set.cuh:82-798): `insert` returns the number of new keys and synchronises,
`insert_async` / `find` / `contains` / ... only enqueue work, sentinels are given at construction.
Only the explicit instantiations listed in include/cuco_b200.h are reachable from here; anything
else needs the C++ headers directly.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import torch

from . import _cabi


@dataclass(frozen=True)
class _Kind:
    kind: int
    key: torch.dtype
    value: torch.dtype | None
    probing: str
    cg_size: int
    window_size: int
    hash: str
    multi: bool = False  # static_multiset: equal keys are stored repeatedly


KINDS = {
    k.kind: k
    for k in (
        _Kind(_cabi.SET_I32_DH4, torch.int32, None, "double_hashing", 4, 1, "xxhash_32"),
        _Kind(_cabi.MAP_I64_LP1, torch.int64, torch.int64, "linear_probing", 1, 1, "xxhash_32"),
        _Kind(_cabi.MAP_I64_DH8, torch.int64, torch.int64, "double_hashing", 8, 1, "xxhash_32"),
        _Kind(_cabi.MAP_I32_LP4, torch.int32, torch.int32, "linear_probing", 4, 1, "xxhash_32"),
        _Kind(_cabi.MAP_I64_LP4, torch.int64, torch.int64, "linear_probing", 4, 1, "xxhash_32"),
        _Kind(_cabi.SET_I64_DH4, torch.int64, None, "double_hashing", 4, 1, "xxhash_32"),
        _Kind(_cabi.MAP_I64_LP1_W2, torch.int64, torch.int64, "linear_probing", 1, 2, "xxhash_32"),
        _Kind(_cabi.MAP_I32_DH2_W2_MM, torch.int32, torch.int32, "double_hashing", 2, 2,
              "murmurhash3_32"),
        _Kind(_cabi.MAP_I32I64_LP1, torch.int32, torch.int64, "linear_probing", 1, 1, "xxhash_32"),
        _Kind(_cabi.MAP_I64_DH8_X64, torch.int64, torch.int64, "double_hashing", 8, 1, "xxhash_64"),
        _Kind(_cabi.MULTISET_I32_DH4_W2, torch.int32, None, "double_hashing", 4, 2, "xxhash_32", True),
        _Kind(_cabi.MULTISET_I64_LP1_W2, torch.int64, None, "linear_probing", 1, 2, "xxhash_32", True),
        _Kind(_cabi.MULTIMAP_I64_LP4, torch.int64, torch.int64, "linear_probing", 4, 1, "xxhash_32", True),
        _Kind(_cabi.MAP_I64_LP1_X64, torch.int64, torch.int64, "linear_probing", 1, 1, "xxhash_64"),
    )
}


def find_kind(key, value, probing, cg_size, window_size, hash="xxhash_32", multi=False) -> int:
    for k in KINDS.values():
        if (k.key, k.value, k.probing, k.cg_size, k.window_size, k.hash, k.multi) == (
                key, value, probing, cg_size, window_size, hash, multi):
            return k.kind
    raise ValueError(
        f"no explicit instantiation for key={key} value={value} {probing}<{cg_size}> "
        f"storage<{window_size}> {hash}; see include/cuco_b200.h")


def _ptr(t: torch.Tensor | None):
    return None if t is None else C.c_void_p(t.data_ptr())


class _Table:
    """Shared plumbing of static_map / static_set."""

    def __init__(self, kind, size, load_factor, empty_key, empty_value, erased_key, device, library):
        self._handle = None
        self._lib = library or _cabi.native()
        self.kind = KINDS[kind]
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("the table lives in GPU memory; pass a cuda device")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.empty_key_sentinel = int(empty_key)
        self.empty_value_sentinel = int(empty_value)
        self.erased_key_sentinel = self.empty_key_sentinel if erased_key is None else int(erased_key)
        handle = C.c_void_p()
        with torch.cuda.device(self.device):
            self._lib.check(self._lib.create(
                kind, int(size), float(load_factor or 0.0), int(empty_key), int(empty_value),
                0 if erased_key is None else 1, 0 if erased_key is None else int(erased_key),
                self._stream(), C.byref(handle)))
        self._handle = handle

    # -- helpers --------------------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _check_keys(self, keys: torch.Tensor, what="keys") -> torch.Tensor:
        if keys.device != self.device:
            raise ValueError(f"{what} must live on {self.device}, got {keys.device}")
        if keys.dtype != self.kind.key:
            raise TypeError(f"{what} must be {self.kind.key}, got {keys.dtype}")
        return keys.contiguous()

    def _input(self, keys, values):
        """Returns (keys tensor, values tensor or None, n) for insert-like calls."""
        if self.kind.value is None:
            k = self._check_keys(keys)
            return k, None, k.numel()
        if values is None:
            # AoS: [n, 2] tensor of the key dtype == cuco::pair<Key, T> when key and payload match
            if keys.dim() != 2 or keys.shape[1] != 2 or self.kind.key != self.kind.value:
                raise ValueError(
                    "pass an [n, 2] pair tensor (same key/payload dtype) or keys and values")
            k = self._check_keys(keys, "pairs")
            if k.data_ptr() % (2 * k.element_size()) != 0:
                raise ValueError("pair tensor must be aligned to the pair size")
            return k, None, k.shape[0]
        k = self._check_keys(keys)
        if (values.dtype != self.kind.value or values.device != self.device
                or values.numel() != k.numel()):
            raise TypeError(
                f"values must be {self.kind.value} on {self.device} with one entry per key")
        return k, values.contiguous(), k.numel()

    def _call(self, fn, *args):
        with torch.cuda.device(self.device):
            self._lib.check(fn(self._handle, *args))

    def _payload_dtype(self):
        return self.kind.value if self.kind.value is not None else self.kind.key

    # -- API --------------------------------------------------------------------------------------
    def capacity(self) -> int:
        return int(self._lib.capacity(self._handle))

    def size(self) -> int:
        out = C.c_int64()
        self._call(self._lib.size, self._stream(), C.byref(out))
        return out.value

    def clear(self) -> None:
        self.clear_async()
        torch.cuda.current_stream(self.device).synchronize()

    def clear_async(self) -> None:
        self._call(self._lib.clear, self._stream())

    def insert(self, keys, values=None) -> int:
        k, v, n = self._input(keys, values)
        out = C.c_int64()
        self._call(self._lib.insert, _ptr(k), _ptr(v), n, self._stream(), C.byref(out))
        return out.value

    def insert_async(self, keys, values=None) -> None:
        k, v, n = self._input(keys, values)
        self._call(self._lib.insert, _ptr(k), _ptr(v), n, self._stream(), None)

    def insert_if(self, keys, stencil, values=None) -> int:
        k, v, n = self._input(keys, values)
        st = stencil.to(torch.uint8).contiguous()
        out = C.c_int64()
        self._call(self._lib.insert_if, _ptr(k), _ptr(v), _ptr(st), n, self._stream(), C.byref(out))
        return out.value

    def insert_if_async(self, keys, stencil, values=None) -> None:
        k, v, n = self._input(keys, values)
        st = stencil.to(torch.uint8).contiguous()
        self._call(self._lib.insert_if, _ptr(k), _ptr(v), _ptr(st), n, self._stream(), None)

    def find(self, keys, out=None) -> torch.Tensor:
        """out[i] = payload (map) / stored key (set) or the empty sentinel. Stream-ordered."""
        k = self._check_keys(keys)
        if out is None:
            out = torch.empty(k.numel(), dtype=self._payload_dtype(), device=self.device)
        self._call(self._lib.find, _ptr(k), _ptr(out), k.numel(), self._stream())
        return out

    find_async = find

    def contains(self, keys, out=None) -> torch.Tensor:
        k = self._check_keys(keys)
        if out is None:
            out = torch.empty(k.numel(), dtype=torch.bool, device=self.device)
        self._call(self._lib.contains, _ptr(k), _ptr(out), k.numel(), self._stream())
        return out

    contains_async = contains

    def contains_if(self, keys, stencil, out=None) -> torch.Tensor:
        k = self._check_keys(keys)
        st = stencil.to(torch.uint8).contiguous()
        if out is None:
            out = torch.empty(k.numel(), dtype=torch.bool, device=self.device)
        self._call(self._lib.contains_if, _ptr(k), _ptr(st), _ptr(out), k.numel(), self._stream())
        return out

    def insert_and_find(self, keys, values=None):
        """Returns (found, inserted): resident payload/key per element and whether it created it."""
        k, v, n = self._input(keys, values)
        found = torch.empty(n, dtype=self._payload_dtype(), device=self.device)
        inserted = torch.empty(n, dtype=torch.bool, device=self.device)
        self._call(self._lib.insert_and_find, _ptr(k), _ptr(v), _ptr(found), _ptr(inserted), n,
                   self._stream())
        return found, inserted

    insert_and_find_async = insert_and_find

    # -- host-buffer variants (cuco_b200_*_host): chunked, copies overlapped with the kernels --------
    def _host_keys(self, keys: torch.Tensor, what="keys") -> torch.Tensor:
        if keys.device.type != "cpu":
            raise ValueError(f"{what} must be a CPU tensor (pinned for full PCIe speed)")
        if keys.dtype != self.kind.key:
            raise TypeError(f"{what} must be {self.kind.key}, got {keys.dtype}")
        return keys.contiguous()

    def insert_host(self, keys, values=None) -> None:
        """insert_async of a batch living in host memory. Stream-ordered."""
        if self.kind.value is not None and values is None:
            if keys.dim() != 2 or keys.shape[1] != 2 or self.kind.key != self.kind.value:
                raise ValueError("pass an [n, 2] pair tensor (same key/payload dtype) or keys and values")
            k, v, n = self._host_keys(keys, "pairs"), None, keys.shape[0]
        else:
            k = self._host_keys(keys)
            v, n = (None if values is None else values.contiguous()), k.numel()
            if v is not None and (v.dtype != self.kind.value or v.device.type != "cpu" or v.numel() != n):
                raise TypeError(f"values must be a CPU tensor of {self.kind.value} with one entry per key")
        self._call(self._lib.insert_host, _ptr(k), _ptr(v), n, self._stream())

    def find_host(self, keys, out=None) -> torch.Tensor:
        """find_async from host keys into a host output tensor. Stream-ordered: synchronise the
        current stream before reading `out`."""
        k = self._host_keys(keys)
        if out is None:
            out = torch.empty(k.numel(), dtype=self._payload_dtype(), pin_memory=True)
        self._call(self._lib.find_host, _ptr(k), _ptr(out), k.numel(), self._stream())
        return out

    def contains_host(self, keys, out=None) -> torch.Tensor:
        k = self._host_keys(keys)
        if out is None:
            out = torch.empty(k.numel(), dtype=torch.bool, pin_memory=True)
        self._call(self._lib.contains_host, _ptr(k), _ptr(out), k.numel(), self._stream())
        return out

    def erase(self, keys) -> None:
        k = self._check_keys(keys)
        self._call(self._lib.erase, _ptr(k), k.numel(), self._stream())

    erase_async = erase

    def rehash(self, capacity: int | None = None) -> None:
        self._call(self._lib.rehash, -1 if capacity is None else int(capacity), self._stream())

    def close(self) -> None:
        if getattr(self, "_handle", None):
            with torch.cuda.device(self.device):
                self._lib.destroy(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


class static_map(_Table):
    """cuco::static_map<Key, T, extent<size_t>, thread_scope_device, equal_to, Probing, .., storage<W>>.

    Exactly one of `capacity` or (`n`, `load_factor`) sizes the table, as with the reference
    constructors (static_map.cuh:160-260)."""

    def __init__(self, capacity=None, *, n=None, load_factor=None, key_dtype=torch.int64,
                 value_dtype=torch.int64, empty_key=-1, empty_value=-1, erased_key=None,
                 probing="linear_probing", cg_size=4, window_size=1, hash="xxhash_32",
                 device=None, _library=None):
        if (capacity is None) == (n is None):
            raise ValueError("give either capacity or n (+ load_factor)")
        if n is not None and load_factor is None:
            raise ValueError("n needs a load_factor")
        kind = find_kind(key_dtype, value_dtype, probing, cg_size, window_size, hash)
        super().__init__(kind, capacity if n is None else n, load_factor if n is not None else 0.0,
                         empty_key, empty_value, erased_key, device, _library)

    def insert_or_assign(self, keys, values=None) -> None:
        k, v, n = self._input(keys, values)
        self._call(self._lib.insert_or_assign, _ptr(k), _ptr(v), n, self._stream())

    insert_or_assign_async = insert_or_assign

    def insert_or_apply(self, keys, values=None, *, op="plus", init=None) -> None:
        """payload[key] = fold of op over all payloads carrying key (cuco::reduce::plus/min/max)."""
        k, v, n = self._input(keys, values)
        code = {"plus": _cabi.PLUS, "min": _cabi.MIN, "max": _cabi.MAX}[op]
        self._call(self._lib.insert_or_apply, _ptr(k), _ptr(v), n, code,
                   0 if init is None else 1, 0 if init is None else int(init), self._stream())

    insert_or_apply_async = insert_or_apply

    def retrieve_all(self):
        cap = self.capacity()
        keys = torch.empty(cap, dtype=self.kind.key, device=self.device)
        vals = torch.empty(cap, dtype=self.kind.value, device=self.device)
        n = C.c_int64()
        self._call(self._lib.retrieve_all, _ptr(keys), _ptr(vals), C.byref(n), self._stream())
        return keys[: n.value], vals[: n.value]


class static_set(_Table):
    """cuco::static_set<Key, extent<size_t>, thread_scope_device, equal_to, Probing, .., storage<W>>."""

    def __init__(self, capacity=None, *, n=None, load_factor=None, key_dtype=torch.int32,
                 empty_key=-1, erased_key=None, probing="double_hashing", cg_size=4, window_size=1,
                 hash="xxhash_32", device=None, _library=None):
        if (capacity is None) == (n is None):
            raise ValueError("give either capacity or n (+ load_factor)")
        if n is not None and load_factor is None:
            raise ValueError("n needs a load_factor")
        kind = find_kind(key_dtype, None, probing, cg_size, window_size, hash)
        super().__init__(kind, capacity if n is None else n, load_factor if n is not None else 0.0,
                         empty_key, 0, erased_key, device, _library)

    def retrieve_all(self):
        cap = self.capacity()
        keys = torch.empty(cap, dtype=self.kind.key, device=self.device)
        n = C.c_int64()
        self._call(self._lib.retrieve_all, _ptr(keys), None, C.byref(n), self._stream())
        return keys[: n.value]

    def retrieve(self, keys):
        """static_set::retrieve (static_set.cuh:620): (probe keys that have a match, matched keys),
        row order unspecified."""
        k = self._check_keys(keys)
        probed = torch.empty_like(k)
        matched = torch.empty_like(k)
        n = C.c_int64()
        self._call(self._lib.retrieve, _ptr(k), k.numel(), 0, _ptr(probed), _ptr(matched),
                   C.byref(n), self._stream())
        return probed[: n.value], matched[: n.value]


class static_multimap(_Table):
    """cuco::experimental::static_multimap<Key, T, ...> (static_multimap.cuh:45-549): a key may map to
    any number of payloads. Bulk API of the reference class: `insert[_async]`, `insert_if`,
    `contains`, `contains_if`, `count`."""

    def __init__(self, capacity=None, *, n=None, load_factor=None, key_dtype=torch.int64,
                 value_dtype=torch.int64, empty_key=-1, empty_value=-1, probing="linear_probing",
                 cg_size=4, window_size=1, hash="xxhash_32", device=None, _library=None):
        if (capacity is None) == (n is None):
            raise ValueError("give either capacity or n (+ load_factor)")
        if n is not None and load_factor is None:
            raise ValueError("n needs a load_factor")
        kind = find_kind(key_dtype, value_dtype, probing, cg_size, window_size, hash, multi=True)
        super().__init__(kind, capacity if n is None else n, load_factor if n is not None else 0.0,
                         empty_key, empty_value, None, device, _library)

    def count(self, keys) -> int:
        """Total number of stored pairs whose key equals a probe key. Synchronises."""
        k = self._check_keys(keys)
        out = C.c_int64()
        self._call(self._lib.count, _ptr(k), k.numel(), 0, self._stream(), C.byref(out))
        return out.value


class static_multiset(_Table):
    """cuco::static_multiset<Key, extent<size_t>, thread_scope_device, equal_to, Probing, ..,
    storage<W>> (static_multiset.cuh:81-729): equal keys are stored as often as they are inserted.
    `insert` stores every element; `count` / `retrieve` enumerate all matches of the probe keys."""

    def __init__(self, capacity=None, *, n=None, load_factor=None, key_dtype=torch.int32,
                 empty_key=-1, probing="double_hashing", cg_size=4, window_size=2,
                 hash="xxhash_32", device=None, _library=None):
        if (capacity is None) == (n is None):
            raise ValueError("give either capacity or n (+ load_factor)")
        if n is not None and load_factor is None:
            raise ValueError("n needs a load_factor")
        kind = find_kind(key_dtype, None, probing, cg_size, window_size, hash, multi=True)
        super().__init__(kind, capacity if n is None else n, load_factor if n is not None else 0.0,
                         empty_key, 0, None, device, _library)

    def count(self, keys, outer=False) -> int:
        """Total number of stored elements equal to the probe keys; `outer`: a key without matches
        counts as one (count / count_outer, static_multiset.cuh:615,661). Synchronises."""
        k = self._check_keys(keys)
        out = C.c_int64()
        self._call(self._lib.count, _ptr(k), k.numel(), 1 if outer else 0, self._stream(),
                   C.byref(out))
        return out.value

    def retrieve(self, keys, outer=False):
        """(probe keys, matched elements), one row per match, row order unspecified; `outer` adds
        {key, empty key sentinel} for keys without matches (retrieve / retrieve_outer,
        static_multiset.cuh:506,593). Sized with `count`. Synchronises."""
        k = self._check_keys(keys)
        rows = self.count(k, outer)
        probed = torch.empty(rows, dtype=self.kind.key, device=self.device)
        matched = torch.empty(rows, dtype=self.kind.key, device=self.device)
        n = C.c_int64()
        self._call(self._lib.retrieve, _ptr(k), k.numel(), 1 if outer else 0, _ptr(probed),
                   _ptr(matched), C.byref(n), self._stream())
        assert n.value == rows, (n.value, rows)
        return probed, matched
