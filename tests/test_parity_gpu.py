"""GPU parity tests: the sm_100a table (through the C ABI) against
  * the CPU oracle (oracle/cuco_oracle.c, sequential restatement of the reference algorithm), and
  * cuco itself (oracle/_ref/libcuco_ref.so: the reference headers behind the same shim)
on the same seeded inputs. Integer work: every comparison is bit-exact.

Mirrors the reference's own suites (tests/static_map/{unique_sequence,insert_and_find,
insert_or_assign,insert_or_apply,duplicate_keys,erase,rehash}_test.cu and tests/static_set/*).
"""
import numpy as np
import pytest
import torch

import cucollections_b200 as cb
from cucollections_b200 import _cabi
from cucollections_b200 import partitioned as cbp
from oracle import oracle

pytestmark = pytest.mark.gpu

MAP_KINDS = [_cabi.MAP_I64_LP1, _cabi.MAP_I64_DH8, _cabi.MAP_I32_LP4, _cabi.MAP_I64_LP4,
             _cabi.MAP_I64_LP1_W2, _cabi.MAP_I32_DH2_W2_MM, _cabi.MAP_I32I64_LP1,
             _cabi.MAP_I64_DH8_X64, _cabi.MAP_I64_LP1_X64]
SET_KINDS = [_cabi.SET_I32_DH4, _cabi.SET_I64_DH4]
TUNINGS = [  # (keys_per_thread, cas_first, sector_chunks, force_generic, coherent_loads, waves)
    (12, 0, 1, 0, 0, 0), (1, 0, 0, 0, 1, 1), (4, 1, 1, 0, 0, 2), (4, 0, 1, 0, 1, 0), (2, 1, 1, 1, 0, 0),
    (1, 1, 1, 0, 0, 1), (2, 0, 1, 0, 0, 1)]


def make(kind, lib, **kw):
    k = cb.KINDS[kind]
    common = dict(key_dtype=k.key, probing=k.probing, cg_size=k.cg_size, window_size=k.window_size,
                  hash=k.hash, _library=lib, **kw)
    if k.value is None:
        return cb.static_set(**common)
    return cb.static_map(value_dtype=k.value, **common)


def dev(a, dtype):
    return torch.from_numpy(np.ascontiguousarray(a)).to("cuda").to(dtype)


def keyset(kind, n, seed, hi=None):
    rng = np.random.default_rng(seed)
    k = cb.KINDS[kind]
    top = hi or (2**31 - 2 if k.key == torch.int32 else 2**62)
    return rng.integers(0, top, size=n, dtype=np.int64)


@pytest.fixture(autouse=True)
def reset_tuning(native_lib):
    yield
    native_lib.set_tuning(12, 0, 1, 0, 0, 1, 0)


@pytest.mark.parametrize("tuning", TUNINGS)
@pytest.mark.parametrize("kind", MAP_KINDS)
def test_map_insert_find_contains_match_oracle(kind, tuning, native_lib):
    native_lib.set_tuning(tuning[0], tuning[1], tuning[2], tuning[5], tuning[3], 1, tuning[4])
    k = cb.KINDS[kind]
    n = 20_000
    keys = keyset(kind, n, 1, hi=n)          # duplicates: ~63 % distinct
    vals = keys * 7 + 3 if k.value == torch.int64 else (keys * 7 + 3) % 100_000
    queries = np.concatenate([keys[: n // 2], keyset(kind, n // 2, 2, hi=4 * n)])
    for lf in (0.5, 0.8):
        t = make(kind, native_lib, n=n, load_factor=lf)
        ref = oracle.Table.for_kind(kind, n, lf)
        assert t.capacity() == ref.capacity()
        got = t.insert(dev(keys, k.key), dev(vals, k.value))
        assert got == ref.insert(keys, vals)
        assert t.size() == ref.size()
        assert np.array_equal(t.find(dev(queries, k.key)).cpu().numpy(), ref.find(queries))
        assert np.array_equal(t.contains(dev(queries, k.key)).cpu().numpy(), ref.contains(queries))
        # second insert of the same stream adds nothing
        assert t.insert(dev(keys, k.key), dev(vals, k.value)) == 0
        assert t.size() == ref.size()
        t.close()


@pytest.mark.parametrize("tuning", TUNINGS[:3])
@pytest.mark.parametrize("kind", SET_KINDS)
def test_set_insert_find_contains_match_oracle(kind, tuning, native_lib):
    native_lib.set_tuning(tuning[0], tuning[1], tuning[2], tuning[5], tuning[3], 1, tuning[4])
    k = cb.KINDS[kind]
    n = 30_000
    keys = keyset(kind, n, 3, hi=n)
    queries = np.concatenate([keys[: n // 2], keyset(kind, n // 2, 4, hi=4 * n)])
    t = make(kind, native_lib, n=n, load_factor=0.5)
    ref = oracle.Table.for_kind(kind, n, 0.5)
    assert t.capacity() == ref.capacity()
    assert t.insert(dev(keys, k.key)) == ref.insert(keys)
    assert t.size() == ref.size()
    assert np.array_equal(t.find(dev(queries, k.key)).cpu().numpy(), ref.find(queries))
    assert np.array_equal(t.contains(dev(queries, k.key)).cpu().numpy(), ref.contains(queries))
    found, inserted = t.insert_and_find(dev(queries, k.key))
    rf, ri = ref.insert_and_find(queries)
    assert np.array_equal(found.cpu().numpy(), rf)
    assert int(inserted.sum().item()) == int(ri.sum())
    assert t.size() == ref.size()
    t.close()


@pytest.mark.parametrize("kind", MAP_KINDS + SET_KINDS)
def test_every_op_matches_cuco_itself(kind, native_lib, reference_lib):
    """Same calls into our build and into cuco's own build of the same shim."""
    k = cb.KINDS[kind]
    is_map = k.value is not None
    n = 50_000
    keys = keyset(kind, n, 5, hi=n // 2)       # heavy duplication
    vals = keys * 3 + 1 if k.value == torch.int64 else (keys * 3 + 1) % 1_000_000
    queries = np.concatenate([keys[::2], keyset(kind, n // 2, 6, hi=3 * n)])
    stencil = (np.arange(n) % 3 != 0)
    dk, dq = dev(keys, k.key), dev(queries, k.key)
    dv = dev(vals, k.value) if is_map else None
    dst = torch.from_numpy(stencil).to("cuda")
    results = {}
    for name, lib in (("ours", native_lib), ("cuco", reference_lib)):
        t = make(kind, lib, n=n, load_factor=0.6)
        r = {"capacity": t.capacity()}
        r["insert_if"] = t.insert_if(dk, dst, dv)
        r["size_after_insert_if"] = t.size()
        r["contains_if"] = t.contains_if(dq, torch.from_numpy(np.arange(dq.numel()) % 2 == 0).to("cuda")).cpu().numpy()
        r["insert"] = t.insert(dk, dv)
        r["size"] = t.size()
        r["find"] = t.find(dq).cpu().numpy()
        r["contains"] = t.contains(dq).cpu().numpy()
        if not (lib is reference_lib and kind == _cabi.MAP_I64_LP1_W2):
            f, ins = t.insert_and_find(dq, dev(queries % 1000, k.value) if is_map else None)
            # which duplicate creates the entry is unspecified; the count and, for keys already
            # present, the payload are not
            r["iaf_new"] = int(ins.sum().item())
            r["iaf_found_existing"] = f.cpu().numpy()[: n // 2]
            r["size_after_iaf"] = t.size()
        if is_map:
            t.insert_or_assign(dk, dev(vals + 5, k.value))
            r["after_assign"] = t.find(dk).cpu().numpy()
            t.clear()
            ones = torch.ones(n, dtype=k.value, device="cuda")
            t.insert_or_apply(dk, ones, op="plus", init=None)
            rk, rv = t.retrieve_all()
            order = torch.argsort(rk)
            r["apply_keys"] = rk[order].cpu().numpy()
            r["apply_vals"] = rv[order].cpu().numpy()
        t.close()
        results[name] = r
    ours, cuco = results["ours"], results["cuco"]
    assert ours.keys() >= cuco.keys()
    for key, want in cuco.items():
        got = ours[key]
        if isinstance(want, np.ndarray):
            assert np.array_equal(got, want), key
        else:
            assert got == want, key


def test_insert_or_apply_sums_like_the_reference_test(native_lib):
    """tests/static_map/insert_or_apply_test.cu:36-272: 10 000 rows / 100 distinct, plus, value 1."""
    for kind in (_cabi.MAP_I64_LP1, _cabi.MAP_I32_LP4, _cabi.MAP_I64_DH8):
        k = cb.KINDS[kind]
        n, distinct = 10_000, 100
        keys = np.arange(n) % distinct
        for sentinel, init in ((0, 0), (0, None), (-1, 0), (-1, None)):
            t = make(kind, native_lib, capacity=2 * distinct, empty_value=sentinel)
            ref = oracle.Table.for_kind(kind, 2 * distinct, empty_value=sentinel)
            t.insert_or_apply(dev(keys, k.key), torch.ones(n, dtype=k.value, device="cuda"), op="plus", init=init)
            ref.insert_or_apply(keys, np.ones(n, dtype=np.int64), oracle.PLUS, init)
            assert t.size() == distinct
            rk, rv = t.retrieve_all()
            order = torch.argsort(rk)
            ok, ov = ref.retrieve_all()
            oo = np.argsort(ok)
            assert np.array_equal(rk[order].cpu().numpy(), ok[oo])
            assert np.array_equal(rv[order].cpu().numpy(), ov[oo])
            t.close()
        for op, code in (("min", oracle.MIN), ("max", oracle.MAX)):
            vals = (np.arange(n) * 7919) % 1000
            t = make(kind, native_lib, capacity=2 * distinct)
            ref = oracle.Table.for_kind(kind, 2 * distinct)
            t.insert_or_apply(dev(keys, k.key), dev(vals, k.value), op=op)
            ref.insert_or_apply(keys, vals, code)
            assert np.array_equal(t.find(dev(np.arange(distinct), k.key)).cpu().numpy(), ref.find(np.arange(distinct)))
            t.close()


def test_empty_and_tiny_inputs(native_lib):
    for kind in (_cabi.MAP_I64_LP1, _cabi.SET_I32_DH4, _cabi.MAP_I64_DH8):
        k = cb.KINDS[kind]
        is_map = k.value is not None
        for cap in (0, 1, 3):
            t = make(kind, native_lib, capacity=cap)
            ref = oracle.Table.for_kind(kind, cap)
            assert t.capacity() == ref.capacity()
            assert t.size() == 0
            empty = torch.empty(0, dtype=k.key, device="cuda")
            assert t.insert(empty, torch.empty(0, dtype=k.value, device="cuda") if is_map else None) == 0
            assert t.find(empty).numel() == 0
            q = dev(np.array([0, 1, 2, 12345]), k.key)
            assert not t.contains(q).any().item()
            sentinel = t.empty_value_sentinel if is_map else t.empty_key_sentinel
            assert (t.find(q) == sentinel).all().item()
            one = dev(np.array([7]), k.key)
            assert t.insert(one, dev(np.array([70]), k.value) if is_map else None) == 1
            assert t.contains(one).all().item() and t.size() == 1
            t.close()


def test_capacity_gold_values(native_lib):
    """tests/static_map/capacity_test.cu: 0 -> 4, 400 -> 422 (cg1 w2), 412 (cg2 w2), (400, 0.8) -> 502."""
    t = make(_cabi.MAP_I64_LP1_W2, native_lib, capacity=0); assert t.capacity() == 4; t.close()
    t = make(_cabi.MAP_I64_LP1_W2, native_lib, capacity=400); assert t.capacity() == 422; t.close()
    t = make(_cabi.MAP_I32_DH2_W2_MM, native_lib, capacity=400); assert t.capacity() == 412; t.close()
    t = make(_cabi.MAP_I64_LP1_W2, native_lib, n=400, load_factor=0.8); assert t.capacity() == 502; t.close()


def test_error_conventions(native_lib):
    with pytest.raises(cb.CucoError):  # load factor must be in (0, 1]
        make(_cabi.MAP_I64_LP1, native_lib, n=100, load_factor=1.5)
    with pytest.raises(cb.CucoError):
        make(_cabi.MAP_I64_LP1, native_lib, n=100, load_factor=-0.5)
    with pytest.raises(cb.CucoError):  # erased sentinel must differ from empty
        make(_cabi.MAP_I64_LP1, native_lib, capacity=100, erased_key=-1)
    t = make(_cabi.MAP_I64_LP1, native_lib, capacity=100)
    with pytest.raises(cb.CucoError):  # erase needs an erased-key sentinel
        t.erase(dev(np.array([1, 2]), torch.int64))
    t.close()


@pytest.mark.parametrize("kind", [_cabi.MAP_I64_LP1, _cabi.MAP_I32_LP4, _cabi.SET_I32_DH4, _cabi.MAP_I64_DH8])
def test_erase_rehash_retrieve(kind, native_lib):
    """tests/static_map/erase_test.cu, rehash_test.cu: erase then lookups, re-insert, rehash keeps content."""
    k = cb.KINDS[kind]
    is_map = k.value is not None
    n = 4_000
    keys = np.random.default_rng(9).permutation(3 * n)[:n] + 1
    vals = keys * 2
    dk = dev(keys, k.key)
    dv = dev(vals, k.value) if is_map else None
    t = make(kind, native_lib, capacity=2 * n, erased_key=-2)
    ref = oracle.Table.for_kind(kind, 2 * n, erased_key=-2)
    assert t.insert(dk, dv) == ref.insert(keys, vals if is_map else None) == n
    half = keys[: n // 2]
    t.erase(dev(half, k.key)); ref.erase(half)
    assert t.size() == ref.size() == n - n // 2
    assert np.array_equal(t.contains(dk).cpu().numpy(), ref.contains(keys))
    assert np.array_equal(t.find(dk).cpu().numpy(), ref.find(keys))
    # erased keys can come back (they are no longer anywhere in the table)
    assert t.insert(dev(half, k.key), dev(half * 2, k.value) if is_map else None) == n // 2
    ref.insert(half, half * 2 if is_map else None)
    assert t.size() == ref.size() == n
    assert np.array_equal(t.find(dk).cpu().numpy(), ref.find(keys))
    t.erase(dev(half, k.key))
    t.rehash()          # same extent, tombstones dropped
    assert t.size() == n - n // 2
    t.rehash(8 * n)     # grow
    assert t.capacity() >= 8 * n and t.size() == n - n // 2
    ref.erase(half)
    assert np.array_equal(t.contains(dk).cpu().numpy(), ref.contains(keys))
    got = t.retrieve_all()
    got_keys = (got[0] if is_map else got).sort().values.cpu().numpy()
    assert np.array_equal(got_keys, np.sort(keys[n // 2:]))
    t.close()


def test_aos_pair_input_equals_separate_arrays(native_lib):
    n = 10_000
    keys = np.random.default_rng(11).integers(1, n, size=n)
    pairs = dev(np.stack([keys, keys + 1], axis=1), torch.int64)
    a = make(_cabi.MAP_I64_LP1, native_lib, n=n, load_factor=0.5)
    b = make(_cabi.MAP_I64_LP1, native_lib, n=n, load_factor=0.5)
    assert a.insert(pairs) == b.insert(dev(keys, torch.int64), dev(keys + 1, torch.int64))
    q = dev(np.arange(2 * n), torch.int64)
    assert torch.equal(a.find(q), b.find(q))
    a.close(); b.close()


def test_full_size_round_trip_properties(native_lib):
    """BASELINE config sizes (scaled to what fits comfortably: 100 M pairs, LF 0.5 and 0.8):
    size-independent properties instead of an oracle run."""
    n = 100_000_000
    keys = torch.randperm(n, device="cuda", dtype=torch.int64)
    pairs = torch.stack([keys, keys ^ 0x5555], dim=1).contiguous()
    absent = keys + n
    for probing, cg, lf in (("linear_probing", 1, 0.5), ("double_hashing", 8, 0.8)):
        t = cb.static_map(n=n, load_factor=lf, probing=probing, cg_size=cg, _library=native_lib)
        assert t.insert(pairs) == n                      # all unique -> all new
        assert t.size() == n
        assert t.insert(pairs) == 0                      # idempotent
        assert torch.equal(t.find(keys), keys ^ 0x5555)  # every payload comes back
        assert bool(t.contains(keys).all().item())
        assert not bool(t.contains(absent).any().item())
        assert bool((t.find(absent) == -1).all().item())
        t.close()
        del t
    torch.cuda.empty_cache()


BLOCKED_VARIANTS = [
    # (region MiB or -KiB, keys per thread, cas_first, prefetch, tile_route, stream_probe, slots)
    (1, 2, 0, 1, 0, 0, 2), (-16, 1, 1, 1, 0, 0, 2), (-64, 4, 1, 0, 0, 0, 2), (-16, 4, 0, 1, 0, 0, 2),
    (-256, 2, 1, 1, 0, 0, 2),
    # round-2 kernels: bulk-copy fed router and warp-persistent probe stream, alone and together
    (-16, 4, 1, 1, 1, 0, 2), (-16, 4, 1, 1, 0, 1, 2), (-64, 4, 1, 1, 1, 1, 2), (-16, 4, 1, 0, 1, 1, 1),
    (1, 4, 1, 1, 1, 1, 1), (-256, 4, 1, 1, 1, 1, 2)]


@pytest.mark.parametrize("variant", BLOCKED_VARIANTS)
@pytest.mark.parametrize("kind", [_cabi.MAP_I64_LP1, _cabi.MAP_I64_DH8, _cabi.MAP_I32_LP4, _cabi.SET_I32_DH4])
def test_l2_blocked_mutations_match_oracle(kind, variant, native_lib):
    """The routed (L2-blocked) insert / insert_or_assign / insert_or_apply path, forced on for small
    inputs (2 to ~120 regions), including a fully skewed batch that overflows a region segment."""
    k = cb.KINDS[kind]
    is_map = k.value is not None
    n = 60_000
    try:
        native_lib.set_blocking(1, variant[0])  # always on
        native_lib.set_blocking_variant(variant[1], variant[2], variant[3])
        native_lib.set_stream_variant(variant[4], variant[5], variant[6])
        aos = bool(variant[4]) and is_map  # the bulk-copy router takes arrays of slot images (AoS pairs)
        streams = {
            "uniform": keyset(kind, n, 21, hi=n),
            "skewed": np.concatenate([np.full(n - 100, 7, dtype=np.int64), np.arange(100, dtype=np.int64) + 100]),
            "unique": np.random.default_rng(5).permutation(4 * n)[:n].astype(np.int64),
        }
        for name, keys in streams.items():
            vals = (keys * 3 + 2) % 1_000_000
            t = make(kind, native_lib, n=n, load_factor=0.5)
            ref = oracle.Table.for_kind(kind, n, 0.5)
            dk = dev(keys, k.key)
            dv = dev(vals, k.value) if is_map else None

            def batch(values):
                # (keys, values) arguments of a bulk call: one [n, 2] array of pairs, or two arrays
                if aos:
                    return (torch.stack([dk, dev(values, k.value)], dim=1).contiguous(), None)
                return (dk, dev(values, k.value) if is_map else None)

            assert t.insert(*batch(vals)) == ref.insert(keys, vals if is_map else None), name
            assert t.size() == ref.size(), name
            q = np.concatenate([keys[::3], keyset(kind, 1000, 22, hi=8 * n)])
            assert np.array_equal(t.find(dev(q, k.key)).cpu().numpy(), ref.find(q)), name
            assert t.insert(*batch(vals)) == 0, name  # second pass through the blocked path: all present
            if is_map:
                t.insert_or_assign(*batch(vals + 9)); ref.insert_or_assign(keys, vals + 9)
                assert np.array_equal(t.find(dk).cpu().numpy(), ref.find(keys)), name
                t.clear(); ref.clear()
                ones = np.ones(n, dtype=np.int64)
                t.insert_or_apply(*batch(ones), op="plus"); ref.insert_or_apply(keys, ones, oracle.PLUS)
                assert t.size() == ref.size()
                assert np.array_equal(t.find(dk).cpu().numpy(), ref.find(keys)), name
            t.close()
    finally:
        native_lib.set_blocking(-1, 16)
        native_lib.set_blocking_variant(4, 1, 1)
        native_lib.set_stream_variant(1, 0, 2)  # the defaults of tuning_t


def test_l2_blocked_large_batch_properties(native_lib):
    """The blocked path at its intended size (auto mode picks it): 50 M uniform pairs with duplicates
    at load factor 0.8 must give the same table as the direct path (size, per-key payloads)."""
    n = 50_000_000
    keys = torch.randint(1, n, (n,), device="cuda", dtype=torch.int64)
    pairs = torch.stack([keys, keys * 7 + 3], dim=1).contiguous()
    sizes, sums = [], []
    try:
        for mode in (0, 1):
            native_lib.set_blocking(mode, 16)
            t = cb.static_map(n=n, load_factor=0.8, probing="linear_probing", cg_size=1, _library=native_lib)
            new = t.insert(pairs)
            assert new == t.size()
            assert torch.equal(t.find(keys), keys * 7 + 3)
            assert not bool(t.contains(keys + n).any().item())
            sizes.append(new)
            t.close()
            del t
    finally:
        native_lib.set_blocking(-1, 16)
    assert sizes[0] == sizes[1] == int(torch.unique(keys).numel())
    torch.cuda.empty_cache()


def _run_device_checks(exe):
    import subprocess
    res = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith(("PASS", "FAIL"))]
    return res.returncode, lines


def test_device_checks_match_reference_headers():
    """tests/device_checks.cu uses only the public cuco:: C++ API (bulk calls with fancy iterators,
    device refs with every op tag, shared-memory tables, CTAD, heterogeneous lookup). The binary built
    against include/ must pass every check, and print exactly what the build against the reference's
    own headers prints."""
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    native = root / "tests" / "_build" / "device_checks_native"
    assert native.exists(), "build it with __graft_entry__.build()"
    rc, lines = _run_device_checks(native)
    assert [ln for ln in lines if ln.startswith("FAIL")] == []
    assert rc == 0 and len(lines) > 100
    ref = root / "oracle" / "_ref" / "device_checks_ref"
    if ref.exists():
        ref_rc, ref_lines = _run_device_checks(ref)
        assert ref_lines == lines
        assert ref_rc == 0


def test_dynamic_map_checks_match_reference_headers():
    """tests/dynamic_map_checks.cu (cuco::experimental::dynamic_map through its public API, after
    tests/dynamic_map/unique_sequence_test_experimental.cu plus a growth case): the build against
    include/ must pass; the build against the reference's headers must print the same lines. The
    reference's `reserve` relies on an out-of-range float -> unsigned conversion, so a reference binary
    that does not finish cleanly is reported as a skip of the comparison, not as our failure."""
    import subprocess
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    native = root / "tests" / "_build" / "dynamic_map_checks_native"
    assert native.exists(), "build it with __graft_entry__.build()"
    rc, lines = _run_device_checks(native)
    assert [ln for ln in lines if ln.startswith("FAIL")] == []
    assert rc == 0 and len(lines) >= 10
    ref = root / "oracle" / "_ref" / "dynamic_map_checks_ref"
    if ref.exists():
        try:
            ref_rc, ref_lines = _run_device_checks(ref)
        except subprocess.TimeoutExpired:
            pytest.skip("reference dynamic_map binary did not finish")
        if ref_rc != 0 and len(ref_lines) < len(lines):
            pytest.skip(f"reference dynamic_map binary exited with {ref_rc}")
        assert ref_lines == lines


@pytest.mark.parametrize("kind", [_cabi.MAP_I64_LP1, _cabi.MAP_I32_LP4, _cabi.SET_I32_DH4])
def test_host_buffer_entry_points_match_device_calls(kind, native_lib):
    """cuco_b200_{insert,find,contains}_host: chunked + overlapped copies must give exactly what the
    device-pointer calls give (several chunks: CUCO_B200_HOST_CHUNK defaults to 4 Mi elements)."""
    k = cb.KINDS[kind]
    is_map = k.value is not None
    n = 9_000_001  # three chunks, the last one ragged
    keys = torch.randint(1, n // 2, (n,), dtype=torch.int64).to(k.key)
    vals = (keys.to(torch.int64) * 5 + 1).to(k.value) if is_map else None
    queries = torch.cat([keys[: n // 2], keys[: n // 2] + n])
    a = make(kind, native_lib, n=n, load_factor=0.5)
    b = make(kind, native_lib, n=n, load_factor=0.5)
    hk, hq = keys.pin_memory(), queries.pin_memory()
    if is_map:
        a.insert_host(hk, vals.pin_memory())
        b.insert_async(keys.cuda(), vals.cuda())
    else:
        a.insert_host(hk)
        b.insert_async(keys.cuda())
    found = a.find_host(hq)
    present = a.contains_host(hq)
    torch.cuda.synchronize()
    assert a.size() == b.size()
    assert torch.equal(found, b.find(queries.cuda()).cpu())
    assert torch.equal(present, b.contains(queries.cuda()).cpu())
    a.close(); b.close()


def test_host_buffer_insert_of_aos_pairs(native_lib):
    n = 5_000_000
    keys = torch.randperm(n, dtype=torch.int64)
    pairs = torch.stack([keys, keys * 3], dim=1).contiguous().pin_memory()
    t = make(_cabi.MAP_I64_LP1, native_lib, n=n, load_factor=0.8)
    t.insert_host(pairs)
    out = t.find_host(keys.pin_memory())
    torch.cuda.synchronize()
    assert t.size() == n
    assert torch.equal(out, keys * 3)
    t.close()


class _SimulatedRank:
    """One 'rank' of the fused exchange path, with every rank living on the same GPU: peer pointers
    are ordinary device pointers, so the routing / probe / return kernels run exactly as they do over
    NVLink (the multi-process version is cucollections_b200/partitioned.py::FusedExchange)."""

    def __init__(self, lib, kind, n_total, n_batch, P, me):
        import ctypes as C
        self.C, self.lib, self.P, self.me = C, lib, P, me
        self.table = make(kind, lib, n=n_total // P + n_total // (8 * P) + 64, load_factor=0.5)
        r, cap, sp = C.c_uint32(), C.c_uint32(), C.c_uint32()
        lib.check(lib.exchange_plan(self.table._handle, n_batch, P, C.byref(r), C.byref(cap), C.byref(sp)))
        self.R, self.cap, self.spill_cap = r.value, cap.value, sp.value
        k = cb.KINDS[kind]
        self.slot_bytes = k.key.itemsize + (k.value.itemsize if k.value is not None else 0)
        seg = P * self.R * self.cap
        z = dict(device="cuda")
        self.segments = torch.zeros(seg * self.slot_bytes, dtype=torch.uint8, **z)
        self.counts = torch.zeros(self.R * P, dtype=torch.int32, **z)
        self.flags = torch.zeros(P, dtype=torch.int32, **z)
        self.results = torch.zeros(seg * 8, dtype=torch.uint8, **z)
        self.counts_local = torch.zeros(P * self.R, dtype=torch.int32, **z)
        self.position_local = torch.zeros(max(1, n_batch), dtype=torch.int32, **z)
        self.spill = torch.zeros(self.spill_cap * self.slot_bytes, dtype=torch.uint8, **z)
        self.spill_index = torch.zeros(self.spill_cap, dtype=torch.int32, **z)
        self.spill_count = torch.zeros(1, dtype=torch.int32, **z)

    def peers(self, ranks, attr):
        C = self.C
        return (C.c_void_p * self.P)(*[getattr(r, attr).data_ptr() for r in ranks])

    def route(self, ranks, keys, values, keys_only):
        C = self.C
        vp = lambda t: None if t is None else C.c_void_p(t.data_ptr())  # noqa: E731
        self.lib.check(self.lib.exchange_route(
            self.table._handle, vp(keys), vp(values), keys.shape[0], int(keys_only), self.R, self.cap,
            self.spill_cap, self.P, self.me, cbp.DEFAULT_SALT, self.peers(ranks, "segments"),
            self.peers(ranks, "counts"), self.peers(ranks, "flags"), vp(self.counts_local), vp(self.position_local),
            vp(self.spill), vp(self.spill_index), vp(self.spill_count), None))

    def mutate(self, op=-1):
        C = self.C
        self.lib.check(self.lib.exchange_mutate(self.table._handle, C.c_void_p(self.segments.data_ptr()),
                                                C.c_void_p(self.counts.data_ptr()), self.R, self.cap, self.P, op,
                                                None))

    def lookup(self, ranks, what):
        C = self.C
        self.lib.check(self.lib.exchange_lookup(self.table._handle, C.c_void_p(self.segments.data_ptr()),
                                                C.c_void_p(self.counts.data_ptr()), self.peers(ranks, "results"),
                                                self.R, self.cap, self.P, self.me, what, None))

    def unpermute(self, out, what):
        C = self.C
        self.lib.check(self.lib.exchange_unpermute(self.table._handle, C.c_void_p(self.results.data_ptr()),
                                                   C.c_void_p(self.position_local.data_ptr()), out.shape[0],
                                                   C.c_void_p(out.data_ptr()), what, None))


@pytest.mark.parametrize("kind,P,region_kib", [
    (_cabi.MAP_I64_LP1, 3, 64), (_cabi.MAP_I64_DH8, 2, 64), (_cabi.MAP_I32_LP4, 4, 64), (_cabi.SET_I64_DH4, 2, 64),
    (_cabi.MAP_I64_LP1, 4, 4),   # > 512 (owner, region) buckets: owner-only routing + local regrouping
    (_cabi.MAP_I32_LP4, 8, 2)])
def test_fused_exchange_kernels_with_simulated_ranks(kind, P, region_kib, native_lib):
    """Routing by (owner, region) into the owners' buffers, region-ordered probe of the received
    segments, lookups answered into the sources' result buffers and un-permuted: the union of the
    shards must behave like ONE table (the oracle) holding every rank's batch."""
    k = cb.KINDS[kind]
    is_map = k.value is not None
    n = 50_000  # per rank
    try:
        native_lib.set_blocking(1, -region_kib)  # small regions: dozens to hundreds per shard
        ranks = [_SimulatedRank(native_lib, kind, n * P, n, P, me) for me in range(P)]
        assert (ranks[0].R == 1) == (region_kib < 64)
        ref = oracle.Table.for_kind(kind, 2 * n * P, 0.0)
        batches = []
        for me in range(P):
            keys = keyset(kind, n, 40 + me, hi=2 * n)  # overlaps between ranks and duplicates within
            vals = keys * 7 + 3
            batches.append((keys, vals))
            ref.insert(keys, vals if is_map else None)
            ranks[me].route(ranks, dev(keys, k.key), dev(vals, k.value) if is_map else None, False)
        torch.cuda.synchronize()
        for r in ranks:
            assert int(r.flags.sum().item()) == 0  # nothing spilled
            r.mutate()
        torch.cuda.synchronize()
        assert sum(r.table.size() for r in ranks) == ref.size()
        sizes = [r.table.size() for r in ranks]
        assert min(sizes) > 0.7 * ref.size() / P  # owner hash balances the shards
        # lookups: half present, half absent, per rank
        queries = [np.concatenate([batches[me][0][: n // 2], keyset(kind, n // 2, 60 + me, hi=2 * n) + 4 * n])
                   for me in range(P)]
        for what in (0, 1):
            for me in range(P):
                ranks[me].route(ranks, dev(queries[me], k.key), None, True)
            torch.cuda.synchronize()
            for r in ranks:
                r.lookup(ranks, what)
            torch.cuda.synchronize()
            for me in range(P):
                if what == 0:
                    out = torch.full((queries[me].shape[0],), -7, dtype=k.value if is_map else k.key, device="cuda")
                    ranks[me].unpermute(out, 0)
                    assert np.array_equal(out.cpu().numpy().astype(np.int64), ref.find(queries[me])), (what, me)
                else:
                    out = torch.full((queries[me].shape[0],), 9, dtype=torch.uint8, device="cuda")
                    ranks[me].unpermute(out, 1)
                    assert np.array_equal(out.cpu().numpy().astype(bool), ref.contains(queries[me])), (what, me)
        if is_map:  # aggregate variant: sum of ones per key over every rank's batch
            for r in ranks:
                r.table.clear()
            agg = oracle.Table.for_kind(kind, 2 * n * P, 0.0, empty_value=0)
            aggs = [make(kind, native_lib, n=n * P, load_factor=0.5, empty_value=0) for _ in range(P)]
            for me in range(P):
                ranks[me].table.close()
                ranks[me].table = aggs[me]
                ones = np.ones(n, dtype=np.int64)
                agg.insert_or_apply(batches[me][0], ones, oracle.PLUS)
                ranks[me].route(ranks, dev(batches[me][0], k.key), dev(ones, k.value), False)
            torch.cuda.synchronize()
            for r in ranks:
                r.mutate(_cabi.PLUS)
            torch.cuda.synchronize()
            got = {}
            for r in ranks:
                ks, vs = r.table.retrieve_all()
                got.update(zip(ks.cpu().tolist(), vs.cpu().tolist()))
            wk, wv = agg.retrieve_all()
            assert got == dict(zip(wk.tolist(), wv.tolist()))
        for r in ranks:
            r.table.close()
    finally:
        native_lib.set_blocking(-1, 16)


def test_fused_exchange_spills_when_a_segment_overflows(native_lib):
    """A batch made of one hot key overflows its (owner, region) segment: the surplus must land in the
    spill list with its source indices, and the flags must tell every peer."""
    P, n = 2, 40_000
    ranks = [_SimulatedRank(native_lib, _cabi.MAP_I64_LP1, n * P, n, P, me) for me in range(P)]
    keys = np.full(n, 12345, dtype=np.int64)
    ranks[0].route(ranks, dev(keys, torch.int64), None, True)
    torch.cuda.synchronize()
    spilled = int(ranks[0].spill_count.item())
    assert spilled == n - ranks[0].cap
    assert [int(r.flags[0].item()) for r in ranks] == [spilled, spilled]
    routed = int(torch.clamp(ranks[0].counts_local, max=ranks[0].cap).sum().item())
    assert routed == ranks[0].cap
    idx = ranks[0].spill_index[:spilled].cpu().numpy()
    assert len(set(idx.tolist())) == spilled and idx.min() >= 0 and idx.max() < n
    marked = (ranks[0].position_local.cpu().numpy().view(np.uint32) == 0xFFFFFFFF).nonzero()[0]
    assert sorted(marked.tolist()) == sorted(idx.tolist())
    for r in ranks:
        r.table.close()


def test_partitioned_table_on_all_visible_gpus():
    """tests/multi_gpu_check.py under torchrun, one rank per visible GPU (needs >= 2): fused P2P
    exchange == all_to_all routing == ONE table over the union of all ranks' batches."""
    import socket
    import subprocess
    import sys
    from pathlib import Path
    gpus = torch.cuda.device_count()
    if gpus < 2:
        pytest.skip("needs at least 2 GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    root = Path(__file__).resolve().parent.parent
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={gpus}",
                          "--master-addr", "127.0.0.1", "--master-port", str(port),
                          str(root / "tests" / "multi_gpu_check.py"), "1000000"],
                         capture_output=True, text=True, timeout=900)
    assert "MULTI_GPU_CHECK PASS" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]
    assert res.returncode == 0
