"""CPU: pins the C oracle (oracle/cuco_oracle.c) to every known-answer vector the reference's own
tests hold for this path, and checks its table semantics on the scenarios those tests describe."""
import numpy as np
import pytest

from oracle import oracle as o


def b(value, dtype):
    return np.asarray(value, dtype=dtype).tobytes()


# tests/utility/hash_test.cu:166-184 (xxhash_32), :102-121 (xxhash_64)
XXH32 = [(b(0, np.int8), 0, 3479547966), (b(42, np.int8), 0, 3774771295), (b(0, np.int8), 42, 2099223482),
         (b(0, np.int32), 0, 148298089), (b(0, np.int32), 42, 2132181312), (b(42, np.int32), 0, 1161967057),
         (b(123456789, np.int32), 0, 2987034094), (b(0, np.int64), 0, 3736311059),
         (b(0, np.int64), 42, 1076387279), (b(42, np.int64), 0, 2332451213),
         (b(123456789, np.int64), 0, 1561711919),
         ((123456789).to_bytes(16, "little"), 0, 1846633701),
         (b([123456789] * 32, np.int32), 0, 3715432378)]
XXH64 = [(b(0, np.int8), 0, 16804241149081757544), (b(42, np.int8), 0, 765293966243412708),
         (b(0, np.int8), 42, 9486749600008296231), (b(0, np.int32), 0, 4246796580750024372),
         (b(0, np.int32), 42, 3614696996920510707), (b(42, np.int32), 0, 15516826743637085169),
         (b(123456789, np.int32), 0, 9462334144942111946), (b(0, np.int64), 0, 3803688792395291579),
         (b(0, np.int64), 42, 13194218611613725804), (b(42, np.int64), 0, 13066772586158965587),
         (b(123456789, np.int64), 0, 14662639848940634189),
         ((123456789).to_bytes(16, "little"), 0, 7986913354431084250),
         (b([123456789] * 32, np.int32), 0, 2031761887105658523)]
SEQ = list(range(1, 17))
# tests/utility/hash_test.cu:251-323
M64 = [(b(0, np.int32), 0, [14961230494313510588, 6383328099726337777]),
       (b(9, np.int32), 0, [1779292183511753683, 16298496441448380334]),
       (b(42, np.int32), 0, [2913627637088662735, 16344193523890567190]),
       (b(42, np.int32), 42, [2248879576374326886, 18006515275339376488]),
       (b([2, 2], np.int32), 0, [12221386834995143465, 6690950894782946573]),
       (b([1, 4, 9], np.int32), 42, [299140022350411792, 9891903873182035274]),
       (b([42, 64, 108, 1024], np.int32), 63, [4333511168876981289, 4659486988434316416]),
       (b(SEQ, np.int32), 1024, [3302412811061286680, 7070355726356610672]),
       (b([2, 2], np.int64), 0, [8554944597931919519, 14938998000509429729]),
       (b([1, 4, 9], np.int64), 42, [13442629947720186435, 7061727494178573325]),
       (b([42, 64, 108, 1024], np.int64), 63, [8786399719555989948, 14954183901757012458]),
       (b(SEQ, np.int64), 1024, [15409921801541329777, 10546487400963404004])]
M86 = [(b(0, np.int32), 0, [3422973727, 2656139328, 2656139328, 2656139328]),
       (b(9, np.int32), 0, [2808089785, 314604614, 314604614, 314604614]),
       (b(42, np.int32), 0, [3611919118, 1962256489, 1962256489, 1962256489]),
       (b(42, np.int32), 42, [3399017053, 732469929, 732469929, 732469929]),
       (b([2, 2], np.int32), 0, [1234494082, 1431451587, 431049201, 431049201]),
       (b([1, 4, 9], np.int32), 42, [2516796247, 2757675829, 778406919, 2453259553]),
       (b([42, 64, 108, 1024], np.int32), 63, [2686265656, 591236665, 3797082165, 2731908938]),
       (b(SEQ, np.int32), 1024, [3918256832, 4205523739, 1707810111, 1625952473]),
       (b([2, 2], np.int64), 0, [3811075945, 727160712, 3510740342, 235225510]),
       (b([1, 4, 9], np.int64), 42, [2817194959, 206796677, 3391242768, 248681098]),
       (b([42, 64, 108, 1024], np.int64), 63, [2335912146, 1566515912, 760710030, 452077451]),
       (b(SEQ, np.int64), 1024, [1101169764, 1758958147, 2406511780, 2903571412])]


@pytest.mark.parametrize("data,seed,want", XXH32)
def test_xxhash32_vectors(data, seed, want):
    assert o.xxhash32(data, seed) == want


@pytest.mark.parametrize("data,seed,want", XXH64)
def test_xxhash64_vectors(data, seed, want):
    assert o.xxhash64(data, seed) == want


@pytest.mark.parametrize("data,seed,want", M64)
def test_murmur3_x64_128_vectors(data, seed, want):
    assert o.murmur3_x64_128(data, seed) == want


@pytest.mark.parametrize("data,seed,want", M86)
def test_murmur3_x86_128_vectors(data, seed, want):
    assert o.murmur3_x86_128(data, seed) == want


def test_murmur3_32_matches_published_vectors():
    # MurmurHash3_x86_32 known answers (SMHasher verification style inputs)
    assert o.murmur3_32(b"", 0) == 0
    assert o.murmur3_32(b"", 1) == 0x514E28B7
    assert o.murmur3_32(b"\xff\xff\xff\xff", 0) == 0x76293B50
    assert o.murmur3_32(b"!Ce\x87", 0) == 0xF55B516B
    assert o.murmur3_32(b"Hello, world!", 0x9747B28C) == 0x24884CBA


def test_capacity_golds():
    """tests/static_map/capacity_test.cu:21-182, static_set/capacity_test.cu, utility/extent_test.cu:27."""
    L = o.lib()
    assert L.oracle_num_windows(0, 1, 2) * 2 == 4
    assert L.oracle_num_windows(-10, 1, 2) * 2 == 4
    assert L.oracle_num_windows(400, 1, 2) * 2 == 422
    assert L.oracle_num_windows(400, 2, 2) * 2 == 412
    assert L.oracle_num_windows(L.oracle_ceil_div_lf(400, 0.8), 1, 2) * 2 == 502
    assert L.oracle_num_windows(1234, 2, 4) == 314
    # table sizes SURVEY.md §8 derives for the BASELINE configs
    assert L.oracle_num_windows(2_000_000, 4, 1) == 524_347 * 4
    assert L.oracle_num_windows(200_000_000, 1, 1) == 200_039_789
    assert L.oracle_num_windows(125_000_000, 1, 1) == 125_057_561
    assert L.oracle_num_windows(200_000_000, 8, 1) == 25_037_357 * 8
    assert L.oracle_num_windows(125_000_000, 8, 1) == 15_730_417 * 8
    assert L.oracle_num_windows(1 << 40, 1, 1) == 0  # "Invalid input extent"


def test_create_rejects_bad_arguments():
    with pytest.raises(ValueError):
        o.Table.for_kind(1, 100, load_factor=1.5)
    with pytest.raises(ValueError):
        o.Table.for_kind(1, 100, load_factor=-1.0)
    with pytest.raises(ValueError):
        o.Table.for_kind(1, 100, erased_key=-1)


def test_probe_sequences_scalar_equals_cg1():
    """tests/utility/probing_scheme_test.cu:79-107 and the stride structure of both schemes."""
    lp = o.Table(8, 8, 1, 2, o.LINEAR, o.XXHASH32, 10)
    s = lp.probe_sequence(42, 0, 8)
    n = lp.capacity() // 2
    assert np.array_equal(s, (s[0] + np.arange(8)) % n)
    assert s[0] == o.xxhash32(b(42, np.int64)) % n
    dh = o.Table(8, 8, 1, 1, o.DOUBLE, o.XXHASH32, 10)
    s = dh.probe_sequence(42, 0, 8)
    n = dh.capacity()
    step = o.xxhash32(b(42, np.int64), 1) % (n - 1) + 1
    assert np.array_equal(s, (s[0] + step * np.arange(8)) % n)
    cg = o.Table(8, 8, 4, 1, o.DOUBLE, o.XXHASH32, 1000)
    n = cg.capacity()
    base = cg.probe_sequence(7, 0, 4)
    for r in range(4):
        assert np.array_equal(cg.probe_sequence(7, r, 4), (base + r) % n)
    assert (base[1] - base[0]) % 4 == 0


@pytest.mark.parametrize("kind", [k for k in o.KIND_GEOMETRY if k not in o.MULTI_KINDS])
def test_table_semantics_against_python_dict(kind):
    """Set/map semantics of tests/static_{map,set}/unique_sequence_test.cu, duplicate_keys_test.cu,
    insert_and_find_test.cu, insert_or_assign_test.cu on every geometry."""
    is_map = o.KIND_GEOMETRY[kind][1] != 0
    rng = np.random.default_rng(kind)
    n = 3000
    keys = rng.integers(0, n // 2, size=n)
    vals = keys * 5 + 1
    t = o.Table.for_kind(kind, n, 0.7)
    assert t.size() == 0
    assert not t.contains(keys).any()
    truth = {}
    for k, v in zip(keys, vals):
        truth.setdefault(int(k), int(v))
    assert t.insert(keys, vals if is_map else None) == len(truth)
    assert t.size() == len(truth)
    q = np.arange(n)
    want_present = np.array([int(x) in truth for x in q])
    assert np.array_equal(t.contains(q), want_present)
    sentinel = -1
    want_found = np.array([(truth[int(x)] if is_map else int(x)) if int(x) in truth else sentinel for x in q])
    assert np.array_equal(t.find(q), want_found)
    found, inserted = t.insert_and_find(q, q * 5 + 1 if is_map else None)
    assert inserted.sum() == n - len(truth)
    assert np.array_equal(inserted, ~want_present)
    assert t.size() == n
    if is_map:
        t.insert_or_assign(keys, vals + 100)
        assert np.array_equal(t.find(keys), vals + 100)
        assert t.size() == n


def test_insert_if_and_contains_if():
    t = o.Table.for_kind(1, 1000)
    keys = np.arange(400)
    stencil = keys % 2 == 0
    assert t.insert_if(keys, stencil, keys) == 200          # unique_sequence_test.cu: even stencil
    assert t.size() == 200
    assert np.array_equal(t.contains(keys), stencil)
    assert np.array_equal(t.contains(keys, stencil=keys % 4 == 0), keys % 4 == 0)


def test_insert_or_apply_rules():
    """tests/static_map/insert_or_apply_test.cu: 10 000 rows / 100 distinct, plus, value 1 -> 100 each;
    with a 16-byte slot and init == sentinel the first arrival combines onto the sentinel."""
    keys = np.arange(10_000) % 100
    ones = np.ones(10_000, dtype=np.int64)
    for kind in (1, 3):
        for sentinel, init, want in ((0, 0, 100), (0, None, 100), (-1, None, 100)):
            t = o.Table.for_kind(kind, 300, empty_value=sentinel)
            t.insert_or_apply(keys, ones, o.PLUS, init)
            assert t.size() == 100
            assert (t.find(np.arange(100)) == want).all()
    t = o.Table.for_kind(1, 300, empty_value=-1)      # 16-byte slot, init == sentinel == -1
    t.insert_or_apply(keys, ones, o.PLUS, -1)
    assert (t.find(np.arange(100)) == 99).all()
    t = o.Table.for_kind(3, 300, empty_value=-1)      # 8-byte slot: packed CAS stores the value
    t.insert_or_apply(keys, ones, o.PLUS, -1)
    assert (t.find(np.arange(100)) == 100).all()
    vals = (np.arange(10_000) * 7919) % 1000
    for op, fn in ((o.MIN, min), (o.MAX, max)):
        t = o.Table.for_kind(1, 300)
        t.insert_or_apply(keys, vals, op)
        want = [fn(int(v) for v in vals[keys == k]) for k in range(100)]
        assert np.array_equal(t.find(np.arange(100)), np.array(want))


def test_erase_semantics():
    """tests/static_map/erase_test.cu: erased keys are gone, can be re-inserted, others stay."""
    t = o.Table.for_kind(1, 2000, erased_key=-2)
    keys = np.arange(1, 1001)
    assert t.insert(keys, keys) == 1000
    t.erase(keys[:500])
    assert t.size() == 500
    assert np.array_equal(t.contains(keys), keys > 500)
    assert t.insert(keys[:500], keys[:500]) == 500
    assert t.size() == 1000 and t.contains(keys).all()
    t.erase(keys)
    assert t.size() == 0
    k, v = t.retrieve_all()
    assert k.size == 0
