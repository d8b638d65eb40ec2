"""with_predicate on the CPU side: the oracle's restatement of src/row_group_filter.rs / bloom_filter.rs / statistics.rs
against the reference's own expectations, and the library's host half (statistics and Bloom-filter parsing, verdicts,
the combined plan) against the oracle on fixtures and generated files.  No kernels are launched here."""
import datetime
import decimal
import os
import random
import zlib

import pyarrow as pa
import pyarrow.orc as paorc
import pytest

import kat_vectors as kv
from conftest import GOLDEN
from oracle import orc_oracle as oo

INTEGRATION = os.path.join(GOLDEN, "ref_integration")
OPS = {"eq": 0, "ne": 1, "lt": 2, "le": 3, "gt": 4, "ge": 5}


@pytest.fixture(scope="module")
def ob():
    import orc_rust_b200 as m
    m.lib()
    return m


def to_api(ob, p):
    """the oracle's tuple form -> orc_rust_b200.Predicate"""
    P, V = ob.Predicate, ob.PredicateValue
    if p[0] == "cmp":
        return P.comparison(p[1], OPS[p[2]], V(p[3][0], p[3][1]))
    if p[0] == "is_null":
        return P.is_null(p[1])
    if p[0] == "is_not_null":
        return P.is_not_null(p[1])
    if p[0] == "and":
        return P.and_([to_api(ob, c) for c in p[1]])
    if p[0] == "or":
        return P.or_([to_api(ob, c) for c in p[1]])
    return P.not_(to_api(ob, p[1]))


# ---------------------------------------------------------------------------------------------------------------------
# the oracle against the reference's expectations
# ---------------------------------------------------------------------------------------------------------------------
def test_oracle_bloom_filter_predicate_prunes():
    of = oo.OracleFile(open(os.path.join(INTEGRATION, "bloom_filter.orc"), "rb").read())
    assert sum(b.num_rows for b in of.read()) == 204
    for pred, rows in kv.BLOOM_FILTER_PRUNES:
        assert sum(b.num_rows for b in of.read(predicate=pred)) == rows, pred


def test_oracle_row_group_filter_unit_cases():
    for name, groups, pred, expect in kv.ROW_GROUP_FILTER:
        entries = []
        for n, has_null, mm, bloom_values in groups:
            stats = None
            if n is not None:
                stats = {"n": n, "has_null": has_null, "type": ("int", mm[0], mm[1]) if (mm and n) else None}
            bloom = None
            if bloom_values is not None:
                bloom = (3, [0, 0])
                for v in bloom_values:
                    oo.bloom_add_hash(bloom, oo.bloom_hash_long(v))
            entries.append({"stats": stats, "bloom": bloom})
        index, cols = {1: entries}, [("age", 1)]
        if expect is None:
            with pytest.raises(oo.OracleError):
                oo.evaluate_predicate(pred, index, cols, len(groups))
        else:
            assert oo.evaluate_predicate(pred, index, cols, len(groups)) == expect, name
    # :1047-1061: a column of the schema without a row index is an error as well
    with pytest.raises(oo.OracleError):
        oo.evaluate_predicate(("cmp", "age", "gt", ("Int32", 10)), {}, [("age", 1)], 1)


def test_oracle_string_comparison_table():
    for lo, hi, op, v, emin, emax, keep in kv.STRING_COMPARISON:
        assert oo.compare_strings(lo, hi, op, v, emin, emax) == keep, (lo, hi, op, v, emin, emax)


def test_oracle_bloom_filter_unit_cases():
    # src/bloom_filter.rs:281-293 (hit and miss) and :295-312 (utf8bitset = the words' little-endian bytes)
    bloom = (3, [0, 0])
    for v in (b"abc", b"def"):
        oo.bloom_add_hash(bloom, oo.bloom_hash_bytes(v))
    assert oo.bloom_test_hash(bloom, oo.bloom_hash_bytes(b"abc"))
    assert not oo.bloom_test_hash(bloom, oo.bloom_hash_bytes(b"xyz"))
    one = (2, [0])
    oo.bloom_add_hash(one, oo.bloom_hash_bytes(b"foo"))
    msg = bytes([0x08, 2, 0x1A, 8]) + one[1][0].to_bytes(8, "little")  # numHashFunctions = 2, utf8bitset
    assert oo.parse_bloom_filter(msg) == (2, one[1])
    assert oo.parse_bloom_filter(b"") is None
    assert oo.parse_bloom_filter(bytes([0x11]) + (5).to_bytes(8, "little")) == (3, [5])  # k defaults to 3


def test_oracle_row_selection_helpers():
    for verdict, stride, rows, expect in kv.FROM_ROW_GROUP_FILTER:
        assert [tuple(x) for x in oo.from_row_group_filter(verdict, stride, rows)] == expect
    for first, second, expect in kv.AND_THEN:
        assert [tuple(x) for x in oo.selection_and_then(first, second)] == expect
    with pytest.raises(oo.OracleError):
        oo.selection_and_then([(False, 5)], [(False, 10)])


def test_oracle_predicate_integration_files():
    # tests/integration/main.rs:163-261: never more rows than without, and no failure without an index
    of = oo.OracleFile(open(os.path.join(INTEGRATION, "TestOrcFile.testPredicatePushdown.orc"), "rb").read())
    full = sum(b.num_rows for b in of.read())
    assert sum(b.num_rows for b in of.read(predicate=("cmp", "int1", "gt", ("Int32", 2000)))) <= full
    got = of.read(predicate=("and", [("cmp", "int1", "ge", ("Int32", 1000)), ("cmp", "int1", "le", ("Int32", 5000))]))
    assert got and sum(b.num_rows for b in got) == 1000  # int1 = 300 * row: only the first group of 1000 rows
    of = oo.OracleFile(open(os.path.join(INTEGRATION, "TestOrcFile.testWithoutIndex.orc"), "rb").read())
    assert sum(b.num_rows for b in of.read(predicate=("cmp", "int1", "gt", ("Int32", 1000)))) == of.number_of_rows


# ---------------------------------------------------------------------------------------------------------------------
# the library's host half against the oracle
# ---------------------------------------------------------------------------------------------------------------------
def test_bloom_hashes_match_oracle(ob):
    L = ob.lib()
    rng = random.Random(11)
    for v in [0, 1, -1, 2**63 - 1, -2**63, 10, 20, 50] + [rng.randrange(-2**63, 2**63) for _ in range(300)]:
        assert L.orcb_bloom_hash_long(v) == oo.bloom_hash_long(v), v
    for n in list(range(0, 40)) + [255, 1000]:
        b = bytes(rng.randrange(256) for _ in range(n))
        assert L.orcb_bloom_hash_bytes(b, len(b)) == oo.bloom_hash_bytes(b), n


def _table(n, seed):
    rng = random.Random(seed)
    base = datetime.date(2020, 1, 1)
    def maybe(v, p=0.05):
        return None if rng.random() < p else v
    cols = {
        "i": [maybe(k // 7 + rng.randrange(50)) for k in range(n)],
        "l": [maybe((k * 1000003) % 5000 - 2500) for k in range(n)],
        "d": [maybe(k / 10.0 + rng.random()) for k in range(n)],
        "f": [maybe(float(k % 977)) for k in range(n)],
        "s": [maybe("key_%06d" % (k // 3)) for k in range(n)],
        "t": [maybe("v%d" % rng.randrange(40)) for k in range(n)],
        "b": [maybe(k % 1000 < 500 if k < n // 2 else True) for k in range(n)],
        "dt": [maybe(base + datetime.timedelta(days=k // 100)) for k in range(n)],
        "ts": [maybe(datetime.datetime(2021, 1, 1) + datetime.timedelta(seconds=k)) for k in range(n)],
        "dec": [maybe(decimal.Decimal(k) / 100) for k in range(n)],
        "bin": [maybe(bytes([k % 251, 1, 2])) for k in range(n)],
        "allnull": [None if (k // 1000) % 2 == 0 else k for k in range(n)],
    }
    types = {"i": pa.int32(), "l": pa.int64(), "d": pa.float64(), "f": pa.float32(), "s": pa.string(), "t": pa.string(),
             "b": pa.bool_(), "dt": pa.date32(), "ts": pa.timestamp("ns"), "dec": pa.decimal128(12, 2), "bin": pa.binary(),
             "allnull": pa.int64()}
    return pa.table({k: pa.array(v, types[k]) for k, v in cols.items()})


@pytest.fixture(scope="module")
def generated(tmp_path_factory):
    d = tmp_path_factory.mktemp("pred")
    out = []
    for name, n, kw in [("plain", 12_500, dict(compression="uncompressed", row_index_stride=1000)),
                        ("snappy_bloom", 9_000, dict(compression="snappy", row_index_stride=2000,
                                                     bloom_filter_columns=[1, 2, 3, 5, 6, 7, 8])),
                        ("lz4_stripes", 30_000, dict(compression="lz4", row_index_stride=1000, stripe_size=64 * 1024,
                                                     bloom_filter_columns=[2, 6], bloom_filter_fpp=0.01))]:
        path = str(d / (name + ".orc"))
        paorc.write_table(_table(n, zlib.crc32(name.encode())), path, **kw)
        out.append(path)
    return out


def _random_value(rng, column):
    kinds = {
        "i": lambda: ("Int32", rng.choice([-5, 0, 3, 700, 1785, 1800, 5000])),
        "l": lambda: (rng.choice(["Int8", "Int16", "Int32", "Int64"]), rng.choice([-2500, -100, 0, 7, 100, 2499, 2500, 9999])),
        "d": lambda: (rng.choice(["Float32", "Float64"]), rng.choice([-1.0, 0.5, 100.25, 899.0, 1250.5, 5000.0])),
        "f": lambda: ("Float32", rng.choice([-1.0, 0.0, 12.0, 976.0, 977.0])),
        "s": lambda: ("Utf8", rng.choice(["", "key_000000", "key_000333", "key_002999", "key_004000", "zzz", "key_0001"])),
        "t": lambda: ("Utf8", rng.choice(["v0", "v39", "v40", "v7", "a", "w"])),
        "b": lambda: ("Boolean", rng.random() < 0.5),
        "dt": lambda: (rng.choice(["Int32", "Int64", "Int16"]), rng.choice([18262, 18270, 18300, 18386, 18387, 19000])),
        "ts": lambda: (rng.choice(["Int64", "Int32"]), rng.choice([0, 1609459200000, 1609459205000, 1609469200000, 2 * 10**12])),
        "dec": lambda: ("Utf8", rng.choice(["0", "1.5", "12.34", "45", "99.99", "124.99", "9"])),
        "bin": lambda: ("Utf8", "x"),
        "allnull": lambda: ("Int64", rng.choice([0, 1500, 3500])),
        "nope": lambda: ("Int32", 1),
    }
    t, v = kinds[column]()
    if rng.random() < 0.03:
        v = None
    if rng.random() < 0.03:
        t, v = "Utf8", "mismatch"
    return (t, v)


def _random_predicate(rng, depth=0):
    cols = ["i", "l", "d", "f", "s", "t", "b", "dt", "ts", "dec", "bin", "allnull"]
    r = rng.random()
    if depth >= 3 or r < 0.55:
        c = rng.choice(cols + (["nope"] if rng.random() < 0.05 else []))
        if c == "allnull" and rng.random() < 0.7:
            c = "l"
        q = rng.random()
        if q < 0.1:
            return ("is_null", c)
        if q < 0.2:
            return ("is_not_null", c)
        return ("cmp", c, rng.choice(list(OPS)), _random_value(rng, c))
    if r < 0.7:
        return ("not", _random_predicate(rng, depth + 1))
    kids = [_random_predicate(rng, depth + 1) for _ in range(rng.randrange(0, 4))]
    return ("and" if r < 0.85 else "or", kids)


def _oracle_verdict(of, si, pred, columns=None):
    return of.predicate_selection(si, pred, columns)[1]


def test_row_group_verdicts_match_oracle(ob, generated):
    files = generated + [os.path.join(INTEGRATION, f) for f in
                         ("TestOrcFile.testPredicatePushdown.orc", "orc_split_elim.orc", "TestOrcFile.testWithoutIndex.orc",
                          "orc_index_int_string.orc", "TestOrcFile.testSnappy.orc", "TestVectorOrcFile.testLz4.orc")]
    rng = random.Random(5)
    pruned = fallbacks = 0
    for path in files:
        of = oo.OracleFile(open(path, "rb").read())
        if of.compression not in (0, 2, 4):  # the device path only opens uncompressed, Snappy and LZ4 files
            continue
        names = [n for n, _ in of.columns]
        fh = ob._File(path)
        for _ in range(120 if path in generated else 25):
            if path in generated:
                pred = _random_predicate(rng)
            else:
                c = rng.choice(names)
                pred = ("cmp", c, rng.choice(list(OPS)), rng.choice([("Int32", rng.choice([0, 2, 3000, 100000])), ("Int64", 5),
                                                                     ("Utf8", rng.choice(["a", "foo", "zebra"])), ("Float64", 1.0)]))
            proj = None if rng.random() < 0.8 else rng.sample(names, max(1, len(names) // 2))
            si = rng.randrange(len(of.stripes))
            try:
                got = ob.predicate_row_groups(fh, si, to_api(ob, pred), projection=proj)
            except ob.OrcError as e:
                assert e.code in (9, 18, 22)  # a projected column no reader can be built for (orc_split_elim's decimal(0, 0))
                continue
            exp = _oracle_verdict(of, si, pred, proj)
            assert got == exp, (path, si, pred, proj)
            fallbacks += exp is None
            pruned += exp is not None and not all(exp)
    assert pruned > 50 and fallbacks > 10  # both branches were exercised


def test_reader_plan_matches_oracle(ob, generated):
    """Predicate + row selection + batch size -> the ranges every stripe yields (src/arrow_reader.rs:250-321)."""
    rng = random.Random(9)
    checked = panics = 0
    for path in generated + [os.path.join(INTEGRATION, "TestOrcFile.testPredicatePushdown.orc")]:
        of = oo.OracleFile(open(path, "rb").read())
        stripes = [s.number_of_rows for s in of.stripes]
        names = [n for n, _ in of.columns]
        for _ in range(60):
            if path in generated:
                pred = _random_predicate(rng)
            else:
                pred = ("cmp", "int1", rng.choice(list(OPS)), ("Int32", rng.choice([0, 3000, 400000, 2000000])))
            bs = rng.choice([100, 1000, 2000, 8192])
            sel = None
            if rng.random() < 0.4:
                sel = [(rng.random() < 0.5, rng.choice([0, 10, 999, 1000, 4000, 20000])) for _ in range(rng.randrange(1, 6))]
            b = ob.ArrowReaderBuilder.try_new(path).with_predicate(to_api(ob, pred)).with_batch_size(bs)
            if sel is not None:
                b = b.with_row_selection(sel)
            reader = b.build()
            try:
                psel = [of.predicate_selection(i, pred)[0] for i in range(len(stripes))]
                exp = oo.selection_views(sel, stripes, bs, psel)
            except oo.OracleError as e:
                assert "panic" in str(e)
                with pytest.raises(ob.OrcError):  # where the reference panics, the plan fails
                    reader.plan()
                panics += 1
                continue
            got = reader.plan()
            assert [None if p is None else [tuple(v) for v in p] for p in got] == \
                   [None if p is None else [tuple(map(int, v)) for v in p] for p in exp], (path, pred, sel, bs)
            checked += 1
    assert checked > 100 and panics > 0


def test_predicate_argument_checks(ob):
    path = os.path.join(INTEGRATION, "TestOrcFile.testPredicatePushdown.orc")
    with pytest.raises(TypeError):
        ob.ArrowReaderBuilder.try_new(path).with_predicate("int1 > 3")
    with pytest.raises(TypeError):
        ob.Predicate.eq("int1", 3)
    with pytest.raises(ValueError):
        ob.PredicateValue("Int128", 3)
    # a malformed node array is refused by the C entry point itself
    import ctypes
    bad = (ob._PredicateNodeC * 1)()
    bad[0].kind = 5  # NOT without its child
    bad[0].n_children = 1
    n, ev = ctypes.c_size_t(0), ctypes.c_int(0)
    fh = ob._File(path)
    rc = ob.lib().orcb_predicate_row_groups(fh._h, 0, None, ctypes.addressof(bad), 1, None, 0, ctypes.byref(n), ctypes.byref(ev))
    assert rc == 21  # InvalidArgument


def test_reader_plan_with_byte_range_and_projection(ob, generated):
    """The predicate is evaluated for the stripes the reader visits only (with_file_byte_range), over the projected
    columns only: a predicate on a column left out of the projection keeps every stripe whole."""
    path = generated[2]  # many small stripes
    of = oo.OracleFile(open(path, "rb").read())
    assert len(of.stripes) >= 4
    mid = of.stripes[len(of.stripes) // 2].offset
    end = of.stripes[-1].offset + 1
    chosen = [i for i, s in enumerate(of.stripes) if mid <= s.offset < end]
    rows = [of.stripes[i].number_of_rows for i in chosen]
    pred = ("cmp", "l", "eq", ("Int64", 7))
    for proj in (None, ["l", "s"], ["s", "t"]):
        b = ob.ArrowReaderBuilder.try_new(path).with_file_byte_range(mid, end).with_predicate(to_api(ob, pred))
        if proj:
            b = b.with_projection(proj)
        got = b.build().plan()
        psel = [of.predicate_selection(i, pred, proj)[0] for i in chosen]
        exp = oo.selection_views(None, rows, 8192, psel)
        assert [None if p is None else [tuple(v) for v in p] for p in got] == \
               [None if p is None else [tuple(map(int, v)) for v in p] for p in exp], proj
        if proj == ["s", "t"]:  # "Column 'l' not found in schema": select_all for every stripe
            assert all(p == [(0, r)] for p, r in zip(got, rows))


def test_verdicts_on_edited_row_indexes_match_oracle(ob, tmp_path):
    """Predicate pushdown over row indexes whose protobuf FIELDS were edited (tools/fuzz_struct.py: statistics with extreme
    or missing values, positions, counts, duplicated / deleted / halved entries): the library and the oracle reach the same
    verdict - the same row groups kept, or the same fall-back to reading the whole stripe (src/arrow_reader.rs:281-291).
    Where the reference would panic (bucket statistics without a count, src/statistics.rs) the library reports
    Unexpected."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tools"))
    import fuzz_struct as fs
    rng = random.Random(2)
    files = []
    for d in fs.seeds(str(tmp_path)):
        try:
            files.append(fs.Orc(d))
        except AssertionError:
            pass
    same = fallbacks = pruned = panics = 0
    for it in range(4000):
        o = rng.choice(files)
        idx, _sf = o.last_stripe_parts()
        part = fs.as_message(idx)
        if part is None or not fs.edit(part, rng):
            continue
        new = fs.rc.pb_build(part)
        if it % 8 and len(new) != len(idx):
            continue  # mostly edits that keep the index streams where the stripe footer says they are
        data = o.build(index=new)
        of = oo.OracleFile(data)
        names = [n for n, _ in of.columns]
        si = len(of.stripes) - 1
        for _ in range(3):
            c = rng.choice(names)
            val = rng.choice([("Int64", rng.choice([0, 5, 1000, 10**6, -3])), ("Int32", rng.choice([0, 7, 50000])),
                              ("Utf8", rng.choice(["a", "m", "w3", "zebra", ""])), ("Float64", rng.choice([0.0, 1.5, 1e9])), ("Boolean", True)])
            pred = rng.choice([("cmp", c, rng.choice(list(OPS)), val), ("is_null", c), ("is_not_null", c),
                               ("not", ("cmp", c, rng.choice(list(OPS)), val))])
            try:
                got = ob.predicate_row_groups(data, si, to_api(ob, pred))
            except ob.OrcError as e:
                assert e.variant == "Unexpected" and "index out of bounds" in str(e), (pred, str(e))
                panics += 1
                continue
            exp = of.predicate_selection(si, pred, None)[1]
            assert got == exp, (pred, got, exp)
            same += 1
            fallbacks += exp is None
            pruned += exp is not None and not all(exp)
    assert same > 1000 and fallbacks > 20 and pruned > 40, (same, fallbacks, pruned, panics)
