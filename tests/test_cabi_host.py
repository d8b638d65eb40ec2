"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/orc_b200.h
declares, and the host half (file tail, schema mapping, stripe selection, launch planning, error mapping)
behaves like the reference's builder.  No kernels are launched here."""
import glob
import os
import re

import pyarrow as pa
import numpy as np
import pytest

from conftest import GOLDEN

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ob():
    import orc_rust_b200 as m
    m.lib()
    return m


def test_library_exports_every_declared_symbol(ob):
    hdr = open(os.path.join(ROOT, "include", "orc_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(orcb_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 30
    L = ob.lib()
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, f"library lacks {missing}"
    assert set(declared) == set(ob.EXPORTED_SYMBOLS)
    assert b"sm_100a" in L.orcb_build_info()
    # the reference-side binding shown in INTEGRATION.md (the Rust `extern "C"` block) covers the whole header
    integ = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    unbound = [s for s in declared if not re.search(r"\bfn " + s + r"\(", integ)]
    assert not unbound, f"INTEGRATION.md does not bind {unbound}"


def test_no_cpu_decode_fallback(ob):
    """Without a CUDA device the decode entry points must fail loudly (status Cuda), never decode on the host."""
    if ob.device_available():
        pytest.skip("a GPU is present")
    with pytest.raises(ob.OrcError) as e:
        ob.decode_int_rle(bytes([0x0A, 0x27, 0x10]), 5, version=2, signed=False)
    assert e.value.variant == "Cuda"
    path = os.path.join(GOLDEN, "ref_basic", "test.orc")
    with pytest.raises(ob.OrcError) as e:
        next(iter(ob.ArrowReaderBuilder.try_new(path).build()))
    assert e.value.variant == "Cuda"


def test_file_metadata_matches_oracle_and_pyarrow(ob):
    import pyarrow.orc as po
    from oracle import orc_oracle as oo
    for f in sorted(glob.glob(os.path.join(GOLDEN, "ref_*", "*.orc"))):
        data = open(f, "rb").read()
        try:
            of = oo.OracleFile(data)
        except oo.OracleError:
            continue
        if of.compression in (3, 5) or not of.is_flat() or os.path.basename(f) == "orc_split_elim.orc":
            continue
        b = ob.ArrowReaderBuilder.try_new(data)
        fm = b.file_metadata()
        assert fm.number_of_rows == of.number_of_rows
        assert fm.num_stripes == len(of.stripes)
        assert fm.compression == of.compression
        assert fm.column_names == [n for n, _ in of.columns]
        for i, s in enumerate(of.stripes):
            si = fm.stripe_info(i)
            assert (si["offset"], si["index_length"], si["data_length"], si["footer_length"], si["number_of_rows"]) == \
                (s.offset, s.index_length, s.data_length, s.footer_length, s.number_of_rows)
        sch = b.schema()
        assert sch.equals(of.schema(), check_metadata=False), f
        ref = po.ORCFile(f).schema
        for name in sch.names:  # same ORC -> Arrow mapping as the Apache reader for flat types
            assert sch.field(name).type == ref.field(name).type or pa.types.is_timestamp(sch.field(name).type)
            assert sch.field(name).nullable


def test_schema_options(ob):
    path = os.path.join(GOLDEN, "ref_basic", "pyarrow_timestamps.orc")
    b = ob.ArrowReaderBuilder.try_new(path)
    assert b.schema().field("timestamp_notz").type == pa.timestamp("ns")
    assert b.schema().field("timestamp_utc").type == pa.timestamp("ns", tz="UTC")
    b.with_timestamp_precision(ob.TimestampPrecision.Microsecond)
    assert b.schema().field("timestamp_notz").type == pa.timestamp("us")
    b = ob.ArrowReaderBuilder.try_new(os.path.join(GOLDEN, "ref_basic", "test.orc")).with_projection(["a", "d"])
    assert b.schema().names == ["a", "d"]


def test_error_mapping(ob):
    with pytest.raises(ob.OrcError) as e:
        ob.ArrowReaderBuilder.try_new(b"")
    assert e.value.variant == "EmptyFile"
    # every compression kind of the format opens: the file tails of Zstandard / LZO files are decoded on the host
    for name, rows in (("alltypes.zstd.orc", 11), ("alltypes.lzo.orc", 11), ("patched_int.orc", None)):
        b = ob.ArrowReaderBuilder.try_new(os.path.join(GOLDEN, "ref_basic", name))
        assert rows is None or b.file_metadata().number_of_rows == rows, name
    # nested types map as in src/schema.rs:530-577
    sch = ob.ArrowReaderBuilder.try_new(os.path.join(GOLDEN, "ref_basic", "nested_map.orc")).schema()
    t = sch.field("map").type
    assert pa.types.is_map(t) and t.key_type == pa.string()  # (pyarrow renames the entries' fields on import)
    with pytest.raises(ob.OrcError) as e:
        ob.ArrowReaderBuilder.try_new(os.path.join(GOLDEN, "ref_integration", "orc_no_format.orc"))
    assert e.value.variant == "OutOfSpec"
    with pytest.raises(ob.OrcError) as e:
        ob.ArrowReaderBuilder.try_new(b"not an orc file at all, just bytes")
    assert e.value.variant in ("OutOfSpec", "DecodeProto")
    with pytest.raises(ob.OrcError) as e:
        ob.ArrowReaderBuilder.try_new("/nonexistent/file.orc")
    assert e.value.variant == "IoError"


def test_launch_plan_statistics(ob, tmp_path):
    """The host planner batches every stream of every projected column of every stripe into one plan."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_orc
    p = str(tmp_path / "li.orc")
    gen_orc.write(gen_orc.lineitem_table(30_000, 0), p, stripe_size=16 << 20, row_index_stride=1000)
    from oracle import orc_oracle as oo
    of = oo.OracleFile(open(p, "rb").read())
    job = ob.DecodeJob([p]).plan()
    st = job.stats()
    assert st["n_stripes"] == len(of.stripes) and st["n_rows"] == of.number_of_rows and st["n_columns"] == 16
    assert st["n_batches"] == sum((s.number_of_rows + 8191) // 8192 for s in of.stripes)
    # algorithmic input bytes = stored bytes of the projected non-index streams
    exp = 0
    for s in of.stripes:
        streams, _, _ = of._stripe_footer(s)
        exp += sum(x.length for x in streams if x.kind in (0, 1, 2, 3, 5) and x.column != 0)
    assert st["input_bytes"] == exp
    # one (stream, row group) segment per row-proportional stream and row group when the row index is used
    groups = sum((s.number_of_rows + 999) // 1000 for s in of.stripes)
    assert st["n_segments"] >= 20 * groups
    no_idx = ob.DecodeJob([p], use_row_index=False).plan().stats()
    assert no_idx["n_segments"] < st["n_segments"] and no_idx["n_segments"] >= 20 * len(of.stripes)
    proj = ob.DecodeJob([p], projection=["l_orderkey", "l_comment"]).plan().stats()
    assert proj["n_columns"] == 2 and proj["input_bytes"] < st["input_bytes"]


def test_stripe_sharding_partitions_stripes(ob, tmp_path):
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_orc
    p = str(tmp_path / "li.orc")
    gen_orc.write(gen_orc.lineitem_table(40_000, 1), p, stripe_size=2 << 20)
    total = ob.DecodeJob([p]).plan().stats()
    assert total["n_stripes"] >= 4
    for n in (2, 3, 8):
        parts = [ob.DecodeJob([p], shard=(r, n)).plan().stats() for r in range(n)]
        assert sum(x["n_stripes"] for x in parts) == total["n_stripes"]
        assert sum(x["n_rows"] for x in parts) == total["n_rows"]
        assert max(x["n_stripes"] for x in parts) - min(x["n_stripes"] for x in parts) <= 1  # round robin


# ---- row selection: host logic (no GPU) -----------------------------------------------------------------------
def test_selection_plan_matches_oracle_restatement():
    """orcb_selection_plan (C++) against the oracle's independent restatement of RowSelection::from / split_off /
    try_advance_stripe / next_with_row_selection, on the reference's own cases and on random selections."""
    import random
    import orc_rust_b200 as ob
    from oracle import orc_oracle as oo
    import kat_vectors as kv

    def norm(plan):
        return [None if p is None else [tuple(map(int, v)) for v in p] for p in plan]

    for name, _file, sel, _proj, rows in kv.ROW_SELECTION:
        got = norm(ob.selection_plan(sel, [10_000 if "large" in name else 5], 8192))
        exp = norm(oo.selection_views(sel, [10_000 if "large" in name else 5], 8192))
        assert got == exp, name
        flat = [r for a, k in exp[0] for r in range(a, a + k)] if exp[0] is not None else None
        assert flat == rows, name
    rng = random.Random(3)
    for _ in range(300):
        stripes = [rng.choice([0, 1, 5, 100, 1000, 8192, 8193, 20000]) for _ in range(rng.randrange(1, 6))]
        sel = [(rng.random() < 0.5, rng.choice([0, 1, 3, 50, 999, 8192, 9000, 30000])) for _ in range(rng.randrange(0, 9))]
        bs = rng.choice([1, 7, 1000, 8192])
        assert norm(ob.selection_plan(sel, stripes, bs)) == norm(oo.selection_views(sel, stripes, bs)), (sel, stripes, bs)
    # RowSelection mirrors: normalisation and from_consecutive_ranges (src/row_selection.rs:158-199, :466-482)
    s = ob.RowSelection([(True, 2), (True, 3), (False, 0), (False, 4)])
    assert [(x.skip, x.row_count) for x in s.selectors] == [(True, 5), (False, 4)]
    s = ob.RowSelection.from_consecutive_ranges([(10, 20), (30, 40)], 50)
    assert [(x.skip, x.row_count) for x in s.selectors] == [(True, 10), (False, 10), (True, 10), (False, 10), (True, 10)]
    assert s.row_count() == 50 and s.selected_row_count() == 20


def test_oracle_row_selection_golden():
    """The oracle under the reference's row-selection cases reads exactly the rows the reference's tests expect."""
    from oracle import orc_oracle as oo
    import kat_vectors as kv
    for name, fname, sel, proj, rows in kv.ROW_SELECTION:
        of = oo.OracleFile(open(os.path.join(GOLDEN, "ref_basic", fname), "rb").read())
        import pyarrow as pa
        full = pa.Table.from_batches(of.read(columns=proj))
        got = of.read(columns=proj, selection=sel)
        n = sum(b.num_rows for b in got)
        assert n == len(rows), name
        if n:
            assert pa.Table.from_batches(got).equals(full.take(pa.array(rows, pa.int64()))), name


def _varint(v):
    out = bytearray()
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def test_crafted_tail_lengths_do_not_wrap(ob):
    """File-controlled 64-bit lengths whose sums wrap (footerLength = 2^64 - 5 and friends) are an error, never an
    out-of-bounds read: the reference returns OutOfSpec / IoError here (src/reader/metadata.rs:199-236)."""
    for comp in (0, 2, 1, 4):
        for footer_len, meta_len in ((2**64 - 5, 0), (2**64 - 1, 2**64 - 1), (40, 2**64 - 30), (2**63, 2**63), (10**6, 0)):
            ps = b"\x08" + _varint(footer_len) + b"\x10" + _varint(comp) + b"\x18" + _varint(262144) + b"\x28" + _varint(meta_len)
            body = b"ORC" + bytes(88 - 3 - len(ps) - 1) if len(ps) + 4 <= 88 else b"ORC"
            data = body + ps + bytes([len(ps)])
            with pytest.raises(ob.OrcError) as e:
                ob.ArrowReaderBuilder.try_new(data)
            assert e.value.variant in ("OutOfSpec", "IoError", "DecodeProto"), (comp, footer_len, meta_len, e.value)
    # stripe and stream lengths of a real file pushed past 2^64
    path = os.path.join(GOLDEN, "ref_basic", "test.orc")
    data0 = open(path, "rb").read()
    rng = np.random.default_rng(5)
    for _ in range(300):  # random damage in the tail: footer, postscript
        data = bytearray(data0)
        for _ in range(int(rng.integers(1, 4))):
            data[int(rng.integers(len(data) - 400, len(data)))] = int(rng.integers(0, 256))
        try:
            b = ob.ArrowReaderBuilder.try_new(bytes(data))
            b.schema()
            ob.DecodeJob([bytes(data)]).plan()  # host planning walks stripe footers, streams and row indexes
        except ob.OrcError:
            pass


def test_host_section_decoders(ob):
    """The host decoders of metadata sections (footer, stripe footers, row indexes) for every compression kind,
    against the data and the oracle.  The Zstandard / LZO parsing is the code the device decoder runs as well
    (csrc/zstd_dec.h), so this also checks it without a GPU."""
    import sys
    import zlib
    import numpy as np
    import pyarrow as pa
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tools"))
    import lzcodec
    from oracle import orc_oracle as oo
    rng = np.random.default_rng(4)
    words = [bytes(rng.integers(97, 123, rng.integers(2, 10), dtype=np.uint8)) for _ in range(400)]
    cases = {
        "empty": b"", "one": b"a", "rle": b"z" * 100_000,
        "text": b" ".join(words[i] for i in rng.integers(0, 400, 40_000)),
        "noise": bytes(rng.integers(0, 256, 60_000, dtype=np.uint8)),
        "lowent": bytes(rng.integers(0, 4, 150_000, dtype=np.uint8)),
        "skewed": bytes(np.minimum(rng.geometric(0.05, 150_000), 255).astype(np.uint8)),
        "ints": np.cumsum(rng.integers(0, 100, 50_000)).astype("<i8").tobytes(),
        "big": b"".join(words[i] for i in rng.integers(0, 400, 120_000))[:500_000],
    }
    def framed(c):
        return (len(c) << 1).to_bytes(3, "little") + c
    for name, d in cases.items():
        for level in (-5, 1, 3, 9, 19):
            f = framed(pa.Codec("zstd", compression_level=level).compress(d, asbytes=True))
            assert ob.host_decompress_section(5, f, 1 << 20) == d, f"zstd {name} level {level}"
            assert bytes(oo.decompress_stream(5, f, 1 << 20)) == d
        f = framed(lzcodec.compress_block("lzo", d))
        assert ob.host_decompress_section(3, f, 1 << 20) == d, f"lzo {name}"
        assert bytes(oo.decompress_stream(3, f, 1 << 20)) == d
        for kind, code in (("snappy", 2), ("lz4", 4)):
            f = framed(lzcodec.compress_block(kind, d)) if d else b""
            assert ob.host_decompress_section(code, f, 1 << 20) == d, f"{kind} {name}"
        co = zlib.compressobj(6, zlib.DEFLATED, -15)
        assert ob.host_decompress_section(1, framed(co.compress(d) + co.flush()), 1 << 20) == d, f"zlib {name}"
    # damaged Zstandard / LZO sections: an error or the bytes the oracle gets
    for code, c in ((5, pa.Codec("zstd").compress(cases["text"], asbytes=True)), (3, lzcodec.compress_block("lzo", cases["text"]))):
        for it in range(80):
            bad = bytearray(c)
            for _ in range(int(rng.integers(1, 4))):
                bad[int(rng.integers(4 if code == 5 else 0, len(bad)))] = int(rng.integers(0, 256))
            try:
                exp = bytes(oo.decompress_stream(code, framed(bytes(bad)), 1 << 20))
            except oo.OracleError:
                exp = None
            try:
                got = ob.host_decompress_section(code, framed(bytes(bad)), 1 << 20)
            except ob.OrcError:
                got = None
            if code == 3:
                assert got == exp, f"lzo damaged #{it}"
            elif exp is not None and got is not None:
                # (libzstd's fast Huffman decoder does not check that a literal stream ends where it should; this decoder
                # does, so some damaged frames libzstd turns into bytes are errors here - never the other way round)
                assert got == exp, f"zstd damaged #{it}"
            else:
                assert got is None, f"zstd damaged #{it}: libzstd fails, this decoder returns bytes"


def test_mutated_files_plan_or_fail_cleanly(ob):
    """Host robustness: every fixture with 1-3 random bytes overwritten (in the postscript / footer area, in the last
    4 KB, or anywhere) either opens and plans, or fails with an OrcError - no crash, no hang, no giant allocation.
    (What the device then decodes from damaged data streams is covered by the -m gpu corruption tests.)"""
    import glob
    import numpy as np
    files = [f for sub in ("ref_basic", "ref_integration") for f in sorted(glob.glob(os.path.join(GOLDEN, sub, "*.orc")))
             if 0 < os.path.getsize(f) < 400_000]
    rng = np.random.default_rng(11)
    ok = err = 0
    for f in files:
        data0 = open(f, "rb").read()
        n = len(data0)
        for it in range(45):
            bad = bytearray(data0)
            span = (min(n, 400), min(n, 4000), n)[it % 3]
            for _ in range(int(rng.integers(1, 4))):
                bad[n - 1 - int(rng.integers(0, span))] = int(rng.integers(0, 256))
            try:
                b = ob.ArrowReaderBuilder.try_new(bytes(bad))
                b.schema()
                ob.DecodeJob([bytes(bad)]).plan()
                ok += 1
            except ob.OrcError:
                err += 1
    assert ok > 500 and err > 500


def test_row_selection_api(ob):
    """The reference's own unit tests of RowSelection (src/row_selection.rs:497-722), against the mirror."""
    RS, S = ob.RowSelection, ob.RowSelector
    sel, skip = S.select, S.skip_rows
    assert (sel(100).row_count, sel(100).skip) == (100, False) and (skip(50).row_count, skip(50).skip) == (50, True)
    s = RS.from_consecutive_ranges([(5, 10), (15, 20)], 25)
    assert s.selectors == [skip(5), sel(5), skip(5), sel(5), skip(5)]
    assert (s.row_count(), s.selected_row_count(), s.skipped_row_count()) == (25, 10, 15)
    assert RS([skip(5), skip(5), sel(10), sel(5)]).selectors == [skip(10), sel(15)]
    a = RS.select_all(100)
    assert (a.row_count(), a.selected_row_count(), a.skipped_row_count(), a.selects_any()) == (100, 100, 0, True)
    a = RS.skip_all(100)
    assert (a.row_count(), a.selected_row_count(), a.skipped_row_count(), a.selects_any()) == (100, 0, 100, False)
    s = RS.from_consecutive_ranges([(10, 30), (40, 60)], 100)
    first = s.split_off(35)
    assert (first.row_count(), s.row_count(), first.selected_row_count(), s.selected_row_count()) == (35, 65, 20, 20)
    assert first.selectors == [skip(10), sel(20), skip(5)] and s.selectors == [skip(5), sel(20), skip(40)]
    whole = RS.select_all(10)
    assert whole.split_off(50).selectors == [sel(10)] and whole.selectors == []
    r = RS.from_consecutive_ranges([(5, 15)], 20).and_then(RS.from_consecutive_ranges([(2, 7)], 10))
    assert r.selectors == [skip(7), sel(5), skip(8)] and (r.row_count(), r.selected_row_count()) == (20, 5)
    with pytest.raises(ValueError):
        RS.select_all(5).and_then(RS.select_all(6))
    assert RS.from_filters([[False, False, True, True, False]]).selectors == [skip(2), sel(2), skip(1)]
    assert RS.from_filters([pa.array([True, False]), pa.array([False, True, True])]).selectors == [sel(1), skip(2), sel(2)]
    e = RS()
    assert (e.row_count(), e.selected_row_count(), e.selects_any()) == (0, 0, False)
    with pytest.raises(ValueError):
        RS.from_consecutive_ranges([(10, 20), (5, 15)], 25)
    f = RS.from_row_group_filter
    s = f([False, True, False], 10000, 30000)
    assert s.selectors == [skip(10000), sel(10000), skip(10000)]
    assert (s.row_count(), s.selected_row_count(), s.skipped_row_count()) == (30000, 10000, 20000)
    assert f([True, True, True], 10000, 30000).selectors == [sel(30000)]
    assert f([False, False, False], 10000, 30000).selectors == [skip(30000)]
    assert f([False, False, True, True, False], 10000, 50000).selectors == [skip(20000), sel(20000), skip(10000)]
    assert f([True, False], 10000, 25000).selectors == [sel(10000), skip(15000)]
    assert f([], 10000, 50000).selectors == [skip(50000)]
    assert list(s.iter()) == s.selectors


def test_projection_mask(ob):
    """ProjectionMask::{all, roots, named_roots} (src/projection.rs:24-80): root columns by ORC column index or name."""
    path = os.path.join(GOLDEN, "ref_basic", "nested_struct.orc")
    b = ob.ArrowReaderBuilder.try_new(path)
    f = b.file_metadata()
    names, ids = f.column_names, f.column_ids
    assert len(names) == len(ids) and ids == sorted(ids) and ids[0] == 1
    assert ob.ProjectionMask.all().names is None
    assert ob.ProjectionMask.roots(f, [ids[0], 999]).names == [names[0]]
    assert ob.ProjectionMask.named_roots(b, [names[-1], "no such column"]).names == [names[-1]]
    path = os.path.join(GOLDEN, "ref_basic", "test.orc")
    b = ob.ArrowReaderBuilder.try_new(path)
    f = b.file_metadata()
    m = ob.ProjectionMask.roots(f, [f.column_ids[0], f.column_ids[3]])
    assert b.with_projection(m).schema().names == [f.column_names[0], f.column_names[3]]
    assert ob.ArrowReaderBuilder.try_new(path).with_projection(ob.ProjectionMask.all()).schema().names == f.column_names


def test_writer_zone_tables_match_zoneinfo(ob):
    """The tables the device searches when a stripe was written in a zone other than UTC (src/array_decoder/timestamp.rs:
    128-147, 242-286: the reference asks chrono-tz for the offset at each instant) against CPython's zoneinfo, an independent
    reader of the same IANA data: the offset at thousands of instants from 1850 to 2400 - before the first transition, at
    both sides of every transition, and past 2037, where the table continues from the TZif footer rule - and the ORC
    epoch (2015-01-01 00:00 on the zone's wall clock)."""
    import datetime
    import zoneinfo
    import numpy as np
    zones = ["America/Los_Angeles", "America/New_York", "Europe/London", "Europe/Berlin", "Europe/Moscow", "Asia/Kolkata",
             "Asia/Shanghai", "Asia/Tokyo", "Asia/Kathmandu", "Australia/Sydney", "Australia/Lord_Howe", "Pacific/Auckland",
             "Pacific/Apia", "America/Sao_Paulo", "America/St_Johns", "Africa/Cairo", "Africa/Casablanca", "Asia/Tehran",
             "Europe/Dublin", "CET", "EET", "MST", "US/Pacific", "Asia/Calcutta", "Japan", "GB"]
    rng = np.random.default_rng(5)
    utc = datetime.timezone.utc
    epoch0 = datetime.datetime(1970, 1, 1, tzinfo=utc)
    lo, hi = -3786825600, 13569465600  # 1850-01-01 .. 2400-01-01
    checked = 0
    for name in zones:
        try:
            zi = zoneinfo.ZoneInfo(name)
        except zoneinfo.ZoneInfoNotFoundError:
            continue
        at, off, first, orc_epoch = ob.zone_table(name)
        assert np.all(np.diff(at) >= 0)
        probes = np.concatenate([rng.integers(lo, hi, 1500), at - 1, at, at + 1, at + 3600])
        probes = probes[(probes >= lo) & (probes < hi)]
        k = np.searchsorted(at, probes, side="right")
        got = np.where(k == 0, first, np.append(off, first)[np.maximum(k - 1, 0)]) if len(at) else np.full(len(probes), first)
        for t, g in zip(probes.tolist(), got.tolist()):
            exp = (epoch0 + datetime.timedelta(seconds=t)).astimezone(zi).utcoffset().total_seconds()
            assert g == exp, f"{name} at {t}: table says {g}, zoneinfo {exp}"
            checked += 1
        wall = datetime.datetime(2015, 1, 1, tzinfo=zi)
        assert orc_epoch == int((wall - epoch0).total_seconds()), name
    assert checked > 20000
    with pytest.raises(ob.OrcError):
        ob.zone_table("Not/AZone")


def test_try_new_async(ob):
    """`ArrowReaderBuilder::try_new_async` (src/async_arrow_reader.rs:292-296), the entry of the reference's async tests
    (tests/basic/main.rs:44-63): metadata and schema without a device."""
    import asyncio

    async def go():
        b = await ob.ArrowReaderBuilder.try_new_async(os.path.join(GOLDEN, "ref_basic", "test.orc"))
        return b.file_metadata().number_of_rows, b.schema().names[:3], b.with_file_byte_range(100, 2000).build().total_row_count()

    rows, names, total = asyncio.run(go())
    assert (rows, names, total) == (5, ["a", "b", "str_direct"], 5)


def test_file_format_version_and_user_metadata(ob):
    """FileMetadata::file_format_version / user_custom_metadata (src/reader/metadata.rs:112-127) on every fixture, against
    the oracle's footer parse and pyarrow (`file_version`, `metadata`)."""
    import glob
    import pyarrow.orc as po
    from oracle import orc_oracle as oo
    n = with_md = 0
    for f in sorted(glob.glob(os.path.join(GOLDEN, "ref_*", "*.orc"))):
        data = open(f, "rb").read()
        try:
            of = oo.OracleFile(data)
        except oo.OracleError:
            continue
        fm = ob.ArrowReaderBuilder.try_new(data).file_metadata()
        assert fm.user_custom_metadata == {k: bytes(v) for k, v in of.user_metadata.items()}, f
        try:
            pf = po.ORCFile(f)
        except Exception:
            continue
        assert fm.file_format_version == pf.file_version, f
        assert {k.encode(): v for k, v in fm.user_custom_metadata.items()} == dict(pf.metadata or {}), f
        n += 1
        with_md += bool(fm.user_custom_metadata)
    assert n > 40 and with_md >= 1


def test_c_program_against_the_header(ob, tmp_path):
    """include/orc_b200.h is plain C (C99, -pedantic) and the library links from C: tests/c_abi_host.c opens a fixture,
    reads metadata and schema, plans a job, and sees statuses (not aborts, not a CPU fallback) where a device or a valid
    file is missing."""
    import shutil
    import subprocess
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    libdir = os.path.dirname(ob.lib()._name)
    exe = str(tmp_path / "c_abi_host")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c_abi_host.c"), "-o", exe, "-L", libdir, "-l:liborc_b200.so",
                           "-Wl,-rpath," + libdir])
    out = subprocess.run([exe, os.path.join(GOLDEN, "ref_basic", "test.orc")], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.splitlines()
    assert lines[0] == "rows=5 stripes=1 compression=0 version=0.12 columns=20"
    assert lines[1] == "schema=+s children=20 first=a:f"
    assert lines[2].startswith("planned stripes=1 rows=5 batches=1 segments=")
    if not ob.device_available():
        assert "next_without_device=20" in lines  # ORCB_CUDA
    assert lines[-1].startswith("open_truncated=3 ")  # ORCB_OUT_OF_SPEC


@pytest.mark.parametrize("kind", ["lz4", "snappy", "zstd", "lzo", "lzo-plain"])
def test_recompressed_files_are_valid_orc(ob, tmp_path, kind):
    """tools/orc_recompress.py (the source of the bench's LZ4 / LZO / Zstandard sets and of config 3's identical chunking,
    SURVEY.md §8(d)) without a GPU: the rewritten file reads back equal to the original through Apache ORC C++ (pyarrow,
    an independent reader) and through the oracle, a fair share of its chunks is really compressed, and the planner
    accepts the rewritten row-index positions (as many (stream, row group) segments as for the uncompressed file)."""
    import sys
    import pyarrow.orc as po
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import gen_orc
    import orc_recompress
    from oracle import orc_oracle as oo
    from parity_util import assert_batches_identical
    for name, table in (("li", gen_orc.lineitem_table(20_000, 21)), ("nh", gen_orc.nullheavy_table(40_000, 5))):
        src = gen_orc.write(table, str(tmp_path / f"{name}.orc"), row_index_stride=2000)
        dst = str(tmp_path / f"{name}.{kind}.orc")
        st = orc_recompress.recompress(src, dst, kind, 65536)
        assert st["compressed"] > 0 and os.path.getsize(dst) < os.path.getsize(src)
        assert orc_recompress.compressed_chunk_fraction(dst)["fraction"] > 0.15
        if kind != "lzo":
            # (LZO: the in-repo compressor emits the M1 instructions of LZO1X - a 2-byte match after 1-3 literals, a 3-byte
            # match at 2049..3072 after a literal run - on purpose, and Apache ORC C++ 's decoder places those 2048 bytes
            # further back than minilzo / the published format do (and rejects a 3-byte M4 match, opcode 0x11): found with this
            # test, see DESIGN.md §6.  The oracle
            # and the planner checks below still run for LZO; "lzo-plain" is the same compressor without those two
            # instructions, which Apache reads.)
            assert po.read_table(dst).equals(po.read_table(src)), f"{name} {kind}: pyarrow reads different data"
        assert_batches_identical(oo.OracleFile(open(dst, "rb").read()).read(), oo.OracleFile(open(src, "rb").read()).read(), f"{name} {kind}")
        a, b = ob.DecodeJob([dst]).plan().stats(), ob.DecodeJob([src]).plan().stats()
        assert a["n_segments"] == b["n_segments"] and a["n_rows"] == b["n_rows"]
        assert a["input_bytes"] < b["input_bytes"]


def test_in_repo_compressors_against_the_libraries(ob):
    """The block compressors that make the decoders' test inputs (tools/lzcodec.c: every Snappy element form, every LZ4
    sequence shape) produce streams the libraries themselves decode to the same bytes - snappy and liblz4 through pyarrow's
    codecs, the C counterparts of the reference's `snap` and `lz4_flex` - and the host decoders of metadata sections
    agree with the libraries in both directions (library-compressed blocks decode here, blocks compressed here decode
    there)."""
    import sys
    import numpy as np
    import pyarrow as pa
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import lzcodec
    from oracle import orc_oracle as oo
    rng = np.random.default_rng(9)
    words = [bytes(rng.integers(97, 123, rng.integers(1, 12), dtype=np.uint8)) for _ in range(300)]
    cases = {
        "text": b" ".join(words[i] for i in rng.integers(0, 300, 30_000)),
        "runs": b"".join(bytes([int(v)]) * int(k) for v, k in zip(rng.integers(0, 256, 400), rng.integers(1, 700, 400))),
        "noise": bytes(rng.integers(0, 256, 70_000, dtype=np.uint8)),
        "far": (lambda a: a + bytes(rng.integers(0, 256, 40_000, dtype=np.uint8)) + a)(bytes(rng.integers(0, 256, 5_000, dtype=np.uint8))),
        "ints": np.cumsum(rng.integers(0, 50, 30_000)).astype("<i4").tobytes(),
        "tiny": b"abc",
    }
    for name, d in cases.items():
        for kind, codec, code in (("snappy", "snappy", 2), ("lz4", "lz4_raw", 4)):
            ours = lzcodec.compress_block(kind, d)
            assert pa.Codec(codec).decompress(ours, decompressed_size=len(d), asbytes=True) == d, f"{kind} {name}: the library rejects our block"
            theirs = pa.Codec(codec).compress(d, asbytes=True)
            for blk in (ours, theirs):
                framed = (len(blk) << 1).to_bytes(3, "little") + blk
                assert ob.host_decompress_section(code, framed, 1 << 20) == d, f"{kind} {name}"
                assert bytes(oo.decompress_stream(code, framed, 1 << 20)) == d, f"oracle {kind} {name}"


def test_async_stream_error_state_without_device(ob):
    """`ArrowStreamReader` (src/async_arrow_reader.rs:252-290) when a request fails: the error reaches the awaiting task
    from the library's thread, and the stream stays in the error state (StreamState::Error, :262-277) - here the failure
    is the missing device, which also shows that the async path has no CPU fallback; an empty stream (a byte range
    without stripes) simply ends."""
    import asyncio
    if ob.device_available():
        pytest.skip("needs a machine without a CUDA device")
    path = os.path.join(GOLDEN, "ref_basic", "test.orc")

    async def failing():
        stream = ob.ArrowReaderBuilder.try_new(path).build_async()
        assert stream.schema().names[:2] == ["a", "b"]
        seen = []
        for _ in range(2):
            try:
                await stream.__anext__()
                seen.append("batch")
            except ob.OrcError as e:
                seen.append(e.variant)
            except StopAsyncIteration:
                seen.append("end")
        return seen

    async def empty():
        return [b async for b in ob.ArrowReaderBuilder.try_new(path).with_file_byte_range(100, 2000).build_async()]

    first, second = asyncio.run(failing())
    assert first == "Cuda" and second in ("Cuda", "Unexpected")  # the reader stays in the error state
    assert asyncio.run(empty()) == []


def test_host_inflate_on_damaged_and_truncated_streams(ob):
    """The host inflate of metadata sections against zlib used the way flate2's read decoder behaves (the oracle,
    oracle/codecs.c zlib_block; src/compression.rs:142-149): same verdict and same bytes on damaged streams, and a
    stream whose input ends early yields what was decoded up to there instead of an error."""
    import zlib
    import numpy as np
    from oracle import orc_oracle as oo
    rng = np.random.default_rng(3)
    words = [bytes(rng.integers(97, 123, rng.integers(2, 10), dtype=np.uint8)) for _ in range(300)]
    cases = {"text": b" ".join(words[i] for i in rng.integers(0, 300, 6000)), "noise": bytes(rng.integers(0, 256, 4000, dtype=np.uint8)),
             "low": bytes(rng.integers(0, 4, 20_000, dtype=np.uint8))}
    both_ok = both_err = 0
    for name, d in cases.items():
        for lvl, strategy in ((1, zlib.Z_DEFAULT_STRATEGY), (9, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_FIXED), (0, zlib.Z_DEFAULT_STRATEGY)):
            co = zlib.compressobj(lvl, zlib.DEFLATED, -15, 9, strategy)
            c = co.compress(d) + co.flush()
            for it in range(150):
                bad = bytearray(c)
                for _ in range(int(rng.integers(0, 4))):
                    bad[int(rng.integers(0, len(bad)))] = int(rng.integers(0, 256))
                if it % 3 == 0:
                    bad = bad[: int(rng.integers(1, len(bad)))]
                f = (len(bad) << 1).to_bytes(3, "little") + bytes(bad)
                try:
                    exp = bytes(oo.decompress_stream(1, f, 1 << 20))
                except oo.OracleError:
                    exp = None
                try:
                    got = ob.host_decompress_section(1, f, 1 << 20)
                except ob.OrcError:
                    got = None
                assert got == exp, f"{name} level {lvl} #{it}: {None if exp is None else len(exp)} vs {None if got is None else len(got)}"
                both_ok += exp is not None
                both_err += exp is None
    assert both_ok > 500 and both_err > 100


@pytest.mark.parametrize("kind,code,codec", [("snappy", 2, "snappy"), ("lz4", 4, "lz4_raw")])
def test_host_lz_decoders_on_damaged_streams(ob, kind, code, codec):
    """The host Snappy / LZ4 decoders of metadata sections against the oracle's (two separate restatements of
    `snap` / `lz4_flex`, src/compression.rs:161-195) on damaged and truncated blocks from the in-repo compressors and
    from the libraries: same verdict, same bytes.  For Snappy the library's verdict is the same as well."""
    import sys
    import numpy as np
    import pyarrow as pa
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import lzcodec
    from oracle import orc_oracle as oo
    rng = np.random.default_rng(8)
    words = [bytes(rng.integers(97, 123, rng.integers(2, 10), dtype=np.uint8)) for _ in range(300)]
    cases = {"text": b" ".join(words[i] for i in rng.integers(0, 300, 5000)), "noise": bytes(rng.integers(0, 256, 3000, dtype=np.uint8)),
             "low": bytes(rng.integers(0, 4, 15_000, dtype=np.uint8))}
    ok = err = 0
    for name, d in cases.items():
        for c in (lzcodec.compress_block(kind, d), pa.Codec(codec).compress(d, asbytes=True)):
            for it in range(200):
                bad = bytearray(c)
                for _ in range(int(rng.integers(0, 4))):
                    bad[int(rng.integers(0, len(bad)))] = int(rng.integers(0, 256))
                if it % 3 == 0:
                    bad = bad[: int(rng.integers(1, len(bad)))]
                f = (len(bad) << 1).to_bytes(3, "little") + bytes(bad)
                try:
                    exp = bytes(oo.decompress_stream(code, f, 1 << 20))
                except oo.OracleError:
                    exp = None
                try:
                    got = ob.host_decompress_section(code, f, 1 << 20)
                except ob.OrcError:
                    got = None
                assert got == exp, f"{kind} {name} #{it}"
                if kind == "snappy":
                    try:
                        lib = pa.Codec(codec).decompress(bytes(bad), decompressed_size=len(got) if got is not None else len(d), asbytes=True)
                    except Exception:
                        lib = None
                    assert lib == got, f"snappy {name} #{it}: the library disagrees"
                ok += got is not None
                err += got is None
    assert ok > 200 and err > 200


def test_snappy_preamble_cannot_size_an_allocation(ob):
    """Regression (found by tools/fuzz_host.sh): the uncompressed-length preamble of a Snappy block in a metadata section
    used to reserve that many bytes before anything was decoded - 32 GiB for five bytes of input.  No element makes more
    than 64 bytes out of 3, so a preamble the input cannot meet is an error at once."""
    import time
    for preamble in (b"\xff\xff\xff\xff\x7f", b"\xff\xff\xff\xff\x0f", b"\x80\x80\x80\x80\x08"):
        blk = preamble + b"\x00a"
        framed = (len(blk) << 1).to_bytes(3, "little") + blk
        t0 = time.time()
        with pytest.raises(ob.OrcError) as e:
            ob.host_decompress_section(2, framed, 1 << 18)
        assert e.value.variant == "BuildSnappyDecoder" and time.time() - t0 < 1.0
    # a preamble the input can meet still decodes
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import lzcodec
    d = b"ab" * 40_000
    blk = lzcodec.compress_block("snappy", d)
    assert len(blk) * 16 < len(d)
    assert ob.host_decompress_section(2, (len(blk) << 1).to_bytes(3, "little") + blk, 1 << 18) == d


def test_damaged_stripe_row_count_with_predicate(ob, tmp_path):
    """Regression (found by tools/fuzz_host.sh): with_predicate sizes its row-group verdict vector from the stripe's row
    count; a footer that claims 2^60 rows used to ask for that much memory.  Such a stripe is refused (the decoder takes
    stripes of up to 2^32 rows), quickly, and so is planning it."""
    import sys
    import time
    import pyarrow as pa
    import pyarrow.orc as po
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import orc_recompress as rc
    p = str(tmp_path / "t.orc")
    po.write_table(pa.table({"x": pa.array(range(5000), pa.int64())}), p, compression="uncompressed", row_index_stride=1000)
    data = open(p, "rb").read()
    n = len(data)
    ps_len = data[-1]
    ps = rc.pb_parse(data[n - 1 - ps_len:n - 1])
    fl = rc.pb_get(ps, 1)
    footer = rc.pb_parse(data[n - 1 - ps_len - fl:n - 1 - ps_len])
    for fld in footer:
        if fld[0] == 3:  # StripeInformation: numberOfRows = 5
            si = rc.pb_parse(fld[2])
            for g in si:
                if g[0] == 5:
                    g[2] = 1 << 60
            fld[2] = rc.pb_build(si)
    new_footer = rc.pb_build(footer)
    for fld in ps:
        if fld[0] == 1:
            fld[2] = len(new_footer)
    new_ps = rc.pb_build(ps)
    bad = data[:n - 1 - ps_len - fl] + new_footer + new_ps + bytes([len(new_ps)])
    t0 = time.time()
    with pytest.raises(ob.OrcError) as e:
        ob.predicate_row_groups(bad, 0, ob.Predicate.eq("x", ob.PredicateValue.Int64(5)))
    assert e.value.variant == "NotImplemented"
    with pytest.raises(ob.OrcError):
        ob.DecodeJob([bad]).plan()
    assert time.time() - t0 < 2.0


def test_crafted_type_trees(ob, tmp_path):
    """A footer whose type list is not a tree - a type that is its own child, a child that is an ancestor or the root, a
    child id past the list, a struct with fewer children than names - is refused when the file is opened (as
    RootDataType::from_proto refuses it); it cannot send the schema export or the planner into a loop."""
    import sys
    import pyarrow as pa
    import pyarrow.orc as po
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import orc_recompress as rc
    p = str(tmp_path / "n.orc")
    po.write_table(pa.table({"s": pa.array([{"a": 1, "b": [1, 2]}, {"a": 2, "b": []}]), "x": pa.array([1, 2])}), p, compression="uncompressed")
    data = open(p, "rb").read()
    n, ps_len = len(data), data[-1]

    def with_subtypes(type_idx, subs):
        ps = rc.pb_parse(data[n - 1 - ps_len:n - 1])
        fl = rc.pb_get(ps, 1)
        footer = rc.pb_parse(data[n - 1 - ps_len - fl:n - 1 - ps_len])
        k = 0
        for f in footer:
            if f[0] == 4:  # Footer.types
                if k == type_idx:
                    ty = rc.pb_parse(f[2])
                    for g in ty:
                        if g[0] == 2:  # Type.subtypes (packed)
                            g[2] = bytes(subs)
                    f[2] = rc.pb_build(ty)
                k += 1
        nf = rc.pb_build(footer)
        for f in ps:
            if f[0] == 1:
                f[2] = len(nf)
        nps = rc.pb_build(ps)
        return data[:n - 1 - ps_len - fl] + nf + nps + bytes([len(nps)])

    # types of the file: 0 struct<s, x>, 1 struct<a, b>, 2 a, 3 list, 4 item, 5 x
    assert ob.ArrowReaderBuilder.try_new(with_subtypes(3, [4])).schema().names == ["s", "x"]  # the rewrite itself is sound
    for what, (ti, subs) in {"a list that is its own child": (3, [3]), "a list whose child is its parent": (3, [1]),
                             "a struct with the root as a child": (1, [0, 3]), "a root that contains itself": (0, [0, 5]),
                             "a child id past the type list": (3, [99]), "fewer children than names": (1, [2])}.items():
        with pytest.raises(ob.OrcError):
            b = ob.ArrowReaderBuilder.try_new(with_subtypes(ti, subs))
            b.schema()
            ob.DecodeJob([with_subtypes(ti, subs)]).plan()
            pytest.fail(what + " was accepted")


def test_non_utf8_strings_in_the_footer(ob, tmp_path):
    """prost checks `string` fields while it decodes the footer, so a field name or a user-metadata key that is not UTF-8
    fails the open with DecodeProto in the reference; so it does here and in the oracle (found by tools/fuzz_struct.py:
    such a name used to be exported in the Arrow schema as it was)."""
    import sys
    import pyarrow as pa
    import pyarrow.orc as po
    from oracle import orc_oracle as oo
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import orc_recompress as rc
    p = str(tmp_path / "u.orc")
    po.write_table(pa.table({"name": pa.array([1, 2, 3])}).replace_schema_metadata({"key": "v"}), p, compression="uncompressed")
    data = open(p, "rb").read()
    n, ps_len = len(data), data[-1]

    def rebuilt(edit):
        ps = rc.pb_parse(data[n - 1 - ps_len:n - 1])
        fl = rc.pb_get(ps, 1)
        footer = rc.pb_parse(data[n - 1 - ps_len - fl:n - 1 - ps_len])
        assert edit(footer)
        nf = rc.pb_build(footer)
        for f in ps:
            if f[0] == 1:
                f[2] = len(nf)
        nps = rc.pb_build(ps)
        return data[:n - 1 - ps_len - fl] + nf + nps + bytes([len(nps)])

    def bad_field_name(footer):
        for f in footer:
            if f[0] == 4:  # types[0]: the root struct
                ty = rc.pb_parse(f[2])
                for g in ty:
                    if g[0] == 3:
                        g[2] = b"na\xffe"
                        f[2] = rc.pb_build(ty)
                        return True
        return False

    def good_metadata_key(footer):  # UserMetadataItem { name = 1 (string), value = 2 (bytes: anything goes) }
        footer.append([5, 2, rc.pb_build([[1, 2, "clé".encode()], [2, 2, b"\xff\x00"]])])
        return True

    def bad_metadata_key(footer):
        footer.append([5, 2, rc.pb_build([[1, 2, b"\xc3\x28"], [2, 2, b"v"]])])
        return True

    assert ob.ArrowReaderBuilder.try_new(rebuilt(lambda f: True)).schema().names == ["name"]
    assert ob.ArrowReaderBuilder.try_new(rebuilt(good_metadata_key)).file_metadata().user_custom_metadata == {"clé": b"\xff\x00"}
    for edit in (bad_field_name, bad_metadata_key):
        bad = rebuilt(edit)
        with pytest.raises(ob.OrcError) as e:
            ob.ArrowReaderBuilder.try_new(bad)
        assert e.value.variant == "DecodeProto"
        with pytest.raises(oo.OracleError) as e2:
            oo.OracleFile(bad)
        assert e2.value.variant == "DecodeProto"


def test_unparsable_row_index_does_not_fail_the_plan(ob, tmp_path):
    """The reference never reads ROW_INDEX streams when it decodes, so a row index that does not even parse as protobuf
    must not fail the decode: the planner falls back to a sequential decode of that column (found by
    tools/fuzz_struct.py; with_predicate, which does read the index in the reference, keeps every row in that case,
    src/arrow_reader.rs:281-291)."""
    import sys
    import pyarrow as pa
    import pyarrow.orc as po
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import fuzz_struct as fs
    p = str(tmp_path / "i.orc")
    po.write_table(pa.table({"a": pa.array(range(6000), pa.int64()), "b": pa.array([f"s{i % 50}" for i in range(6000)])}), p,
                   compression="uncompressed", row_index_stride=1000, dictionary_key_size_threshold=0.8)
    o = fs.Orc(open(p, "rb").read())
    idx, _ = o.last_stripe_parts()
    whole = ob.DecodeJob([o.build()]).plan().stats()
    damaged = 0
    for at in range(4, len(idx), 7):
        bad = bytearray(idx)
        bad[at] = 0x07  # wire type 7 does not exist
        bad[at + 1] = 0xFF
        data = o.build(index=bytes(bad))
        st = ob.DecodeJob([data]).plan().stats()  # does not raise
        assert st["n_rows"] == whole["n_rows"]
        damaged += st["n_segments"] < whole["n_segments"]
        keep = ob.predicate_row_groups(data, 0, ob.Predicate.eq("a", ob.PredicateValue.Int64(5)))
        assert keep is None or len(keep) == 6
    assert damaged > 3


def test_callback_feed_on_stripes_without_an_index_area(ob):
    """Regression (found by tools/fuzz_host.sh under AddressSanitizer): the per-stripe ranges a callback-fed file keeps in
    memory touch each other, and a stripe without an index area starts its data exactly where the range of the stripe
    before it ends - the lookup of "the loaded range that holds this offset" took the range that merely ended there, so
    the stripe footer (and the staged data) of every stripe but the first came from beyond the wrong buffer:
    TestOrcFile.testWithoutIndex.orc failed to plan through a ChunkReader.  Every multi-stripe fixture plans through
    callbacks exactly as it does from memory, with one read per stripe outside the tail."""
    import glob
    checked = no_index = 0
    for p in sorted(glob.glob(os.path.join(GOLDEN, "ref_*", "*.orc"))):
        try:
            f = ob._File(p)
        except ob.OrcError:
            continue
        if f.num_stripes < 2:
            continue
        try:
            want = ob.DecodeJob([f]).plan().stats()
        except ob.OrcError:
            continue
        cr = ob.FileChunkReader(p)
        got = ob.DecodeJob([ob._File(cr)]).plan().stats()
        for k in ("n_stripes", "n_rows", "n_segments", "input_bytes", "staged_bytes", "n_batches"):
            assert got[k] == want[k], (os.path.basename(p), k)
        assert len(cr.calls) <= f.num_stripes + 2
        checked += 1
        no_index += any(f.stripe_info(i)["index_length"] == 0 for i in range(f.num_stripes))
    assert checked >= 8 and no_index >= 3
