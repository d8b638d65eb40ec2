"""Pins the file-level oracle (oracle/orc_oracle.py) against
  (a) the reference's own expected_arrow feather goldens (tests/integration/main.rs:35-70), and
  (b) pyarrow.orc (Apache ORC C++), an independent reader, on the reference's fixture files.
Fixture files are copies of /root/reference/tests/{basic,integration}/data (data only, no source)."""
import glob
import os

import pyarrow as pa
import pyarrow.feather as feather
import pyarrow.orc as po
import pytest

from oracle import orc_oracle as oo

from conftest import GOLDEN


def _oracle_table(path, **kw):
    of = oo.OracleFile(open(path, "rb").read())
    batches = of.read(**kw)
    return of, (pa.Table.from_batches(batches, schema=of.schema()) if batches else of.schema().empty_table())


def _flat_files(sub):
    out = []
    for f in sorted(glob.glob(os.path.join(GOLDEN, sub, "*.orc"))):
        try:
            of = oo.OracleFile(open(f, "rb").read())
        except oo.OracleError:
            continue
        if of.is_flat():
            out.append(f)
    return out


FEATHER = [f for f in sorted(glob.glob(os.path.join(GOLDEN, "ref_expected_arrow", "*.feather")))]


@pytest.mark.parametrize("fpath", FEATHER, ids=[os.path.basename(f) for f in FEATHER])
def test_oracle_vs_reference_feather(fpath):
    name = os.path.basename(fpath)[: -len(".feather")]
    orc = os.path.join(GOLDEN, "ref_integration", name + ".orc")
    if name == "orc-file-11-format":
        pytest.skip("0.11 file without metadataLength: ignored by the reference (tests/integration/main.rs:334-337)")
    of = oo.OracleFile(open(orc, "rb").read())
    if name == "orc_split_elim":
        pytest.skip("DECIMAL(0,0): ignored by the reference itself (tests/integration/main.rs:347-351)")
    _, got = _oracle_table(orc)
    exp = feather.read_table(fpath)
    assert got.num_rows == exp.num_rows
    for c in got.column_names:
        a = got[c].combine_chunks()
        b = exp[c].combine_chunks()
        if not of.is_flat([c]):
            # nested: the feathers were written by pyarrow (map fields key / value, dense unions; the reference's test
            # normalises the same way, tests/integration/main.rs:96-130): compare the values
            assert a.to_pylist() == b.to_pylist(), f"column {c} differs"
            continue
        if a.type != b.type:
            b = b.cast(a.type)
        assert a.equals(b), f"column {c} differs"


def _nested_files():
    out = []
    for sub in ("ref_basic", "ref_integration"):
        for f in sorted(glob.glob(os.path.join(GOLDEN, sub, "*.orc"))):
            try:
                of = oo.OracleFile(open(f, "rb").read())
            except oo.OracleError:
                continue
            if not of.is_flat():
                out.append(f)
    return out


NESTED = _nested_files()


@pytest.mark.parametrize("fpath", NESTED, ids=[os.path.basename(f) for f in NESTED])
def test_oracle_nested_vs_pyarrow(fpath):
    """struct / list / map / union columns (src/array_decoder/{struct_decoder,list,map,union}.rs) against the Apache
    reader, by value; the batch layout is checked separately below."""
    of, got = _oracle_table(fpath)
    exp = po.read_table(fpath)
    assert got.num_rows == exp.num_rows
    for c in got.column_names:
        assert got[c].to_pylist() == exp[c].to_pylist(), f"column {c} differs"


def test_oracle_nested_layout():
    """Arrow types of nested columns (src/schema.rs:530-577) and per-batch validity omission for nested nodes."""
    of = oo.OracleFile(open(os.path.join(GOLDEN, "ref_basic", "nested_map.orc"), "rb").read())
    t = of.schema().field("map").type
    assert t.key_field.name == "keys" and not t.key_field.nullable and t.item_field.name == "values" and t.item_field.nullable
    of = oo.OracleFile(open(os.path.join(GOLDEN, "ref_basic", "nested_array.orc"), "rb").read())
    t = of.schema().field("value").type
    assert t.value_field.name == "item" and t.value_field.nullable
    of = oo.OracleFile(open(os.path.join(GOLDEN, "ref_integration", "TestOrcFile.testUnionAndTimestamp.orc"), "rb").read())
    t = of.schema().field("union").type
    assert t.mode == "sparse" and [t.field(i).name for i in range(t.num_fields)] == ["_union_0", "_union_1"]
    for b in of.read(batch_size=1000):
        u = b.column("union")
        assert len(u.field(0)) == len(u) and len(u.field(1)) == len(u)  # sparse: every child spans the batch


ALL_FLAT = _flat_files("ref_basic") + _flat_files("ref_integration")


@pytest.mark.parametrize("fpath", ALL_FLAT, ids=[os.path.basename(f) for f in ALL_FLAT])
def test_oracle_vs_pyarrow(fpath):
    name = os.path.basename(fpath)
    if name in ("orc_split_elim.orc",):
        pytest.skip("DECIMAL(0,0)")
    try:
        _, got = _oracle_table(fpath)
    except oo.OracleError as e:
        # files the reference itself rejects (tests/basic/main.rs:588-592 overflowing timestamps, ...)
        if name in ("overflowing_timestamps.orc", "timestamps_0001.orc", "decimal64_v2.orc",
                    "decimal64_v2_cplusplus.orc"):
            pytest.skip(f"reference errors on this file too: {e}")
        raise
    exp = po.read_table(fpath)
    assert got.num_rows == exp.num_rows
    for c in got.column_names:
        a = got[c].combine_chunks()
        b = exp[c].combine_chunks()
        if a.type != b.type:
            b = b.cast(a.type)
        assert a.equals(b), f"column {c} differs"


def test_batch_layout_conventions():
    """Physical conventions the reference implements but arrow's logical == does not pin
    (SURVEY §8(b) 'Batch semantics to preserve')."""
    path = os.path.join(GOLDEN, "ref_integration", "nulls-at-end-snappy.orc")
    of = oo.OracleFile(open(path, "rb").read())
    batches = of.read(batch_size=8192)
    assert all(b.num_rows <= 8192 for b in batches)
    assert sum(b.num_rows for b in batches) == of.number_of_rows
    saw_no_validity = saw_validity = False
    for b in batches:
        for col in b.columns:
            bufs = col.buffers()
            if col.null_count == 0:
                assert bufs[0] is None  # null buffer omitted when the batch has no nulls
                saw_no_validity = True
            else:
                assert bufs[0] is not None
                saw_validity = True
            if pa.types.is_string(col.type) or pa.types.is_binary(col.type):
                import numpy as np
                offs = np.frombuffer(bufs[1], dtype=np.int32, count=len(col) + 1)
                assert offs[0] == 0  # offsets restart per batch
    assert saw_no_validity and saw_validity
