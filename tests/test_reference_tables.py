"""The reference's own end-to-end tests (tests/basic/main.rs), restated: every expected table those tests hold
(tests/golden/ref_basic_tables.json, extracted by tools/extract_ref_tables.py) and every count they assert, against
the oracle on the CPU and against the CUDA path through the C ABI on a GPU."""
import json
import os

import pyarrow as pa
import pytest

from conftest import GOLDEN
from ref_tables_util import assert_table_matches

TABLES = json.load(open(os.path.join(GOLDEN, "ref_basic_tables.json"), encoding="utf-8"))
BASIC = os.path.join(GOLDEN, "ref_basic")


def _check(name, spec, batches, schema, total_row_count):
    """`batches`: the RecordBatches the reader yielded, in order."""
    if "n_batches" in spec:
        assert len(batches) == spec["n_batches"], name
    if "first_batch_len" in spec:
        assert batches[0].num_rows == spec["first_batch_len"], name
    if "batch_lens" in spec:
        assert [b.num_rows for b in batches] == spec["batch_lens"], name
    rows = sum(b.num_rows for b in batches)
    if "total_rows" in spec:
        assert rows == spec["total_rows"], name
    if spec.get("total_rows_is_total_row_count"):
        assert rows == total_row_count, name
    if "schema" in spec:
        assert {f.name: str(f.type) for f in schema} == spec["schema"], name
        assert all(f.nullable for f in schema), name
        assert all(b.schema.equals(schema) for b in batches), name
    if "expected" in spec:
        if "tail" in spec:  # the last rows of the last batch
            last = batches[-1]
            table = pa.Table.from_batches([last.slice(last.num_rows - spec["tail"], spec["tail"])], schema=schema)
        else:
            table = pa.Table.from_batches(batches, schema=schema)
        assert_table_matches(table, spec["expected"], name)


@pytest.mark.parametrize("name", sorted(TABLES))
def test_oracle_against_reference_basic_tests(name):
    from oracle import orc_oracle as oo
    spec = TABLES[name]
    of = oo.OracleFile(open(os.path.join(BASIC, spec["file"]), "rb").read())
    fields = spec.get("fields")
    stripes = None
    if "byte_range" in spec:  # with_file_byte_range: stripes whose offset lies in the range (src/arrow_reader.rs:358-372)
        lo, hi = spec["byte_range"]
        stripes = [i for i, s in enumerate(of.stripes) if lo <= s.offset < hi]
    if spec.get("is_err"):
        with pytest.raises(oo.OracleError):
            of.read(columns=fields)
        return
    batches = of.read(columns=fields, stripes=stripes)
    _check(name, spec, batches, of.schema(fields), sum(s.number_of_rows for s in of.stripes))


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(TABLES))
def test_gpu_against_reference_basic_tests(name):
    import orc_rust_b200 as ob
    assert ob.device_available(), "no CUDA device: the product path has no CPU fallback"
    spec = TABLES[name]
    b = ob.ArrowReaderBuilder.try_new(os.path.join(BASIC, spec["file"]))
    if spec.get("fields"):
        b = b.with_projection(spec["fields"])
    if "byte_range" in spec:
        b = b.with_file_byte_range(*spec["byte_range"])
    reader = b.build()
    total = reader.total_row_count()
    if spec.get("is_err"):
        with pytest.raises(ob.OrcError):
            list(reader)
        return
    _check(name, spec, list(reader), reader.schema(), total)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["TestOrcFile.testWithoutIndex.orc", "TestOrcFile.testMemoryManagementV11.orc",
                                  "TestOrcFile.testMemoryManagementV12.orc"])
def test_gpu_callback_feed_without_index_area(name):
    """Multi-stripe files whose stripes have no index area, seen only through a ChunkReader: the data area of stripe k+1
    starts where the loaded range of stripe k ends (see tests/test_cabi_host.py::
    test_callback_feed_on_stripes_without_an_index_area); the batches are the oracle's, byte for byte."""
    import orc_rust_b200 as ob
    from oracle import orc_oracle as oo
    from parity_util import assert_batches_identical
    assert ob.device_available(), "no CUDA device: the product path has no CPU fallback"
    path = os.path.join(GOLDEN, "ref_integration", name)
    exp = oo.OracleFile(open(path, "rb").read()).read()
    got = list(ob.ArrowReaderBuilder.try_new(ob.FileChunkReader(path)).build())
    assert_batches_identical(got, exp, name)
