"""Byte-level comparison of Arrow batches: product (CUDA, via the C ABI) vs oracle (CPU restatement).

Logical equality is not enough: the reference fixes physical conventions (null slots zeroed, null buffer
omitted when a batch has no nulls, offsets restarting at 0 per batch), so buffers are compared by content
over their meaningful extent."""
import numpy as np
import pyarrow as pa


def _buf(b, n):
    if b is None:
        return None
    return np.frombuffer(b, dtype=np.uint8, count=n)


def assert_array_identical(a: pa.Array, b: pa.Array, what: str):
    assert a.type == b.type, f"{what}: type {a.type} != {b.type}"
    assert len(a) == len(b), f"{what}: length {len(a)} != {len(b)}"
    assert a.offset == 0 and b.offset == 0
    assert a.null_count == b.null_count, f"{what}: null_count {a.null_count} != {b.null_count}"
    n = len(a)
    ba, bb = a.buffers(), b.buffers()
    assert len(ba) == len(bb), f"{what}: buffer count"
    # validity
    assert (ba[0] is None) == (bb[0] is None), f"{what}: validity presence {ba[0] is not None} != {bb[0] is not None}"
    if ba[0] is not None:
        va = np.unpackbits(_buf(ba[0], (n + 7) // 8), bitorder="little")[:n]
        vb = np.unpackbits(_buf(bb[0], (n + 7) // 8), bitorder="little")[:n]
        assert np.array_equal(va, vb), f"{what}: validity bits differ at {np.flatnonzero(va != vb)[:5]}"
    t = a.type
    if pa.types.is_boolean(t):
        xa = np.unpackbits(_buf(ba[1], (n + 7) // 8), bitorder="little")[:n]
        xb = np.unpackbits(_buf(bb[1], (n + 7) // 8), bitorder="little")[:n]
        assert np.array_equal(xa, xb), f"{what}: boolean values differ at {np.flatnonzero(xa != xb)[:5]}"
    elif pa.types.is_string(t) or pa.types.is_binary(t):
        oa = np.frombuffer(ba[1], dtype=np.int32, count=n + 1)
        ob = np.frombuffer(bb[1], dtype=np.int32, count=n + 1)
        assert np.array_equal(oa, ob), f"{what}: offsets differ at {np.flatnonzero(oa != ob)[:5]}"
        total = int(oa[-1])
        if total:
            da = _buf(ba[2], total)
            db = _buf(bb[2], total)
            assert np.array_equal(da, db), f"{what}: string bytes differ at {np.flatnonzero(da != db)[:5]}"
    else:
        w = t.bit_width // 8
        da = _buf(ba[1], n * w)
        db = _buf(bb[1], n * w)
        if not np.array_equal(da, db):
            bad = np.flatnonzero(da != db)[:5] // w
            raise AssertionError(f"{what}: values differ at rows {bad}: {a[int(bad[0])]} vs {b[int(bad[0])]}")


def assert_batches_identical(got, exp, what=""):
    assert len(got) == len(exp), f"{what}: {len(got)} batches vs {len(exp)}"
    for i, (g, e) in enumerate(zip(got, exp)):
        assert g.schema.names == e.schema.names, f"{what} batch {i}: names"
        assert g.num_rows == e.num_rows, f"{what} batch {i}: rows {g.num_rows} vs {e.num_rows}"
        for name in g.schema.names:
            assert_array_identical(g.column(name), e.column(name), f"{what} batch {i} col {name}")
