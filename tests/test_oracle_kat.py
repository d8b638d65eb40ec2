"""Pins the oracle's stream codecs to the reference's own known-answer vectors (tests/kat_vectors.py)."""
import numpy as np
import pytest

import kat_vectors as kv
from oracle import orc_oracle as oo


@pytest.mark.parametrize("name,data,signed,expected", kv.RLE_V2, ids=[k[0] for k in kv.RLE_V2])
def test_rle_v2_kat(name, data, signed, expected):
    out = oo.rle_v2(bytes(data), len(expected), signed, 8)
    assert out.tolist() == expected


@pytest.mark.parametrize("name,data,signed,expected", kv.RLE_V1, ids=[k[0] for k in kv.RLE_V1])
def test_rle_v1_kat(name, data, signed, expected):
    out = oo.rle_v1(bytes(data), len(expected), signed, 8)
    assert out.tolist() == expected


@pytest.mark.parametrize("name,data,expected", kv.BYTE_RLE, ids=[k[0] for k in kv.BYTE_RLE])
def test_byte_rle_kat(name, data, expected):
    assert oo.byte_rle(bytes(data), len(expected)).tolist() == expected


@pytest.mark.parametrize("name,data,expected", kv.BOOL_RLE, ids=[k[0] for k in kv.BOOL_RLE])
def test_bool_rle_kat(name, data, expected):
    assert oo.bool_rle(bytes(data), len(expected)).tolist() == expected


def test_varint_kat():
    for data, expected in kv.VARINT_U64:
        # a one-element RLEv1 literal run (header 0xff) carries exactly one read_varint::<i64>
        assert oo.rle_v1(bytes([0xFF] + data), 1, False, 8).tolist() == [expected]
    with pytest.raises(oo.OracleError) as e:
        oo.rle_v1(bytes([0xFF] + kv.VARINT_TOO_LARGE), 1, False, 8)
    assert e.value.variant == "VarintTooLarge"
    with pytest.raises(oo.OracleError) as e:
        oo.rle_v1(bytes([0xFF] + kv.VARINT_TRUNCATED), 1, False, 8)
    assert e.value.variant == "IoError"


def test_decimal_varint_kat():
    for data, expected in kv.DECIMAL_VARINT:
        out = oo.varint_i128(bytes(data), len(expected))
        got = [int(lo) | (int(hi) << 64) for lo, hi in out.tolist()]
        got = [g - (1 << 128) if g >> 127 else g for g in got]
        assert got == expected
    with pytest.raises(oo.OracleError):
        oo.varint_i128(bytes([0x00, 0x02, 0x01]), 4)  # EOF


def test_chunk_header_kat():
    for data, expected in kv.CHUNK_HEADER:
        assert oo.chunk_header(bytes(data)) == expected


def test_delta_reference_unit_tests():
    """delta.rs:194-284 — the reference builds these with its encoder; bytes here are hand-encoded."""
    # fixed delta +10, len 100, base 0 (unsigned): header 0xc0 | len-1=99 -> [0xc0, 99], base 0, delta zz(10)=20
    out = oo.rle_v2(bytes([0xC0, 99, 0x00, 20]), 100, False, 8)
    assert out.tolist() == [i * 10 for i in range(100)]
    # fixed delta -63, len 150, base 10000 (unsigned varint 0x90 0x4e), delta zz(-63)=125
    out = oo.rle_v2(bytes([0xC0, 149, 0x90, 0x4E, 125]), 150, False, 8)
    assert out.tolist() == [10000 - i * 63 for i in range(150)]


def test_i16_i32_range_checks():
    """integer/mod.rs:236-313 and delta.rs:286-310: checked add/sub in the target width."""
    # i32 delta run overflowing i32::MAX: base = 2^31-1 (signed zigzag varint), delta +1, len 2
    base_zz = (2**31 - 1) << 1
    vb = []
    v = base_zz
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            vb.append(b | 0x80)
        else:
            vb.append(b)
            break
    with pytest.raises(oo.OracleError) as e:
        oo.rle_v2(bytes([0xC0, 1] + vb + [2]), 2, True, 4)
    assert e.value.variant == "OutOfSpec"
    assert oo.rle_v2(bytes([0xC0, 1] + vb + [2]), 2, True, 8).tolist() == [2**31 - 1, 2**31]
    # DIRECT width 32 into i16 is out of spec (direct.rs:47-52)
    with pytest.raises(oo.OracleError):
        oo.rle_v2(bytes([0x40 | (27 << 1), 0, 0, 0, 0, 1]), 1, True, 2)
    # SHORT_REPEAT width 3 bytes into i16 (short_repeat.rs:44-51)
    with pytest.raises(oo.OracleError):
        oo.rle_v2(bytes([(2 << 3) | 0, 0, 0, 1]), 3, True, 2)


def test_with_schema_timestamp_goldens():
    """The reference's with_schema tests (tests/basic/main.rs:594-770): other units and Decimal128(38, 9), with and
    without a writer time zone."""
    import os
    from conftest import GOLDEN
    from oracle import orc_oracle as oo
    import kat_vectors as kv
    for name, rel, units, bs, col, expected in kv.TS_SCHEMA:
        of = oo.OracleFile(open(os.path.join(GOLDEN, rel), "rb").read())
        batch = of.read(batch_size=bs, ts_unit=units, stripes=[0])[0]
        assert batch.column(batch.schema.get_field_index(col)).to_pylist()[: len(expected)] == expected, name
