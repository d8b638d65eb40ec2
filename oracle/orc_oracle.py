"""ORACLE — TEST INFRASTRUCTURE ONLY (not the product path; the CUDA library never imports this).

File-level CPU restatement of orc-rust v0.8.0's ArrowReader decode path: file tail -> stripes ->
per-column decoders -> per-batch Arrow buffers with the reference's physical conventions
(8192-row batches that never span stripes, null slots zero-filled, null buffer omitted when a
batch has no nulls, string offsets restarting at 0 per batch).  The byte-level codecs live in
oracle/codecs.c; this module restates the planner half:

  read_metadata              src/reader/metadata.rs:180-263
  Stripe::new stream offsets src/stripe.rs:128-182
  Column::rle_version etc.   src/column.rs:40-59
  array_decoder_factory      src/array_decoder/mod.rs:390-511
  NaiveStripeDecoder batches src/array_decoder/mod.rs:371-387, 514-564
  string decoders            src/array_decoder/string.rs:51-153, 205-224
  decimal / timestamp        src/array_decoder/decimal.rs:36-166, timestamp.rs:51-314
  schema mapping             src/schema.rs:390-577

Parity pinning: tests/test_oracle_kat.py (reference unit-test vectors), tests/test_oracle_files.py
(reference expected_arrow feather goldens + pyarrow.orc as an independent reader) and
tests/test_reference_tables.py (the expected tables of the reference's tests/basic/main.rs).
Nested columns (struct / list / map / union) are restated as a recursive walk (OracleFile._decode_node).
"""
from __future__ import annotations

import ctypes
import datetime as _dt
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import build as _build

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(_build.build())
        _lib.orc_oracle_decompress_stream.restype = ctypes.c_int64
        _lib.orc_oracle_decompress_bound.restype = ctypes.c_int64
    return _lib


# OrcError ordinal + 1 (src/error.rs:31-174)
ERR_NAMES = {
    1: "IoError", 2: "EmptyFile", 3: "OutOfSpec", 4: "DecodeFloat", 5: "DecodeTimestamp", 6: "OffsetOverflow",
    7: "DecodeProto", 8: "NoTypes", 9: "UnsupportedTypeVariant", 10: "MismatchedSchema",
    11: "ConvertRecordBatch", 12: "VarintTooLarge", 13: "Unexpected", 14: "BuildZstdDecoder",
    15: "BuildSnappyDecoder", 16: "BuildLzoDecoder", 17: "BuildLz4Decoder", 18: "Arrow",
}


class OracleError(Exception):
    def __init__(self, code: int, msg: str = ""):
        self.code = code
        self.variant = ERR_NAMES.get(code, f"code{code}")
        super().__init__(f"{self.variant}: {msg}")


def _check(rc: int, what: str = ""):
    if rc != 0:
        raise OracleError(rc, what)


def _u8p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


# ------------------------------------------------------------------------------------------------
# stream-level wrappers around codecs.c (also used directly by the KAT tests)
# ------------------------------------------------------------------------------------------------
def _as_u8(buf) -> np.ndarray:
    a = np.frombuffer(buf, dtype=np.uint8) if not isinstance(buf, np.ndarray) else buf
    return np.ascontiguousarray(a)


def rle_v2(buf, n: int, signed: bool, nbytes: int = 8) -> np.ndarray:
    a = _as_u8(buf)
    out = np.empty(n, dtype=np.int64)
    _check(lib().orc_oracle_rle_v2(_u8p(a), ctypes.c_size_t(a.size), int(signed), nbytes, _u8p(out),
                                   ctypes.c_size_t(n), None), "rle_v2")
    return out


def rle_v1(buf, n: int, signed: bool, nbytes: int = 8) -> np.ndarray:
    a = _as_u8(buf)
    out = np.empty(n, dtype=np.int64)
    _check(lib().orc_oracle_rle_v1(_u8p(a), ctypes.c_size_t(a.size), int(signed), nbytes, _u8p(out),
                                   ctypes.c_size_t(n), None), "rle_v1")
    return out


def int_rle(buf, n, signed, nbytes, version):
    return rle_v2(buf, n, signed, nbytes) if version == 2 else rle_v1(buf, n, signed, nbytes)


def byte_rle(buf, n: int) -> np.ndarray:
    a = _as_u8(buf)
    out = np.empty(n, dtype=np.uint8)
    _check(lib().orc_oracle_byte_rle(_u8p(a), ctypes.c_size_t(a.size), _u8p(out), ctypes.c_size_t(n), None),
           "byte_rle")
    return out


def bool_rle(buf, n: int) -> np.ndarray:
    a = _as_u8(buf)
    out = np.empty(n, dtype=np.uint8)
    _check(lib().orc_oracle_bool_rle(_u8p(a), ctypes.c_size_t(a.size), _u8p(out), ctypes.c_size_t(n)), "bool_rle")
    return out


def varint_i128(buf, n: int) -> np.ndarray:
    """Returns an (n, 2) uint64 array: little-endian (lo, hi) halves of each i128."""
    a = _as_u8(buf)
    out = np.empty((n, 2), dtype=np.uint64)
    _check(lib().orc_oracle_varint_i128(_u8p(a), ctypes.c_size_t(a.size), _u8p(out), ctypes.c_size_t(n), None),
           "varint_i128")
    return out


def chunk_header(b3: bytes) -> Tuple[int, bool]:
    ln = ctypes.c_uint32()
    orig = ctypes.c_int()
    lib().orc_oracle_chunk_header(bytes(b3), ctypes.byref(ln), ctypes.byref(orig))
    return ln.value, bool(orig.value)


def decompress_stream(kind: int, buf, block_size: int, stats: Optional[np.ndarray] = None) -> np.ndarray:
    a = _as_u8(buf)
    if kind == 0:
        return a
    bound = lib().orc_oracle_decompress_bound(kind, _u8p(a), ctypes.c_size_t(a.size), ctypes.c_size_t(block_size))
    out = np.empty(max(int(bound), 1), dtype=np.uint8)
    r = lib().orc_oracle_decompress_stream(kind, _u8p(a), ctypes.c_size_t(a.size), ctypes.c_size_t(block_size),
                                           _u8p(out), ctypes.c_size_t(out.size),
                                           _u8p(stats) if stats is not None else None)
    if r < 0:
        raise OracleError(int(-r), "decompress")
    return out[: int(r)]


# ------------------------------------------------------------------------------------------------
# minimal protobuf reader (the reference uses prost-generated src/proto.rs)
# ------------------------------------------------------------------------------------------------
def _pb_varint(b: bytes, p: int) -> Tuple[int, int]:
    r = 0
    s = 0
    while True:
        if p >= len(b):
            raise OracleError(7, "truncated varint")
        c = b[p]
        p += 1
        r |= (c & 0x7F) << s
        s += 7
        if not (c & 0x80):
            return r, p


def pb_fields(b: bytes):
    if isinstance(b, int):  # a varint where a message is expected (prost: invalid wire type)
        raise OracleError(7, "wire type of a message field")
    p = 0
    n = len(b)
    while p < n:
        key, p = _pb_varint(b, p)
        fno, wt = key >> 3, key & 7
        if fno == 0 or key >> 32:  # prost: "invalid tag value: 0" / "invalid key value"
            raise OracleError(7, "invalid protobuf key")
        if wt == 0:
            v, p = _pb_varint(b, p)
        elif wt == 1:
            v = b[p:p + 8]
            if len(v) != 8:
                raise OracleError(7, "truncated fixed64 field")
            p += 8
        elif wt == 2:
            ln, p = _pb_varint(b, p)
            v = b[p:p + ln]
            if len(v) != ln:
                raise OracleError(7, "truncated bytes field")
            p += ln
        elif wt == 5:
            v = b[p:p + 4]
            if len(v) != 4:
                raise OracleError(7, "truncated fixed32 field")
            p += 4
        else:
            raise OracleError(7, f"wire type {wt}")
        yield fno, wt, v


def _packed_u(v, wt) -> List[int]:
    if wt == 0:
        return [v]
    out = []
    p = 0
    while p < len(v):
        x, p = _pb_varint(v, p)
        out.append(x)
    return out


K_BOOLEAN, K_BYTE, K_SHORT, K_INT, K_LONG, K_FLOAT, K_DOUBLE, K_STRING, K_BINARY, K_TIMESTAMP, K_LIST, K_MAP, \
    K_STRUCT, K_UNION, K_DECIMAL, K_DATE, K_VARCHAR, K_CHAR, K_TIMESTAMP_INSTANT = range(19)

S_PRESENT, S_DATA, S_LENGTH, S_DICTIONARY_DATA, S_DICTIONARY_COUNT, S_SECONDARY, S_ROW_INDEX = range(7)


@dataclass
class OrcType:
    kind: int = K_BOOLEAN
    subtypes: List[int] = field(default_factory=list)
    field_names: List[str] = field(default_factory=list)
    maximum_length: int = 0
    precision: int = 0
    scale: int = 0


@dataclass
class StripeInfo:
    offset: int = 0
    index_length: int = 0
    data_length: int = 0
    footer_length: int = 0
    number_of_rows: int = 0


@dataclass
class StreamInfo:
    kind: int
    column: int
    length: int
    offset: int


ORC_EPOCH = 1_420_070_400  # array_decoder/timestamp.rs:51


class OracleFile:
    """ArrowReaderBuilder::try_new + build + drain, restated (src/arrow_reader.rs:202-346)."""

    def __init__(self, data: bytes):
        self.data = bytes(data)
        n = len(self.data)
        if n == 0:
            raise OracleError(2, "empty file")
        ps_len = self.data[-1]
        if n - 1 < ps_len:
            raise OracleError(3, "File too small for given postscript length")
        ps = self.data[n - 1 - ps_len:n - 1]
        footer_len = meta_len = None
        self.compression = 0
        self.block_size = 256 * 1024  # compression.rs:31
        for f, wt, v in pb_fields(ps):
            if f == 1:
                footer_len = v
            elif f == 2:
                self.compression = v
            elif f == 3:
                self.block_size = v
            elif f == 5:
                meta_len = v
        if footer_len is None:
            raise OracleError(3, "Footer length is empty")
        if meta_len is None:
            raise OracleError(3, "Metadata length is empty")
        fend = n - 1 - ps_len
        footer_raw = self.data[fend - footer_len:fend]
        footer = bytes(decompress_stream(self.compression, footer_raw, self.block_size))
        self.types: List[OrcType] = []
        self.stripes: List[StripeInfo] = []
        self.user_metadata: Dict[str, bytes] = {}
        self.number_of_rows = 0
        self.row_index_stride = None
        for f, wt, v in pb_fields(footer):
            if f == 3:
                s = StripeInfo()
                for g, _, w in pb_fields(v):
                    if g == 1:
                        s.offset = w
                    elif g == 2:
                        s.index_length = w
                    elif g == 3:
                        s.data_length = w
                    elif g == 4:
                        s.footer_length = w
                    elif g == 5:
                        s.number_of_rows = w
                self.stripes.append(s)
            elif f == 4:
                t = OrcType()
                for g, gw, w in pb_fields(v):
                    if g == 1:
                        t.kind = w
                    elif g == 2:
                        t.subtypes.extend(_packed_u(w, gw))
                    elif g == 3:
                        t.field_names.append(_pb_str(w, gw))
                    elif g == 4:
                        t.maximum_length = w
                    elif g == 5:
                        t.precision = w
                    elif g == 6:
                        t.scale = w
                self.types.append(t)
            elif f == 5:
                name, val = "", b""
                for g, gw, w in pb_fields(v):
                    if g == 1:
                        name = _pb_str(w, gw)
                    elif g == 2:
                        val = bytes(w)
                self.user_metadata[name] = val
            elif f == 6:
                self.number_of_rows = v
            elif f == 8:
                self.row_index_stride = v
        if not self.types:
            raise OracleError(8, "No types found")
        root = self.types[0]
        if root.kind != K_STRUCT:
            raise OracleError(13, "non-struct root")
        self._check_type_tree(0, 0)
        self.columns: List[Tuple[str, int]] = list(zip(root.field_names, root.subtypes))

    def _check_type_tree(self, cid: int, depth: int):
        """RootDataType::from_proto / DataType::from_proto (src/schema.rs:154-162, 206-234, 390-487): every type reachable
        from the root must exist and have the children its kind takes.  (The reference recurses without a visited set: a
        type that contains itself overflows its stack; here that is an Unexpected error like the other shapes.)"""
        if depth > 256 or cid >= len(self.types):
            raise OracleError(13, f"Column index out of bounds: {cid}")
        t = self.types[cid]
        n = len(t.subtypes)
        if t.kind == K_STRUCT and n != len(t.field_names):
            raise OracleError(13, f"Struct type for column index {cid} must have matching lengths for subtypes and field names lists")
        if t.kind == K_LIST and n != 1:
            raise OracleError(13, f"List type for column index {cid} must have 1 sub type, found {n}")
        if t.kind == K_MAP and n != 2:
            raise OracleError(13, f"Map type for column index {cid} must have 2 sub types, found {n}")
        if t.kind == K_UNION and n > 127:
            raise OracleError(13, f"Union type for column index {cid} cannot exceed 127 variants, found {n}")
        if t.kind in (K_STRUCT, K_LIST, K_MAP, K_UNION):
            for c in t.subtypes:
                self._check_type_tree(c, depth + 1)

    # --- schema (src/schema.rs:503-577) ---------------------------------------------------------
    def arrow_type(self, col_id: int, ts_unit: str = "ns"):
        import pyarrow as pa
        t = self.types[col_id]
        k = t.kind
        simple = {K_BOOLEAN: pa.bool_(), K_BYTE: pa.int8(), K_SHORT: pa.int16(), K_INT: pa.int32(),
                  K_LONG: pa.int64(), K_FLOAT: pa.float32(), K_DOUBLE: pa.float64(), K_STRING: pa.utf8(),
                  K_VARCHAR: pa.utf8(), K_CHAR: pa.utf8(), K_BINARY: pa.binary(), K_DATE: pa.date32()}
        if k in simple:
            return simple[k]
        if k == K_DECIMAL:
            # arrow-rs validates the type when the array is built (array_decoder/decimal.rs:99): precision 1..=38, scale <= precision
            if t.precision == 0 or t.precision > 38 or t.scale > t.precision:
                raise OracleError(18, f"invalid Decimal128 precision / scale ({t.precision}, {t.scale})")
            return pa.decimal128(t.precision, t.scale)
        if k in (K_TIMESTAMP, K_TIMESTAMP_INSTANT) and ts_unit == "dec":
            return pa.decimal128(38, 9)  # with_schema: Decimal128(38, 9) nanoseconds (array_decoder/timestamp.rs:150-232)
        if k == K_TIMESTAMP:
            return pa.timestamp(ts_unit)
        if k == K_TIMESTAMP_INSTANT:
            return pa.timestamp(ts_unit, tz="UTC")
        # nested types (src/schema.rs:530-577)
        if k == K_STRUCT:
            return pa.struct([pa.field(n, self.arrow_type(c, ts_unit), True) for n, c in zip(t.field_names, t.subtypes)])
        if k == K_LIST:
            return pa.list_(pa.field("item", self.arrow_type(t.subtypes[0], ts_unit), True))
        if k == K_MAP:
            return pa.map_(pa.field("keys", self.arrow_type(t.subtypes[0], ts_unit), False),
                           pa.field("values", self.arrow_type(t.subtypes[1], ts_unit), True))
        if k == K_UNION:
            return pa.sparse_union([pa.field(f"_union_{i}", self.arrow_type(c, ts_unit), True) for i, c in enumerate(t.subtypes)],
                                   type_codes=list(range(len(t.subtypes))))
        raise NotImplementedError(f"ORC type kind {k}")

    def schema(self, columns: Optional[List[str]] = None, ts_unit: str = "ns"):
        import pyarrow as pa
        fields = [pa.field(name, self.arrow_type(cid, _unit_of(ts_unit, name)), True) for name, cid in self._projected(columns)]
        md = {k: v.decode("utf-8", "replace") for k, v in self.user_metadata.items()}
        return pa.schema(fields, metadata=md or None)

    def _projected(self, columns):
        if columns is None:
            return list(self.columns)
        return [(n, c) for n, c in self.columns if n in columns]

    def is_flat(self, columns=None) -> bool:
        return all(self.types[c].kind not in (K_LIST, K_MAP, K_STRUCT, K_UNION) for _, c in self._projected(columns))

    # --- stripe (src/stripe.rs:128-182) ---------------------------------------------------------
    def _stripe_footer(self, s: StripeInfo):
        off = s.offset + s.index_length + s.data_length
        raw = self.data[off:off + s.footer_length]
        b = bytes(decompress_stream(self.compression, raw, self.block_size))
        streams: List[StreamInfo] = []
        encodings: List[Tuple[int, int]] = []
        tz = None
        pos = s.offset
        for f, fw, v in pb_fields(b):
            if f == 1:
                kind = column = length = 0
                for g, _, w in pb_fields(v):
                    if g == 1:
                        kind = w
                    elif g == 2:
                        column = w
                    elif g == 3:
                        length = w
                streams.append(StreamInfo(kind, column, length, pos))
                pos += length
            elif f == 2:
                ek = ds = 0
                for g, _, w in pb_fields(v):
                    if g == 1:
                        ek = w
                    elif g == 2:
                        ds = w
                encodings.append((ek, ds))
            elif f == 3:
                tz = _pb_str(v, fw)
        return streams, encodings, tz

    def _stream(self, smap, col, kind) -> np.ndarray:
        """StreamMap::get (stripe.rs:319-335): missing stream decodes as empty."""
        st = smap.get((col, kind))
        if st is None:
            return np.empty(0, dtype=np.uint8)
        raw = np.frombuffer(self.data, dtype=np.uint8, count=st.length, offset=st.offset)
        return decompress_stream(self.compression, raw, self.block_size)

    def _decode_leaf(self, smap, encodings, tz, cid, nn, name, ts_unit):
        """The value streams of one primitive column: `nn` non-null values (array_decoder_factory, mod.rs:390-463)."""
        t = self.types[cid]
        ek, dict_size = encodings[cid] if cid < len(encodings) else (0, 0)
        ver = 2 if ek in (2, 3) else 1  # column.rs:52-59
        k = t.kind
        if k == K_BOOLEAN:
            vals = bool_rle(self._stream(smap, cid, S_DATA), nn)
            payload = ("bool", vals)
        elif k == K_BYTE:
            payload = ("prim", byte_rle(self._stream(smap, cid, S_DATA), nn).view(np.int8))
        elif k in (K_SHORT, K_INT, K_LONG, K_DATE):
            nb = {K_SHORT: 2, K_INT: 4, K_LONG: 8, K_DATE: 4}[k]
            v = int_rle(self._stream(smap, cid, S_DATA), nn, True, nb, ver)
            payload = ("prim", v.astype({2: np.int16, 4: np.int32, 8: np.int64}[nb]))
        elif k in (K_FLOAT, K_DOUBLE):
            w = 4 if k == K_FLOAT else 8
            raw = self._stream(smap, cid, S_DATA)
            if raw.size < nn * w:
                raise OracleError(1, "float stream too short")  # read_exact -> IoError
            payload = ("prim", raw[: nn * w].copy().view(np.float32 if k == K_FLOAT else np.float64))
        elif k in (K_STRING, K_VARCHAR, K_CHAR, K_BINARY):
            if k != K_BINARY and ek in (1, 3):
                dl = int_rle(self._stream(smap, cid, S_LENGTH), dict_size, False, 8, ver)
                doff = np.empty(dict_size + 1, dtype=np.int32)
                _check(lib().orc_oracle_offsets(_u8p(dl), ctypes.c_size_t(dict_size), _u8p(doff)), "dict offsets")
                ddata = self._stream(smap, cid, S_DICTIONARY_DATA)
                if ddata.size < int(doff[-1]):
                    # the dictionary is built by the same next_byte_batch as direct strings (string.rs:65-74):
                    # offsets past the bytes that could be read fail GenericByteArray::try_new
                    raise OracleError(18, "dictionary data shorter than its lengths")
                _validate_utf8(ddata, doff)  # the dictionary is a StringArray of its own
                keys = int_rle(self._stream(smap, cid, S_DATA), nn, False, 8, ver)
                payload = ("dict", keys, doff, np.ascontiguousarray(ddata))
            else:
                lens = int_rle(self._stream(smap, cid, S_LENGTH), nn, False, 8, ver)
                payload = ("bytes", lens, self._stream(smap, cid, S_DATA))
        elif k == K_DECIMAL:
            v = varint_i128(self._stream(smap, cid, S_DATA), nn)
            sc = int_rle(self._stream(smap, cid, S_SECONDARY), nn, True, 4, ver)
            lib().orc_oracle_decimal_fix_scale(_u8p(v), _u8p(sc), ctypes.c_size_t(nn), ctypes.c_uint32(t.scale))
            payload = ("prim", v)
        elif k in (K_TIMESTAMP, K_TIMESTAMP_INSTANT):
            d = int_rle(self._stream(smap, cid, S_DATA), nn, True, 8, ver)
            sec = int_rle(self._stream(smap, cid, S_SECONDARY), nn, False, 8, ver)
            base = ORC_EPOCH
            zone = None
            if k == K_TIMESTAMP and tz is not None:
                zone = _zone(tz)
                base = int(_dt.datetime(2015, 1, 1, tzinfo=zone).timestamp())
            unit = _unit_of(ts_unit, name)
            moved = zone is not None and tz not in ("UTC", "GMT", "Etc/UTC", "Etc/GMT")
            if unit == "dec":
                # TimestampNanosecondAsDecimalDecoder (+ ...WithTzDecoder, array_decoder/timestamp.rs:316-333)
                o = np.empty((nn, 2), dtype=np.uint64)
                lib().orc_oracle_timestamp_i128(_u8p(d), _u8p(sec), ctypes.c_size_t(nn), ctypes.c_int64(base), _u8p(o))
                if moved:
                    vals = [(int(hi) << 64 | int(lo)) - ((int(hi) >> 63) << 128) for lo, hi in o.tolist()]
                    epoch = _dt.datetime(1970, 1, 1, tzinfo=_dt.timezone.utc)
                    for i, v in enumerate(vals):
                        off = int((epoch + _dt.timedelta(seconds=v // 1_000_000_000)).astimezone(zone).utcoffset().total_seconds())
                        w = (v + off * 1_000_000_000) & ((1 << 128) - 1)
                        o[i, 0], o[i, 1] = w & 0xFFFFFFFFFFFFFFFF, w >> 64
            else:
                unit_ns = {"ns": 1, "us": 1000, "ms": 1_000_000, "s": 1_000_000_000}[unit]
                o = np.empty(nn, dtype=np.int64)
                _check(lib().orc_oracle_timestamp(_u8p(d), _u8p(sec), ctypes.c_size_t(nn), ctypes.c_int64(base),
                                                  ctypes.c_int64(unit_ns), _u8p(o)), "timestamp")
                if moved:
                    # a value the move pushes out of the unit's range becomes a null (try_unary, then unary_opt:
                    # array_decoder/timestamp.rs:277-283)
                    o, over = _tz_to_utc(o, zone, unit_ns)
                    if over.any():
                        return ("prim", o, over)
            payload = ("prim", o)
        else:
            raise NotImplementedError(f"ORC type kind {k}")
        return payload

    def _decode_node(self, smap, encodings, tz, cid, n, parent_present, name, ts_unit):
        """One column over `n` slots of its parent (rows for a root column), recursively.  parent_present: bool array over
        the slots or None.  Returns {"cid", "kind", "n", "present" (merged, derive_present_vec mod.rs:231-252, or None),
        ...}: leaves carry "payload" and "rows"; struct "children"; list / map "lens" (per slot, 0 for nulls) and
        "children" over sum(lens) slots; union "tags" (per slot, 0 for nulls) and "children"."""
        t = self.types[cid]
        k = t.kind
        ek, _ = encodings[cid] if cid < len(encodings) else (0, 0)
        ver = 2 if ek in (2, 3) else 1
        present = parent_present
        if (cid, S_PRESENT) in smap:
            count = n if parent_present is None else int(parent_present.sum())
            own = bool_rle(self._stream(smap, cid, S_PRESENT), count)
            if parent_present is None:
                present = own
            else:  # merge_parent_present (mod.rs:216-229)
                present = np.zeros(n, dtype=own.dtype)
                present[parent_present.astype(bool)] = own
        nn = int(present.sum()) if present is not None else n
        node = {"cid": cid, "kind": k, "n": n, "present": present, "name": name}
        if k == K_STRUCT:  # struct_decoder.rs:59-78
            node["children"] = [self._decode_node(smap, encodings, tz, c, n, present, cn, ts_unit)
                                for cn, c in zip(t.field_names, t.subtypes)]
        elif k in (K_LIST, K_MAP):  # list.rs:63-87, map.rs:74-104
            lens = int_rle(self._stream(smap, cid, S_LENGTH), nn, False, 8, ver)
            rows = _to_rows(present, ("prim", lens), n)
            total = int(rows.sum())
            node["lens"] = rows
            node["children"] = [self._decode_node(smap, encodings, tz, c, total, None, "", ts_unit) for c in t.subtypes]
        elif k == K_UNION:  # union.rs:69-136
            tags = byte_rle(self._stream(smap, cid, S_DATA), nn).view(np.int8)
            rows = _to_rows(present, ("prim", tags), n)
            node["tags"] = rows
            kids = []
            for i, c in enumerate(t.subtypes):
                cp = rows == i
                if i == 0 and present is not None:
                    cp = cp & present.astype(bool)
                kids.append(self._decode_node(smap, encodings, tz, c, n, cp.astype(np.uint8), "", ts_unit))
            node["children"] = kids
        else:
            payload = self._decode_leaf(smap, encodings, tz, cid, nn, name, ts_unit)
            if payload[0] == "prim" and len(payload) == 3:
                # dense values that turned into nulls (writer-zone move out of range): out of the validity they go
                over = payload[2]
                present = np.ones(n, dtype=np.uint8) if present is None else present.copy()
                idx = np.flatnonzero(present)
                present[idx[over]] = 0
                payload = ("prim", payload[1][~over])
                node["present"] = present
            node["payload"] = payload
            node["rows"] = _to_rows(present, node["payload"], n)
        return node

    def decode_stripe_columns(self, si: int, columns=None, ts_unit: str = "ns"):
        """Whole-stripe decode: returns (rows, [node]) with one node per projected root column (see _decode_node)."""
        s = self.stripes[si]
        streams, encodings, tz = self._stripe_footer(s)
        smap = {(st.column, st.kind): st for st in streams}
        n = s.number_of_rows
        return n, [self._decode_node(smap, encodings, tz, cid, n, None, name, ts_unit) for name, cid in self._projected(columns)]

    # --- with_predicate (src/row_index.rs:204-331, src/arrow_reader.rs:256-293) ------------------------------------
    def stripe_row_index(self, si: int, columns=None):
        """{column id: [RowGroupEntry dict]} for the projected top-level columns that have a ROW_INDEX stream."""
        s = self.stripes[si]
        streams, _, _ = self._stripe_footer(s)
        by_key = {}
        for st in streams:
            by_key.setdefault((st.column, st.kind), st)  # the reference's HashMap keeps the last; one per key in practice
        out = {}
        for name, cid in self.columns:
            if columns is not None and name not in columns:
                continue
            st = by_key.get((cid, S_ROW_INDEX))
            if st is None:
                continue
            raw = bytes(decompress_stream(self.compression, self.data[st.offset:st.offset + st.length], self.block_size))
            entries = []
            for f, wt, v in pb_fields(raw):
                if f == 1:
                    stats = None
                    for g, _, w in pb_fields(v):
                        if g == 2:
                            stats = parse_column_statistics(w)
                    entries.append({"stats": stats, "bloom": None})
            out[cid] = entries
        for name, cid in self.columns:
            if cid not in out:
                continue
            st = by_key.get((cid, 7)) or by_key.get((cid, 8))  # BLOOM_FILTER, else BLOOM_FILTER_UTF8
            if st is None:
                continue
            raw = bytes(decompress_stream(self.compression, self.data[st.offset:st.offset + st.length], self.block_size))
            filters = [b for b in (parse_bloom_filter(v) for f, _, v in pb_fields(raw) if f == 1) if b is not None]
            if len(filters) != len(out[cid]):
                raise OracleError(13, "panic: Bloom filter count mismatch")
            for e, b in zip(out[cid], filters):
                e["bloom"] = b
        return out

    def predicate_selection(self, si: int, predicate, columns=None):
        """([skip, count] selectors, row-group verdicts or None): what try_advance_stripe derives for one stripe."""
        rows = self.stripes[si].number_of_rows
        stride = 10_000 if self.row_index_stride is None else self.row_index_stride
        try:
            index = self.stripe_row_index(si, columns)
            cols = [(n, c) for n, c in self.columns if columns is None or n in columns]
            groups = 0 if stride == 0 else -(-rows // stride)
            verdict = evaluate_predicate(predicate, index, cols, groups)
        except OracleError as e:
            if str(e.args[-1]).startswith("panic"):
                raise
            return [[False, rows]], None  # "Keep all rows (maybe)"
        return from_row_group_filter(verdict, stride, rows), verdict

    def read_stripe(self, si: int, batch_size: int = 8192, columns=None, ts_unit: str = "ns", views=None):
        """Batches of one stripe as pyarrow RecordBatches with the reference's physical layout."""
        import pyarrow as pa
        n, cols = self.decode_stripe_columns(si, columns, ts_unit)
        schema = self.schema(columns, ts_unit)
        if views is None:
            views = [(a, min(batch_size, n - a)) for a in range(0, n, batch_size)]
        batches = []
        for a, k in views:
            arrays = [_node_to_arrow(schema.field(node["name"]).type, node, a, a + k) for node in cols]
            if not arrays:
                batches.append(pa.RecordBatch.from_struct_array(pa.array([{}] * k, pa.struct([]))))
            else:
                batches.append(pa.RecordBatch.from_arrays(arrays, schema=schema))
        return batches

    def read(self, batch_size: int = 8192, columns=None, ts_unit: str = "ns", stripes=None, selection=None, predicate=None):
        """`selection`: [(skip, row_count), ...] as ArrowReaderBuilder::with_row_selection takes it; `predicate`: a
        nested tuple as `evaluate_predicate` below takes it (ArrowReaderBuilder::with_predicate)."""
        order = list(range(len(self.stripes)) if stripes is None else stripes)
        psel = None if predicate is None else [self.predicate_selection(si, predicate, columns)[0] for si in order]
        plan = None
        if selection is not None or psel is not None:
            plan = selection_views(selection, [self.stripes[si].number_of_rows for si in order], batch_size, psel)
        out = []
        for k, si in enumerate(order):
            views = None
            if plan is not None and plan[k] is not None:
                views = plan[k]
                if not views:
                    continue
            out.extend(self.read_stripe(si, batch_size, columns, ts_unit, views))
        return out


# ------------------------------------------------------------------------------------------------
# predicate pushdown: statistics, Bloom filters, row-group verdicts
# ------------------------------------------------------------------------------------------------
def _zz(v: int) -> int:
    return (v >> 1) ^ -(v & 1)


def _pb_str(v, wt) -> str:
    if wt != 2:
        raise OracleError(7, "wire type of a string field")
    try:
        return bytes(v).decode("utf-8")
    except UnicodeDecodeError:
        raise OracleError(7, "invalid string value: data is not UTF-8 encoded")


def parse_column_statistics(b: bytes):
    """TryFrom<&proto::ColumnStatistics> (src/statistics.rs:77-143): {"n", "has_null", "type": (variant, ...) | None}."""
    import struct
    n, has_null, sub = 0, False, {}
    for f, wt, v in pb_fields(b):
        if f == 1:
            n = v
        elif f == 10:
            has_null = bool(v)
        elif 2 <= f <= 9 or f == 12:
            if wt != 2:
                raise OracleError(7, "wire type of a statistics message")
            sub.setdefault(f, []).append(bytes(v))
    def fields(fno):
        d = {}
        for part in sub[fno]:  # prost merges repeated occurrences of a message field
            for g, gw, w in pb_fields(part):
                d.setdefault(g, []).append((gw, w))
        return d
    # strings are checked while decoding, whether or not the variant is used afterwards
    for fno, str_fields in ((4, (1, 2, 4, 5)), (6, (1, 2, 3))):
        if fno in sub:
            d = fields(fno)
            for g in str_fields:
                for gw, w in d.get(g, []):
                    _pb_str(w, gw)
    typ = None
    if n == 0:
        typ = None
    elif 2 in sub:
        d = fields(2)
        typ = ("int", _zz(d[1][-1][1]) if 1 in d else 0, _zz(d[2][-1][1]) if 2 in d else 0)
    elif 3 in sub:
        d = fields(3)
        dbl = lambda g: struct.unpack("<d", d[g][-1][1])[0] if g in d else 0.0
        typ = ("double", dbl(1), dbl(2))
    elif 4 in sub:
        d = fields(4)
        txt = lambda g: _pb_str(d[g][-1][1], d[g][-1][0]) if g in d else None
        lo, hi = txt(1), txt(2)
        typ = ("string", lo if lo is not None else (txt(4) or ""), hi if hi is not None else (txt(5) or ""),
               lo is not None, hi is not None)
    elif 5 in sub:
        d = fields(5)
        counts = []
        for gw, w in d.get(1, []):
            counts.extend(_packed_u(w, gw))
        if not counts:
            raise OracleError(13, "panic: index out of bounds (bucket statistics without a count)")
        typ = ("bucket", counts[0])
    elif 6 in sub:
        d = fields(6)
        txt = lambda g: _pb_str(d[g][-1][1], d[g][-1][0]) if g in d else ""
        typ = ("decimal", txt(1), txt(2))
    elif 7 in sub:
        d = fields(7)
        z32 = lambda g: int(np.int32(np.uint32(_zz(d[g][-1][1] & 0xFFFFFFFF) & 0xFFFFFFFF))) if g in d else 0
        typ = ("date", z32(1), z32(2))
    elif 8 in sub:
        typ = ("binary",)
    elif 9 in sub:
        d = fields(9)
        typ = ("timestamp", _zz(d[3][-1][1]) if 3 in d else 0, _zz(d[4][-1][1]) if 4 in d else 0)  # minimumUtc / maximumUtc
    elif 12 in sub:
        typ = ("collection",)
    return {"n": n, "has_null": has_null, "type": typ}


def parse_bloom_filter(b: bytes):
    """BloomFilter::try_from_proto (src/bloom_filter.rs:35-75): (k, [u64 words]) or None for an empty message."""
    k, words, utf8 = 0, [], None
    for f, wt, v in pb_fields(b):
        if f == 1:
            k = v
        elif f == 2:
            raw = bytes(v)
            if wt not in (1, 2) or len(raw) % 8:
                raise OracleError(7, "BloomFilter.bitset")
            words.extend(int.from_bytes(raw[i:i + 8], "little") for i in range(0, len(raw), 8))
        elif f == 3:
            utf8 = bytes(v)
    if words and utf8 is not None:
        raise OracleError(13, "panic: Bloom filter proto has both bitset and utf8bitset populated")
    if not words and utf8 is None:
        return None
    if not words:
        words = [int.from_bytes(utf8[i:i + 8], "little") for i in range(0, len(utf8), 8)]
    return (k if k else 3, words)


_M64 = (1 << 64) - 1


def _sar(x: int, n: int) -> int:
    """arithmetic shift right of a 64-bit pattern"""
    if x >> 63:
        x -= 1 << 64
    return (x >> n) & _M64


def bloom_hash_long(v: int) -> int:
    """BloomFilter::hash_long (src/bloom_filter.rs:136-149), Thomas Wang's mix on an i64"""
    key = v & _M64
    key = ((~key & _M64) + (key << 21)) & _M64
    key ^= _sar(key, 24)
    key = (key + (key << 3) + (key << 8)) & _M64
    key ^= _sar(key, 14)
    key = (key + (key << 2) + (key << 4)) & _M64
    key ^= _sar(key, 28)
    return (key + (key << 31)) & _M64


def bloom_hash_bytes(data: bytes) -> int:
    """murmur3_64_orc (src/bloom_filter.rs:182-242): Murmur3 x64, h1 only, seed 104729"""
    c1, c2 = 0x87C37B91114253D5, 0x4CF5AD432745937F
    rotl = lambda x, r: ((x << r) | (x >> (64 - r))) & _M64
    h1 = 104729
    nb = len(data) // 8
    for i in range(nb):
        k1 = int.from_bytes(data[8 * i:8 * i + 8], "little")
        k1 = rotl(k1 * c1 & _M64, 31) * c2 & _M64
        h1 = rotl(h1 ^ k1, 27)
        h1 = (h1 * 5 + 1390208809) & _M64
    tail = data[8 * nb:]
    if tail:
        k1 = int.from_bytes(tail, "little")
        k1 = rotl(k1 * c1 & _M64, 31) * c2 & _M64
        h1 ^= k1
    h1 ^= len(data)
    h1 ^= h1 >> 33
    h1 = h1 * 0xFF51AFD7ED558CCD & _M64
    h1 ^= h1 >> 33
    h1 = h1 * 0xC4CEB9FE1A85EC53 & _M64
    h1 ^= h1 >> 33
    return h1


def bloom_test_hash(bloom, h: int) -> bool:
    """BloomFilter::test_hash (src/bloom_filter.rs:109-133)"""
    k, words = bloom
    bits = len(words) * 64
    if bits == 0:
        return True
    as_i32 = lambda x: ((x & 0xFFFFFFFF) ^ 0x80000000) - 0x80000000
    h1, h2 = as_i32(h), as_i32(h >> 32)
    for i in range(1, k + 1):
        combined = as_i32(h1 + as_i32(i * h2))
        if combined < 0:
            combined = ~combined
        bit = combined % bits
        if not (words[bit // 64] >> (bit % 64)) & 1:
            return False
    return True


def bloom_add_hash(bloom, h: int):
    """BloomFilter::add_hash (src/bloom_filter.rs:85-107), for the tests that build filters"""
    k, words = bloom
    bits = len(words) * 64
    as_i32 = lambda x: ((x & 0xFFFFFFFF) ^ 0x80000000) - 0x80000000
    h1, h2 = as_i32(h), as_i32(h >> 32)
    for i in range(1, k + 1):
        combined = as_i32(h1 + as_i32(i * h2))
        if combined < 0:
            combined = ~combined
        bit = combined % bits
        words[bit // 64] |= 1 << (bit % 64)


_NEGATED = {"eq": "ne", "ne": "eq", "lt": "ge", "le": "gt", "gt": "le", "ge": "lt"}
_INT_TYPES = ("Int8", "Int16", "Int32", "Int64")


def _cmp_numbers(lo, hi, op, v):
    return {"eq": lo <= v <= hi, "ne": not (lo == v and hi == v), "lt": lo < v, "le": lo <= v, "gt": hi > v, "ge": hi >= v}[op]


def _cmp_floats(lo, hi, op, v):
    eps = 1e-9
    if op == "eq":
        return (lo - eps) <= v <= (hi + eps)
    if op == "ne":
        return not (abs(lo - v) < eps and abs(hi - v) < eps)
    return {"lt": lo < v, "le": lo <= v, "gt": hi > v, "ge": hi >= v}[op]


def compare_strings(lo: str, hi: str, op: str, v: str, exact_min: bool, exact_max: bool) -> bool:
    """evaluate_string_comparison (src/row_group_filter.rs:500-530); Rust orders str by bytes"""
    lo_b, hi_b, v_b = lo.encode(), hi.encode(), v.encode()
    min_le = lo_b < v_b or (lo_b == v_b and exact_min)
    max_ge = hi_b > v_b or (hi_b == v_b and exact_max)
    if op == "eq":
        return min_le and max_ge
    if op == "le":
        return min_le
    if op == "ge":
        return max_ge
    if op == "lt":
        return lo_b < v_b
    if op == "gt":
        return hi_b > v_b
    return not (lo_b == hi_b and exact_min and exact_max and lo_b == v_b)


def _stats_verdict(stats, op, value):
    """evaluate_comparison_with_stats (src/row_group_filter.rs:170-360); value = (type name, python value | None)"""
    vt, v = value
    t = stats["type"]
    if t is None:
        raise OracleError(13, "Statistics missing type-specific information")
    kind = t[0]
    def need(ok, what):
        if not ok or v is None:
            raise OracleError(13, "Type mismatch: expected " + what)
    if kind == "int":
        need(vt in _INT_TYPES, "integer value")
        return _cmp_numbers(t[1], t[2], op, v)
    if kind == "double":
        need(vt in ("Float32", "Float64"), "float value")
        return _cmp_floats(t[1], t[2], op, float(np.float32(v)) if vt == "Float32" else float(v))
    if kind == "string":
        need(vt == "Utf8", "string value")
        return compare_strings(t[1], t[2], op, v, t[3], t[4])
    if kind == "date":
        need(vt in ("Int32", "Int64"), "integer value for date")
        return _cmp_numbers(t[1], t[2], op, v)
    if kind == "timestamp":
        need(vt == "Int64", "integer value for timestamp")
        return _cmp_numbers(t[1], t[2], op, v)
    if kind == "decimal":
        need(vt == "Utf8", "string value for decimal")
        return compare_strings(t[1], t[2], op, v, True, True)
    if kind == "bucket":
        need(vt == "Boolean", "boolean value")
        trues, falses = t[1], (stats["n"] - t[1]) & _M64
        if op == "eq":
            return trues > 0 if v else falses > 0
        if op == "ne":
            return falses > 0 if v else trues > 0
        return True
    return True


def _bloom_verdict(entry, op, value):
    """row_group_might_match_bloom + bloom_value_hash64 (src/row_group_filter.rs:362-406)"""
    import struct
    vt, v = value
    if op != "eq" or entry["bloom"] is None or v is None:
        return True
    if vt == "Utf8":
        h = bloom_hash_bytes(v.encode())
    elif vt in ("Float32", "Float64"):
        d = float(np.float32(v)) if vt == "Float32" else float(v)
        h = bloom_hash_long(struct.unpack("<q", struct.pack("<d", d))[0])
    elif vt == "Boolean":
        h = bloom_hash_long(1 if v else 0)
    else:
        h = bloom_hash_long(int(v))
    return bloom_test_hash(entry["bloom"], h)


def evaluate_predicate(pred, index, cols, groups: int):
    """evaluate_predicate (src/row_group_filter.rs:24-114).  `pred`: ("cmp", column, op, (type, value)) with op in
    eq ne lt le gt ge, ("is_null", column), ("is_not_null", column), ("and", [..]), ("or", [..]), ("not", pred);
    `index`: stripe_row_index(); `cols`: projected (name, column id).  Raises what the reference returns as Err."""
    def entries_of(column):
        for name, cid in cols:
            if name == column:
                if cid not in index:
                    raise OracleError(13, f"Row index not found for column '{column}'")
                return index[cid]
        raise OracleError(13, f"Column '{column}' not found in schema")

    def compare(column, op, value, result):
        for g, e in enumerate(entries_of(column)[:len(result)]):
            if e["stats"] is not None and not _stats_verdict(e["stats"], op, value):
                result[g] = False
            else:
                result[g] = _bloom_verdict(e, op, value)

    def null_test(column, want_null, result):
        for g, e in enumerate(entries_of(column)[:len(result)]):
            if e["stats"] is None:
                result[g] = True
            else:
                result[g] = e["stats"]["has_null"] if want_null else e["stats"]["n"] > 0

    def rec(p, result):
        tag = p[0]
        if tag == "cmp":
            compare(p[1], p[2], p[3], result)
        elif tag == "is_null":
            null_test(p[1], True, result)
        elif tag == "is_not_null":
            null_test(p[1], False, result)
        elif tag == "and":
            for c in p[1]:
                tmp = [True] * len(result)
                rec(c, tmp)
                for g in range(len(result)):
                    result[g] = result[g] and tmp[g]
        elif tag == "or":
            tmps = []
            for c in p[1]:
                tmp = [True] * len(result)
                rec(c, tmp)
                tmps.append(tmp)
            for g in range(len(result)):
                result[g] = any(t[g] for t in tmps)
        elif tag == "not":
            q = p[1]
            if q[0] == "not":
                rec(q[1], result)
            elif q[0] == "is_null":
                null_test(q[1], False, result)
            elif q[0] == "is_not_null":
                null_test(q[1], True, result)
            elif q[0] == "cmp":
                compare(q[1], _NEGATED[q[2]], q[3], result)
            elif q[0] == "and":
                rec(("or", [("not", c) for c in q[1]]), result)
            else:
                rec(("and", [("not", c) for c in q[1]]), result)
        else:
            raise ValueError(tag)

    result = [True] * groups
    rec(pred, result)
    return result


def from_row_group_filter(verdict, stride: int, rows: int):
    """RowSelection::from_row_group_filter (src/row_selection.rs:348-390) as [skip, count] lists"""
    if not verdict:
        return [[True, rows]]
    sel = []
    for keep in verdict:
        if sel and sel[-1][0] == (not keep):
            sel[-1][1] += stride
        else:
            sel.append([not keep, stride])
    covered = len(verdict) * stride
    if covered < rows:
        if sel[-1][0]:
            sel[-1][1] += rows - covered
        else:
            sel.append([True, rows - covered])
    return sel


def selection_and_then(first, second):
    """RowSelection::and_then (src/row_selection.rs:401-463) on [skip, count] lists: `second` picks among the rows
    `first` selects.  The reference panics when the two do not fit; that is OracleError(13, "panic: ...") here."""
    first = [list(x) for x in first]
    second = [list(x) for x in second]
    out, to_skip, a, b = [], 0, 0, 0
    while b < len(second):
        if a >= len(first):
            raise OracleError(13, "panic: selection exceeds the number of selected rows")
        if second[b][1] == 0:
            b += 1
            continue
        if first[a][1] == 0:
            a += 1
            continue
        if first[a][0]:
            to_skip += first[a][1]
            a += 1
            continue
        k = min(first[a][1], second[b][1])
        first[a][1] -= k
        second[b][1] -= k
        if second[b][0]:
            to_skip += k
        else:
            if to_skip:
                out.append([True, to_skip])
                to_skip = 0
            out.append([False, k])
    for sk, c in first[a:]:
        if c:
            if not sk:
                raise OracleError(13, "panic: selection contains less than the number of selected rows")
            to_skip += c
    if to_skip:
        out.append([True, to_skip])
    return out


def selection_views(selection, stripe_rows, batch_size, predicate_selections=None):
    """Row ranges each stripe yields under a row selection (None = the stripe is read whole).  Follows
    RowSelection::from(Vec<RowSelector>) (src/row_selection.rs:466-482), ArrowReader::try_advance_stripe
    (src/arrow_reader.rs:296-309: `split_off(stripe_rows)` while the selection still has rows, no selection at all
    afterwards) and NaiveStripeDecoder::next_with_row_selection (src/array_decoder/mod.rs:313-364), including its
    habit of staying on a selector until one step has covered the selector's whole row_count."""
    sel = []
    for skip, count in (selection or []):
        if count == 0:
            continue
        if sel and sel[-1][0] == bool(skip):
            sel[-1][1] += count
        else:
            sel.append([bool(skip), count])
    out = []
    for k_stripe, rows in enumerate(stripe_rows):
        # try_advance_stripe (src/arrow_reader.rs:256-309): the predicate's selection first, then the caller's while
        # it still has rows, combined as caller.and_then(predicate's)
        head = None if predicate_selections is None else [list(x) for x in predicate_selections[k_stripe]]
        if selection is not None and sum(c for _, c in sel) > 0:
            # split_off(rows)
            mine, acc, idx = [], 0, None
            for i, (sk, c) in enumerate(sel):
                acc += c
                if acc > rows:
                    idx = i
                    break
            if idx is None:
                mine, sel = sel, []
            else:
                mine, rest = [list(x) for x in sel[:idx]], [list(x) for x in sel[idx:]]
                overflow = acc - rows
                if rest[0][1] != overflow:
                    mine.append([rest[0][0], rest[0][1] - overflow])
                rest[0][1] = overflow
                sel = rest
            head = mine if head is None else selection_and_then(mine, head)
        if head is None:
            out.append(None)
            continue
        # next_with_row_selection
        views, index, si = [], 0, 0
        while index < rows and si < len(head):
            skip, count = head[si]
            remaining = rows - index
            k = min(count, remaining) if skip else min(count, batch_size, remaining)
            if k == 0:
                si += 1
                continue
            if not skip:
                views.append((index, k))
            index += k
            if k >= count:
                si += 1
        out.append(views)
    return out


# old link names ("backward" file of the tz database) that minimal tzdata installs leave out; chrono-tz knows them
_ZONE_LINKS = {"US/Pacific": "America/Los_Angeles", "US/Eastern": "America/New_York", "US/Central": "America/Chicago",
               "US/Mountain": "America/Denver", "US/Alaska": "America/Anchorage", "US/Hawaii": "Pacific/Honolulu",
               "US/Arizona": "America/Phoenix", "Asia/Calcutta": "Asia/Kolkata", "Japan": "Asia/Tokyo", "PRC": "Asia/Shanghai",
               "GB": "Europe/London", "Eire": "Europe/Dublin", "NZ": "Pacific/Auckland", "Singapore": "Asia/Singapore"}


def _unit_of(ts_unit, name: str) -> str:
    """`ts_unit`: one unit for every timestamp column, or {column: unit}; "dec" = Decimal128(38, 9) (with_schema)."""
    if isinstance(ts_unit, dict):
        return ts_unit.get(name, "ns")
    return ts_unit


def _zone(name: str):
    import zoneinfo
    try:
        return zoneinfo.ZoneInfo(name)
    except zoneinfo.ZoneInfoNotFoundError:
        if name in _ZONE_LINKS:
            return zoneinfo.ZoneInfo(_ZONE_LINKS[name])
        raise


def _tz_to_utc(vals: np.ndarray, zone, unit_ns: int) -> np.ndarray:
    """TimestampOffsetArrayDecoder (array_decoder/timestamp.rs:242-286): wall clock of the instant in the
    writer zone, re-read as UTC = instant + utcoffset(instant)."""
    per_s = 1_000_000_000 // unit_ns
    epoch = _dt.datetime(1970, 1, 1, tzinfo=_dt.timezone.utc)
    secs = np.floor_divide(vals, per_s)
    uniq, inv = np.unique(secs, return_inverse=True)  # one zone lookup per distinct second
    offs = np.empty(uniq.size, dtype=np.int64)
    for i, sv in enumerate(uniq.tolist()):
        offs[i] = int((epoch + _dt.timedelta(seconds=sv)).astimezone(zone).utcoffset().total_seconds())
    with np.errstate(over="ignore"):
        out = vals + offs[inv] * per_s  # wraps where the sum leaves i64: those values become nulls (caller)
    over = np.zeros(vals.size, dtype=bool)
    sus = np.flatnonzero((np.abs(vals.astype(np.float64)) > 9.2e18))  # only values near the ends of i64 can overflow
    for i in sus.tolist():
        w = int(vals[i]) + int(offs[inv[i]]) * per_s
        over[i] = not (-(1 << 63) <= w < (1 << 63))
    out[over] = 0
    return out, over


def _to_rows(present, payload, n):
    """decode_spaced (encoding/mod.rs:64-91): place dense values at valid rows, zeros elsewhere."""
    kind = payload[0]
    if kind in ("prim", "bool"):
        v = payload[1]
        if present is None:
            return v
        out = np.zeros((n,) + v.shape[1:], dtype=v.dtype)
        out[present.astype(bool)] = v
        return out
    if kind == "dict":
        keys = payload[1]
        if present is None:
            return keys
        out = np.zeros(n, dtype=np.int64)
        out[present.astype(bool)] = keys
        return out
    if kind == "bytes":
        lens = payload[1]
        if present is None:
            return lens
        out = np.zeros(n, dtype=np.int64)
        out[present.astype(bool)] = lens
        return out
    raise AssertionError(kind)


def _validity(present, a, b):
    """derive_present_vec (array_decoder/mod.rs:231-252): None when the batch has no nulls."""
    import pyarrow as pa
    if present is None:
        return None, 0
    p = present[a:b]
    nulls = int((p == 0).sum())
    if nulls == 0:
        return None, 0
    return pa.py_buffer(np.packbits(p, bitorder="little").tobytes()), nulls


def _slice_to_arrow(typ, present, payload, rows, a, b):
    import pyarrow as pa
    n = b - a
    vbuf, nulls = _validity(present, a, b)
    kind = payload[0]
    if kind == "prim":
        data = np.ascontiguousarray(rows[a:b])
        return pa.Array.from_buffers(typ, n, [vbuf, pa.py_buffer(data.tobytes())], null_count=nulls)
    if kind == "bool":
        bits = np.packbits(rows[a:b], bitorder="little")
        return pa.Array.from_buffers(typ, n, [vbuf, pa.py_buffer(bits.tobytes())], null_count=nulls)
    if kind == "bytes":
        lens = np.ascontiguousarray(rows[a:b])
        offs = np.empty(n + 1, dtype=np.int32)
        _check(lib().orc_oracle_offsets(_u8p(lens), ctypes.c_size_t(n), _u8p(offs)), "offsets")
        start = int(rows[:a].sum())
        total = int(offs[-1])
        data = payload[2][start:start + total]
        if data.size < total:
            raise OracleError(18, "string data shorter than offsets")  # GenericByteArray::try_new
        if pa.types.is_string(typ):
            _validate_utf8(data, offs)
        return pa.Array.from_buffers(typ, n, [vbuf, pa.py_buffer(offs.tobytes()), pa.py_buffer(data.tobytes())],
                                     null_count=nulls)
    if kind == "dict":
        keys = np.ascontiguousarray(rows[a:b])
        doff, ddata = payload[2], payload[3]
        pb = None
        if vbuf is not None:
            pb = np.ascontiguousarray(present[a:b])
        offs = np.empty(n + 1, dtype=np.int32)
        total = ctypes.c_int64()
        _check(lib().orc_oracle_dict_gather(_u8p(keys), _u8p(pb) if pb is not None else None, ctypes.c_size_t(n),
                                            _u8p(doff), ctypes.c_size_t(doff.size - 1), _u8p(ddata), _u8p(offs), None,
                                            ctypes.byref(total)), "dict gather")
        data = np.empty(max(total.value, 1), dtype=np.uint8)
        _check(lib().orc_oracle_dict_gather(_u8p(keys), _u8p(pb) if pb is not None else None, ctypes.c_size_t(n),
                                            _u8p(doff), ctypes.c_size_t(doff.size - 1), _u8p(ddata), _u8p(offs),
                                            _u8p(data), ctypes.byref(total)), "dict gather")
        return pa.Array.from_buffers(typ, n, [vbuf, pa.py_buffer(offs.tobytes()),
                                              pa.py_buffer(data[: total.value].tobytes())], null_count=nulls)
    raise AssertionError(kind)


def _node_to_arrow(typ, node, a, b):
    """Slots [a, b) of a decoded column as the Arrow array the reference builds for that batch."""
    import pyarrow as pa
    k = node["kind"]
    n = b - a
    if k not in (K_STRUCT, K_LIST, K_MAP, K_UNION):
        return _slice_to_arrow(typ, node["present"], node["payload"], node["rows"], a, b)
    vbuf, nulls = _validity(node["present"], a, b)
    if k == K_STRUCT:
        kids = [_node_to_arrow(typ.field(i).type, c, a, b) for i, c in enumerate(node["children"])]
        return pa.Array.from_buffers(typ, n, [vbuf], null_count=nulls, children=kids)
    if k in (K_LIST, K_MAP):
        lens = np.ascontiguousarray(node["lens"][a:b])
        offs = np.empty(n + 1, dtype=np.int32)
        _check(lib().orc_oracle_offsets(_u8p(lens), ctypes.c_size_t(n), _u8p(offs)), "offsets")
        start = int(node["lens"][:a].sum())
        end = start + int(offs[-1])
        if k == K_LIST:
            child = _node_to_arrow(typ.value_type, node["children"][0], start, end)
            return pa.Array.from_buffers(typ, n, [vbuf, pa.py_buffer(offs.tobytes())], null_count=nulls, children=[child])
        keys = _node_to_arrow(typ.key_type, node["children"][0], start, end)
        items = _node_to_arrow(typ.item_type, node["children"][1], start, end)
        entries = pa.Array.from_buffers(pa.struct([typ.key_field, typ.item_field]), end - start, [None], null_count=0,
                                        children=[keys, items])
        return pa.Array.from_buffers(typ, n, [vbuf, pa.py_buffer(offs.tobytes())], null_count=nulls, children=[entries])
    # sparse union: type ids + one child per variant, each over every slot (union.rs:118-124)
    tags = np.ascontiguousarray(node["tags"][a:b]).astype(np.int8)
    kids = [_node_to_arrow(typ.field(i).type, c, a, b) for i, c in enumerate(node["children"])]
    return pa.UnionArray.from_sparse(pa.array(tags, pa.int8()), kids, [typ.field(i).name for i in range(typ.num_fields)],
                                     list(range(typ.num_fields)))


def _validate_utf8(data: np.ndarray, offs: np.ndarray):
    """GenericByteArray::<Utf8>::try_new (string.rs:150-151): the bytes the offsets cover are one valid UTF-8
    string and every offset falls on a character boundary."""
    used = data[:int(offs[-1])]
    try:
        used.tobytes().decode("utf-8")
    except UnicodeDecodeError as e:
        raise OracleError(18, f"invalid utf-8: {e}")
    inner = np.asarray(offs[:-1], dtype=np.int64)
    inner = inner[inner < used.size]
    if inner.size and np.any((used[inner] & 0xC0) == 0x80):
        raise OracleError(18, "invalid utf-8: value starts inside a character")
