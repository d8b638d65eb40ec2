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

Parity pinning: tests/test_oracle_kat.py (reference unit-test vectors) and tests/test_oracle_files.py
(reference expected_arrow feather goldens + pyarrow.orc as an independent reader).
Only flat (non-nested) root columns are restated so far; nested columns raise NotImplementedError.
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
    p = 0
    n = len(b)
    while p < n:
        key, p = _pb_varint(b, p)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            v, p = _pb_varint(b, p)
        elif wt == 1:
            v = b[p:p + 8]
            p += 8
        elif wt == 2:
            ln, p = _pb_varint(b, p)
            v = b[p:p + ln]
            if len(v) != ln:
                raise OracleError(7, "truncated bytes field")
            p += ln
        elif wt == 5:
            v = b[p:p + 4]
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
                        t.field_names.append(w.decode("utf-8"))
                    elif g == 4:
                        t.maximum_length = w
                    elif g == 5:
                        t.precision = w
                    elif g == 6:
                        t.scale = w
                self.types.append(t)
            elif f == 5:
                name, val = "", b""
                for g, _, w in pb_fields(v):
                    if g == 1:
                        name = w.decode("utf-8")
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
        self.columns: List[Tuple[str, int]] = list(zip(root.field_names, root.subtypes))

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
            return pa.decimal128(t.precision, t.scale)
        if k in (K_TIMESTAMP, K_TIMESTAMP_INSTANT) and ts_unit == "dec":
            return pa.decimal128(38, 9)  # with_schema: Decimal128(38, 9) nanoseconds (array_decoder/timestamp.rs:150-232)
        if k == K_TIMESTAMP:
            return pa.timestamp(ts_unit)
        if k == K_TIMESTAMP_INSTANT:
            return pa.timestamp(ts_unit, tz="UTC")
        raise NotImplementedError(f"nested ORC type kind {k} not restated in the oracle yet")

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
        for f, _, v in pb_fields(b):
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
                tz = v.decode("utf-8")
        return streams, encodings, tz

    def _stream(self, smap, col, kind) -> np.ndarray:
        """StreamMap::get (stripe.rs:319-335): missing stream decodes as empty."""
        st = smap.get((col, kind))
        if st is None:
            return np.empty(0, dtype=np.uint8)
        raw = np.frombuffer(self.data, dtype=np.uint8, count=st.length, offset=st.offset)
        return decompress_stream(self.compression, raw, self.block_size)

    def decode_stripe_columns(self, si: int, columns=None, ts_unit: str = "ns"):
        """Whole-stripe decode: returns list of (name, col_id, present_bools|None, payload)."""
        s = self.stripes[si]
        streams, encodings, tz = self._stripe_footer(s)
        smap = {(st.column, st.kind): st for st in streams}
        n = s.number_of_rows
        out = []
        for name, cid in self._projected(columns):
            t = self.types[cid]
            ek, dict_size = encodings[cid] if cid < len(encodings) else (0, 0)
            ver = 2 if ek in (2, 3) else 1  # column.rs:52-59
            present = None
            if (cid, S_PRESENT) in smap:
                present = bool_rle(self._stream(smap, cid, S_PRESENT), n)
            nn = int(present.sum()) if present is not None else n
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
                        o = _tz_to_utc(o, zone, unit_ns)
                payload = ("prim", o)
            else:
                raise NotImplementedError(f"nested ORC type kind {k} not restated in the oracle yet")
            out.append((name, cid, present, payload))
        return n, out

    def read_stripe(self, si: int, batch_size: int = 8192, columns=None, ts_unit: str = "ns", views=None):
        """Batches of one stripe as pyarrow RecordBatches with the reference's physical layout."""
        import pyarrow as pa
        n, cols = self.decode_stripe_columns(si, columns, ts_unit)
        schema = self.schema(columns, ts_unit)
        # spaced (row-domain) representation per column
        spaced = []
        for name, cid, present, payload in cols:
            spaced.append(_to_rows(present, payload, n))
        if views is None:
            views = [(a, min(batch_size, n - a)) for a in range(0, n, batch_size)]
        batches = []
        for a, k in views:
            arrays = []
            for (name, cid, present, payload), rows in zip(cols, spaced):
                arrays.append(_slice_to_arrow(schema.field(name).type, present, payload, rows, a, a + k))
            if not arrays:
                batches.append(pa.RecordBatch.from_struct_array(pa.array([{}] * k, pa.struct([]))))
            else:
                batches.append(pa.RecordBatch.from_arrays(arrays, schema=schema))
        return batches

    def read(self, batch_size: int = 8192, columns=None, ts_unit: str = "ns", stripes=None, selection=None):
        """`selection`: [(skip, row_count), ...] as ArrowReaderBuilder::with_row_selection takes it."""
        order = list(range(len(self.stripes)) if stripes is None else stripes)
        plan = None if selection is None else selection_views(selection, [self.stripes[si].number_of_rows for si in order], batch_size)
        out = []
        for k, si in enumerate(order):
            views = None
            if plan is not None and plan[k] is not None:
                views = plan[k]
                if not views:
                    continue
            out.extend(self.read_stripe(si, batch_size, columns, ts_unit, views))
        return out


def selection_views(selection, stripe_rows, batch_size):
    """Row ranges each stripe yields under a row selection (None = the stripe is read whole).  Follows
    RowSelection::from(Vec<RowSelector>) (src/row_selection.rs:466-482), ArrowReader::try_advance_stripe
    (src/arrow_reader.rs:296-309: `split_off(stripe_rows)` while the selection still has rows, no selection at all
    afterwards) and NaiveStripeDecoder::next_with_row_selection (src/array_decoder/mod.rs:313-364), including its
    habit of staying on a selector until one step has covered the selector's whole row_count."""
    sel = []
    for skip, count in selection:
        if count == 0:
            continue
        if sel and sel[-1][0] == bool(skip):
            sel[-1][1] += count
        else:
            sel.append([bool(skip), count])
    out = []
    for rows in stripe_rows:
        if sum(c for _, c in sel) == 0:
            out.append(None)
            continue
        # split_off(rows)
        head, acc, idx = [], 0, None
        for i, (sk, c) in enumerate(sel):
            acc += c
            if acc > rows:
                idx = i
                break
        if idx is None:
            head, sel = sel, []
        else:
            head, rest = [list(x) for x in sel[:idx]], [list(x) for x in sel[idx:]]
            overflow = acc - rows
            if rest[0][1] != overflow:
                head.append([rest[0][0], rest[0][1] - overflow])
            rest[0][1] = overflow
            sel = rest
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
    return vals + offs[inv] * per_s


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
