"""orc_rust_b200 — B200-native ORC stripe decoder behind orc-rust's reader API.

Host-side mirror of the reference's public interface for the decode path
(`ArrowReaderBuilder` / `ArrowReader`, src/arrow_reader.rs:39-347 of datafusion-contrib/orc-rust),
bound over the C ABI in include/orc_b200.h.  Every decode goes through the CUDA library
(`liborc_b200.so`, built in-tree by `orc_rust_b200.build`); there is no CPU fallback: if the
library or a CUDA device is missing, calls fail loudly.
"""
from __future__ import annotations

import ctypes
import os
from typing import Iterator, List, Optional, Sequence

from . import build as _build

__all__ = ["ArrowReaderBuilder", "ArrowReader", "DecodeJob", "ChunkReader", "FileChunkReader", "OrcError", "lib", "device_available",
           "TimestampPrecision", "decode_int_rle", "decode_byte_rle", "decode_bool_rle", "decode_varint128",
           "decompress_stream"]

STATUS_NAMES = {
    1: "IoError", 2: "EmptyFile", 3: "OutOfSpec", 4: "DecodeFloat", 5: "DecodeTimestamp", 6: "OffsetOverflow",
    7: "DecodeProto", 8: "NoTypes", 9: "UnsupportedTypeVariant", 10: "MismatchedSchema", 11: "ConvertRecordBatch",
    12: "VarintTooLarge", 13: "Unexpected", 14: "BuildZstdDecoder", 15: "BuildSnappyDecoder", 16: "BuildLzoDecoder",
    17: "BuildLz4Decoder", 18: "Arrow", 19: "UnsupportedDeviceCodec", 20: "Cuda", 21: "InvalidArgument",
    22: "NotImplemented", 23: "DeviceHeapOverflow",
}


class OrcError(Exception):
    """Mirror of `OrcError` (src/error.rs:31-174) + the device-path variants."""

    def __init__(self, code: int, msg: str):
        self.code = code
        self.variant = STATUS_NAMES.get(code, f"status{code}")
        super().__init__(f"{self.variant}: {msg}")


class TimestampPrecision:
    """src/schema.rs:31-39"""
    Nanosecond = 0
    Microsecond = 1
    Millisecond = 2  # reachable in the reference through with_schema
    Second = 3


class _ReadOptions(ctypes.Structure):
    _fields_ = [
        ("device", ctypes.c_int32), ("batch_size", ctypes.c_uint32),
        ("projection_names", ctypes.POINTER(ctypes.c_char_p)), ("n_projection", ctypes.c_uint32),
        ("range_start", ctypes.c_uint64), ("range_end", ctypes.c_uint64),
        ("timestamp_unit", ctypes.c_int32), ("use_row_index", ctypes.c_int32), ("device_resident", ctypes.c_int32),
        ("max_stripes_per_launch", ctypes.c_uint32), ("cuda_stream", ctypes.c_void_p), ("flags", ctypes.c_uint32),
        ("stripe_shard_index", ctypes.c_uint32), ("stripe_shard_count", ctypes.c_uint32),
        ("waves", ctypes.c_uint32), ("reserved0", ctypes.c_uint32),
    ]


class JobStats(ctypes.Structure):
    _fields_ = [(n, ctypes.c_uint64) for n in (
        "n_stripes", "n_rows", "n_columns", "input_bytes", "staged_bytes", "output_bytes", "device_bytes",
        "n_segments", "n_kernel_launches", "n_batches", "d2h_meta_bytes", "aliased_output_bytes", "n_waves")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class KernelStat(ctypes.Structure):
    _fields_ = [("name", ctypes.c_char * 32), ("ms", ctypes.c_double), ("alg_bytes", ctypes.c_uint64),
                ("work_items", ctypes.c_uint64)]


class _ArrowArray(ctypes.Structure):
    pass


_ArrowArray._fields_ = [
    ("length", ctypes.c_int64), ("null_count", ctypes.c_int64), ("offset", ctypes.c_int64),
    ("n_buffers", ctypes.c_int64), ("n_children", ctypes.c_int64), ("buffers", ctypes.POINTER(ctypes.c_void_p)),
    ("children", ctypes.POINTER(ctypes.POINTER(_ArrowArray))), ("dictionary", ctypes.POINTER(_ArrowArray)),
    ("release", ctypes.c_void_p), ("private_data", ctypes.c_void_p),
]


class _ArrowSchema(ctypes.Structure):
    pass


_ArrowSchema._fields_ = [
    ("format", ctypes.c_char_p), ("name", ctypes.c_char_p), ("metadata", ctypes.c_void_p),
    ("flags", ctypes.c_int64), ("n_children", ctypes.c_int64),
    ("children", ctypes.POINTER(ctypes.POINTER(_ArrowSchema))), ("dictionary", ctypes.POINTER(_ArrowSchema)),
    ("release", ctypes.c_void_p), ("private_data", ctypes.c_void_p),
]


class _ArrowDeviceArray(ctypes.Structure):
    _fields_ = [("array", _ArrowArray), ("device_id", ctypes.c_int64), ("device_type", ctypes.c_int32),
                ("sync_event", ctypes.c_void_p), ("reserved", ctypes.c_int64 * 3)]


_lib = None
BATCH_DONE = ctypes.CFUNCTYPE(None, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_char_p)
READ_AT = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint64, ctypes.POINTER(ctypes.c_uint8))

EXPORTED_SYMBOLS = [
    "orcb_open_memory", "orcb_open_path", "orcb_open_callbacks", "orcb_file_io_stats", "orcb_file_clone", "orcb_file_free", "orcb_file_num_rows", "orcb_file_num_stripes",
    "orcb_file_compression", "orcb_file_compression_block_size", "orcb_file_row_index_stride",
    "orcb_file_num_root_columns", "orcb_file_root_column_name", "orcb_file_root_column_id", "orcb_file_stripe_info", "orcb_schema",
    "orcb_reader_new", "orcb_reader_new_with_selection", "orcb_reader_new_ex", "orcb_reader_build", "orcb_reader_plan", "orcb_predicate_row_groups", "orcb_bloom_hash_long", "orcb_bloom_hash_bytes", "orcb_reader_counters", "orcb_selection_plan", "orcb_reader_free", "orcb_reader_total_row_count", "orcb_reader_next",
    "orcb_reader_next_device", "orcb_reader_next_async", "orcb_reader_drain", "orcb_job_new", "orcb_job_free", "orcb_job_plan", "orcb_job_stage",
    "orcb_job_launch", "orcb_job_finish", "orcb_job_stats", "orcb_job_restage", "orcb_job_kernel_stats", "orcb_job_num_batches", "orcb_job_export_batch",
    "orcb_job_export_batch_device", "orcb_decode_int_rle", "orcb_decode_byte_rle", "orcb_decode_bool_rle",
    "orcb_decode_varint128", "orcb_decompress_stream", "orcb_host_decompress_section", "orcb_last_error", "orcb_index_retries", "orcb_layout_retries", "orcb_build_info",
    "orcb_device_available", "orcb_zone_table", "orcb_file_format_version", "orcb_file_num_user_metadata", "orcb_file_user_metadata",
]


def lib() -> ctypes.CDLL:
    """Loads (building if stale) the CUDA library.  Raises if it cannot be built or loaded."""
    global _lib
    if _lib is None:
        path = _build.build()
        L = ctypes.CDLL(path)
        L.orcb_last_error.restype = ctypes.c_char_p
        L.orcb_build_info.restype = ctypes.c_char_p
        L.orcb_index_retries.restype = ctypes.c_uint64
        L.orcb_layout_retries.restype = ctypes.c_uint64
        L.orcb_file_num_rows.restype = ctypes.c_uint64
        L.orcb_file_compression_block_size.restype = ctypes.c_uint64
        L.orcb_file_row_index_stride.restype = ctypes.c_int64
        L.orcb_file_root_column_name.restype = ctypes.c_char_p
        L.orcb_reader_total_row_count.restype = ctypes.c_uint64
        L.orcb_job_num_batches.restype = ctypes.c_uint64
        for name in ("orcb_file_num_rows", "orcb_file_num_stripes", "orcb_file_compression",
                     "orcb_file_compression_block_size", "orcb_file_row_index_stride",
                     "orcb_file_num_root_columns", "orcb_file_free", "orcb_reader_free", "orcb_job_free",
                     "orcb_reader_total_row_count", "orcb_job_num_batches"):
            getattr(L, name).argtypes = [ctypes.c_void_p]
        L.orcb_file_root_column_name.argtypes = [ctypes.c_void_p, ctypes.c_uint32]
        L.orcb_file_root_column_id.argtypes = [ctypes.c_void_p, ctypes.c_uint32]
        L.orcb_file_root_column_id.restype = ctypes.c_uint32
        L.orcb_selection_plan.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint64,
                                          ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
        L.orcb_reader_plan.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_size_t, ctypes.c_void_p]
        L.orcb_reader_build.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.orcb_predicate_row_groups.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32,
                                                ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p]
        L.orcb_bloom_hash_long.argtypes = [ctypes.c_int64]
        L.orcb_bloom_hash_long.restype = ctypes.c_uint64
        L.orcb_bloom_hash_bytes.argtypes = [ctypes.c_char_p, ctypes.c_size_t]
        L.orcb_bloom_hash_bytes.restype = ctypes.c_uint64
        L.orcb_reader_new_ex.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_int,
                                         ctypes.c_void_p, ctypes.c_void_p]
        L.orcb_reader_new_with_selection.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32,
                                                     ctypes.c_void_p]
        L.orcb_file_stripe_info.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.POINTER(ctypes.c_uint64)]
        L.orcb_open_memory.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_void_p)]
        L.orcb_open_path.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p)]
        L.orcb_file_clone.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p)]
        L.orcb_open_callbacks.argtypes = [ctypes.c_uint64, READ_AT, ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p)]
        L.orcb_file_io_stats.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint64)]
        L.orcb_schema.argtypes = [ctypes.c_void_p, ctypes.POINTER(_ReadOptions), ctypes.c_void_p]
        L.orcb_reader_new.argtypes = [ctypes.c_void_p, ctypes.POINTER(_ReadOptions), ctypes.POINTER(ctypes.c_void_p)]
        L.orcb_reader_next.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_int)]
        L.orcb_reader_next_device.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_int)]
        L.orcb_reader_drain.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint64)]
        L.orcb_reader_next_async.argtypes = [ctypes.c_void_p, ctypes.c_void_p, BATCH_DONE, ctypes.c_void_p]
        L.orcb_job_new.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_uint32, ctypes.POINTER(_ReadOptions),
                                   ctypes.POINTER(ctypes.c_void_p)]
        L.orcb_job_kernel_stats.argtypes = [ctypes.c_void_p, ctypes.POINTER(KernelStat), ctypes.c_uint32,
                                            ctypes.POINTER(ctypes.c_uint32)]
        for name in ("orcb_job_plan", "orcb_job_stage", "orcb_job_launch", "orcb_job_finish", "orcb_job_restage"):
            getattr(L, name).argtypes = [ctypes.c_void_p]
        L.orcb_job_stats.argtypes = [ctypes.c_void_p, ctypes.POINTER(JobStats)]
        L.orcb_job_export_batch.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p]
        L.orcb_job_export_batch_device.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p]
        L.orcb_decode_int_rle.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int,
                                          ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t]
        L.orcb_decode_byte_rle.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t]
        L.orcb_decode_bool_rle.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t]
        L.orcb_decode_varint128.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t]
        L.orcb_decompress_stream.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t,
                                             ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t)]
        _lib = L
    return _lib


def _check(rc: int):
    if rc != 0:
        raise OrcError(rc, lib().orcb_last_error().decode("utf-8", "replace"))


def index_retries() -> int:
    """Jobs decoded a second time without the row index (positions that did not agree with the streams)."""
    return int(lib().orcb_index_retries())


def layout_retries() -> int:
    """Jobs decoded a second time because a compressed chunk in the middle of a stream did not fill its block."""
    return int(lib().orcb_layout_retries())


def device_available() -> bool:
    return bool(lib().orcb_device_available())


class ChunkReader:
    """Mirror of the reference's `ChunkReader` trait (src/reader/mod.rs:27-46): subclass with `len()` and
    `get_bytes(offset_from_start, length) -> bytes`.  `FileChunkReader` serves a local file."""

    def len(self) -> int:
        raise NotImplementedError

    def get_bytes(self, offset_from_start: int, length: int) -> bytes:
        raise NotImplementedError


class FileChunkReader(ChunkReader):
    def __init__(self, path):
        self._f = open(path, "rb")
        self._n = os.fstat(self._f.fileno()).st_size
        self.calls = []

    def len(self):
        return self._n

    def get_bytes(self, offset_from_start, length):
        self.calls.append((offset_from_start, length))
        return os.pread(self._f.fileno(), length, offset_from_start)


class _File:
    """FileMetadata handle (src/reader/metadata.rs:63-178)."""

    def __init__(self, source):
        self._h = ctypes.c_void_p()
        self._keep = None
        if isinstance(source, ChunkReader):
            # ChunkReader (src/reader/mod.rs:27-46): len() + get_bytes(offset, length) behind a C callback
            def read_at(ctx, offset, length, dst, _src=source):
                try:
                    b = _src.get_bytes(int(offset), int(length))
                    if len(b) != length:
                        return 2
                    ctypes.memmove(dst, bytes(b), length)
                    return 0
                except Exception:
                    return 1
            cb = READ_AT(read_at)
            self._keep = (cb, source)
            _check(lib().orcb_open_callbacks(source.len(), cb, None, ctypes.byref(self._h)))
        elif isinstance(source, (str, os.PathLike)):
            _check(lib().orcb_open_path(os.fspath(source).encode(), ctypes.byref(self._h)))
        else:
            data = bytes(source) if not isinstance(source, bytes) else source
            self._keep = data  # borrowed by the library for the life of the handle
            buf = ctypes.cast(ctypes.c_char_p(data), ctypes.c_void_p)
            _check(lib().orcb_open_memory(buf, len(data), ctypes.byref(self._h)))

    def __del__(self):
        if getattr(self, "_h", None) and self._h.value and _lib is not None:
            _lib.orcb_file_free(self._h)
            self._h = ctypes.c_void_p()

    def io_stats(self) -> dict:
        """read_at calls and bytes so far (files opened through a ChunkReader)."""
        out = (ctypes.c_uint64 * 2)()
        _check(lib().orcb_file_io_stats(self._h, out))
        return {"reads": int(out[0]), "bytes": int(out[1])}

    def clone(self) -> "_File":
        """A second handle on the same host bytes (a bulk job stages each handle's stripes separately)."""
        c = object.__new__(_File)
        c._h = ctypes.c_void_p()
        c._keep = self  # the original owns the bytes
        _check(lib().orcb_file_clone(self._h, ctypes.byref(c._h)))
        return c

    @property
    def number_of_rows(self) -> int:
        return lib().orcb_file_num_rows(self._h)

    @property
    def num_stripes(self) -> int:
        return lib().orcb_file_num_stripes(self._h)

    @property
    def compression(self) -> int:
        return lib().orcb_file_compression(self._h)

    @property
    def compression_block_size(self) -> int:
        return lib().orcb_file_compression_block_size(self._h)

    @property
    def file_format_version(self) -> str:
        L = lib()
        L.orcb_file_format_version.restype = ctypes.c_char_p
        L.orcb_file_format_version.argtypes = [ctypes.c_void_p]
        return L.orcb_file_format_version(self._h).decode()

    @property
    def user_custom_metadata(self) -> dict:
        L = lib()
        L.orcb_file_num_user_metadata.argtypes = [ctypes.c_void_p]
        out = {}
        for i in range(L.orcb_file_num_user_metadata(self._h)):
            k, v, n = ctypes.c_char_p(), ctypes.POINTER(ctypes.c_uint8)(), ctypes.c_size_t(0)
            _check(L.orcb_file_user_metadata(self._h, ctypes.c_uint32(i), ctypes.byref(k), ctypes.byref(v), ctypes.byref(n)))
            out[k.value.decode()] = ctypes.string_at(v, n.value) if n.value else b""
        return out

    @property
    def row_index_stride(self) -> Optional[int]:
        v = lib().orcb_file_row_index_stride(self._h)
        return None if v < 0 else v

    @property
    def column_names(self) -> List[str]:
        n = lib().orcb_file_num_root_columns(self._h)
        return [lib().orcb_file_root_column_name(self._h, i).decode() for i in range(n)]

    @property
    def column_ids(self) -> List[int]:
        """ORC column index of every root column (`DataType::column_index`, src/schema.rs:325)."""
        n = lib().orcb_file_num_root_columns(self._h)
        return [int(lib().orcb_file_root_column_id(self._h, i)) for i in range(n)]

    def stripe_info(self, i: int):
        out = (ctypes.c_uint64 * 5)()
        _check(lib().orcb_file_stripe_info(self._h, i, out))
        return dict(zip(("offset", "index_length", "data_length", "footer_length", "number_of_rows"), list(out)))


def _make_options(device=0, batch_size=8192, projection=None, byte_range=None, timestamp_precision=0,
                  use_row_index=True, device_resident=False, max_stripes_per_launch=0, cuda_stream=None,
                  shard=None, waves=0):
    o = _ReadOptions()
    o.waves = waves
    o.device = device
    o.batch_size = batch_size
    keep = None
    if projection is not None:
        arr = (ctypes.c_char_p * max(len(projection), 1))(*[p.encode() for p in projection])
        o.projection_names = arr
        o.n_projection = len(projection)
        keep = arr
    if byte_range is not None:
        o.range_start, o.range_end = byte_range
    o.timestamp_unit = timestamp_precision
    o.flags = 0 if use_row_index else 1
    o.device_resident = 1 if device_resident else 0
    o.max_stripes_per_launch = max_stripes_per_launch
    o.cuda_stream = cuda_stream
    if shard is not None:
        o.stripe_shard_index, o.stripe_shard_count = shard
    return o, keep


def _import_schema(fh, opts):
    import pyarrow as pa
    cs = _ArrowSchema()
    _check(lib().orcb_schema(fh, ctypes.byref(opts), ctypes.byref(cs)))
    return pa.Schema._import_from_c(ctypes.addressof(cs))


class _RowSelectorC(ctypes.Structure):
    _fields_ = [("row_count", ctypes.c_uint64), ("skip", ctypes.c_int32), ("reserved", ctypes.c_int32)]


class RowSelector:
    """src/row_selection.rs:32-57"""

    def __init__(self, row_count: int, skip: bool):
        self.row_count, self.skip = int(row_count), bool(skip)

    @classmethod
    def select(cls, row_count: int) -> "RowSelector":
        return cls(row_count, False)

    @classmethod
    def skip_rows(cls, row_count: int) -> "RowSelector":  # `RowSelector::skip` (the attribute has the name)
        return cls(row_count, True)

    def __repr__(self):
        return f"RowSelector({'skip' if self.skip else 'select'} {self.row_count})"

    def __eq__(self, other):
        return isinstance(other, RowSelector) and (self.row_count, self.skip) == (other.row_count, other.skip)


class RowSelection:
    """src/row_selection.rs:90-460 (the parts a reader needs): a list of selectors, normalised on construction
    as `RowSelection::from(Vec<RowSelector>)` does (:466-482)."""

    def __init__(self, selectors=()):
        if isinstance(selectors, RowSelection):
            selectors = selectors.selectors
        out = []
        for x in selectors:
            if not isinstance(x, RowSelector):
                skip, count = x
                x = RowSelector(count, skip)
            if x.row_count == 0:
                continue
            if out and out[-1].skip == x.skip:
                out[-1] = RowSelector(out[-1].row_count + x.row_count, x.skip)
            else:
                out.append(RowSelector(x.row_count, x.skip))
        self.selectors = out

    @classmethod
    def from_consecutive_ranges(cls, ranges, total_rows: int) -> "RowSelection":
        """:158-199: ascending, non-overlapping [start, end) ranges to read out of `total_rows`."""
        sel, last = [], 0
        for a, b in ranges:
            if b <= a:
                continue
            if a < last:
                raise ValueError("ranges must be provided in order and must not overlap")
            if a > last:
                sel.append(RowSelector(a - last, True))
            sel.append(RowSelector(b - a, False))
            last = b
        if last < total_rows:
            sel.append(RowSelector(total_rows - last, True))
        return cls(sel)

    @classmethod
    def select_all(cls, row_count: int) -> "RowSelection":
        return cls([RowSelector(row_count, False)])

    @classmethod
    def skip_all(cls, row_count: int) -> "RowSelection":
        return cls([RowSelector(row_count, True)])

    @classmethod
    def from_filters(cls, filters) -> "RowSelection":
        """:105-142: boolean arrays (pyarrow BooleanArray or sequences of bool, no nulls), one after the other."""
        ranges, offset = [], 0
        for f in filters:
            if hasattr(f, "null_count"):
                assert f.null_count == 0, "filter arrays must not contain nulls"
                f = f.to_pylist()
            start = None
            for idx, v in enumerate(f):
                if v and start is None:
                    start = idx
                elif not v and start is not None:
                    ranges.append((start + offset, idx + offset))
                    start = None
            if start is not None:
                ranges.append((start + offset, len(f) + offset))
            offset += len(f)
        return cls.from_consecutive_ranges(ranges, offset)

    @classmethod
    def from_row_group_filter(cls, row_group_filter, rows_per_group: int, total_rows: int) -> "RowSelection":
        """:348-392: one bool per row group; rows behind the last group are skipped."""
        if not len(row_group_filter):
            return cls.skip_all(total_rows)
        sel = [RowSelector(rows_per_group, not keep) for keep in row_group_filter]
        covered = len(row_group_filter) * rows_per_group
        if covered < total_rows:
            sel.append(RowSelector(total_rows - covered, True))
        return cls(sel)

    def row_count(self) -> int:
        return sum(x.row_count for x in self.selectors)

    def selected_row_count(self) -> int:
        return sum(x.row_count for x in self.selectors if not x.skip)

    def skipped_row_count(self) -> int:
        return sum(x.row_count for x in self.selectors if x.skip)

    def selects_any(self) -> bool:
        return any(not x.skip for x in self.selectors)

    def iter(self):
        return iter(self.selectors)

    def __iter__(self):
        return iter(self.selectors)

    def __eq__(self, other):
        return isinstance(other, RowSelection) and [(x.row_count, x.skip) for x in self.selectors] == \
            [(x.row_count, x.skip) for x in other.selectors]

    def __repr__(self):
        return f"RowSelection({self.selectors})"

    def split_off(self, row_count: int) -> "RowSelection":
        """:278-317: returns the first `row_count` rows, keeps the rest."""
        total = 0
        for idx, x in enumerate(self.selectors):
            total += x.row_count
            if total > row_count:
                break
        else:
            first, self.selectors = self.selectors, []
            return RowSelection(first)
        first, rest = self.selectors[:idx], [RowSelector(x.row_count, x.skip) for x in self.selectors[idx:]]
        overflow = total - row_count
        if rest[0].row_count != overflow:
            first.append(RowSelector(rest[0].row_count - overflow, rest[0].skip))
        rest[0].row_count = overflow
        self.selectors = rest
        out = RowSelection()
        out.selectors = first
        return out

    def and_then(self, other: "RowSelection") -> "RowSelection":
        """:401-463: `other` addresses the rows this selection selects; the result addresses all rows."""
        out, to_skip = [], 0
        first = [[x.row_count, x.skip] for x in self.selectors]
        second = [[x.row_count, x.skip] for x in other.selectors]
        i = j = 0
        while j < len(second):
            b = second[j]
            if i >= len(first):
                raise ValueError("selection exceeds the number of selected rows")
            a = first[i]
            if b[0] == 0:
                j += 1
                continue
            if a[0] == 0:
                i += 1
                continue
            if a[1]:
                to_skip += a[0]
                i += 1
                continue
            n = min(a[0], b[0])
            a[0] -= n
            b[0] -= n
            if b[1]:
                to_skip += n
            else:
                if to_skip:
                    out.append(RowSelector(to_skip, True))
                    to_skip = 0
                out.append(RowSelector(n, False))
        for cnt, skip in first[i:]:
            if cnt:
                if not skip:
                    raise ValueError("selection contains less than the number of selected rows")
                to_skip += cnt
        if to_skip:
            out.append(RowSelector(to_skip, True))
        res = RowSelection()
        res.selectors = out
        return res


class _PredicateNodeC(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int32), ("op", ctypes.c_int32), ("value_type", ctypes.c_int32),
                ("value_is_null", ctypes.c_int32), ("i64", ctypes.c_int64), ("f64", ctypes.c_double),
                ("column", ctypes.c_char_p), ("str", ctypes.c_char_p), ("str_len", ctypes.c_uint64),
                ("n_children", ctypes.c_uint32), ("reserved", ctypes.c_uint32)]


class _ReaderBuildC(ctypes.Structure):
    _fields_ = [("options", ctypes.c_void_p), ("selectors", ctypes.c_void_p), ("n_selectors", ctypes.c_uint32),
                ("has_selection", ctypes.c_int32), ("schema", ctypes.c_void_p), ("predicate", ctypes.c_void_p),
                ("n_predicate_nodes", ctypes.c_uint32), ("reserved", ctypes.c_uint32)]


class ComparisonOp:
    """src/predicate.rs:55-77"""
    Equal, NotEqual, LessThan, LessThanOrEqual, GreaterThan, GreaterThanOrEqual = range(6)


class PredicateValue:
    """`PredicateValue` / `ScalarValue` (src/predicate.rs:28-48): a typed scalar, `None` standing for X(None)."""
    _TYPES = ("Boolean", "Int8", "Int16", "Int32", "Int64", "Float32", "Float64", "Utf8")

    def __init__(self, type_name: str, value):
        if type_name not in self._TYPES:
            raise ValueError(f"unknown predicate value type {type_name}")
        self.type_name, self.value = type_name, value

    def __repr__(self):
        return f"PredicateValue.{self.type_name}({self.value!r})"


for _t in PredicateValue._TYPES:
    setattr(PredicateValue, _t, staticmethod(lambda value=None, _t=_t: PredicateValue(_t, value)))
ScalarValue = PredicateValue


class Predicate:
    """`Predicate` (src/predicate.rs:79-190): comparisons and null tests on top-level columns, combined with
    AND / OR / NOT.  Row groups whose statistics or Bloom filters rule the predicate out are not decoded."""
    _COMPARISON, _IS_NULL, _IS_NOT_NULL, _AND, _OR, _NOT = range(6)

    def __init__(self, kind, column=None, op=None, value=None, children=()):
        self.kind, self.column, self.op, self.value, self.children = kind, column, op, value, list(children)

    @classmethod
    def comparison(cls, column: str, op: int, value: PredicateValue) -> "Predicate":
        if not isinstance(value, PredicateValue):
            raise TypeError("value must be a PredicateValue")
        return cls(cls._COMPARISON, column, op, value)

    @classmethod
    def eq(cls, column, value): return cls.comparison(column, ComparisonOp.Equal, value)

    @classmethod
    def ne(cls, column, value): return cls.comparison(column, ComparisonOp.NotEqual, value)

    @classmethod
    def lt(cls, column, value): return cls.comparison(column, ComparisonOp.LessThan, value)

    @classmethod
    def lte(cls, column, value): return cls.comparison(column, ComparisonOp.LessThanOrEqual, value)

    @classmethod
    def gt(cls, column, value): return cls.comparison(column, ComparisonOp.GreaterThan, value)

    @classmethod
    def gte(cls, column, value): return cls.comparison(column, ComparisonOp.GreaterThanOrEqual, value)

    @classmethod
    def is_null(cls, column): return cls(cls._IS_NULL, column)

    @classmethod
    def is_not_null(cls, column): return cls(cls._IS_NOT_NULL, column)

    @classmethod
    def and_(cls, predicates): return cls(cls._AND, children=predicates)

    @classmethod
    def or_(cls, predicates): return cls(cls._OR, children=predicates)

    @classmethod
    def not_(cls, predicate): return cls(cls._NOT, children=[predicate])

    def _flatten(self, out, keep):
        n = _PredicateNodeC()
        n.kind = self.kind
        n.n_children = len(self.children)
        if self.column is not None:
            raw_name = self.column.encode("utf-8")
            keep.append(raw_name)
            n.column = raw_name
        if self.kind == self._COMPARISON:
            v = self.value
            n.op = self.op
            n.value_type = PredicateValue._TYPES.index(v.type_name)
            n.value_is_null = 1 if v.value is None else 0
            if v.value is not None:
                if v.type_name == "Utf8":
                    raw = v.value.encode("utf-8")
                    keep.append(raw)
                    n.str = raw
                    n.str_len = len(raw)
                elif v.type_name == "Float32":
                    import numpy as np
                    n.f64 = float(np.float32(v.value))  # `*v as f64`
                elif v.type_name == "Float64":
                    n.f64 = float(v.value)
                else:
                    n.i64 = int(v.value)
        out.append(n)
        for c in self.children:
            c._flatten(out, keep)

    def _to_c(self):
        nodes, keep = [], []
        self._flatten(nodes, keep)
        arr = (_PredicateNodeC * len(nodes))(*nodes)
        keep.append(nodes)
        return arr, keep


class ProjectionMask:
    """src/projection.rs:24-80: which root columns to read (a root column brings its whole subtree)."""

    def __init__(self, names=None):
        self.names = None if names is None else list(names)

    @classmethod
    def all(cls) -> "ProjectionMask":
        return cls(None)

    @classmethod
    def roots(cls, file, indices) -> "ProjectionMask":
        """By ORC column index of the root columns (:37-50); indices that name no root column are ignored."""
        f = file.file_metadata() if hasattr(file, "file_metadata") else file
        want = set(int(i) for i in indices)
        return cls([n for n, cid in zip(f.column_names, f.column_ids) if cid in want])

    @classmethod
    def named_roots(cls, file, names) -> "ProjectionMask":
        """By name (:53-69); names the file does not have are ignored."""
        f = file.file_metadata() if hasattr(file, "file_metadata") else file
        want = set(names)
        return cls([n for n in f.column_names if n in want])


class ArrowReaderBuilder:
    """Mirror of `ArrowReaderBuilder` (src/arrow_reader.rs:39-231) with one extra option, `with_device`."""

    def __init__(self, file: _File):
        self._file = file
        self._batch_size = 8192  # DEFAULT_BATCH_SIZE, src/arrow_reader.rs:37
        self._projection: Optional[List[str]] = None
        self._byte_range = None
        self._ts = TimestampPrecision.Nanosecond
        self._device = 0
        self._use_row_index = True
        self._device_resident = False
        self._max_stripes = 0

    @classmethod
    def try_new(cls, source) -> "ArrowReaderBuilder":
        """`source`: path or bytes-like (the two `ChunkReader` impls, src/reader/mod.rs:48-76)."""
        return cls(_File(source))

    @classmethod
    async def try_new_async(cls, source) -> "ArrowReaderBuilder":
        """`ArrowReaderBuilder::try_new_async` (src/async_arrow_reader.rs:292-296): the tail is read and parsed off the
        event loop (an `AsyncChunkReader` maps to a `ChunkReader` whose `get_bytes` blocks on a worker thread)."""
        import asyncio
        return await asyncio.get_running_loop().run_in_executor(None, cls.try_new, source)

    def file_metadata(self) -> _File:
        return self._file

    def with_batch_size(self, batch_size: int) -> "ArrowReaderBuilder":
        self._batch_size = batch_size
        return self

    def with_projection(self, names) -> "ArrowReaderBuilder":
        """A ProjectionMask, or root column names (= ProjectionMask::named_roots, src/projection.rs:53-69)."""
        if isinstance(names, ProjectionMask):
            self._projection = None if names.names is None else list(names.names)
        else:
            self._projection = list(names)
        return self

    def with_file_byte_range(self, start: int, end: int) -> "ArrowReaderBuilder":
        self._byte_range = (start, end)
        return self

    def with_timestamp_precision(self, precision: int) -> "ArrowReaderBuilder":
        self._ts = precision
        return self

    def with_device(self, ordinal: int, resident: bool = False) -> "ArrowReaderBuilder":
        self._device = ordinal
        self._device_resident = resident
        return self

    def with_row_index(self, enabled: bool) -> "ArrowReaderBuilder":
        self._use_row_index = enabled
        return self

    def with_max_stripes_per_launch(self, n: int) -> "ArrowReaderBuilder":
        self._max_stripes = n
        return self

    def with_row_selection(self, selection) -> "ArrowReaderBuilder":
        """`ArrowReaderBuilder::with_row_selection` (src/arrow_reader.rs:113-116).  `selection`: a RowSelection or an
        iterable of RowSelector / (skip, row_count) pairs, counted over the stripes the reader visits."""
        self._selection = RowSelection(selection)
        return self

    def with_schema(self, schema) -> "ArrowReaderBuilder":
        """`ArrowReaderBuilder::with_schema` (src/arrow_reader.rs:80-83): a pyarrow.Schema with one field per projected
        column.  Types must be the default ones, except that timestamps may be asked for in any unit or as
        Decimal128(38, 9); the batches carry this schema (its field names included)."""
        self._schema_override = schema
        return self

    def with_predicate(self, predicate: "Predicate") -> "ArrowReaderBuilder":
        """`ArrowReaderBuilder::with_predicate` (src/arrow_reader.rs:140-176): row groups are pruned by their
        statistics and Bloom filters, stripe by stripe; combined with a row selection as the reference combines them."""
        if not isinstance(predicate, Predicate):
            raise TypeError("predicate must be a Predicate")
        self._predicate = predicate
        return self

    def _options(self):
        return _make_options(self._device, self._batch_size, self._projection, self._byte_range, self._ts,
                             self._use_row_index, self._device_resident, self._max_stripes)

    def schema(self):
        o, keep = self._options()
        return _import_schema(self._file._h, o)

    def build(self) -> "ArrowReader":
        return ArrowReader(self)

    def build_async(self) -> "ArrowStreamReader":
        """`ArrowReaderBuilder::build_async` (src/async_arrow_reader.rs:292-321)."""
        return ArrowStreamReader(ArrowReader(self))


class ArrowReader:
    """Mirror of `ArrowReader` (src/arrow_reader.rs:233-347): an iterator of RecordBatches."""

    def __init__(self, b: ArrowReaderBuilder):
        self._file = b._file
        self._opts, self._keep = b._options()
        self._schema = _import_schema(self._file._h, self._opts)
        self._h = ctypes.c_void_p()
        sel = getattr(b, "_selection", None)
        override = getattr(b, "_schema_override", None)
        pred = getattr(b, "_predicate", None)
        if sel is None and override is None and pred is None:
            _check(lib().orcb_reader_new(self._file._h, ctypes.byref(self._opts), ctypes.byref(self._h)))
        else:
            n_sel = len(sel.selectors) if sel is not None else 0
            arr = (_RowSelectorC * max(n_sel, 1))()
            for i in range(n_sel):
                arr[i].row_count = sel.selectors[i].row_count
                arr[i].skip = 1 if sel.selectors[i].skip else 0
            c_schema = None
            if override is not None:
                c_schema = _ArrowSchema()
                override._export_to_c(ctypes.addressof(c_schema))
            bc = _ReaderBuildC()
            bc.options = ctypes.addressof(self._opts)
            bc.selectors = ctypes.addressof(arr)
            bc.n_selectors = n_sel
            bc.has_selection = 1 if sel is not None else 0
            bc.schema = ctypes.addressof(c_schema) if c_schema is not None else None
            if pred is not None:
                nodes, keep = pred._to_c()
                bc.predicate = ctypes.addressof(nodes)
                bc.n_predicate_nodes = len(nodes)
            try:
                _check(lib().orcb_reader_build(self._file._h, ctypes.byref(bc), ctypes.byref(self._h)))
            finally:
                if c_schema is not None and c_schema.release:  # the library only reads it: release the export here
                    ctypes.CFUNCTYPE(None, ctypes.c_void_p)(c_schema.release)(ctypes.addressof(c_schema))
            if override is not None:
                self._schema = override

    def __del__(self):
        if getattr(self, "_h", None) and self._h.value and _lib is not None:
            _lib.orcb_reader_free(self._h)
            self._h = ctypes.c_void_p()

    def schema(self):
        return self._schema

    def total_row_count(self) -> int:
        return lib().orcb_reader_total_row_count(self._h)

    def plan(self):
        """Host-only: per visited stripe, None (read whole) or the (first row, rows) ranges it will yield."""
        n, ns = ctypes.c_size_t(0), ctypes.c_size_t(0)
        _check(lib().orcb_reader_plan(self._h, None, 0, ctypes.byref(ns), None, 0, ctypes.byref(n)))
        applies = (ctypes.c_int32 * max(ns.value, 1))()
        tri = (ctypes.c_uint64 * (3 * max(n.value, 1)))()
        _check(lib().orcb_reader_plan(self._h, applies, ns.value, ctypes.byref(ns), tri, n.value, ctypes.byref(n)))
        out = [([] if applies[i] else None) for i in range(ns.value)]
        for k in range(n.value):
            out[int(tri[3 * k])].append((int(tri[3 * k + 1]), int(tri[3 * k + 2])))
        return out

    def counters(self) -> dict:
        """Segments planned / stripe tasks staged so far (shrinks with a row selection: partial decode)."""
        out = (ctypes.c_uint64 * 2)()
        _check(lib().orcb_reader_counters(self._h, out))
        return {"segments": int(out[0]), "stripe_tasks": int(out[1])}

    def __iter__(self) -> Iterator:
        return self

    def __next__(self):
        import pyarrow as pa
        arr = _ArrowArray()
        eos = ctypes.c_int(0)
        _check(lib().orcb_reader_next(self._h, ctypes.byref(arr), ctypes.byref(eos)))
        if eos.value:
            raise StopIteration
        return pa.RecordBatch._import_from_c(ctypes.addressof(arr), self._schema)

    def drain(self):
        """Consumes every remaining batch inside the library (`for _ in reader {}`): returns (batches, rows)."""
        out = (ctypes.c_uint64 * 2)()
        _check(lib().orcb_reader_drain(self._h, out))
        return int(out[0]), int(out[1])

    def read_all(self):
        import pyarrow as pa
        batches = list(self)
        return pa.Table.from_batches(batches, schema=self._schema)


class ArrowStreamReader:
    """Mirror of `ArrowStreamReader` (src/async_arrow_reader.rs:283-290): an async stream of RecordBatches over
    `orcb_reader_next_async`: the batch is produced on a thread of the library and the awaiting task is woken from its
    completion callback, so the event loop stays free while the GPU works; after an error the stream stays in the
    error state, like `StreamState::Error` (:262-277)."""

    def __init__(self, reader: "ArrowReader"):
        self._reader = reader

    def schema(self):
        return self._reader.schema()

    def __aiter__(self):
        return self

    async def __anext__(self):
        import asyncio
        import pyarrow as pa
        loop = asyncio.get_running_loop()
        fut = loop.create_future()
        arr = _ArrowArray()

        def done(ctx, status, eos, error):
            msg = (error or b"").decode("utf-8", "replace")
            loop.call_soon_threadsafe(fut.set_result, (status, eos, msg))

        cb = BATCH_DONE(done)
        _check(lib().orcb_reader_next_async(self._reader._h, ctypes.byref(arr), cb, None))
        status, eos, msg = await fut
        if status:
            raise OrcError(status, msg)
        if eos:
            raise StopAsyncIteration
        return pa.RecordBatch._import_from_c(ctypes.addressof(arr), self._reader._schema)


class DecodeJob:
    """One device launch plan over many stripes of many files (bulk API used by bench.py)."""

    def __init__(self, sources, *, device=0, batch_size=8192, projection=None, use_row_index=True,
                 cuda_stream=None, shard=None, timestamp_precision=0, waves=0):
        self._files = [s if isinstance(s, _File) else _File(s) for s in sources]
        self._opts, self._keep = _make_options(device, batch_size, projection, None, timestamp_precision,
                                               use_row_index, True, 0, cuda_stream, shard, waves)
        arr = (ctypes.c_void_p * len(self._files))(*[f._h for f in self._files])
        self._h = ctypes.c_void_p()
        _check(lib().orcb_job_new(arr, len(self._files), ctypes.byref(self._opts), ctypes.byref(self._h)))
        self._schema = None

    def __del__(self):
        if getattr(self, "_h", None) and self._h.value and _lib is not None:
            _lib.orcb_job_free(self._h)
            self._h = ctypes.c_void_p()

    def plan(self):
        _check(lib().orcb_job_plan(self._h))
        return self

    def stage(self):
        _check(lib().orcb_job_stage(self._h))
        return self

    def launch(self):
        _check(lib().orcb_job_launch(self._h))
        return self

    def finish(self):
        _check(lib().orcb_job_finish(self._h))
        return self

    def restage(self):
        _check(lib().orcb_job_restage(self._h))
        return self

    def kernel_stats(self):
        arr = (KernelStat * 32)()
        n = ctypes.c_uint32(0)
        _check(lib().orcb_job_kernel_stats(self._h, arr, 32, ctypes.byref(n)))
        return [dict(name=arr[i].name.decode(), ms=arr[i].ms, alg_bytes=int(arr[i].alg_bytes),
                     work_items=int(arr[i].work_items)) for i in range(n.value)]

    def stats(self) -> dict:
        s = JobStats()
        _check(lib().orcb_job_stats(self._h, ctypes.byref(s)))
        return s.as_dict()

    @property
    def num_batches(self) -> int:
        return lib().orcb_job_num_batches(self._h)

    def schema(self):
        if self._schema is None:
            self._schema = _import_schema(self._files[0]._h, self._opts)
        return self._schema

    def batch(self, i: int):
        import pyarrow as pa
        arr = _ArrowArray()
        _check(lib().orcb_job_export_batch(self._h, i, ctypes.byref(arr)))
        return pa.RecordBatch._import_from_c(ctypes.addressof(arr), self.schema())

    def batches(self):
        return [self.batch(i) for i in range(self.num_batches)]


def predicate_row_groups(source, stripe: int, predicate: "Predicate", projection=None):
    """Host-only: the row groups of one stripe a predicate keeps, or None where the reference falls back to reading the
    whole stripe (src/arrow_reader.rs:281-291)."""
    f = source if isinstance(source, _File) else _File(source)
    opts, keep_opts = _make_options(0, 8192, projection)
    nodes, keep = predicate._to_c()
    n = ctypes.c_size_t(0)
    ev = ctypes.c_int(0)
    cap = 1 << 16
    buf = (ctypes.c_uint8 * cap)()
    _check(lib().orcb_predicate_row_groups(f._h, stripe, ctypes.addressof(opts), ctypes.addressof(nodes), len(nodes),
                                           ctypes.addressof(buf), cap, ctypes.byref(n), ctypes.byref(ev)))
    if not ev.value:
        return None
    return [bool(buf[i]) for i in range(n.value)]


def selection_plan(selection, stripe_rows, batch_size=8192):
    """Host-only view of `orcb_selection_plan`: per stripe None (read whole) or the list of (first row, rows)."""
    raw = list(selection.selectors) if isinstance(selection, RowSelection) else [
        x if isinstance(x, RowSelector) else RowSelector(x[1], x[0]) for x in selection]
    arr = (_RowSelectorC * max(len(raw), 1))()
    for i, x in enumerate(raw):  # un-normalised on purpose: the library normalises like RowSelection::from
        arr[i].row_count, arr[i].skip = x.row_count, 1 if x.skip else 0
    rows = (ctypes.c_uint64 * max(len(stripe_rows), 1))(*stripe_rows)
    applies = (ctypes.c_int32 * max(len(stripe_rows), 1))()
    n = ctypes.c_size_t(0)
    _check(lib().orcb_selection_plan(ctypes.addressof(arr), len(raw), ctypes.addressof(rows), len(stripe_rows), batch_size,
                                     ctypes.addressof(applies), None, 0, ctypes.addressof(n)))
    tri = (ctypes.c_uint64 * max(3 * n.value, 1))()
    _check(lib().orcb_selection_plan(ctypes.addressof(arr), len(raw), ctypes.addressof(rows), len(stripe_rows), batch_size,
                                     ctypes.addressof(applies), ctypes.addressof(tri), n.value, ctypes.addressof(n)))
    out = [([] if applies[s] else None) for s in range(len(stripe_rows))]
    for k in range(n.value):
        out[tri[3 * k]].append((tri[3 * k + 1], tri[3 * k + 2]))
    return out


# ---- stream-level entry points (parity tests against the reference's unit-test vectors) -------------
def decode_int_rle(data: bytes, n: int, *, version: int = 2, signed: bool = True, nbytes: int = 8, device: int = 0):
    import numpy as np
    out = np.empty(n, dtype=np.int64)
    buf = ctypes.cast(ctypes.c_char_p(bytes(data)), ctypes.c_void_p)
    _check(lib().orcb_decode_int_rle(device, buf, len(data), 1 if version == 2 else 0, int(signed), nbytes,
                                     out.ctypes.data_as(ctypes.c_void_p), n))
    return out


def decode_byte_rle(data: bytes, n: int, device: int = 0):
    import numpy as np
    out = np.empty(n, dtype=np.uint8)
    buf = ctypes.cast(ctypes.c_char_p(bytes(data)), ctypes.c_void_p)
    _check(lib().orcb_decode_byte_rle(device, buf, len(data), out.ctypes.data_as(ctypes.c_void_p), n))
    return out


def decode_bool_rle(data: bytes, n: int, device: int = 0):
    """Returns the LSB-first Arrow bitmap unpacked to one uint8 per value."""
    import numpy as np
    bm = np.zeros((n + 7) // 8 + 8, dtype=np.uint8)
    buf = ctypes.cast(ctypes.c_char_p(bytes(data)), ctypes.c_void_p)
    _check(lib().orcb_decode_bool_rle(device, buf, len(data), bm.ctypes.data_as(ctypes.c_void_p), n))
    return np.unpackbits(bm, bitorder="little")[:n]


def decode_varint128(data: bytes, n: int, device: int = 0):
    """Returns an (n, 2) uint64 array of little-endian (lo, hi) halves."""
    import numpy as np
    out = np.empty((n, 2), dtype=np.uint64)
    buf = ctypes.cast(ctypes.c_char_p(bytes(data)), ctypes.c_void_p)
    _check(lib().orcb_decode_varint128(device, buf, len(data), out.ctypes.data_as(ctypes.c_void_p), n))
    return out


def host_decompress_section(kind: int, data: bytes, block_size: int) -> bytes:
    """The host decoder of metadata sections (footers, row indexes); works without a GPU.  Data streams never use it."""
    import numpy as np
    cap = 1 << 16
    while True:
        out = np.empty(cap, dtype=np.uint8)
        out_len = ctypes.c_size_t(0)
        buf = ctypes.cast(ctypes.c_char_p(bytes(data)), ctypes.c_void_p)
        rc = lib().orcb_host_decompress_section(kind, buf, len(data), block_size, out.ctypes.data_as(ctypes.c_void_p), cap,
                                                ctypes.byref(out_len))
        if rc == 21 and out_len.value > cap:
            cap = out_len.value
            continue
        _check(rc)
        return out[: out_len.value].tobytes()


def zone_table(name: str):
    """Host-only: the writer-zone table the device searches (transition instants, offsets from each on, the offset before
    the first, the ORC epoch on the zone's wall clock); works without a GPU."""
    import numpy as np
    n, first, epoch = ctypes.c_size_t(0), ctypes.c_int32(0), ctypes.c_int64(0)
    L = lib()
    _check(L.orcb_zone_table(name.encode(), None, None, ctypes.c_size_t(0), ctypes.byref(n), ctypes.byref(first), ctypes.byref(epoch)))
    at, off = np.zeros(max(n.value, 1), dtype=np.int64), np.zeros(max(n.value, 1), dtype=np.int32)
    _check(L.orcb_zone_table(name.encode(), at.ctypes.data_as(ctypes.c_void_p), off.ctypes.data_as(ctypes.c_void_p),
                             ctypes.c_size_t(n.value), ctypes.byref(n), ctypes.byref(first), ctypes.byref(epoch)))
    return at[: n.value], off[: n.value], int(first.value), int(epoch.value)


def decompress_stream(kind: int, data: bytes, block_size: int, device: int = 0) -> bytes:
    import numpy as np
    n_chunks_max = len(data) // 3 + 1
    cap = min(n_chunks_max, max(1, len(data))) * block_size + 64
    # exact chunk count: walk the headers (framing only)
    p = 0
    chunks = 0
    while p + 3 <= len(data):
        h = data[p] | (data[p + 1] << 8) | (data[p + 2] << 16)
        p += 3 + (h >> 1)
        chunks += 1
    cap = chunks * block_size + 64
    out = np.empty(cap, dtype=np.uint8)
    out_len = ctypes.c_size_t(0)
    buf = ctypes.cast(ctypes.c_char_p(bytes(data)), ctypes.c_void_p)
    _check(lib().orcb_decompress_stream(device, kind, buf, len(data), block_size,
                                        out.ctypes.data_as(ctypes.c_void_p), cap, ctypes.byref(out_len)))
    return out[: out_len.value].tobytes()
