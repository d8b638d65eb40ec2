"""ctypes front of tools/lzcodec.c (LZ4-block / Snappy-raw / LZO1X compressors for test and bench inputs; Zstandard
frames come from pyarrow's codec)."""
import ctypes
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(HERE, "_liblzcodec.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(HERE, "lzcodec.c")
        if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
            subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", _SO, src])
        L = ctypes.CDLL(_SO)
        if not hasattr(L, "lzc_lzo_use_m1"):  # a library from before that entry point (copied trees do not keep mtimes)
            tmp = _SO + f".{os.getpid()}.tmp"
            subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", tmp, src])
            os.replace(tmp, _SO)
            L = ctypes.CDLL(_SO)
        for f in (L.lzc_lz4_compress, L.lzc_snappy_compress, L.lzc_lzo_compress):
            f.restype = ctypes.c_size_t
            f.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p]
        L.lzc_bound.restype = ctypes.c_size_t
        L.lzc_bound.argtypes = [ctypes.c_size_t]
        _lib = L
    return _lib


def compress_block(kind: str, data: bytes) -> bytes:
    L = lib()
    buf = ctypes.create_string_buffer(L.lzc_bound(len(data)))
    if kind in ("zstd", "lz4-lib", "snappy-lib"):
        # the libraries themselves (libzstd, liblz4, snappy) through pyarrow's codecs
        import pyarrow as pa
        codec = pa.Codec("zstd", compression_level=3) if kind == "zstd" else pa.Codec("lz4_raw" if kind == "lz4-lib" else "snappy")
        return codec.compress(bytes(data), asbytes=True) if len(data) else b""
    fn = {"lz4": L.lzc_lz4_compress, "snappy": L.lzc_snappy_compress, "lzo": L.lzc_lzo_compress, "lzo-plain": L.lzc_lzo_compress}[kind]
    L.lzc_lzo_use_m1(0 if kind == "lzo-plain" else 1)  # "lzo-plain": without the M1 instructions (see lzcodec.c)
    n = fn(bytes(data), len(data), buf)
    return buf.raw[:n]


def orc_frame(data: bytes, kind: str, block_size: int, keep_if_smaller: bool = True) -> bytes:
    """`data` as an ORC stream: <= block_size chunks, each compressed (or stored when that is not smaller) behind the
    3-byte header (len << 1) | is_original (reference src/compression.rs:113-123)."""
    out = bytearray()
    for p in range(0, len(data), block_size):
        chunk = data[p:p + block_size]
        c = compress_block(kind, chunk)
        if keep_if_smaller and len(c) >= len(chunk):
            out += ((len(chunk) << 1) | 1).to_bytes(3, "little") + chunk
        else:
            out += (len(c) << 1).to_bytes(3, "little") + c
    return bytes(out)
