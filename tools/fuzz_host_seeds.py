"""Seed corpus for tools/fuzz_host.cc: framed compressed sections of every codec, one file per seed, named
<kind-code>_<name>.sec (kind-code = proto CompressionKind).  Usage: python tools/fuzz_host_seeds.py <out-dir>"""
import os
import sys
import zlib

import numpy as np
import pyarrow as pa

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import lzcodec  # noqa: E402


def framed(c: bytes) -> bytes:
    return (len(c) << 1).to_bytes(3, "little") + c


def main(out: str) -> None:
    os.makedirs(out, exist_ok=True)
    rng = np.random.default_rng(4)
    words = [bytes(rng.integers(97, 123, rng.integers(2, 10), dtype=np.uint8)) for _ in range(400)]
    cases = {
        "one": b"a", "rle": b"z" * 20_000,
        "text": b" ".join(words[i] for i in rng.integers(0, 400, 6_000)),
        "noise": bytes(rng.integers(0, 256, 5_000, dtype=np.uint8)),
        "lowent": bytes(rng.integers(0, 4, 30_000, dtype=np.uint8)),
        "skewed": bytes(np.minimum(rng.geometric(0.05, 30_000), 255).astype(np.uint8)),
        "ints": np.cumsum(rng.integers(0, 100, 5_000)).astype("<i8").tobytes(),
    }
    for name, d in cases.items():
        for level in (1, 3, 19):
            open(os.path.join(out, f"5_{name}_l{level}.sec"), "wb").write(
                framed(pa.Codec("zstd", compression_level=level).compress(d, asbytes=True)))
        open(os.path.join(out, f"3_{name}.sec"), "wb").write(framed(lzcodec.compress_block("lzo", d)))
        open(os.path.join(out, f"2_{name}.sec"), "wb").write(framed(lzcodec.compress_block("snappy", d)))
        open(os.path.join(out, f"4_{name}.sec"), "wb").write(framed(lzcodec.compress_block("lz4", d)))
        for lvl in (1, 6, 9):
            co = zlib.compressobj(lvl, zlib.DEFLATED, -15)
            open(os.path.join(out, f"1_{name}_l{lvl}.sec"), "wb").write(framed(co.compress(d) + co.flush()))
        co = zlib.compressobj(6, zlib.DEFLATED, -15, 9, zlib.Z_FIXED)
        open(os.path.join(out, f"1_{name}_fixed.sec"), "wb").write(framed(co.compress(d) + co.flush()))
    # generated ORC files: the BASELINE configs at toy sizes (several stripes, row index every 1000 rows), NONE and every
    # codec (pyarrow's writers, plus the re-compressor for real LZ4 / LZO chunks)
    import gen_orc
    import orc_recompress
    import pyarrow.orc as orc
    tables = {"cfg1": gen_orc.config1_table(6000), "lineitem": gen_orc.lineitem_table(1500), "nullheavy": gen_orc.nullheavy_table(5000)}
    for name, t in tables.items():
        for comp in ("uncompressed", "snappy", "zlib", "zstd"):
            orc.write_table(t, os.path.join(out, f"gen_{name}_{comp}.orc"), compression=comp, stripe_size=64 << 10,
                            row_index_stride=1000, compression_block_size=64 << 10, dictionary_key_size_threshold=0.8)
        for kind in ("lz4", "lzo"):
            orc_recompress.recompress(os.path.join(out, f"gen_{name}_uncompressed.orc"), os.path.join(out, f"gen_{name}_{kind}.orc"), kind, 16 << 10)


if __name__ == "__main__":
    main(sys.argv[1])
