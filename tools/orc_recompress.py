"""ORC re-compressor: rewrites an UNCOMPRESSED ORC file as a Snappy- or LZ4-compressed one with real compressed
chunks of `block_size` bytes (SURVEY.md §7 "LZ4 test data", §8(d) config 3).

pyarrow's ORC writer stores every LZ4 chunk "original", so files that exercise on-device LZ4 decompression - and Snappy
files with exactly the same chunking - are made here: every stream (data, ROW_INDEX, Bloom filters), every stripe
footer, the metadata section and the file footer are cut into chunks of at most `block_size` bytes behind the 3-byte
header of src/compression.rs:113-123 (a chunk is stored as is when compression does not make it smaller, as writers
do); stream lengths, stripe offsets / lengths, the PostScript (compression kind, block size, footer and metadata
lengths) are rewritten, and every row-index position `byte` becomes the pair (start of the chunk inside the compressed
stream, offset inside the decompressed chunk) that compressed files record (src/row_index.rs:37-51; layouts per
column type as in orc_rust_b200/csrc/plan.cc).  Column statistics and everything else are carried over untouched.

    python tools/orc_recompress.py in.orc out.orc --kind lz4 [--block-size 262144]
"""
from __future__ import annotations

import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import lzcodec  # noqa: E402

KIND_CODE = {"snappy": 2, "snappy-lib": 2, "lzo": 3, "lzo-plain": 3, "lz4": 4, "lz4-lib": 4, "zstd": 5}
# ORC TypeKind numbers (orc_proto.proto)
T_BOOLEAN, T_BYTE, T_SHORT, T_INT, T_LONG, T_FLOAT, T_DOUBLE, T_STRING, T_BINARY, T_TIMESTAMP, T_LIST, T_MAP, T_STRUCT, \
    T_UNION, T_DECIMAL, T_DATE, T_VARCHAR, T_CHAR, T_TIMESTAMP_INSTANT = range(19)
S_PRESENT, S_DATA, S_LENGTH, S_DICTIONARY_DATA, S_DICTIONARY_COUNT, S_SECONDARY, S_ROW_INDEX, S_BLOOM, S_BLOOM_UTF8 = range(9)


# ---- protobuf, generic: a message is a list of [field, wire, value] -------------------------------------------
def _varint(buf, p):
    v = 0
    s = 0
    while True:
        b = buf[p]
        p += 1
        v |= (b & 0x7F) << s
        s += 7
        if b < 0x80:
            return v, p


def pb_parse(buf):
    out = []
    p = 0
    n = len(buf)
    while p < n:
        key, p = _varint(buf, p)
        f, w = key >> 3, key & 7
        if w == 0:
            v, p = _varint(buf, p)
        elif w == 1:
            v = bytes(buf[p:p + 8])
            p += 8
        elif w == 2:
            ln, p = _varint(buf, p)
            v = bytes(buf[p:p + ln])
            p += ln
        elif w == 5:
            v = bytes(buf[p:p + 4])
            p += 4
        else:
            raise ValueError("unsupported wire type %d" % w)
        out.append([f, w, v])
    return out


def _enc_varint(v):
    out = bytearray()
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def pb_build(fields):
    out = bytearray()
    for f, w, v in fields:
        out += _enc_varint((f << 3) | w)
        if w == 0:
            out += _enc_varint(v)
        elif w == 2:
            out += _enc_varint(len(v)) + v
        else:
            out += v
    return bytes(out)


def pb_get(fields, f, default=None):
    for ff, _, v in fields:
        if ff == f:
            return v
    return default


def pb_packed(v):
    out = []
    p = 0
    while p < len(v):
        x, p = _varint(v, p)
        out.append(x)
    return out


# ---- chunking ------------------------------------------------------------------------------------------------
def frame(data: bytes, kind: str, block_size: int):
    """-> (framed bytes, [compressed offset of every chunk header], compressed chunks, chunks)"""
    out = bytearray()
    starts = []
    ncomp = 0
    for p in range(0, len(data), block_size):
        chunk = data[p:p + block_size]
        starts.append(len(out))
        c = lzcodec.compress_block(kind, chunk)
        if len(c) >= len(chunk):
            out += ((len(chunk) << 1) | 1).to_bytes(3, "little") + chunk
        else:
            out += (len(c) << 1).to_bytes(3, "little") + c
            ncomp += 1
    return bytes(out), starts, ncomp, len(starts)


def _position_layout(kind, encoding_kind, has_present):
    """[(stream kind, extra positions after the byte offset)] in the order writers record them."""
    out = []
    if has_present:
        out.append((S_PRESENT, 2))
    dict_enc = encoding_kind in (1, 3)
    if kind == T_BOOLEAN:
        out.append((S_DATA, 2))
    elif kind in (T_BYTE, T_SHORT, T_INT, T_LONG, T_DATE):
        out.append((S_DATA, 1))
    elif kind in (T_FLOAT, T_DOUBLE):
        out.append((S_DATA, 0))
    elif kind in (T_STRING, T_VARCHAR, T_CHAR, T_BINARY):
        if dict_enc and kind != T_BINARY:
            out.append((S_DATA, 1))
        else:
            out += [(S_DATA, 0), (S_LENGTH, 1)]
    elif kind == T_DECIMAL:
        out += [(S_DATA, 0), (S_SECONDARY, 1)]
    elif kind in (T_TIMESTAMP, T_TIMESTAMP_INSTANT):
        out += [(S_DATA, 1), (S_SECONDARY, 1)]
    elif kind in (T_LIST, T_MAP):
        out.append((S_LENGTH, 1))
    elif kind == T_UNION:
        out.append((S_DATA, 1))
    return out


def recompress(src_path: str, dst_path: str, kind: str = "lz4", block_size: int = 256 << 10):
    data = open(src_path, "rb").read()
    n = len(data)
    ps_len = data[-1]
    ps = pb_parse(data[n - 1 - ps_len:n - 1])
    if pb_get(ps, 2, 0) != 0:
        raise ValueError("input must be an uncompressed ORC file")
    footer_len, meta_len = pb_get(ps, 1), pb_get(ps, 5, 0)
    fend = n - 1 - ps_len
    footer = pb_parse(data[fend - footer_len:fend])
    metadata = data[fend - footer_len - meta_len:fend - footer_len]
    types = [pb_parse(v) for f, _, v in footer if f == 4]
    type_kind = [pb_get(t, 1, 0) for t in types]
    stats = {"chunks": 0, "compressed": 0}

    def fr(b):
        out, starts, nc, nt = frame(b, kind, block_size)
        stats["chunks"] += nt
        stats["compressed"] += nc
        return out, starts

    out = bytearray(b"ORC")
    new_stripes = []
    for f, w, v in footer:
        if f != 3:
            continue
        si = pb_parse(v)
        off, ilen, dlen, flen = pb_get(si, 1), pb_get(si, 2, 0), pb_get(si, 3, 0), pb_get(si, 4, 0)
        sf = pb_parse(data[off + ilen + dlen:off + ilen + dlen + flen])
        streams = [pb_parse(sv) for sf_f, _, sv in sf if sf_f == 1]
        encodings = [pb_get(pb_parse(ev), 1, 0) for sf_f, _, ev in sf if sf_f == 2]
        # pass 1: data streams -> framed bytes + chunk starts
        pos = off
        raw = []
        for st in streams:
            ln = pb_get(st, 3, 0)
            raw.append(data[pos:pos + ln])
            pos += ln
        framed = [None] * len(streams)
        chunk_starts = {}
        raw_len = {}
        for i, st in enumerate(streams):
            sk, col = pb_get(st, 1, 0), pb_get(st, 2, 0)
            if sk in (S_ROW_INDEX, S_BLOOM, S_BLOOM_UTF8):
                continue
            framed[i], starts = fr(raw[i])
            chunk_starts[(col, sk)] = (starts, len(framed[i]))
            raw_len[(col, sk)] = len(raw[i])
        # pass 2: index streams, positions rewritten
        for i, st in enumerate(streams):
            sk, col = pb_get(st, 1, 0), pb_get(st, 2, 0)
            if sk in (S_BLOOM, S_BLOOM_UTF8):
                framed[i], _ = fr(raw[i])
                continue
            if sk != S_ROW_INDEX:
                continue
            has_present = (col, S_PRESENT) in chunk_starts
            layout = _position_layout(type_kind[col], encodings[col] if col < len(encodings) else 0, has_present)
            ri = pb_parse(raw[i])
            for ent in ri:
                if ent[0] != 1:
                    continue
                e = pb_parse(ent[2])
                for fld in e:
                    if fld[0] != 1:
                        continue
                    p_old = pb_packed(fld[2]) if fld[1] == 2 else [fld[2]]
                    lay = layout
                    expect = sum(1 + x for _, x in lay)
                    if len(p_old) != expect and not has_present:
                        # writers may keep the positions of a suppressed PRESENT stream: carry them over as zeros
                        lay = [(S_PRESENT, 2)] + layout
                        expect = sum(1 + x for _, x in lay)
                    if len(p_old) != expect:
                        raise ValueError("unexpected row-index layout for column %d: %d positions, expected %d" % (col, len(p_old), expect))
                    p_new = []
                    q = 0
                    for skind, extra in lay:
                        byte = p_old[q]
                        q += 1
                        starts, clen = chunk_starts.get((col, skind), ([], 0))
                        k = byte // block_size
                        if k < len(starts):
                            p_new += [starts[k], byte - k * block_size]
                        else:
                            p_new += [clen, 0]  # a position at the very end of the stream names no chunk
                        p_new += p_old[q:q + extra]
                        q += extra
                    fld[1] = 2
                    fld[2] = b"".join(_enc_varint(x) for x in p_new)
                ent[2] = pb_build(e)
            framed[i], _ = fr(pb_build(ri))
        # new stripe footer: stream lengths
        for i, st in enumerate(streams):
            for fld in st:
                if fld[0] == 3:
                    fld[2] = len(framed[i])
        k = 0
        for fld in sf:
            if fld[0] == 1:
                fld[2] = pb_build(streams[k])
                k += 1
        new_off = len(out)
        new_ilen = new_dlen = 0
        for i, st in enumerate(streams):
            if pb_get(st, 1, 0) in (S_ROW_INDEX, S_BLOOM, S_BLOOM_UTF8):
                new_ilen += len(framed[i])
            else:
                new_dlen += len(framed[i])
            out += framed[i]
        sfb, _ = fr(pb_build(sf))
        out += sfb
        for fld in si:
            if fld[0] == 1:
                fld[2] = new_off
            elif fld[0] == 2:
                fld[2] = new_ilen
            elif fld[0] == 3:
                fld[2] = new_dlen
            elif fld[0] == 4:
                fld[2] = len(sfb)
        new_stripes.append(pb_build(si))
    content_len = len(out)
    k = 0
    for fld in footer:
        if fld[0] == 3:
            fld[2] = new_stripes[k]
            k += 1
        elif fld[0] == 2:
            fld[2] = content_len
    mb, _ = fr(metadata) if metadata else (b"", [])
    fb, _ = fr(pb_build(footer))
    out += mb
    out += fb
    has_bs = False
    for fld in ps:
        if fld[0] == 1:
            fld[2] = len(fb)
        elif fld[0] == 2:
            fld[2] = KIND_CODE[kind]
        elif fld[0] == 3:
            fld[2] = block_size
            has_bs = True
        elif fld[0] == 5:
            fld[2] = len(mb)
    if pb_get(ps, 2) is None:
        ps.insert(1, [2, 0, KIND_CODE[kind]])
    if not has_bs:
        ps.insert(2, [3, 0, block_size])
    psb = pb_build(ps)
    out += psb
    out.append(len(psb))
    tmp = dst_path + ".tmp%d" % os.getpid()
    with open(tmp, "wb") as f:
        f.write(out)
    os.replace(tmp, dst_path)
    return stats


def _one(args):
    src, dst, kind, bs = args
    if not os.path.exists(dst):
        recompress(src, dst, kind, bs)
    return dst


def recompress_dataset(files, kind: str, block_size: int = 256 << 10, procs: int = 0):
    """Re-compresses every file next to its source (name + .<kind>); existing outputs are reused."""
    import concurrent.futures as cf
    import multiprocessing as mp
    lzcodec.lib()  # build once, before forking
    jobs = [(f, f[:-4] + f".re{kind}.orc", kind, block_size) for f in files]
    procs = procs or min(len(jobs), os.cpu_count() or 1)
    if procs <= 1 or all(os.path.exists(j[1]) for j in jobs):
        return [_one(j) for j in jobs]
    with cf.ProcessPoolExecutor(procs, mp_context=mp.get_context("fork")) as ex:
        return list(ex.map(_one, jobs))


def compressed_chunk_fraction(path: str):
    """Share of a file's data-stream chunks that are really compressed (and the byte share they decode to)."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    data = open(path, "rb").read()
    n = len(data)
    ps_len = data[-1]
    ps = pb_parse(data[n - 1 - ps_len:n - 1])
    comp = pb_get(ps, 2, 0)
    if comp == 0:
        return {"chunks": 0, "compressed": 0, "fraction": 0.0}
    total = compd = 0
    # chunk headers can be walked without decompressing anything once the stripe layout is known; the layout itself
    # is compressed, so it comes from the library's own metadata reader
    import orc_rust_b200 as ob
    f = ob._File(path)
    for s in range(f.num_stripes):
        si = f.stripe_info(s)
        p = si["offset"] + si["index_length"]
        end = p + si["data_length"]
        while p + 3 <= end:
            h = data[p] | (data[p + 1] << 8) | (data[p + 2] << 16)
            total += 1
            compd += 0 if (h & 1) else 1
            p += 3 + (h >> 1)
    return {"chunks": total, "compressed": compd, "fraction": round(compd / max(total, 1), 4)}


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("src")
    ap.add_argument("dst")
    ap.add_argument("--kind", default="lz4", choices=["lz4", "lz4-lib", "snappy", "snappy-lib", "zstd", "lzo", "lzo-plain"])
    ap.add_argument("--block-size", type=int, default=256 << 10)
    a = ap.parse_args()
    st = recompress(a.src, a.dst, a.kind, a.block_size)
    print(f"{a.src} ({os.path.getsize(a.src)} B) -> {a.dst} ({os.path.getsize(a.dst)} B), {st['compressed']}/{st['chunks']} chunks compressed")
