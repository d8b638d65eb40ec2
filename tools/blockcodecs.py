"""Tiny pure-Python LZ4-block and Snappy-raw COMPRESSORS (test tooling: pyarrow's ORC writer stores every
LZ4 chunk uncompressed, so real LZ4 blocks with matches have to be produced here) plus the ORC chunk framing
(3-byte little-endian header, (len << 1) | is_original, reference src/compression.rs:113-123)."""
from __future__ import annotations


def _matches(data: bytes, min_match: int = 4, max_dist: int = 65535):
    """Greedy hash-table match finder.  Yields (literal_start, literal_end, match_dist, match_len)."""
    n = len(data)
    table = {}
    i = 0
    lit = 0
    while i + min_match <= n:
        key = data[i:i + min_match]
        j = table.get(key)
        table[key] = i
        if j is not None and i - j <= max_dist:
            ml = min_match
            while i + ml < n and data[j + ml] == data[i + ml]:
                ml += 1
            yield lit, i, i - j, ml
            i += ml
            lit = i
        else:
            i += 1
    yield lit, n, 0, 0


def lz4_compress_block(data: bytes) -> bytes:
    out = bytearray()
    n = len(data)

    def put_len(v):
        while v >= 255:
            out.append(255)
            v -= 255
        out.append(v)

    def final_literals(start):
        ll = n - start
        out.append(min(ll, 15) << 4)
        if ll >= 15:
            put_len(ll - 15)
        out.extend(data[start:n])

    for ls, le, dist, ml in _matches(data):
        # LZ4 end-of-block rules: the last 5 bytes are literals, the last match starts >= 12 bytes before the end
        last = False
        if ml and (le > n - 12 or le + ml > n - 5):
            ml = min(ml, n - 5 - le) if le <= n - 12 else 0
            if ml < 4:
                ml = 0
            last = True
        if ml == 0:
            final_literals(ls)
            return bytes(out)
        ll = le - ls
        out.append((min(ll, 15) << 4) | min(ml - 4, 15))
        if ll >= 15:
            put_len(ll - 15)
        out.extend(data[ls:le])
        out.extend((dist & 255, dist >> 8))
        if ml - 4 >= 15:
            put_len(ml - 4 - 15)
        if last:
            final_literals(le + ml)
            return bytes(out)
    return bytes(out)


def snappy_compress_block(data: bytes) -> bytes:
    out = bytearray()
    v = len(data)
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            break

    def literal(chunk):
        ll = len(chunk)
        if ll == 0:
            return
        if ll <= 60:
            out.append((ll - 1) << 2)
        elif ll <= 256:
            out.append(60 << 2)
            out.append(ll - 1)
        elif ll <= 65536:
            out.append(61 << 2)
            out.extend((ll - 1).to_bytes(2, "little"))
        else:
            out.append(62 << 2)
            out.extend((ll - 1).to_bytes(3, "little"))
        out.extend(chunk)

    for ls, le, dist, ml in _matches(data):
        literal(data[ls:le])
        while ml > 0:
            take = min(ml, 64)
            if ml - take in (1, 2, 3):  # leave a copy of at least 4 for the next round
                take = ml - 4 if ml - 4 >= 4 else take
            if 4 <= take <= 11 and dist < 2048:
                out.append(1 | ((take - 4) << 2) | ((dist >> 8) << 5))
                out.append(dist & 255)
            else:
                out.append(2 | ((take - 1) << 2))
                out += dist.to_bytes(2, "little")
            ml -= take
    return bytes(out)


def orc_frame(data: bytes, kind: str, block_size: int, keep_if_smaller: bool = True) -> bytes:
    """Splits `data` into <= block_size chunks, compresses each (kind = 'lz4' | 'snappy') and adds the ORC header."""
    comp = lz4_compress_block if kind == "lz4" else snappy_compress_block
    out = bytearray()
    for p in range(0, len(data), block_size):
        chunk = data[p:p + block_size]
        c = comp(chunk)
        if keep_if_smaller and len(c) >= len(chunk):
            hdr = (len(chunk) << 1) | 1
            out += hdr.to_bytes(3, "little") + chunk
        else:
            hdr = len(c) << 1
            out += hdr.to_bytes(3, "little") + c
    return bytes(out)
