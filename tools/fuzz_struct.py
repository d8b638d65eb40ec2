"""Structure-aware companion of tools/fuzz_host.sh: instead of overwriting bytes it edits protobuf FIELDS of an
ORC file - the file footer, the last stripe's footer and its row-index streams - and rebuilds the file around
the edit (section lengths in the footer / postscript follow; sections of compressed files are stored back as original chunks), so every input still frames correctly and the edit reaches
the code behind the parser: a stream length of 2^63, a dictionary size of 2^32-1, a row-index position past the stream,
a stripe of 2^60 rows, a stride of 0, a duplicated or missing stream ...  Each input goes through open, schema, planning
with and without the row index, predicate evaluation and a reader with a selection.  Run it on the sanitizer build:

    tools/asan_pytest.sh --version >/dev/null   # builds /tmp/orcb_fuzz/liborc_b200_asan.so
    G=$(dirname $(gcc -print-file-name=libasan.so)); LD_PRELOAD=$G/libasan.so:$G/libubsan.so \\
      ASAN_OPTIONS=detect_leaks=0:allocator_may_return_null=1 python tools/fuzz_struct.py --asan [iterations] [seed]
"""
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import orc_recompress as rc  # noqa: E402  (pb_parse / pb_build)

EXTREMES = [0, 1, 2, 3, 7, 127, 128, 255, 256, 1000, 65535, 65536, (1 << 24) - 1, (1 << 31) - 1, 1 << 31, (1 << 32) - 1, 1 << 32,
            (1 << 40) + 5, (1 << 62) + 1, (1 << 63) - 1, 1 << 63, (1 << 64) - 1]


def as_message(b: bytes):
    """bytes -> field list if they parse as a protobuf message with sane field numbers, else None"""
    if not b:
        return None
    try:
        f = rc.pb_parse(b)
    except Exception:
        return None
    if not f or any(x[0] == 0 or x[0] > 64 for x in f) or rc.pb_build(f) != bytes(b):
        return None
    return f


def edit(fields, rng, depth=0):
    """One random edit somewhere in the field tree (in place); returns True when something changed."""
    if not fields:
        return False
    for _ in range(8):
        i = rng.randrange(len(fields))
        f, w, v = fields[i]
        r = rng.random()
        if w == 0:
            fields[i][2] = rng.choice(EXTREMES) if r < 0.7 else max(0, v + rng.choice((-1, 1, -8, 8, 1000)))
            return True
        if w == 2:
            sub = as_message(v) if depth < 4 else None
            if sub is not None and r < 0.75:
                if edit(sub, rng, depth + 1):
                    fields[i][2] = rc.pb_build(sub)
                    return True
                continue
            if r < 0.80:  # packed varints (positions, subtypes, versions): rewrite one of them
                try:
                    vals = rc.pb_packed(v)
                    if vals:
                        vals[rng.randrange(len(vals))] = rng.choice(EXTREMES)
                        fields[i][2] = b"".join(rc._enc_varint(x) for x in vals)
                        return True
                except Exception:
                    pass
            if r < 0.88:
                del fields[i]
                return True
            if r < 0.96:
                fields.insert(i, [f, w, v])
                return True
            fields[i][2] = bytes(v[: len(v) // 2])
            return True
    return False


class Orc:
    """A file cut into the pieces this fuzzer rebuilds.  Compressed files too: a metadata section is decompressed, edited
    and stored back as "original" chunks (header bit 0 set), which every compression kind allows; their row-index area is
    a run of separately framed streams and is left alone."""

    def unframe(self, raw: bytes) -> bytes:
        if self.kind == 0:
            return bytes(raw)
        from oracle import orc_oracle as oo
        return bytes(oo.decompress_stream(self.kind, bytes(raw), self.block))

    def frame(self, b: bytes) -> bytes:
        if self.kind == 0:
            return b
        out = bytearray()
        for p in range(0, len(b), self.block):
            c = b[p:p + self.block]
            out += ((len(c) << 1) | 1).to_bytes(3, "little") + c
        return bytes(out)

    def __init__(self, data: bytes):
        self.data = data
        n = len(data)
        self.ps_len = data[-1]
        self.ps = rc.pb_parse(data[n - 1 - self.ps_len:n - 1])
        self.kind = rc.pb_get(self.ps, 2, 0)
        self.block = rc.pb_get(self.ps, 3, 256 << 10)
        self.fl = rc.pb_get(self.ps, 1)
        self.ml = rc.pb_get(self.ps, 5, 0)
        self.footer = rc.pb_parse(self.unframe(data[n - 1 - self.ps_len - self.fl:n - 1 - self.ps_len]))
        self.meta = data[n - 1 - self.ps_len - self.fl - self.ml:n - 1 - self.ps_len - self.fl]
        self.stripes = [rc.pb_parse(f[2]) for f in self.footer if f[0] == 3]
        self.body_end = n - 1 - self.ps_len - self.fl - self.ml

    def build(self, footer=None, stripe_footer=None, index=None):
        """The file with a new footer, and / or a new footer or index area for its LAST stripe."""
        footer = [list(f) for f in (footer if footer is not None else self.footer)]
        body = self.data[:self.body_end]
        if (stripe_footer is not None or index is not None) and self.stripes:
            si = {g[0]: g[2] for g in self.stripes[-1]}
            off, il, dl, sfl = si.get(1, 0), si.get(2, 0), si.get(3, 0), si.get(4, 0)
            idx = index if index is not None else body[off:off + il]
            sf = self.frame(stripe_footer) if stripe_footer is not None else body[off + il + dl:off + il + dl + sfl]
            body = body[:off] + idx + body[off + il:off + il + dl] + sf
            k = [i for i, f in enumerate(footer) if f[0] == 3][-1]
            s = rc.pb_parse(footer[k][2])
            for g in s:
                if g[0] == 2:
                    g[2] = len(idx)
                if g[0] == 4:
                    g[2] = len(sf)
            footer[k][2] = rc.pb_build(s)
        nf = self.frame(rc.pb_build(footer))
        ps = [list(f) for f in self.ps]
        for f in ps:
            if f[0] == 1:
                f[2] = len(nf)
        nps = rc.pb_build(ps)
        return body + self.meta + nf + nps + bytes([len(nps)])

    def last_stripe_parts(self):
        si = {g[0]: g[2] for g in self.stripes[-1]}
        off, il, dl, sfl = si.get(1, 0), si.get(2, 0), si.get(3, 0), si.get(4, 0)
        return self.data[off:off + il], self.unframe(self.data[off + il + dl:off + il + dl + sfl])


def exercise(ob, data: bytes):
    try:
        b = ob.ArrowReaderBuilder.try_new(data)
        names = b.schema().names
    except ob.OrcError:
        return "open"
    verdict = "ok"
    for idx in (True, False):
        try:
            ob.DecodeJob([data], use_row_index=idx).plan().stats()
        except ob.OrcError:
            verdict = "plan"
    for name in names[:4]:
        try:
            ob.predicate_row_groups(data, 0, ob.Predicate.or_([ob.Predicate.eq(name, ob.PredicateValue.Int64(5)),
                                                              ob.Predicate.lt(name, ob.PredicateValue.Utf8("m")),
                                                              ob.Predicate.is_null(name)]))
        except ob.OrcError:
            verdict = "predicate"
    try:
        r = ob.ArrowReaderBuilder.try_new(data).with_row_selection(ob.RowSelection.from_consecutive_ranges([(3, 900), (2500, 2600)], 6000)).build()
        r.plan()
    except ob.OrcError:
        verdict = "selection"
    return verdict


def seeds(tmp):
    import pyarrow as pa
    import pyarrow.orc as po
    import gen_orc
    out = []
    tables = {"cfg1": gen_orc.config1_table(6000), "lineitem": gen_orc.lineitem_table(1500), "nullheavy": gen_orc.nullheavy_table(5000),
              "nested": pa.table({"s": pa.array([{"a": i, "b": [i, i + 1], "m": [("k", float(i))]} if i % 7 else None for i in range(3000)],
                                                pa.struct([("a", pa.int64()), ("b", pa.list_(pa.int64())), ("m", pa.map_(pa.string(), pa.float64()))])),
                                  "x": pa.array(range(3000))})}
    for name, t in tables.items():
        p = os.path.join(tmp, f"fs_{name}.orc")
        po.write_table(t, p, compression="uncompressed", stripe_size=64 << 10, row_index_stride=1000, dictionary_key_size_threshold=0.8)
        out.append(open(p, "rb").read())
    for comp in ("snappy", "zlib", "zstd"):  # compressed files: chunk tables, (chunk start, offset in chunk) positions
        p = os.path.join(tmp, f"fs_lineitem_{comp}.orc")
        po.write_table(tables["lineitem"], p, compression=comp, stripe_size=64 << 10, row_index_stride=1000, compression_block_size=64 << 10,
                       dictionary_key_size_threshold=0.8)
        out.append(open(p, "rb").read())
    import orc_recompress
    for kind in ("lz4", "lzo"):
        p = os.path.join(tmp, f"fs_nullheavy_{kind}.orc")
        orc_recompress.recompress(os.path.join(tmp, "fs_nullheavy.orc"), p, kind, 16 << 10)
        out.append(open(p, "rb").read())
    # Bloom filters (with_predicate reads them for equality): integer and string columns
    t = pa.table({"k": pa.array([(i * 7919) % 100_000 for i in range(6000)], pa.int64()), "w": pa.array([f"w{(i * 31) % 977}" for i in range(6000)])})
    p = os.path.join(tmp, "fs_bloom.orc")
    po.write_table(t, p, compression="uncompressed", row_index_stride=1000, bloom_filter_columns=[0, 1], bloom_filter_fpp=0.05,
                   dictionary_key_size_threshold=0.0)
    out.append(open(p, "rb").read())
    for f in ("tests/golden/ref_basic/test.orc", "tests/golden/ref_basic/alltypes.none.orc"):
        d = open(os.path.join(ROOT, f), "rb").read()
        out.append(d)
    return out


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    iters = int(args[0]) if args else 2000
    rng = random.Random(int(args[1]) if len(args) > 1 else 1)
    import orc_rust_b200 as ob
    if "--asan" in sys.argv:
        ob._build.build = lambda *a, **k: "/tmp/orcb_fuzz/liborc_b200_asan.so"
    import tempfile
    tally = {}
    with tempfile.TemporaryDirectory() as tmp:
        files = []
        for d in seeds(tmp):
            try:
                files.append(Orc(d))
            except AssertionError:
                pass
        for o in files:
            assert exercise(ob, o.build()) == "ok", "an unedited rebuild must still read"
        for it in range(iters):
            o = rng.choice(files)
            target = rng.choice(("footer", "footer", "stripe_footer", "stripe_footer", "index"))
            if target == "index" and o.kind != 0:
                target = "stripe_footer"
            if target == "footer":
                f = [list(x) for x in o.footer]
                if not edit(f, rng):
                    continue
                data = o.build(footer=f)
            else:
                idx, sf = o.last_stripe_parts()
                part = as_message(sf if target == "stripe_footer" else idx)
                if target == "index":
                    # the index area is a concatenation of RowIndex messages, one per indexed column: edit it as one message
                    part = as_message(idx)
                if part is None or not edit(part, rng):
                    continue
                new = rc.pb_build(part)
                if target == "index" and len(new) != len(idx):
                    continue  # the stripe footer holds the index streams' lengths: only edits that keep them are index-only damage
                data = o.build(stripe_footer=new) if target == "stripe_footer" else o.build(index=new)
            v = exercise(ob, data)
            tally[(target, v)] = tally.get((target, v), 0) + 1
            if "--oracle" in sys.argv:
                # does the oracle (the restatement of the reference) take the same view of the metadata?
                from oracle import orc_oracle as oo
                try:
                    of = oo.OracleFile(data)
                    of.schema()
                    of.read()  # the whole decode: what the product's planner refuses, the reference must fail on as well
                    ov = "ok"
                except oo.OracleError as e:
                    ov = "err:" + e.variant
                except Exception as e:  # noqa: BLE001
                    ov = "exc:" + type(e).__name__
                pv = "ok"
                try:
                    b = ob.ArrowReaderBuilder.try_new(data)
                    b.schema()
                    ob.DecodeJob([data], use_row_index=False).plan()
                except ob.OrcError as e:
                    pv = "err:" + e.variant
                    if ov == "ok" and "--verbose" in sys.argv:
                        print("   oracle reads it, product:", target, e)
                if ov == "ok" and pv != "ok" or ov.startswith("exc"):
                    key = ("DIFF", target, ov, pv)
                    tally[key] = tally.get(key, 0) + 1
                    if tally[key] <= 2:
                        open(f"/tmp/fs_diff_{len(tally)}_{it}.orc", "wb").write(data)
    diffs = {k: n for k, n in tally.items() if k[0] == "DIFF"}
    tally = {k: n for k, n in tally.items() if k[0] != "DIFF"}
    print("fuzz_struct:", sum(tally.values()), "inputs;", ", ".join(f"{t}/{v}: {n}" for (t, v), n in sorted(tally.items())))
    for k, n in sorted(diffs.items(), key=lambda kv: -kv[1]):
        print("  oracle / product differ:", k[1], "oracle", k[2], "product", k[3], "x", n)


if __name__ == "__main__":
    main()
