"""Wall time of the device Snappy decoder per stream of one lineitem stripe (1M rows, 256 KiB chunks): which
streams hold the slow chunks.  Run on a GPU box: python tools/snappy_probe.py"""
import os, sys, time
os.environ["ORCB_STREAM_TIMING"] = "1"
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gen_orc
import orc_rust_b200 as ob
from oracle import orc_oracle as oo

path = "/tmp/snappy_probe.orc"
if not os.path.exists(path):
    gen_orc.write(gen_orc.lineitem_table(250_000, 0), path, compression="snappy", block_size=256 << 10)
data = open(path, "rb").read()
of = oo.OracleFile(data)
streams, _, _ = of._stripe_footer(of.stripes[0])
names = {cid: n for n, cid in of.columns}
ob.decompress_stream(2, data[streams[-1].offset:streams[-1].offset + streams[-1].length], 256 << 10)  # warm-up
for st in streams:
    if st.kind in (6, 7, 8) or st.length < 100_000:
        continue
    raw = data[st.offset:st.offset + st.length]
    exp = bytes(oo.decompress_stream(2, raw, 256 << 10))
    best = 1e9
    for _ in range(2):
        t0 = time.perf_counter()
        got = ob.decompress_stream(2, raw, 256 << 10)
        best = min(best, time.perf_counter() - t0)
    assert bytes(got) == exp
    print(f"{names.get(st.column, '?'):16s} kind {st.kind} in {st.length:9d} out {len(exp):9d}  {best * 1e3:7.2f} ms", file=sys.stderr, flush=True)
