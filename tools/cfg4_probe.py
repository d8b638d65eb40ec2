"""Where the time of a single-stripe compressed file goes: per-kernel times of config 4 (null-heavy, 2 M rows)."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
import gen_orc
import orc_rust_b200 as ob
from oracle import orc_oracle as oo

d = "/tmp/cfg4"; os.makedirs(d, exist_ok=True)
t = gen_orc.nullheavy_table(2_000_000, 1)
for comp in ("uncompressed", "snappy", "zstd"):
    p = os.path.join(d, f"nh_{comp}.orc")
    if not os.path.exists(p):
        gen_orc.write(t, p, compression=comp)
    st = torch.cuda.Stream()
    job = ob.DecodeJob([p], cuda_stream=st.cuda_stream)
    job.plan(); job.stage(); job.launch(); job.finish()
    for _ in range(3):
        job.launch()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(5):
        job.launch()
    e1.record(st); torch.cuda.synchronize()
    job.finish()
    ks = job.kernel_stats()
    print(comp, "file MB %.1f" % (os.path.getsize(p) / 1e6), "ms/launch %.3f" % (e0.elapsed_time(e1) / 5), job.stats()["n_stripes"], "stripes")
    print("   ", {k["name"][:16]: round(k["ms"], 3) for k in ks if k["ms"] > 0.02})
    if comp != "uncompressed":
        of = oo.OracleFile(open(p, "rb").read())
        import collections
        sizes = collections.Counter()
        n_chunks = 0
        for s in of.stripes[:1]:
            sf = of.stripe_footer(0) if hasattr(of, "stripe_footer") else None
        print("   block size", of.block_size)
