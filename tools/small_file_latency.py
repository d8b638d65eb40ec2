"""Latency of the reference-shaped call on small files: ArrowReaderBuilder.try_new(path).build() + read every batch.

    python tools/small_file_latency.py [--device-resident]
"""
import glob
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import orc_rust_b200 as ob

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
names = ["ref_basic/test.orc", "ref_basic/alltypes.snappy.orc", "ref_basic/alltypes.zstd.orc", "ref_basic/nested_struct.orc",
         "ref_integration/TestOrcFile.test1.orc", "ref_integration/decimal.orc", "ref_integration/TestOrcFile.testSnappy.orc",
         "ref_basic/demo-12-zlib.orc"]
resident = "--device-resident" in sys.argv
for rel in names:
    p = os.path.join(GOLDEN, rel)
    if not os.path.exists(p):
        continue
    def once():
        b = ob.ArrowReaderBuilder.try_new(p)
        if resident:
            b = b.with_device(resident=True)
        r = b.build()
        n = 0
        for batch in r:
            n += batch.num_rows
        return n
    rows = once()
    once()
    ts = []
    for _ in range(20):
        t0 = time.perf_counter()
        once()
        ts.append(time.perf_counter() - t0)
    ts.sort()
    print(f"{rel:50s} {os.path.getsize(p):9d} B {rows:8d} rows  median {ts[len(ts)//2]*1e3:7.3f} ms  min {ts[0]*1e3:7.3f} ms")
