"""Per-column cost of the decode kernels: one job per projected column over the bench file set
(run with ORCB_SERIAL=1 so kernel timings do not overlap).

    ORCB_SERIAL=1 python tools/col_cost.py [rows] [files]
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import bench  # noqa: E402
import orc_rust_b200 as ob  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 16_000_000
nfiles = int(sys.argv[2]) if len(sys.argv) > 2 else 8
files, _ = bench._dataset(rows, nfiles, "uncompressed")
names = ob.ArrowReaderBuilder.try_new(files[0]).schema().names
for col in names:
    job = ob.DecodeJob(files, device=0, projection=[col])
    job.plan(); job.stage()
    for _ in range(3):
        job.launch()
    job.finish()
    ks = {k["name"]: round(k["ms"], 3) for k in job.kernel_stats() if k["ms"] > 0.002}
    print(f"{col:18s} total={sum(ks.values()):7.3f} ms  {ks}")
    del job
