"""Where one ArrowReader per file spends its time (host phases), single-threaded: open (file -> pinned memory), build,
drain (plan / stage / launch / wait / export).  ORCB_READER_TIMING=1 adds the library's own phase lines."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gen_orc
import orc_rust_b200 as ob
files = gen_orc.lineitem_dataset(os.environ.get("ORCB_BENCH_DIR", "/tmp/orcb200_bench"), 59_986_052, 32)[:8]
for resident in (True, False):
    for rep in range(2):
        t_open = t_build = t_drain = 0.0
        for p in files:
            t0 = time.perf_counter(); b = ob.ArrowReaderBuilder.try_new(p); t1 = time.perf_counter()
            r = b.with_device(0, resident=resident).build(); t2 = time.perf_counter()
            r.drain(); t3 = time.perf_counter()
            t_open += t1 - t0; t_build += t2 - t1; t_drain += t3 - t2
            del r, b
        n = len(files)
        print(f"resident={resident} pass {rep}: per file open {t_open/n*1e3:.2f} ms, build {t_build/n*1e3:.2f} ms, drain {t_drain/n*1e3:.2f} ms", flush=True)
