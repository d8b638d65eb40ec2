"""Throughput of the reference-facing iterator (ArrowReaderBuilder ... build(), device-resident batches) over the
bench's lineitem files, next to the bulk DecodeJob path bench.py times: what one reader per file costs in planning,
allocation and launch overheads.    python tools/reader_probe.py [rows] [files]"""
import ctypes
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gen_orc
import orc_rust_b200 as ob
import torch

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 59_986_052
n_files = int(sys.argv[2]) if len(sys.argv) > 2 else 32
files = gen_orc.lineitem_dataset(os.environ.get("ORCB_BENCH_DIR", "/tmp/orcb200_bench"), rows, n_files)
L = ob.lib()
release = ctypes.CFUNCTYPE(None, ctypes.c_void_p)
handles = [ob._File(f) for f in files]  # opened once: the file bytes sit in pinned host memory


first_s = 0.0


def one_pass():
    global first_s
    n_rows = n_batches = 0
    first_s = 0.0
    for fh in handles:
        reader = ob.ArrowReaderBuilder(fh).with_device(0, resident=True).build()
        first = True
        while True:
            dev = ob._ArrowDeviceArray()
            eos = ctypes.c_int(0)
            t0 = time.perf_counter()
            ob._check(L.orcb_reader_next_device(reader._h, ctypes.byref(dev), ctypes.byref(eos)))
            if first:  # the call that plans, stages, launches and waits for the file's stripes
                first_s += time.perf_counter() - t0
                first = False
            if eos.value:
                break
            n_rows += dev.array.length
            n_batches += 1
            release(dev.array.release)(ctypes.addressof(dev.array))
    return n_rows, n_batches


one_pass()
torch.cuda.synchronize()
os.environ["ORCB_READER_TIMING"] = "1"
r = ob.ArrowReaderBuilder(handles[0]).with_device(0, resident=True).build()
for _ in r:
    break
r = ob.ArrowReaderBuilder(handles[1]).with_device(0, resident=True).build()
for _ in r:
    break
del r
del os.environ["ORCB_READER_TIMING"]
best = 1e9
for _ in range(3):
    t0 = time.perf_counter()
    n_rows, n_batches = one_pass()
    torch.cuda.synchronize()
    best = min(best, time.perf_counter() - t0)
print(f"reader path: {n_rows} rows, {n_batches} batches, {best * 1e3:.1f} ms per pass over {n_files} files "
      f"({n_rows / best / 1e9:.2f} G rows/s); {first_s * 1e3:.1f} ms of it in the first next() of each file (plan + allocate + "
      f"copy + decode), the rest is per-batch export through ctypes")
