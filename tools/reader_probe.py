"""Throughput of the reference-facing iterator (ArrowReaderBuilder ... build(), device-resident batches), next to the
bulk DecodeJob path bench.py times: what one reader per file costs in planning, allocation and launch overheads, and
what starting the next group of stripes ahead of time buys on a file with several stripes.

    python tools/reader_probe.py            # one reader per file over the bench's 32 SF10 files + the multi-stripe file
    python tools/reader_probe.py --multi    # the multi-stripe file only (ORCB_NO_PREFETCH=1 to compare)
    python tools/reader_probe.py --host     # batches copied back to host memory (`for batch in reader`), 6 files
ORCB_READER_TIMING=1 prints the host phases of every group to stderr."""
import ctypes
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gen_orc
import orc_rust_b200 as ob
import torch

L = ob.lib()
release = ctypes.CFUNCTYPE(None, ctypes.c_void_p)


def drain(reader):
    """all batches of a reader, released at once; returns (rows, batches, seconds spent in the first next())"""
    rows = batches = 0
    first = None
    while True:
        dev = ob._ArrowDeviceArray()
        eos = ctypes.c_int(0)
        t0 = time.perf_counter()
        ob._check(L.orcb_reader_next_device(reader._h, ctypes.byref(dev), ctypes.byref(eos)))
        if first is None:
            first = time.perf_counter() - t0
        if eos.value:
            return rows, batches, first
        rows += dev.array.length
        batches += 1
        release(dev.array.release)(ctypes.addressof(dev.array))


def files_pass(handles):
    rows = batches = 0
    first = 0.0
    for fh in handles:
        r, b, f = drain(ob.ArrowReaderBuilder(fh).with_device(0, resident=True).build())
        rows, batches, first = rows + r, batches + b, first + f
    return rows, batches, first


def best_of(n, fn):
    best, out = 1e9, None
    for _ in range(n):
        t0 = time.perf_counter()
        o = fn()
        torch.cuda.synchronize()
        t = time.perf_counter() - t0
        if t < best:
            best, out = t, o
    return best, out


if "--host" in sys.argv:
    files = gen_orc.lineitem_dataset(os.environ.get("ORCB_BENCH_DIR", "/tmp/orcb200_bench"), 59_986_052, 32)[:6]
    handles = [ob._File(f) for f in files]

    def host_pass():
        return sum(b.num_rows for fh in handles for b in ob.ArrowReaderBuilder(fh).build())

    host_pass()
    t, rows = best_of(3, host_pass)
    print(f"host-resident batches: {rows} rows over {len(files)} files, {t * 1e3 / len(files):.1f} ms per file")
    sys.exit(0)

if "--multi" not in sys.argv:
    files = gen_orc.lineitem_dataset(os.environ.get("ORCB_BENCH_DIR", "/tmp/orcb200_bench"), 59_986_052, 32)
    handles = [ob._File(f) for f in files]  # opened once: the file bytes sit in pinned host memory
    files_pass(handles)
    t, (rows, batches, first) = best_of(3, lambda: files_pass(handles))
    print(f"one reader per file: {rows} rows, {batches} batches, {t * 1e3:.1f} ms per pass over {len(files)} files "
          f"({rows / t / 1e9:.2f} G rows/s); {first * 1e3:.1f} ms of it in the first next() of each file (plan + allocate + "
          f"copy + decode), the rest is per-batch export through ctypes")

# one file with several stripes, one stripe per launch group
multi = "/tmp/orcb200_reader_probe_multi.orc"
if not os.path.exists(multi):
    gen_orc.write(gen_orc.lineitem_table(2_000_000, 5), multi)
fh = ob._File(multi)
mk = lambda: ob.ArrowReaderBuilder(fh).with_device(0, resident=True).with_max_stripes_per_launch(1).build()
drain(mk())
t, (rows, batches, first) = best_of(4, lambda: drain(mk()))
print(f"multi-stripe file: {rows} rows, {fh.num_stripes} stripes, one per group: {t * 1e3:.1f} ms "
      f"(prefetch {'off' if os.environ.get('ORCB_NO_PREFETCH') == '1' else 'on'})")
