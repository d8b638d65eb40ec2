#!/usr/bin/env python
"""bench.py — decoded Arrow GB/s (+ rows/s) of the ORC stripe decode path on TPC-H-lineitem-shaped ORC.

  python bench.py --gpus N --steps K --warmup W            # B200 CUDA path (this repo)
  python bench.py --impl reference --steps K --warmup W    # CPU path (oracle port of orc-rust's decoders)

Workload (BASELINE.json configs[4] = configs[1] at SF100 scale): ONE fixed set of lineitem ORC files, SF100-shaped:
the synthetic SF10 set (32 files, 96 stripes of <= 64 MiB, 59 986 052 rows, tools/gen_orc.py seeds 0..31) tiled
`--tiles` (7) times = 672 stripes / 419.9 M rows / 23.1 GB stored / 72.5 GB of Arrow.  Stripe i of that list goes to
rank i % N (no collective on the data path); a rank decodes its stripes in launch groups of at most 96 stripes (one
DecodeJob each, arenas resident in HBM together).  Scaling is therefore STRONG: the set is the same for every N.
A "step" = one full decode of every stripe of the set.
  value      : whole-job decoded Arrow GB/s with the compressed stripes already resident in HBM
               (CUDA events on the launching stream, max over ranks)
  e2e        : same metric through the C ABI with HOST buffers: pinned H2D of every stripe + decode + D2H of the
               per-batch metadata inside the timed region (decoded Arrow stays in HBM)
  e2e_reader / e2e_reader_host : one ArrowReader per file of the SF10 tile, drained inside the library
               (orcb_reader_drain), batches left in HBM / copied back to pinned host memory
  configs    : the other BASELINE.json configs (1: 1 M rows one stripe; 2: SF10 in one job; 3: Snappy / LZ4;
               4: null-heavy), each with ms, GB/s, roofline fraction and a parity digest against the oracle
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

SF10_ROWS = 59_986_052
METRIC = "decoded Arrow GB/s, lineitem ORC"
GROUP_STRIPES = 96  # stripes per launch group (one SF10 tile at N = 1: ~18.5 GB of arenas)


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def _bench_dir():
    return os.environ.get("ORCB_BENCH_DIR", "/tmp/orcb200_bench")


def _dataset(rows: int, n_files: int, compression: str):
    import gen_orc
    t0 = time.time()
    if compression in ("lz4", "lz4-lib", "snappy-recompressed", "zstd", "lzo"):
        # pyarrow's LZ4 writer stores every chunk "original": real LZ4 (and an identically chunked Snappy variant)
        # comes from the in-repo re-compressor applied to the uncompressed set
        import orc_recompress
        base = gen_orc.lineitem_dataset(_bench_dir(), rows, n_files, compression="uncompressed")
        files = orc_recompress.recompress_dataset(base, "snappy" if compression == "snappy-recompressed" else compression, 256 << 10)
    else:
        files = gen_orc.lineitem_dataset(_bench_dir(), rows, n_files, compression=compression)
    return files, time.time() - t0


def _config(args):
    """The workload, identical for both arms."""
    rows = args.rows * args.tiles
    return {
        "workload": (f"lineitem SF{rows / SF10_ROWS * 10:.3g}-shaped ORC: SF{args.rows / SF10_ROWS * 10:g} set "
                     f"({args.rows} rows, {args.files} files, 64 MiB stripes) tiled {args.tiles}x = {rows} rows; "
                     f"compression {args.compression}; RLEv2 + dictionary strings + decimal128(15,2) + date32; "
                     f"stripe i -> GPU i % N, launch groups of <= {args.group_stripes} stripes"),
        "batch_size": 8192, "rows": rows, "tiles": args.tiles, "files_per_tile": args.files,
        "compression": args.compression, "row_index": not args.no_row_index,
        "l2_policy": "inputs+outputs per launch group (>= 13 GB) far exceed the 126 MB L2",
    }


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled during the timed region (B200_PROFILING.md)."""

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.samples = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.05)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm = sorted(int(s[0]) for s in self.samples if s and s[0].isdigit())
        mx = [int(s[1]) for s in self.samples if len(s) > 1 and s[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for s in self.samples for i in range(4) if len(s) > 2 + i and s[2 + i] == "Active"})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------------
# CPU arm: the oracle port (orc-rust cannot be built here: no Rust toolchain), one reader per stripe
# ---------------------------------------------------------------------------------------------------
def _cpu_decode_stripe(args):
    path, stripe = args
    from oracle import orc_oracle as oo
    of = _cpu_decode_stripe.cache.get(path)
    if of is None:
        of = oo.OracleFile(open(path, "rb").read())
        _cpu_decode_stripe.cache[path] = of
    batches = of.read_stripe(stripe)
    return sum(b.num_rows for b in batches), sum(b.nbytes for b in batches)


_cpu_decode_stripe.cache = {}


def _pyarrow_decode_stripe(args):
    path, stripe = args
    import pyarrow.orc as po
    f = _pyarrow_decode_stripe.cache.get(path)
    if f is None:
        f = po.ORCFile(path)
        _pyarrow_decode_stripe.cache[path] = f
    b = f.read_stripe(stripe)
    return b.num_rows, b.nbytes


_pyarrow_decode_stripe.cache = {}


def _cpu_tasks(files):
    import pyarrow.orc as po
    tasks = []
    for f in files:
        tasks += [(f, i) for i in range(po.ORCFile(f).nstripes)]
    return tasks


def cpu_run(files, steps, warmup, max_stripes=None, fn=_cpu_decode_stripe):
    """Decodes every stripe (or a bounded sample) with one reader per stripe on all host cores."""
    import concurrent.futures as cf
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    tasks = _cpu_tasks(files)
    if max_stripes:
        tasks = tasks[:max_stripes]
    times = []
    rows = nbytes = 0
    with cf.ProcessPoolExecutor(cores, mp_context=mp.get_context("fork")) as ex:
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            res = list(ex.map(fn, tasks, chunksize=1))
            dt = time.perf_counter() - t0
            if it >= warmup:
                times.append(dt)
            rows = sum(r[0] for r in res)
            nbytes = sum(r[1] for r in res)
    return dict(seconds=sum(times) / len(times), rows=rows, arrow_bytes=nbytes, cores=cores, stripes=len(tasks))


def _claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner to fd 1
    when NCCL_DEBUG is set), so fd 1 is pointed at stderr for the whole run and the line goes to the real stdout."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)
    return real


def _emit(real_stdout: int, line: dict):
    sys.stdout.flush()
    os.write(real_stdout, (json.dumps(line) + "\n").encode())


def _log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)


# ---------------------------------------------------------------------------------------------------
# parity digest: sha256 over the meaningful extent of every Arrow buffer of every batch
# ---------------------------------------------------------------------------------------------------
def batches_digest(batches):
    import numpy as np
    import pyarrow as pa
    h = hashlib.sha256()

    def arr(a):
        n = len(a)
        h.update(str(a.type).encode() + b"|%d|%d|" % (n, a.null_count))
        bufs = a.buffers()
        assert a.offset == 0
        if bufs[0] is not None and a.null_count:
            h.update(np.unpackbits(np.frombuffer(bufs[0], np.uint8, (n + 7) // 8), bitorder="little")[:n].tobytes())
        t = a.type
        if pa.types.is_struct(t):
            for i in range(t.num_fields):
                arr(a.field(i))
        elif pa.types.is_list(t) or pa.types.is_map(t):
            o = np.frombuffer(bufs[1], np.int32, n + 1)
            h.update(o.tobytes())
            arr(a.values if pa.types.is_list(t) else pa.StructArray.from_arrays([a.keys, a.items], ["k", "v"]))
        elif pa.types.is_boolean(t):
            h.update(np.unpackbits(np.frombuffer(bufs[1], np.uint8, (n + 7) // 8), bitorder="little")[:n].tobytes())
        elif pa.types.is_string(t) or pa.types.is_binary(t):
            o = np.frombuffer(bufs[1], np.int32, n + 1)
            h.update(o.tobytes())
            if int(o[-1]):
                h.update(np.frombuffer(bufs[2], np.uint8, int(o[-1])).tobytes())
        else:
            h.update(np.frombuffer(bufs[1], np.uint8, n * (t.bit_width // 8)).tobytes())

    for b in batches:
        h.update(b"B%d|" % b.num_rows)
        for i in range(b.num_columns):
            arr(b.column(i))
    return h.hexdigest()


def parity_sample(path, stripe, device=0):
    """GPU vs oracle on one stripe of one file: digests of the byte-normalised batches."""
    import orc_rust_b200 as ob
    from oracle import orc_oracle as oo
    f = ob._File(path)
    stripe = min(stripe, f.num_stripes - 1)
    si = f.stripe_info(stripe)
    got = list(ob.ArrowReaderBuilder(f).with_device(device).with_file_byte_range(si["offset"], si["offset"] + 1).build())
    exp = oo.OracleFile(open(path, "rb").read()).read_stripe(stripe)
    dg, de = batches_digest(got), batches_digest(exp)
    return {"file": os.path.basename(path), "stripe": stripe, "rows": sum(b.num_rows for b in got),
            "gpu_sha256": dg[:16], "oracle_sha256": de[:16], "match": dg == de}


# ---------------------------------------------------------------------------------------------------
# timing helpers (GPU arm)
# ---------------------------------------------------------------------------------------------------
class GroupRunner:
    """The rank's launch groups, each a DecodeJob on its own torch stream; a step forks them from / joins them to the
    timing stream so that groups overlap like the waves inside a group do."""

    def __init__(self, torch, ob, groups, device, use_row_index, shard, waves, n_streams):
        self.torch = torch
        self.main = torch.cuda.Stream()
        self.streams = [torch.cuda.Stream() for _ in range(max(1, min(n_streams, len(groups))))]
        self.jobs = []
        for gi, files in enumerate(groups):
            st = self.streams[gi % len(self.streams)]
            self.jobs.append((ob.DecodeJob(files, device=device, cuda_stream=st.cuda_stream, use_row_index=use_row_index,
                                           shard=shard, waves=waves), st))
        self.fork = torch.cuda.Event()
        self.joins = [torch.cuda.Event() for _ in self.streams]

    def prepare(self):
        for j, _ in self.jobs:
            j.plan()
        for j, _ in self.jobs:
            j.stage()
        for j, _ in self.jobs:
            j.launch()
        for j, _ in self.jobs:
            j.finish()  # surfaces decode errors before anything is timed

    def step(self, restage=False):
        self.fork.record(self.main)
        for s in self.streams:
            s.wait_event(self.fork)
        for j, _ in self.jobs:
            if restage:
                j.restage()
            j.launch()
        for s, e in zip(self.streams, self.joins):
            e.record(s)
            self.main.wait_event(e)

    def finish(self):
        for j, _ in self.jobs:
            j.finish()

    def stats(self):
        tot = {}
        for j, _ in self.jobs:
            for k, v in j.stats().items():
                tot[k] = tot.get(k, 0) + v if k != "n_columns" else v
        return tot

    def kernel_stats(self):
        agg = {}
        for j, _ in self.jobs:
            for k in j.kernel_stats():
                a = agg.setdefault(k["name"], {"name": k["name"], "ms": 0.0, "alg_bytes": 0})
                a["ms"] += k["ms"]
                a["alg_bytes"] += k["alg_bytes"]
        return list(agg.values())


def time_device(torch, runner, steps, warmup, sync_all, clock_gpu=None):
    for _ in range(warmup):
        runner.step()
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(clock_gpu) if clock_gpu is not None else None
    if sampler:
        sampler.__enter__()
    e0.record(runner.main)
    for _ in range(steps):
        runner.step()
    e1.record(runner.main)
    sync_all()
    if sampler:
        sampler.__exit__()
    ms = e0.elapsed_time(e1) / steps
    runner.finish()
    return ms, (sampler.summary() if sampler else None)


def time_e2e(torch, runner, steps, warmup, sync_all):
    def step():
        runner.step(restage=True)
        runner.finish()  # D2H of the per-batch metadata + error words, host sync
    for _ in range(warmup):
        step()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3 / steps


def h2d_ceiling(torch, sync_all, nbytes=1 << 30, reps=3):
    """Pinned host -> device copy rate of this rank while every rank copies (plain torch copies, none of this repo's
    code): the platform's ceiling for the end-to-end figures."""
    src = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    dst = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    dst.copy_(src, non_blocking=True)
    sync_all()
    t0 = time.perf_counter()
    for _ in range(reps):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    del src, dst
    return reps * nbytes / dt / 1e9


def time_readers(ob, files, device, resident, threads, passes=2):
    """One ArrowReader per file (the reference's public API), drained inside the library; `threads` readers in flight."""
    import concurrent.futures as cf

    def one(path):
        b = ob.ArrowReaderBuilder.try_new(path).with_device(device, resident=resident)
        return b.build().drain()

    best, rows = None, 0
    with cf.ThreadPoolExecutor(threads) as ex:
        for _ in range(passes + 1):  # first pass warms the pinned-buffer cache and the memory pool
            t0 = time.perf_counter()
            res = list(ex.map(one, files))
            dt = time.perf_counter() - t0
            rows = sum(r[1] for r in res)
            best = dt if best is None else min(best, dt)
    return best * 1e3, rows


def single_job_case(torch, ob, name, files, device, peak, steps=5, warmup=3, parity=None, waves=0):
    """One DecodeJob over `files`: device-resident ms per launch, GB/s, roofline fraction, optional parity sample."""
    st = torch.cuda.Stream()
    job = ob.DecodeJob(files, device=device, cuda_stream=st.cuda_stream, waves=waves)
    job.plan(); job.stage(); job.launch(); job.finish()
    for _ in range(warmup):
        job.launch()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(steps):
        job.launch()
    e1.record(st)
    torch.cuda.synchronize()
    job.finish()
    ms = e0.elapsed_time(e1) / steps
    s = job.stats()
    ks = sorted(job.kernel_stats(), key=lambda k: -k["ms"])[:4]
    line = {"config": name, "rows": s["n_rows"], "stripes": s["n_stripes"], "input_bytes": s["input_bytes"],
            "arrow_bytes": s["output_bytes"], "ms": round(ms, 4), "value": round(s["output_bytes"] / ms / 1e6, 1),
            "unit": "GB/s", "rows_per_s": round(s["n_rows"] / ms * 1e3),
            "step_frac": round((s["input_bytes"] + s["output_bytes"]) / ms / 1e6 / peak, 4),
            "top_kernels_in_step": {k["name"]: round(k["ms"], 3) for k in ks}}
    del job
    if parity:
        try:
            line["parity"] = parity_sample(*parity, device=device)
        except Exception as e:  # a failed check is reported, never hidden
            line["parity"] = {"match": False, "error": repr(e)[:200]}
    return line


def other_configs(torch, ob, args, device, peak, sf10_files):
    import gen_orc
    out = []
    d = os.path.join(_bench_dir(), "cfg")
    os.makedirs(d, exist_ok=True)
    check = not args.no_cpu_baseline

    def case(name, mk):
        try:
            out.append(mk())
        except Exception as e:
            out.append({"config": name, "error": repr(e)[:300]})
        _log("config done:", name)

    p1 = os.path.join(d, "config1.orc")
    if not os.path.exists(p1):
        gen_orc.write(gen_orc.config1_table(1_000_000, 0), p1, stripe_size=1 << 30, dict_threshold=1.0)
    case("1", lambda: single_job_case(torch, ob, "1: 1M rows, single stripe, int64 DELTA + int64 DIRECT-24 + dictionary string, NONE",
                                      [p1], device, peak, steps=20, parity=(p1, 0) if check else None))
    case("2", lambda: single_job_case(torch, ob, f"2: lineitem SF10 ({args.rows} rows, {args.files} files), NONE, one DecodeJob",
                                      sf10_files, device, peak, steps=10, parity=(sf10_files[0], 1) if check else None))
    for comp, label in (("snappy", "Snappy (pyarrow writer, 256 KiB chunks)"), ("snappy-recompressed", "Snappy (in-repo re-compressor, 256 KiB chunks)"),
                        ("lz4-lib", "LZ4 (liblz4 blocks framed by the in-repo re-compressor, 256 KiB chunks)"),
                        ("zlib", "Zlib (pyarrow writer, 256 KiB chunks; the ORC default of Hive / Java writers)"),
                        ("zstd", "Zstandard level 3 (in-repo re-compressor, 256 KiB chunks)")):
        def mk(comp=comp, label=label):
            files, _ = _dataset(args.rows, args.files, comp)
            line = single_job_case(torch, ob, f"3: lineitem SF10, {label}", files, device, peak, steps=5,
                                   parity=(files[0], 1) if check else None)
            try:
                import orc_recompress
                line["compressed_chunk_fraction"] = orc_recompress.compressed_chunk_fraction(files[0])
            except Exception:
                pass
            return line
        case("3 " + comp, mk)
    # the reference's own bench file (benches/arrow_reader.rs:29-59): demo-12-zlib.orc, 1.92 M rows, Zlib
    demo = os.path.join(ROOT, "tests", "golden", "ref_basic", "demo-12-zlib.orc")
    if os.path.exists(demo):
        case("zlib demo-12", lambda: single_job_case(torch, ob, "reference bench file demo-12-zlib.orc (1 920 800 rows, Zlib chunks inflated on the device)",
                                                     [demo], device, peak, steps=5, parity=(demo, 0) if check else None))
    t = None
    for comp in ("uncompressed", "snappy"):
        p4 = os.path.join(d, f"nullheavy_{comp}.orc")
        if not os.path.exists(p4):
            t = t if t is not None else gen_orc.nullheavy_table(2_000_000, 1)
            gen_orc.write(t, p4, compression=comp)
        case("4 " + comp, lambda p4=p4, comp=comp: single_job_case(
            torch, ob, f"4: 2M rows, 50% nulls, PATCHED_BASE / timestamps / decimal(38,10) / bool / f64, {comp}", [p4], device, peak,
            steps=10, parity=(p4, 0) if check else None))
    return out


def main():
    real_stdout = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=SF10_ROWS, help="rows of one tile (default: SF10)")
    ap.add_argument("--files", type=int, default=32, help="ORC files of one tile")
    ap.add_argument("--tiles", type=int, default=7, help="times the tile is repeated (7 x SF10 = 672 stripes)")
    ap.add_argument("--compression", default="uncompressed", choices=["uncompressed", "snappy", "zlib", "lz4", "lz4-lib", "snappy-recompressed", "zstd", "lzo"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-row-index", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the configs array (1, 2, 3, 4)")
    ap.add_argument("--no-readers", action="store_true", help="skip the ArrowReader end-to-end figures")
    ap.add_argument("--waves", type=int, default=0, help="stripe waves per DecodeJob (0 = library default)")
    ap.add_argument("--group-streams", type=int, default=3, help="launch groups in flight")
    ap.add_argument("--reader-threads", type=int, default=4)
    ap.add_argument("--group-stripes", type=int, default=GROUP_STRIPES, help="stripes per launch group and rank")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    config = _config(args)

    if args.impl == "reference":
        if rank != 0:
            return
        files, gen_s = _dataset(args.rows, args.files, args.compression)
        r = cpu_run(files, args.steps, args.warmup)
        gbs = r["arrow_bytes"] / r["seconds"] / 1e9
        line = {
            "impl": "reference", "metric": METRIC, "value": gbs, "unit": "GB/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["seconds"] * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
            "rows_per_s": r["rows"] / r["seconds"], "config": config,
            "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": r["cores"], "kind": "port",
                             "sample": f"one tile of the set per step: all {r['stripes']} distinct stripes ({r['rows']} rows; the "
                                       f"{args.tiles} tiles are identical), one oracle reader per stripe over {r['cores']} processes "
                                       "(oracle port of orc-rust's decoders; the Rust reference cannot be built in this image)"},
            "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "run": {"dataset_gen_s": round(gen_s, 1)},
        }
        _emit(real_stdout, line)
        return

    import torch
    import torch.distributed as dist
    import orc_rust_b200 as ob

    if not torch.cuda.is_available() or not ob.device_available():
        raise SystemExit("bench.py: no CUDA device — the decode path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    # rank 0 generates (or finds) the file set, everyone else waits
    if rank == 0:
        files, gen_s = _dataset(args.rows, args.files, args.compression)
    if world > 1:
        dist.barrier()
    if rank != 0:
        files, gen_s = _dataset(args.rows, args.files, args.compression)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- the set: tiles x files, stripes sharded i % world; launch groups of whole tiles with <= GROUP_STRIPES stripes per rank
    t_open = time.time()
    base = [ob._File(f) for f in files]   # one pinned host copy per file and rank
    stripes_per_tile = sum(f.num_stripes for f in base)
    tiles_per_group = max(1, min(args.tiles, (args.group_stripes * world) // max(stripes_per_tile, 1)))
    tiles = [base] + [[f.clone() for f in base] for _ in range(args.tiles - 1)]
    groups = []
    for t0 in range(0, args.tiles, tiles_per_group):
        groups.append([f for tile in tiles[t0:t0 + tiles_per_group] for f in tile])
    shard_note = "stripe i % N over the whole list"
    if (stripes_per_tile * tiles_per_group) % world:
        shard_note = "stripe i % N inside each launch group"
    open_s = time.time() - t_open
    runner = GroupRunner(torch, ob, groups, local_rank, not args.no_row_index, (rank, world) if world > 1 else None,
                         args.waves, args.group_streams)
    torch.cuda.set_stream(runner.main)
    runner.prepare()
    st = runner.stats()
    _log(f"rank {rank}: {st['n_stripes']} stripes in {len(groups)} groups, {st['device_bytes'] / 1e9:.1f} GB of arenas, "
         f"{st['n_waves']} waves, open {open_s:.1f}s")

    # ---- device-resident timing
    dev_ms, clocks = time_device(torch, runner, args.steps, args.warmup, sync_all, clock_gpu=local_rank)
    kstats = runner.kernel_stats()
    st = runner.stats()
    # one more pass with one kernel at a time (outside the timed region): per-kernel durations without the
    # contention of the overlapped schedule, reported beside the in-step ones
    kserial = None
    if "ORCB_SERIAL" not in os.environ:
        os.environ["ORCB_SERIAL"] = "1"
        try:
            for j, _ in runner.jobs:
                j.launch()
                j.finish()
            kserial = runner.kernel_stats()
        finally:
            del os.environ["ORCB_SERIAL"]

    # ---- end-to-end through the C ABI: pinned H2D of every stripe + decode + metadata D2H every step
    e2e_ms = time_e2e(torch, runner, max(2, args.steps // 2), 1, sync_all)
    e2e_staged, e2e_meta = st["staged_bytes"], st["d2h_meta_bytes"]

    h2d_gbs = h2d_ceiling(torch, sync_all)
    t = torch.tensor([dev_ms, e2e_ms, -h2d_gbs], device="cuda", dtype=torch.float64)
    tot = torch.tensor([st["output_bytes"], st["input_bytes"], st["n_rows"], st["n_stripes"], st["aliased_output_bytes"],
                        e2e_staged, e2e_meta, st["n_kernel_launches"], st["device_bytes"]], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    dev_ms, e2e_ms, h2d_min = t.tolist()
    h2d_min = -h2d_min  # slowest rank's rate while all ranks copy
    out_bytes, in_bytes, rows, n_stripes, aliased, h2d, d2h, launches, dev_bytes = (int(x) for x in tot.tolist())
    value = out_bytes / (dev_ms / 1e3) / 1e9
    e2e = out_bytes / (e2e_ms / 1e3) / 1e9
    peak, peak_src = _peaks()

    # ---- readers (public API) on this rank's share of the files of one tile
    readers = None
    del runner
    torch.cuda.synchronize()
    if not args.no_readers:
        try:
            mine = files[rank::world]
            ms_dev, rrows = time_readers(ob, mine, local_rank, True, args.reader_threads)
            ms_host, _ = time_readers(ob, mine, local_rank, False, args.reader_threads)
            rt = torch.tensor([ms_dev, ms_host], device="cuda", dtype=torch.float64)
            rr = torch.tensor([rrows], device="cuda", dtype=torch.float64)
            if world > 1:
                dist.all_reduce(rt, op=dist.ReduceOp.MAX)
                dist.all_reduce(rr, op=dist.ReduceOp.SUM)
            ms_dev, ms_host = rt.tolist()
            tile_bytes = out_bytes / args.tiles
            readers = {
                "e2e_reader": {"value": tile_bytes / ms_dev / 1e6, "unit": "GB/s", "ms_per_pass": ms_dev,
                               "note": "ArrowReaderBuilder.try_new(path).with_device(resident).build() per file of one tile, "
                                       f"drained in the library, {args.reader_threads} readers in flight per GPU; file read into pinned "
                                       "memory + H2D + decode inside the timed region, batches stay in HBM"},
                "e2e_reader_host": {"value": tile_bytes / ms_host / 1e6, "unit": "GB/s", "ms_per_pass": ms_host,
                                    "d2h_bytes_per_pass": int(tile_bytes),
                                    "note": "same, host-resident batches (the reference's `for batch in reader`): D2H of the "
                                            "Arrow buffers into pinned memory included"},
                "rows_per_pass": int(rr.item()),
            }
        except Exception as e:
            readers = {"error": repr(e)[:300]}

    # ---- dominant kernel = most device time among the kernels that move data (event pairs on the launching streams)
    top = max((k for k in kstats if k["alg_bytes"]), key=lambda k: k["ms"]) if kstats else None
    roofline = None
    if top:
        # Waves and launch groups overlap, so the event pair around a kernel also spans whatever ran beside it and the
        # pairs of one step add up to several times the step.  The kernel is charged its SHARE of the step:
        # step time x (its event time / sum of all kernels' event times) - the quantity the ncu launch list
        # (profiles/r02_launches_sf10_summary.txt, serialised) reports as well.
        ev_sum = sum(k["ms"] for k in kstats)
        share = top["ms"] / ev_sum
        attributed_ms = dev_ms * share
        ach = top["alg_bytes"] / (attributed_ms / 1e3) / 1e9
        traffic = None
        try:  # dram__bytes_read.sum + dram__bytes_write.sum of one SF10 launch, from the committed ncu launch list
            with open(os.path.join(ROOT, "profiles", "r02_traffic_sf10.json")) as f:
                traffic = json.load(f).get(top["name"])
        except Exception:
            pass
        groups_per_gpu = max(1, len(groups))
        roofline = {"bound": "hbm", "kernel": top["name"], "achieved": ach, "peak": peak, "unit": "GB/s",
                    "frac": ach / peak, "traffic": traffic, "traffic_unit": "bytes per SF10 launch group (ncu)",
                    "peak_source": peak_src,
                    "kernel_share_of_step": share, "kernel_ms": attributed_ms, "kernel_alg_bytes": top["alg_bytes"],
                    "kernel_alg_bytes_per_launch_group": top["alg_bytes"] / groups_per_gpu,
                    "kernel_event_ms_raw": top["ms"], "frac_raw_events": top["alg_bytes"] / (top["ms"] / 1e3) / 1e9 / peak,
                    "step_achieved": (in_bytes + out_bytes) / world / (dev_ms / 1e3) / 1e9,
                    "step_frac": (in_bytes + out_bytes) / world / (dev_ms / 1e3) / 1e9 / peak,
                    "aliased_output_bytes": aliased,
                    "step_frac_written_only": (in_bytes + out_bytes - aliased) / world / (dev_ms / 1e3) / 1e9 / peak,
                    "kernels": [{"name": k["name"], "ms": round(k["ms"], 4), "alg_gb": round(k["alg_bytes"] / 1e9, 4)}
                                for k in kstats]}
        if kserial:
            ks = {k["name"]: k for k in kserial}
            if top["name"] in ks and ks[top["name"]]["ms"] > 0:
                roofline["frac_alone"] = top["alg_bytes"] / (ks[top["name"]]["ms"] / 1e3) / 1e9 / peak
            roofline["kernels_alone"] = [{"name": k["name"], "ms": round(k["ms"], 4),
                                          "gbs": round(k["alg_bytes"] / max(k["ms"], 1e-9) / 1e6, 1)} for k in kserial]
            roofline["note"] = ("rank 0, summed over its launch groups.  kernels: CUDA event pairs inside the timed step, where waves and "
                                "groups overlap (a pair also spans what runs beside the kernel: frac_raw_events); frac: the dominant "
                                "kernel charged its share of the step (kernel_share_of_step x ms_per_step); frac_alone / kernels_alone: "
                                "one extra pass with one kernel at a time; step_frac: (stored stream bytes + Arrow bytes) per GPU / "
                                "step time / peak; aliased_output_bytes are Arrow bytes no kernel writes (direct-string values "
                                "alias the staged stream)")
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    cfgs = None
    if not args.no_configs and args.compression == "uncompressed" and world == 1:
        cfgs = other_configs(torch, ob, args, local_rank, peak, files)
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        r = cpu_run(files, 2, 1)
        cpu = {"value": r["arrow_bytes"] / r["seconds"] / 1e9, "unit": "GB/s", "cores": r["cores"], "kind": "port",
               "rows_per_s": r["rows"] / r["seconds"],
               "sample": f"one tile: {r['stripes']} stripes ({r['rows']} rows), one oracle reader per stripe over {r['cores']} "
                         "processes, 1 warm-up + 2 timed passes"}
        try:
            rp = cpu_run(files, 2, 1, fn=_pyarrow_decode_stripe)
            cpu["pyarrow"] = {"value": rp["arrow_bytes"] / rp["seconds"] / 1e9, "unit": "GB/s", "cores": rp["cores"],
                              "note": "Apache ORC C++ via pyarrow.orc read_stripe (not orc-rust), same stripes, same pool"}
        except Exception as e:
            cpu["pyarrow"] = {"error": repr(e)[:200]}
    line = {
        "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int64",
        "data": "synthetic", "rows_per_s": rows / (dev_ms / 1e3), "config": config,
        "e2e": {"value": e2e, "unit": "GB/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "h2d_gbs_achieved": h2d / (e2e_ms / 1e3) / 1e9,
                "platform_h2d_gbs": {"per_gpu_all_ranks_copying": round(h2d_min, 1), "aggregate": round(h2d_min * world, 1),
                                     "how": "torch pinned -> device copies of 1 GiB on every rank at once, slowest rank"},
                "note": "pinned H2D of all stripes + decode + D2H of per-batch metadata, launch groups pipelined; "
                        "decoded Arrow stays in HBM (north_star: output stays device-resident per rank).  The step is bound "
                        "by host -> device copies: compare h2d_gbs_achieved with platform_h2d_gbs.aggregate"},
        "gpu_launches": launches * args.steps,
        "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
        "run": {"stripes": n_stripes, "launch_groups_per_gpu": len(groups), "stripe_sharding": shard_note,
                "input_bytes": in_bytes, "arrow_bytes": out_bytes, "device_bytes": dev_bytes, "dataset_gen_s": round(gen_s, 1),
                "waves": args.waves, "group_streams": args.group_streams, "index_retries": ob.index_retries(), "layout_retries": ob.layout_retries()},
    }
    if readers:
        line.update(readers)
    if cfgs is not None:
        line["configs"] = cfgs
    _emit(real_stdout, line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
