#!/usr/bin/env python
"""bench.py — decoded Arrow GB/s (+ rows/s) of the ORC stripe decode path on TPC-H-lineitem-shaped ORC.

  python bench.py --gpus N --steps K --warmup W            # B200 CUDA path (this repo)
  python bench.py --impl reference --steps K --warmup W    # CPU path (oracle port of orc-rust's decoders)

Workload (BASELINE.json configs[1]): synthetic lineitem, SF10 rows (59 986 052) per GPU, compression NONE,
64 MiB stripes, dictionary strings, decimals, dates — written with pyarrow.orc from tools/gen_orc.py seeds.
A "step" = one full decode of every stripe of the rank's file set.
  value : whole-job decoded Arrow GB/s with the compressed stripes already resident in HBM
          (CUDA events on the launching stream, max over ranks)
  e2e   : same metric through the C-ABI with HOST buffers: pinned H2D of every stripe + decode +
          D2H of the per-batch metadata inside the timed region (decoded Arrow stays in HBM)
Scaling is weak: every rank decodes its own SF10-sized file set (stripes are independent; no collective).
"""
from __future__ import annotations

import argparse
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


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def _dataset(rows: int, n_files: int, compression: str):
    import gen_orc
    d = os.environ.get("ORCB_BENCH_DIR", "/tmp/orcb200_bench")
    t0 = time.time()
    files = gen_orc.lineitem_dataset(d, rows, n_files, compression=compression)
    return files, time.time() - t0


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
            self._stop.wait(0.1)

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
    rows = sum(b.num_rows for b in batches)
    nbytes = sum(b.nbytes for b in batches)
    return rows, nbytes


_cpu_decode_stripe.cache = {}


def _cpu_tasks(files):
    from oracle import orc_oracle as oo
    tasks = []
    for f in files:
        of = oo.OracleFile(open(f, "rb").read())
        tasks += [(f, i) for i in range(len(of.stripes))]
    return tasks


def cpu_run(files, steps, warmup, max_stripes=None):
    """Decodes every stripe (or a bounded sample) with one oracle reader per stripe on all host cores."""
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
            res = list(ex.map(_cpu_decode_stripe, tasks, chunksize=1))
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


def main():
    real_stdout = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=SF10_ROWS, help="rows per GPU (default: SF10)")
    ap.add_argument("--files", type=int, default=32, help="ORC files the rows are spread over")
    ap.add_argument("--compression", default="uncompressed", choices=["uncompressed", "snappy", "lz4"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-row-index", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    workload = (f"lineitem SF{args.rows / SF10_ROWS * 10:g} per GPU ({args.rows} rows, {args.files} ORC files, "
                f"64 MiB stripes, compression {args.compression}, RLEv2 + dictionary strings + decimal128(15,2) + date32)")

    if args.impl == "reference":
        if rank != 0:
            return
        files, gen_s = _dataset(args.rows, args.files, args.compression)
        r = cpu_run(files, args.steps, args.warmup)
        gbs = r["arrow_bytes"] / r["seconds"] / 1e9
        line = {
            "impl": "reference", "metric": METRIC, "value": gbs, "unit": "GB/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["seconds"] * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
            "rows_per_s": r["rows"] / r["seconds"],
            "config": {"workload": workload, "note": "CPU path; one reader per stripe on all host cores"},
            "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": r["cores"], "kind": "port",
                             "sample": f"all {r['stripes']} stripes per step (oracle port of orc-rust's decoders; "
                                       "the Rust reference cannot be built in this image)"},
            "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
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

    # a real (non-null) torch stream: the library enqueues on it, torch.cuda.Event times it
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    job = ob.DecodeJob(files, device=local_rank, cuda_stream=stream.cuda_stream, use_row_index=not args.no_row_index)
    job.plan()
    job.stage()     # allocates arenas, H2D of every stripe (untimed for `value`)
    job.launch()
    job.finish()    # surfaces decode errors before anything is timed
    st = job.stats()

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident timing
    for _ in range(args.warmup):
        job.launch()
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        e0.record(stream)
        for _ in range(args.steps):
            job.launch()
        e1.record(stream)
        sync_all()
    dev_ms = e0.elapsed_time(e1) / args.steps
    job.finish()
    kstats = job.kernel_stats()
    st = job.stats()
    # one more pass with every kernel on one stream (outside the timed region): per-kernel durations without the
    # contention of the overlapped schedule, reported beside the in-step ones
    kserial = None
    if "ORCB_SERIAL" not in os.environ:
        os.environ["ORCB_SERIAL"] = "1"
        try:
            job.launch()
            job.finish()
            kserial = job.kernel_stats()
        finally:
            del os.environ["ORCB_SERIAL"]

    # ---- end-to-end through the C ABI: H2D + decode + metadata D2H every step.  The file set is split into a
    #      few jobs, each on its own library stream, so the pinned H2D copy of group g+1 overlaps the decode of g.
    del job
    torch.cuda.synchronize()
    n_groups = min(4, len(files))
    groups = [files[i::n_groups] for i in range(n_groups)]
    jobs = [ob.DecodeJob(g, device=local_rank, use_row_index=not args.no_row_index) for g in groups]
    for j in jobs:
        j.plan(); j.stage(); j.launch(); j.finish()
    e2e_staged = sum(j.stats()["staged_bytes"] for j in jobs)
    e2e_meta = sum(j.stats()["d2h_meta_bytes"] for j in jobs)

    def e2e_step():
        for j in jobs:
            j.restage()
            j.launch()
        for j in jobs:
            j.finish()

    for _ in range(2):
        e2e_step()
    sync_all()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps

    t = torch.tensor([dev_ms, e2e_ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = t.tolist()

    out_bytes, in_bytes, rows = st["output_bytes"], st["input_bytes"], st["n_rows"]
    value = out_bytes * world / (dev_ms / 1e3) / 1e9
    e2e = out_bytes * world / (e2e_ms / 1e3) / 1e9
    peak, peak_src = _peaks()
    # dominant kernel = most device time among the kernels that move data (event pairs on the launching streams)
    top = max((k for k in kstats if k["alg_bytes"]), key=lambda k: k["ms"]) if kstats else None
    roofline = None
    if top:
        ach = top["alg_bytes"] / (top["ms"] / 1e3) / 1e9
        traffic = None
        try:  # dram__bytes_read.sum + dram__bytes_write.sum per launch, from the committed ncu capture of this workload
            with open(os.path.join(ROOT, "profiles", "r01_traffic_sf10.json")) as f:
                traffic = json.load(f).get(top["name"])
        except Exception:
            pass
        roofline = {"bound": "hbm", "kernel": top["name"], "achieved": ach, "peak": peak, "unit": "GB/s",
                    "frac": ach / peak, "traffic": traffic, "peak_source": peak_src,
                    "kernel_ms": top["ms"], "kernel_alg_bytes": top["alg_bytes"],
                    "step_achieved": (in_bytes + out_bytes) / (dev_ms / 1e3) / 1e9,
                    "step_frac": (in_bytes + out_bytes) / (dev_ms / 1e3) / 1e9 / peak,
                    "kernels": [{"name": k["name"], "ms": round(k["ms"], 4), "alg_gb": round(k["alg_bytes"] / 1e9, 4)}
                                for k in kstats]}
        if kserial:
            ks = {k["name"]: k for k in kserial}
            if top["name"] in ks and ks[top["name"]]["ms"] > 0:
                roofline["frac_alone"] = top["alg_bytes"] / (ks[top["name"]]["ms"] / 1e3) / 1e9 / peak
            roofline["kernels_alone"] = [{"name": k["name"], "ms": round(k["ms"], 4),
                                          "gbs": round(k["alg_bytes"] / max(k["ms"], 1e-9) / 1e6, 1)} for k in kserial]
            roofline["note"] = ("frac / kernels: event pairs inside the timed, two-stream step (kernels of the other stream "
                                "share the SMs); frac_alone / kernels_alone: one extra pass with all kernels on one stream")
    if rank != 0:
        return
    cpu = None
    if not args.no_cpu_baseline:
        r = cpu_run(files, 1, 0, max_stripes=max(16, 2 * (os.cpu_count() or 1)))
        cpu = {"value": r["arrow_bytes"] / r["seconds"] / 1e9, "unit": "GB/s", "cores": r["cores"], "kind": "port",
               "rows_per_s": r["rows"] / r["seconds"],
               "sample": f"{r['stripes']} stripes ({r['rows']} rows) of the same file set, one oracle reader per stripe "
                         f"over {r['cores']} processes"}
    line = {
        "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64",
        "data": "synthetic", "rows_per_s": rows * world / (dev_ms / 1e3),
        "config": {"workload": workload, "batch_size": 8192, "stripes_per_gpu": st["n_stripes"],
                   "segments_per_gpu": st["n_segments"], "input_bytes_per_gpu": in_bytes,
                   "arrow_bytes_per_gpu": out_bytes, "l2_policy": "inputs+outputs per step (>=13 GB) far exceed the 126 MB L2",
                   "row_index": not args.no_row_index, "dataset_gen_s": round(gen_s, 1)},
        "e2e": {"value": e2e, "unit": "GB/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": e2e_staged,
                "d2h_bytes_per_step": e2e_meta,
                "note": "pinned H2D of all stripes + decode + D2H of per-batch metadata (4 pipelined jobs); "
                        "decoded Arrow stays in HBM"},
        "gpu_launches": st["n_kernel_launches"] * args.steps,
        "roofline": roofline, "cpu_baseline": cpu, "clocks": clk.summary(),
    }
    _emit(real_stdout, line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
