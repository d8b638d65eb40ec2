"""Device-resident decode time of the BASELINE configs other than the bench workload:
config 1 (1 M rows, one stripe), config 3 (lineitem Snappy / LZ4), config 4 (50 % nulls).  One JSON line per
case: rows, stored and Arrow bytes, ms per launch (CUDA events inside the library, kernels on one stream),
GB/s, per-kernel times, and - for cases the CPU oracle finishes in seconds - a byte-level parity check.

    python tools/bench_configs.py [--out-dir /tmp/orcb200_cfg]
"""
import argparse
import json
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gen_orc  # noqa: E402
import orc_rust_b200 as ob  # noqa: E402


def recompress_lz4(src: str, dst: str):
    """pyarrow's LZ4 writer stores every chunk as 'original'; this is only a marker that the case is what it is."""
    raise NotImplementedError


def run_case(name, paths, check_rows=0):
    job = ob.DecodeJob(paths, device=0)
    job.plan()
    job.stage()
    for _ in range(4):
        job.launch()
    job.finish()
    st = job.stats()
    ks = [k for k in job.kernel_stats() if k["ms"] > 0]
    ms = sum(k["ms"] for k in ks)
    line = {"case": name, "rows": st["n_rows"], "stripes": st["n_stripes"], "input_bytes": st["input_bytes"],
            "arrow_bytes": st["output_bytes"], "ms_sum_of_kernels": round(ms, 4),
            "arrow_gbs": round(st["output_bytes"] / ms / 1e6, 1), "alg_gbs": round((st["input_bytes"] + st["output_bytes"]) / ms / 1e6, 1),
            "kernels": {k["name"]: round(k["ms"], 4) for k in ks}}
    if check_rows and st["n_rows"] <= check_rows:
        from oracle import orc_oracle as oo
        from parity_util import assert_batches_identical
        exp = []
        for p in paths:
            exp += oo.OracleFile(open(p, "rb").read()).read()
        got = job.batches()
        assert_batches_identical(got, exp, name)
        line["parity"] = "byte-identical to the oracle (%d batches)" % len(got)
    print(json.dumps(line), flush=True)
    del job


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out-dir", default="/tmp/orcb200_cfg")
    ap.add_argument("--lineitem-rows", type=int, default=8_000_000)
    args = ap.parse_args()
    d = args.out_dir
    os.makedirs(d, exist_ok=True)
    os.environ.setdefault("ORCB_SERIAL", "1")
    p = gen_orc.write(gen_orc.config1_table(1_000_000, 0), os.path.join(d, "config1.orc"), stripe_size=1 << 30, dict_threshold=1.0)
    run_case("config1: 1M rows, one stripe, NONE", [p], check_rows=1_000_000)
    for comp in ("snappy", "lz4"):
        files = gen_orc.lineitem_dataset(os.path.join(d, "li_" + comp), args.lineitem_rows, 4, compression=comp, block_size=256 << 10)
        run_case("config3: lineitem %s 256 KiB chunks (pyarrow writer%s)" % (comp, "; every LZ4 chunk is stored 'original'" if comp == "lz4" else ""), files)
    t = gen_orc.nullheavy_table(2_000_000, 1)
    for comp in ("uncompressed", "snappy"):
        p = gen_orc.write(t, os.path.join(d, "nullheavy_%s.orc" % comp), compression=comp)
        run_case("config4: 2M rows, 50%% nulls, %s" % comp, [p], check_rows=2_000_000 if comp == "uncompressed" else 0)


if __name__ == "__main__":
    main()
