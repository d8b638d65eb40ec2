"""Synthetic ORC inputs for the BASELINE.json configs, written with pyarrow.orc
(the Apache ORC C++ writer).  Deterministic: numpy default_rng(seed).

  config1(n)            1-stripe file: int64 DELTA-ish, int64 DIRECT-24, dictionary string
  lineitem(n, seed)     TPC-H lineitem shape of the reference's scripts/convert_tpch.py:46-63
  nullheavy(n, seed)    50 % nulls: PATCHED_BASE ints, timestamps with nanos, decimal128(38,10), bool, i8, f64
"""
from __future__ import annotations

import os

import numpy as np
import pyarrow as pa
import pyarrow.orc as po

_WORDS = ("furiously carefully quickly blithely slyly regular express special pending ironic final bold unusual even "
          "silent requests deposits packages accounts instructions theodolites dependencies foxes pinto beans ideas "
          "platelets asymptotes courts dolphins excuses frays sleep wake haggle nag cajole boost detect "
          "integrate").split()


def config1_table(n: int = 1_000_000, seed: int = 0) -> pa.Table:
    rng = np.random.default_rng(seed)
    a = np.cumsum(rng.integers(1, 4, n, dtype=np.int64))
    b = rng.integers(0, 1 << 24, n, dtype=np.int64)
    vocab = np.array(["v%02d_%s" % (i, "x" * (3 + i % 9)) for i in range(100)])
    s = vocab[rng.integers(0, 100, n)]
    return pa.table({"a": a, "b": b, "s": pa.array(s, type=pa.utf8())})


def lineitem_table(n_orders: int, seed: int = 0, sf: float = 10.0) -> pa.Table:
    """About 4 rows per order.  Decimal columns are decimal128(15,2) as in convert_tpch.py."""
    rng = np.random.default_rng(seed)
    lines = rng.integers(1, 8, n_orders)
    n = int(lines.sum())
    # dbgen-style sparse order keys: 8 used out of every 32
    idx = np.arange(n_orders, dtype=np.int64) + (seed * n_orders)
    okeys = (idx // 8) * 32 + (idx % 8) + 1
    orderkey = np.repeat(okeys, lines)
    starts = np.cumsum(lines) - lines
    linenumber = (np.arange(n) - np.repeat(starts, lines) + 1).astype(np.int32)
    partkey = rng.integers(1, int(200_000 * sf) + 1, n, dtype=np.int64)
    suppkey = rng.integers(1, int(10_000 * sf) + 1, n, dtype=np.int64)

    def dec(unscaled):
        arr = pa.array(unscaled.astype(np.int64))
        # exact decimal construction from unscaled ints: (int64 -> decimal(15,0)) then rescale via buffers
        lo = unscaled.astype(np.int64)
        buf = np.empty((n, 2), dtype=np.int64)
        buf[:, 0] = lo
        buf[:, 1] = lo >> 63
        return pa.Array.from_buffers(pa.decimal128(15, 2), n, [None, pa.py_buffer(buf.tobytes())])

    quantity = dec(rng.integers(1, 51, n) * 100)
    extendedprice = dec(rng.integers(90_000, 10_500_001, n))
    discount = dec(rng.integers(0, 11, n))
    tax = dec(rng.integers(0, 9, n))
    orderdate = np.repeat(rng.integers(8036, 10441, n_orders), lines)
    shipdate = (orderdate + rng.integers(1, 122, n)).astype(np.int32)
    commitdate = (orderdate + rng.integers(30, 91, n)).astype(np.int32)
    receiptdate = (shipdate + rng.integers(1, 31, n)).astype(np.int32)
    cur = 9298  # 1995-06-17
    def take(vocab, idx):
        return pa.array(vocab, pa.utf8()).take(pa.array(idx.astype(np.int32)))

    rf = take(["R", "A", "N"], np.where(receiptdate <= cur, (rng.random(n) < 0.5).astype(np.int32), 2))
    ls = take(["O", "F"], (shipdate <= cur).astype(np.int32))
    instr = take(["DELIVER IN PERSON", "COLLECT COD", "NONE", "TAKE BACK RETURN"], rng.integers(0, 4, n))
    mode = take(["REG AIR", "AIR", "RAIL", "SHIP", "TRUCK", "MAIL", "FOB"], rng.integers(0, 7, n))
    # comments: 2-6 words from a 40-word vocabulary, assembled directly as Arrow offsets + bytes
    nw = rng.integers(2, 7, n)
    table = (" ".join(_WORDS) + " ").encode()
    wl = np.array([len(x) for x in _WORDS], dtype=np.int64)
    wstart = np.concatenate([[0], np.cumsum(wl + 1)[:-1]])
    total_words = int(nw.sum())
    wid = rng.integers(0, len(_WORDS), total_words)
    last = np.zeros(total_words, dtype=bool)
    last[np.cumsum(nw) - 1] = True
    plen = wl[wid] + 1 - last  # word + trailing space, no space after the last word of a row
    pstart = np.cumsum(plen) - plen
    nchar = int(plen.sum())
    src = np.repeat(wstart[wid] - pstart, plen) + np.arange(nchar, dtype=np.int64)
    cdata = np.frombuffer(table, dtype=np.uint8)[src]
    row_len = np.add.reduceat(plen, np.cumsum(nw) - nw)
    coffs = np.zeros(n + 1, dtype=np.int32)
    np.cumsum(row_len, out=coffs[1:])
    comment = pa.Array.from_buffers(pa.utf8(), n, [None, pa.py_buffer(coffs.tobytes()), pa.py_buffer(cdata.tobytes())])
    return pa.table({
        "l_orderkey": orderkey, "l_partkey": partkey, "l_suppkey": suppkey, "l_linenumber": linenumber,
        "l_quantity": quantity, "l_extendedprice": extendedprice, "l_discount": discount, "l_tax": tax,
        "l_returnflag": rf, "l_linestatus": ls,
        "l_shipdate": pa.array(shipdate, pa.date32()), "l_commitdate": pa.array(commitdate, pa.date32()),
        "l_receiptdate": pa.array(receiptdate, pa.date32()),
        "l_shipinstruct": instr, "l_shipmode": mode,
        "l_comment": comment,
    })


def nullheavy_table(n: int = 500_000, seed: int = 1) -> pa.Table:
    rng = np.random.default_rng(seed)

    def mask():
        return rng.random(n) < 0.5

    pb = rng.integers(0, 1000, n, dtype=np.int64)
    out = rng.random(n) < 0.03
    pb = np.where(out, pb + rng.integers(1 << 30, 1 << 40, n, dtype=np.int64), pb)
    secs = rng.integers(0, 4_000_000_000, n, dtype=np.int64)  # pyarrow writes unusable nanos for pre-epoch sub-second values
    ns = rng.integers(0, 1_000_000_000, n, dtype=np.int64)
    ts = secs * 1_000_000_000 + ns
    ts_ms = secs * 1_000_000_000 + (ns // 1_000_000) * 1_000_000
    dec_lo = rng.integers(-10**17, 10**17, n, dtype=np.int64)
    dbuf = np.empty((n, 2), dtype=np.int64)
    dbuf[:, 0] = dec_lo
    dbuf[:, 1] = dec_lo >> 63
    vocab = np.array(["alpha", "beta", "gamma", "delta", "epsilon"])
    cols = {
        "pb": pa.array(pb, mask=mask()),
        "ts": pa.array(ts, pa.timestamp("ns"), mask=mask()),
        "ts_ms": pa.array(ts_ms, pa.timestamp("ns"), mask=mask()),
        "b": pa.array(rng.random(n) < 0.3, mask=mask()),
        "i8": pa.array(rng.integers(-128, 128, n).astype(np.int8), mask=mask()),
        "f64": pa.array(rng.standard_normal(n), mask=mask()),
        "f32": pa.array(rng.standard_normal(n).astype(np.float32), mask=mask()),
        "i16": pa.array(rng.integers(-30000, 30000, n).astype(np.int16), mask=mask()),
        "s": pa.array(vocab[rng.integers(0, 5, n)], pa.utf8(), mask=mask()),
        "sd": pa.array(np.char.add("row-", rng.integers(0, 10**9, n).astype(str)), pa.utf8(), mask=mask()),
    }
    dec = pa.Array.from_buffers(pa.decimal128(38, 10), n,
                                [pa.py_buffer(np.packbits(~mask(), bitorder="little").tobytes()),
                                 pa.py_buffer(dbuf.tobytes())])
    cols["dec"] = dec
    return pa.table(cols)


def write(table: pa.Table, path: str, compression: str = "uncompressed", stripe_size: int = 64 << 20,
          block_size: int = 256 << 10, dict_threshold: float = 0.8, row_index_stride: int = 10000):
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    po.write_table(table, path, compression=compression, stripe_size=stripe_size,
                   compression_block_size=block_size, dictionary_key_size_threshold=dict_threshold,
                   row_index_stride=row_index_stride)
    return path


def _lineitem_file(args):
    path, n_orders, seed, compression, block_size = args
    if not os.path.exists(path):
        t = lineitem_table(n_orders, seed)
        tmp = path + ".tmp%d" % os.getpid()
        write(t, tmp, compression=compression, block_size=block_size)
        os.replace(tmp, path)
    return path


def lineitem_dataset(out_dir: str, total_rows: int, n_files: int, compression: str = "uncompressed",
                     block_size: int = 256 << 10, procs: int = 0):
    """Writes `n_files` lineitem ORC files (seeds 0..n_files-1, ~total_rows rows in all) in parallel.
    Existing files are reused, so back-to-back bench arms share one generation."""
    import concurrent.futures as cf
    import multiprocessing as mp
    os.makedirs(out_dir, exist_ok=True)
    n_orders = max(1, total_rows // n_files // 4)
    jobs = [(os.path.join(out_dir, f"lineitem_{compression}_{n_orders}_{i:03d}.orc"), n_orders, i, compression,
             block_size) for i in range(n_files)]
    procs = procs or min(n_files, os.cpu_count() or 1)
    if procs <= 1 or all(os.path.exists(j[0]) for j in jobs):
        return [_lineitem_file(j) for j in jobs]
    with cf.ProcessPoolExecutor(procs, mp_context=mp.get_context("fork")) as ex:
        return list(ex.map(_lineitem_file, jobs))


if __name__ == "__main__":
    import argparse
    import time

    ap = argparse.ArgumentParser()
    ap.add_argument("kind", choices=["config1", "lineitem", "nullheavy"])
    ap.add_argument("out")
    ap.add_argument("--rows", type=int, default=1_000_000)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--compression", default="uncompressed")
    a = ap.parse_args()
    t0 = time.time()
    if a.kind == "config1":
        t = config1_table(a.rows, a.seed)
    elif a.kind == "lineitem":
        t = lineitem_table(a.rows // 4, a.seed)
    else:
        t = nullheavy_table(a.rows, a.seed)
    t1 = time.time()
    write(t, a.out, a.compression, stripe_size=(1 << 30) if a.kind == "config1" else (64 << 20))
    print(f"{a.kind}: {t.num_rows} rows, gen {t1 - t0:.1f}s, write {time.time() - t1:.1f}s, "
          f"{os.path.getsize(a.out) / 1e6:.1f} MB")
