"""Summarise an `ncu --metrics ... --csv --log-file` launch list: one line per kernel launch of the chosen step.

    python tools/ncu_launches.py gpurun_out/launches.csv [set_index]
A "set" is one pass of the decode pipeline (starts at k_rle_index or the first kernel name seen).
"""
import collections
import csv
import sys


def load(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ik, im, iv, iid = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    per = collections.OrderedDict()
    for r in rows[1:]:
        per.setdefault(r[iid], {"k": r[ik].split("(")[0]})[r[im]] = float(r[iv].replace(",", ""))
    return list(per.values())


def sets(launches):
    first = launches[0]["k"]
    out = []
    for d in launches:
        if d["k"] == first:
            out.append([])
        out[-1].append(d)
    return out


if __name__ == "__main__":
    ls = sets(load(sys.argv[1]))
    which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    print(f"# {len(ls)} passes in the file; pass {which}:")
    tot = 0
    for d in ls[which]:
        t = d.get("gpu__time_duration.sum", 0) / 1e6
        tot += t
        print(f"{d['k']:22s} {t:8.3f} ms  inst {d.get('smsp__inst_executed.sum', 0) / 1e6:9.1f} M  "
              f"dram rd {d.get('dram__bytes_read.sum', 0) / 1e6:9.1f} MB  wr {d.get('dram__bytes_write.sum', 0) / 1e6:9.1f} MB  "
              f"issue {d.get('smsp__issue_active.avg.pct_of_peak_sustained_active', 0):5.1f} %")
    print(f"total {tot:.3f} ms")
