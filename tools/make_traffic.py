"""Per-kernel DRAM traffic of one full-size decode pass, from an ncu metrics launch list:

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum \
        --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 1
    python tools/make_traffic.py gpurun_out/launches.csv 1 > profiles/r01_traffic_sf10.json

Keys are the kernel-stat names `bench.py` reports (`roofline.kernels[].name`); values are
dram__bytes_read.sum + dram__bytes_write.sum of one launch (summed over the kernels a stat groups)."""
import json
import sys

from ncu_launches import load, sets

GROUPS = {
    "k_int_rle": "k_int_rle(+general,+coop_runs)",
    "k_int_rle_general": "k_int_rle(+general,+coop_runs)",
    "k_coop_runs": "k_int_rle(+general,+coop_runs)",
    "k_dict_prepare": "k_strings(5 kernels)",
    "k_str_tile_sum": "k_strings(5 kernels)",
    "k_str_tile_scan": "k_strings(5 kernels)",
    "k_str_offsets": "k_strings(5 kernels)",
    "k_utf8_bounds": "k_strings(5 kernels)",
}

if __name__ == "__main__":
    passes = sets(load(sys.argv[1]))
    which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    out = {}
    for d in passes[which]:
        name = GROUPS.get(d["k"], d["k"])
        out[name] = out.get(name, 0) + int(d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0))
    print(json.dumps(out, indent=1, sort_keys=True))
