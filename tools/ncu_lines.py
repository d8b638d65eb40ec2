"""Per-source-line instruction counts of one kernel from an .ncu-rep captured with --import-source on.

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep '^k_int_rle$' [top_n]
"""
import csv, subprocess, sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda", "--kernel-name",
                      f"regex:{kern}"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
lines = {}
cur = None
for r in rows:
    if r and r[0] == "Line No":
        hdr = r
        i_inst = hdr.index("Instructions Executed")
        i_samp = hdr.index("# Samples")
        continue
    if hdr is None or len(r) < len(hdr) - 5:
        continue
    if r[0] != "":
        cur = (int(r[0]), r[1].strip())
        lines.setdefault(cur, [0, 0])
        continue
    if cur is None or r[2] in ("...", "-"):
        continue
    try:
        lines[cur][0] += int(r[i_inst]); lines[cur][1] += int(r[i_samp])
    except ValueError:
        pass
tot = sum(v[0] for v in lines.values()) or 1
tots = sum(v[1] for v in lines.values()) or 1
print(f"total warp instructions {tot}, samples {tots}")
for (ln, src), (n, s) in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{ln:5d} {100*n/tot:5.1f}% inst {100*s/tots:5.1f}% samp  {src[:110]}")
