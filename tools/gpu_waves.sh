#!/bin/bash
# e2e of one launch group (what a rank of an 8-GPU run has) against the number of stripe waves
out=gpurun_out/wv; mkdir -p $out
for comp in snappy uncompressed; do
for w in 1 2 3 4 6; do
timeout 600 python bench.py --tiles 1 --steps 6 --warmup 3 --no-configs --no-readers --no-cpu-baseline --compression $comp --waves $w > $out/b_${comp}_$w.json 2> $out/b_${comp}_$w.err
python - "$out/b_${comp}_$w.json" $comp $w <<'PY'
import json,sys
b=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[2], 'waves', sys.argv[3], 'device ms %.2f  e2e ms %.2f  (%.1f GB/s, h2d %.1f GB/s)'%(b['ms_per_step'], b['e2e']['ms_per_step'], b['e2e']['value'], b['e2e']['h2d_gbs_achieved']))
PY
done
done
