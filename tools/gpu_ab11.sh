#!/bin/bash
# A/B: lanes per warp in the header walk under the two-phase schedule
out=gpurun_out/ab11; mkdir -p $out
for v in 8 16 8 16 32; do
ORCB_IDX_LANES=$v timeout 300 python bench.py --tiles 1 --steps 20 --warmup 5 --no-configs --no-readers --no-cpu-baseline > $out/b_$v.json 2> $out/b_$v.err
python - "$out/b_$v.json" $v <<'PY'
import json,sys
b=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=b['roofline']
print('lanes', sys.argv[2], 'SF10 ms %.3f frac %.4f'%(b['ms_per_step'],r['step_frac']))
PY
done
for v in 8 16; do
ORCB_IDX_LANES=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-configs --no-readers --no-cpu-baseline > $out/b70_$v.json 2> $out/b70_$v.err
python - "$out/b70_$v.json" $v <<'PY'
import json,sys
b=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=b['roofline']
print('lanes', sys.argv[2], 'SF70 ms %.3f frac %.4f'%(b['ms_per_step'],r['step_frac']))
PY
done
