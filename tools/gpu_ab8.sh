#!/bin/bash
tag=${1:-ab8}
out=gpurun_out/$tag
mkdir -p $out
for lanes in 8 16 32; do
  ORCB_IDX_LANES=$lanes timeout 600 python bench.py --steps 5 --warmup 3 --no-configs --no-readers --no-cpu-baseline > $out/t7_l$lanes.json 2> $out/t7_l$lanes.err
  python - "$out/t7_l$lanes.json" "$lanes" <<'PY'
import json,sys
try:
    b=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=b['roofline']
    print('SF70 idx lanes',sys.argv[2],'ms %.3f frac %.4f'%(b['ms_per_step'],r['step_frac']))
except Exception as e: print('ERR',sys.argv[2],e)
PY
done
