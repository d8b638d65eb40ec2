#!/bin/bash
tag=${1:-ab7}
out=gpurun_out/$tag
mkdir -p $out
for cfg in "96 3 0" "192 2 0" "192 2 4" "288 2 0" "384 2 4" "672 1 4" "672 1 8"; do
  set -- $cfg
  timeout 600 python bench.py --steps 5 --warmup 3 --no-configs --no-readers --no-cpu-baseline --group-stripes $1 --group-streams $2 --waves $3 > $out/t7_s$1_g$2_w$3.json 2> $out/t7_s$1_g$2_w$3.err
  python - "$out/t7_s$1_g$2_w$3.json" "$cfg" <<'PY'
import json,sys
try:
    b=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=b['roofline']
    print('SF70 group-stripes/streams/waves',sys.argv[2],'ms %.3f frac %.4f e2e %.1f'%(b['ms_per_step'],r['step_frac'],b['e2e']['value']))
except Exception as e: print('ERR',sys.argv[2],e)
PY
done
