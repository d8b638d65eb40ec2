#!/bin/bash
tag=${1:-ab5}
out=gpurun_out/$tag
mkdir -p $out
for prio in 1 2 0; do for w in 1 2; do
  ORCB_AUX_PRIO=$prio timeout 300 python bench.py --tiles 1 --steps 10 --warmup 3 --no-configs --no-readers --no-cpu-baseline --waves $w > $out/b_p${prio}_w$w.json 2> $out/b_p${prio}_w$w.err
  python - "$out/b_p${prio}_w$w.json" $prio $w <<'PY'
import json,sys
try:
    b=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=b['roofline']
    print('prio',sys.argv[2],'waves',sys.argv[3],'ms %.3f frac %.4f'%(b['ms_per_step'],r['step_frac']), {k['name'][:12]:k['ms'] for k in r['kernels']})
except Exception as e: print('ERR',e)
PY
done; done
