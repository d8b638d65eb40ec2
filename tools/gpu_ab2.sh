#!/bin/bash
tag=${1:-ab}
out=gpurun_out/$tag
mkdir -p $out
for lanes in 8 4 2 1; do for w in 1 2 4; do
  ORCB_IDX_LANES=$lanes timeout 300 python bench.py --tiles 1 --steps 10 --warmup 3 --no-configs --no-readers --no-cpu-baseline --waves $w > $out/b_l${lanes}_w$w.json 2> $out/b_l${lanes}_w$w.err
  python - "$out/b_l${lanes}_w$w.json" $lanes $w <<'PY'
import json,sys
try:
    b=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=b['roofline']
    al={k['name'][:12]:k['ms'] for k in r['kernels_alone']}
    print('lanes',sys.argv[2],'waves',sys.argv[3],'ms %.3f frac %.4f'%(b['ms_per_step'],r['step_frac']),'index alone',al.get('k_rle_index'))
except Exception as e: print('ERR',e)
PY
done; done
