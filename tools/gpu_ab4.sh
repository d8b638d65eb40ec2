#!/bin/bash
tag=${1:-ab4}
out=gpurun_out/$tag
mkdir -p $out
for order in kcv kvc vkc ckv vck; do
  ORCB_MAIN_ORDER=$order timeout 300 python bench.py --tiles 1 --steps 10 --warmup 3 --no-configs --no-readers --no-cpu-baseline --waves 1 > $out/b_$order.json 2> $out/b_$order.err
  python - "$out/b_$order.json" $order <<'PY'
import json,sys
try:
    b=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=b['roofline']
    print('order',sys.argv[2],'ms %.3f frac %.4f'%(b['ms_per_step'],r['step_frac']))
except Exception as e: print('ERR',e)
PY
done
for gs in 1 2 3; do for w in 1 2; do
  timeout 600 python bench.py --steps 5 --warmup 3 --no-configs --no-readers --no-cpu-baseline --waves $w --group-streams $gs > $out/t7_g${gs}_w$w.json 2> $out/t7_g${gs}_w$w.err
  python - "$out/t7_g${gs}_w$w.json" $gs $w <<'PY'
import json,sys
try:
    b=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=b['roofline']
    print('SF70 group-streams',sys.argv[2],'waves',sys.argv[3],'ms %.3f frac %.4f'%(b['ms_per_step'],r['step_frac']))
except Exception as e: print('ERR',e)
PY
done; done
