#!/bin/bash
tag=${1:-q}
out=gpurun_out/$tag
mkdir -p $out
for w in 1 0; do
timeout 300 python bench.py --tiles 1 --steps 10 --warmup 3 --no-configs --no-readers --no-cpu-baseline --waves $w > $out/b_w$w.json 2> $out/b_w$w.err
python - "$out/b_w$w.json" <<'PY'
import json,sys
b=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=b['roofline']
print('ms %.3f frac %.4f kernel frac %.3f raw %.3f alone %.3f'%(b['ms_per_step'],r['step_frac'],r['frac'],r['frac_raw_events'],r.get('frac_alone',0)), b['run'])
print({k['name'][:14]:k['ms'] for k in r['kernels_alone']})
PY
done
