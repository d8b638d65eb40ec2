#!/bin/bash
tag=${1:-ab6}
out=gpurun_out/$tag
mkdir -p $out
for bg in 1 0; do for w in 1 2; do
  ORCB_BG=$bg timeout 300 python bench.py --tiles 1 --steps 10 --warmup 3 --no-configs --no-readers --no-cpu-baseline --waves $w > $out/b_bg${bg}_w$w.json 2> $out/b_bg${bg}_w$w.err
  python - "$out/b_bg${bg}_w$w.json" $bg $w <<'PY'
import json,sys
try:
    b=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=b['roofline']
    print('bg',sys.argv[2],'waves',sys.argv[3],'ms %.3f frac %.4f'%(b['ms_per_step'],r['step_frac']), {k['name'][:12]:k['ms'] for k in r['kernels']})
except Exception as e: print('ERR',e)
PY
done; done
bash tools/gpu_tests.sh ${tag}t "fixture_files or synthetic or nested or builder"
