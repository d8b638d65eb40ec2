#!/bin/bash
# usage: tools/gpu_codec.sh <tag> <compression>...   one SF10 pass per compression kind, kernel times
tag=${1:-c}; shift
out=gpurun_out/$tag
mkdir -p $out
for comp in "$@"; do
( time timeout 900 python bench.py --tiles 1 --steps 5 --warmup 3 --no-configs --no-readers --no-cpu-baseline --compression $comp ) > $out/b_$comp.json 2> $out/b_$comp.err
tail -3 $out/b_$comp.err
python - "$out/b_$comp.json" $comp <<'PY'
import json,sys
b=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=b['roofline']
print(sys.argv[2], 'ms %.3f frac %.4f input GB %.3f  e2e %.1f GB/s'%(b['ms_per_step'],r['step_frac'],b['run']['input_bytes']/1e9,b['e2e']['value']), 'gen s', b['run'].get('dataset_gen_s'))
print({k['name'][:14]:round(k['ms'],3) for k in r['kernels_alone']})
PY
done
