#!/bin/bash
tag=${1:-lz}
out=gpurun_out/$tag
mkdir -p $out
( time timeout 600 python -m pytest tests -m gpu -x -q -k "decompress or recompressed or inflate" ) > $out/pytest_lz.log 2>&1
tail -3 $out/pytest_lz.log
for c in snappy snappy-recompressed lz4; do
  timeout 600 python bench.py --tiles 1 --steps 5 --warmup 3 --no-configs --no-readers --no-cpu-baseline --compression $c > $out/bench_$c.json 2> $out/bench_$c.err
  python - "$out/bench_$c.json" $c <<'PY'
import json,sys
try:
    b=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=b['roofline']
    al={k['name'][:12]:k['ms'] for k in r['kernels_alone']}
    print(sys.argv[2],'step ms %.2f'%b['ms_per_step'],'decompress alone %.2f'%al.get('k_decompress',0),'e2e %.1f GB/s'%b['e2e']['value'])
except Exception as e: print('ERR',e)
PY
done
