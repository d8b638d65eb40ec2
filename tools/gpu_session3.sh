#!/bin/bash
tag=${1:-s3}
out=gpurun_out/$tag
mkdir -p $out
( time timeout 600 python -m pytest tests -m gpu -x -q -k "decompress or recompressed or fixture" ) > $out/pytest_lz.log 2>&1
echo "rc=$?" >> $out/pytest_lz.log
tail -8 $out/pytest_lz.log
for c in snappy snappy-recompressed lz4; do
  timeout 600 python bench.py --tiles 1 --steps 5 --warmup 3 --no-configs --no-readers --no-cpu-baseline --compression $c > $out/bench_$c.json 2> $out/bench_$c.err
  tail -2 $out/bench_$c.err
done
for f in $out/bench_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    b=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=b.get('roofline') or {}
    print(' value %.1f GB/s  ms %.3f  step_frac %.4f  e2e %.1f (%.1f ms)' % (b['value'], b['ms_per_step'], r.get('step_frac',0), b['e2e']['value'], b['e2e']['ms_per_step']))
    print('  in-step', {k['name'][:20]:k['ms'] for k in r['kernels']})
    print('  alone  ', {k['name'][:20]:k['ms'] for k in r['kernels_alone']})
except Exception as e: print(' ERR', e)
PY
done
# one ncu capture of the decompression kernel on a smaller set (4 files)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_decompress -s 2 -c 1 -o $out/decomp -f \
  python bench.py --tiles 1 --rows 7500000 --files 4 --steps 1 --warmup 3 --no-configs --no-readers --no-cpu-baseline --compression snappy-recompressed --waves 1 > $out/ncu.log 2>&1
tail -3 $out/ncu.log
