#!/bin/bash
# A/B: two-phase integer path (strings start after their own segments) on / off
out=gpurun_out/ab10; mkdir -p $out
bash tools/gpu_tests.sh ab10/t > /dev/null 2>&1; grep -E "passed|failed" $out/t/pytest.log
for v in 1 0 1 0; do
ORCB_SPLIT_INT=$v timeout 300 python bench.py --tiles 1 --steps 20 --warmup 5 --no-configs --no-readers --no-cpu-baseline > $out/b_$v.json 2> $out/b_$v.err
python - "$out/b_$v.json" $v <<'PY'
import json,sys
b=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=b['roofline']
print('split', sys.argv[2], 'SF10 ms %.3f frac %.4f'%(b['ms_per_step'],r['step_frac']))
PY
done
for v in 1 0; do
ORCB_SPLIT_INT=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-configs --no-readers --no-cpu-baseline > $out/b70_$v.json 2> $out/b70_$v.err
python - "$out/b70_$v.json" $v <<'PY'
import json,sys
b=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=b['roofline']
print('split', sys.argv[2], 'SF70 ms %.3f frac %.4f e2e %.1f'%(b['ms_per_step'],r['step_frac'],b['e2e']['value']))
PY
done
