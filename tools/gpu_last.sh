#!/bin/bash
out=gpurun_out/last; mkdir -p $out
( timeout 600 python -m pytest tests -m gpu -x -q ) > $out/pytest.log 2>&1; grep -E "passed|failed" $out/pytest.log
for v in 8 4; do
ORCB_PEEK_RUNS=$v timeout 200 python bench.py --tiles 1 --steps 20 --warmup 5 --no-configs --no-cpu-baseline --reader-threads 4 > $out/b_$v.json 2> $out/b_$v.err
python - "$out/b_$v.json" $v <<'PY'
import json,sys
b=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=b['roofline']
print('peek', sys.argv[2], 'SF10 ms %.3f frac %.4f'%(b['ms_per_step'],r['step_frac']), {k:(round(b[k]['value'],1), round(b[k]['ms_per_pass'],1)) for k in ('e2e_reader','e2e_reader_host') if k in b})
PY
done
