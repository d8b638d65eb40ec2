#!/bin/bash
tag=${1:-f1}
out=gpurun_out/$tag
mkdir -p $out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > $out/pytest.log 2>&1; echo "rc=$?" >> $out/pytest.log; tail -4 $out/pytest.log
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > $out/ref.json 2> $out/ref.err; tail -3 $out/ref.err
( time timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 ) > $out/bench.json 2> $out/bench.err; tail -4 $out/bench.err
python - "$out/bench.json" "$out/ref.json" <<'PY'
import json,sys
b=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
rf=json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
r=b.get('roofline') or {}
print('ref value %.2f GB/s (%.0f ms/step)'%(rf['value'],rf['ms_per_step']), 'same config:', rf['config']==b['config'])
print('value %.1f GB/s  ms %.3f  step_frac %.4f kernel frac %.3f (%s) traffic %s  e2e %.1f GB/s (%.1f ms) h2d %.1f of %s' % (b['value'], b['ms_per_step'], r.get('step_frac',0), r.get('frac',0), r.get('kernel'), r.get('traffic'), b['e2e']['value'], b['e2e']['ms_per_step'], b['e2e'].get('h2d_gbs_achieved',0), b['e2e'].get('platform_h2d_gbs',{}).get('aggregate')))
print({k:(round(b[k]['value'],1), round(b[k]['ms_per_pass'],1)) for k in ('e2e_reader','e2e_reader_host') if k in b}, b['run'].get('index_retries'))
print('cpu', b.get('cpu_baseline'))
for c in b.get('configs') or []:
    print(' ', c.get('config','')[:70], c.get('ms'), c.get('value'), c.get('step_frac'), (c.get('parity') or {}).get('match'), c.get('error','')[:100])
PY
