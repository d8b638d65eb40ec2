#!/bin/bash
tag=${1:-b}
out=gpurun_out/$tag
mkdir -p $out
( time timeout 1500 python bench.py --steps 10 --warmup 3 ) > $out/bench.json 2> $out/bench.err
tail -4 $out/bench.err
python - "$out/bench.json" <<'PY'
import json,sys
b=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r=b.get('roofline') or {}
print('value %.1f GB/s  ms %.3f  step_frac %.4f  e2e %.1f GB/s (%.1f ms) h2d %.1f of %s' % (b['value'], b['ms_per_step'], r.get('step_frac',0), b['e2e']['value'], b['e2e']['ms_per_step'], b['e2e'].get('h2d_gbs_achieved',0), b['e2e'].get('platform_h2d_gbs')))
print({k:(round(b[k]['value'],1), round(b[k]['ms_per_pass'],1)) for k in ('e2e_reader','e2e_reader_host') if k in b})
print('cpu', b.get('cpu_baseline'))
for c in b.get('configs') or []:
    print(' ', c.get('config','')[:70], c.get('ms'), c.get('value'), c.get('step_frac'), (c.get('parity') or {}).get('match'), c.get('error','')[:100], c.get('compressed_chunk_fraction'))
PY
