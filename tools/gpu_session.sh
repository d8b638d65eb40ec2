#!/bin/bash
# One gpurun session: parity tests, the bench line, and A/B variants.  Outputs under gpurun_out/<tag>/.
tag=${1:-s1}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,memory.total --format=csv > $out/gpu.txt 2>&1
nvidia-smi topo -m >> $out/gpu.txt 2>&1
lscpu | grep -i -E "numa|model name|^cpu\(s\)|socket" >> $out/gpu.txt 2>&1
free -g >> $out/gpu.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > $out/pytest.log 2>&1
echo "pytest rc=$?" >> $out/pytest.log
( time timeout 1200 python bench.py --steps 10 --warmup 3 ) > $out/bench_default.json 2> $out/bench_default.err
for w in 1 2 3 4 6; do
  timeout 300 python bench.py --tiles 1 --steps 10 --warmup 3 --no-configs --no-readers --no-cpu-baseline --waves $w > $out/bench_t1_w$w.json 2> $out/bench_t1_w$w.err
done
ORCB_SORT_SEGS=0 timeout 300 python bench.py --tiles 1 --steps 10 --warmup 3 --no-configs --no-readers --no-cpu-baseline --waves 1 > $out/bench_t1_w1_nosort.json 2> $out/bench_t1_w1_nosort.err
for g in 1 3 4; do
  timeout 600 python bench.py --steps 5 --warmup 3 --no-configs --no-readers --no-cpu-baseline --group-streams $g > $out/bench_t7_g$g.json 2> $out/bench_t7_g$g.err
done
tail -c 600 $out/pytest.log
for f in $out/bench_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    b=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r=b.get('roofline') or {}
    print(' value %.1f GB/s  ms %.3f  step_frac %.4f  e2e %.1f' % (b['value'], b['ms_per_step'], r.get('step_frac',0), b['e2e']['value']))
except Exception as e: print(' ERR', e)
PY
done
