#!/bin/bash
# usage: tools/gpu_multi.sh <tag> <N>
tag=$1; N=$2
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi topo -m > $out/topo.txt 2>&1
lscpu | grep -i -E "numa|model name|^cpu\(s\)|socket" >> $out/topo.txt 2>&1
for d in /sys/bus/pci/devices/*; do if [ -f $d/class ] && grep -q "^0x0302" $d/class; then echo "$d numa $(cat $d/numa_node) $(cat $d/current_link_speed 2>/dev/null) x$(cat $d/current_link_width 2>/dev/null)"; fi; done >> $out/topo.txt 2>&1
free -g >> $out/topo.txt
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps ${STEPS:-20} --warmup ${WARMUP:-5} $EXTRA ) > $out/bench_n$N.json 2> $out/bench_n$N.err
tail -5 $out/bench_n$N.err
python - "$out/bench_n$N.json" <<'PY'
import json,sys
b=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r=b.get('roofline') or {}
print('N',b['n_gpus'],'value %.1f GB/s  ms %.3f  step_frac %.4f  e2e %.1f GB/s (%.1f ms)' % (b['value'], b['ms_per_step'], r.get('step_frac',0), b['e2e']['value'], b['e2e']['ms_per_step']))
print(b.get('run')); print({k:b[k] for k in ('e2e_reader','e2e_reader_host') if k in b})
PY
