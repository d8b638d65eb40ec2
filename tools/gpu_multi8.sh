#!/bin/bash
tag=$1; N=$2
out=gpurun_out/$tag
mkdir -p $out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/h2d_probe.py > $out/h2d_probe.json 2> $out/h2d_probe.err
cat $out/h2d_probe.json
bash tools/gpu_multi.sh $tag $N
