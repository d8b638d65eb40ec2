#!/bin/bash
# usage: tools/gpu_profile_codec.sh <tag> <compression>: one ncu --set full capture of k_decompress on a 16M-row set
tag=${1:-pc}; comp=${2:-zstd}
out=gpurun_out/$tag
mkdir -p $out
B="--tiles 1 --steps 1 --warmup 1 --no-configs --no-readers --no-cpu-baseline --waves 1"
timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:${3:-k_decompress}" -s 1 -c 1 -o $out/decomp_$comp -f \
  python bench.py $B --rows 16000000 --files 8 --compression $comp > $out/k_$comp.log 2>&1
tail -3 $out/k_$comp.log
ls -la $out | head
