#!/bin/bash
# one ncu --set full capture (with source) of the headline's hot kernels on a 16M-row uncompressed set
tag=${1:-ph}
out=gpurun_out/$tag; mkdir -p $out
B="--tiles 1 --steps 1 --warmup 1 --no-configs --no-readers --no-cpu-baseline --waves 1"
ORCB_SPLIT_INT=0 timeout 1200 ncu --set full --clock-control none --import-source on -k 'regex:k_int_rle$|k_str_offsets|k_rle_index' -s 6 -c 3 -o $out/hot -f \
  python bench.py $B --rows 16000000 --files 8 > $out/k.log 2>&1
tail -2 $out/k.log | cut -c1-200
ls -la $out
