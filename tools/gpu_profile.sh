#!/bin/bash
tag=${1:-prof}
out=gpurun_out/$tag
mkdir -p $out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active
B="--tiles 1 --steps 1 --warmup 1 --no-configs --no-readers --no-cpu-baseline --waves 1"
timeout 900 ncu --metrics $M --clock-control none --csv --log-file $out/launches_sf10.csv python bench.py $B > $out/l1.log 2>&1
timeout 900 ncu --metrics $M --clock-control none --csv --log-file $out/launches_sf10_snappy.csv python bench.py $B --compression snappy > $out/l2.log 2>&1
timeout 900 ncu --metrics $M --clock-control none --csv --log-file $out/launches_sf10_lz4.csv python bench.py $B --compression lz4 > $out/l3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_int_rle$|k_rle_index|k_str_offsets|k_varint128|k_int_rle_coop|k_coop_runs|k_str_tile_sum|k_utf8$' -s 14 -c 9 -o $out/kernels -f \
  python bench.py $B --rows 16000000 --files 8 > $out/k.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_decompress' -s 1 -c 1 -o $out/decomp_snappy -f \
  python bench.py $B --rows 16000000 --files 8 --compression snappy > $out/k2.log 2>&1
ls -la $out | head -20
