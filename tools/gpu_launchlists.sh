#!/bin/bash
# ncu launch lists (per-launch time, DRAM bytes, instructions) of one SF10 pass per compression kind
tag=${1:-ll}
out=gpurun_out/$tag
mkdir -p $out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active
B="--tiles 1 --steps 1 --warmup 1 --no-configs --no-readers --no-cpu-baseline --waves 1"
for comp in uncompressed snappy lz4-lib zstd; do
  timeout 900 ncu --metrics $M --clock-control none --csv --log-file $out/launches_sf10_$comp.csv python bench.py $B --compression $comp > $out/l_$comp.log 2>&1
  python tools/ncu_launches.py $out/launches_sf10_$comp.csv > $out/launches_sf10_${comp}_summary.txt 2>&1
  tail -3 $out/launches_sf10_${comp}_summary.txt
done
