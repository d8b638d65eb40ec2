#!/bin/bash
out=gpurun_out/pc4; mkdir -p $out
python tools/cfg4_probe.py > $out/probe.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_decompress' -s 2 -c 1 -o $out/cfg4_snappy -f python tools/cfg4_probe.py > $out/ncu.log 2>&1
tail -3 $out/ncu.log
