#!/bin/bash
out=gpurun_out/san2; mkdir -p $out
( time timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -q -x -k 'zstd or lzo or inflate or (recompressed and (zstd or lzo)) or synthetic or feather' ) > $out/memcheck.log 2>&1
echo "rc=$?" >> $out/memcheck.log
grep -E "passed|failed|ERROR SUMMARY|rc=|Invalid|out of bounds" $out/memcheck.log | head -20
