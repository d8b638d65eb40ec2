#!/bin/bash
# usage: tools/gpu_tests.sh <tag> [pytest -k expression]
tag=${1:-t}
out=gpurun_out/$tag
mkdir -p $out
if [ -n "$2" ]; then
  ( time timeout 900 python -m pytest tests -m gpu -x -q -k "$2" ) > $out/pytest.log 2>&1
else
  ( time timeout 1200 python -m pytest tests -m gpu -x -q ) > $out/pytest.log 2>&1
fi
echo "rc=$?" >> $out/pytest.log
tail -40 $out/pytest.log
