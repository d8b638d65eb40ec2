#!/bin/bash
# usage: tools/gpu_codec2.sh <tag> <pytest -k expr> <compression>...
tag=$1; expr=$2; shift; shift
bash tools/gpu_tests.sh $tag "$expr" | tail -6
bash tools/gpu_codec.sh $tag "$@"
