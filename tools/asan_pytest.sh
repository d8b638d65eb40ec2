#!/bin/bash
# The CPU test suite (-m "not gpu") against an AddressSanitizer + UBSan build of the library: every host path the Python
# mirror reaches (reader plans, selections, predicates, with_schema, zone tables, section decoders ...).  Shares the
# sanitizer objects of tools/fuzz_host.sh (run that first, or this builds them).  No GPU needed.
#   tools/asan_pytest.sh [pytest args]
set -e
cd "$(dirname "$0")/.."
B=/tmp/orcb_fuzz
tools/fuzz_host.sh 1 1 > /dev/null   # builds / refreshes the sanitizer objects
nvcc -shared -o $B/liborc_b200_asan.so $(ls $B/obj/*.o | grep -v fuzz_host) -gencode arch=compute_100a,code=sm_100a -cudart static
cat > $B/run_pytest.py <<PY
import sys
sys.path.insert(0, "$PWD")
import orc_rust_b200
orc_rust_b200._build.build = lambda *a, **k: "$B/liborc_b200_asan.so"
import pytest
sys.exit(pytest.main(["$PWD/tests", "-x", "-q", "-m", "not gpu", "-p", "no:cacheprovider", "-k", "not gloo and not c_program"] + sys.argv[1:]))
PY
G=$(dirname "$(gcc -print-file-name=libasan.so)")
LD_PRELOAD=$G/libasan.so:$G/libubsan.so ASAN_OPTIONS=detect_leaks=0:allocator_may_return_null=1 UBSAN_OPTIONS=print_stacktrace=1 \
  python $B/run_pytest.py "$@"
