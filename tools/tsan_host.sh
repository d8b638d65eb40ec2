#!/bin/bash
# Builds the library with ThreadSanitizer on the host side and runs tools/tsan_host.cc.  No GPU needed.
set -e
cd "$(dirname "$0")/.."
B=/tmp/orcb_tsan; mkdir -p $B/obj
SRC=orc_rust_b200/csrc
SAN="-fsanitize=thread,-fno-omit-frame-pointer"
pids=()
for s in k_int.cu k_streams.cu k_strings.cu k_decompress.cu meta.cc tz.cc schema.cc plan.cc job.cc export.cc selection.cc predicate.cc c_api.cc; do
  if [ ! -f $B/obj/$s.o ] || [ $SRC/$s -nt $B/obj/$s.o ] || [ -n "$(find $SRC -name '*.h' -newer $B/obj/$s.o)" ]; then
    nvcc -gencode arch=compute_100a,code=sm_100a -O1 -g -std=c++17 -Xcompiler "-fPIC,$SAN" -x cu -c $SRC/$s -o $B/obj/$s.o &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
g++ -O1 -g -std=c++17 ${SAN//,/ } -c tools/tsan_host.cc -o $B/obj/tsan_host.o
nvcc -o $B/tsan_host $B/obj/*.o -gencode arch=compute_100a,code=sm_100a -cudart static -Xlinker -ltsan
BIG=$B/big.orc
[ -f $BIG ] || python - <<PY
import sys
sys.path.insert(0, "tools")
import gen_orc
gen_orc.write(gen_orc.lineitem_table(150_000, 5), "$BIG", stripe_size=8 << 20)
PY
TSAN_OPTIONS="halt_on_error=0 second_deadlock_stack=1" $B/tsan_host $BIG tests/golden/ref_integration/TestOrcFile.testSnappy.orc
