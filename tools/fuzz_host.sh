#!/bin/bash
# Builds the library with AddressSanitizer + UBSan on the host side (device code unchanged) and runs tools/fuzz_host.cc
# over the reference fixtures and a corpus of compressed sections.  No GPU needed.
#   tools/fuzz_host.sh [iterations-per-seed=300] [seed=1]
set -e
cd "$(dirname "$0")/.."
ITERS=${1:-300}; SEED=${2:-1}
B=/tmp/orcb_fuzz; mkdir -p $B/obj $B/seeds
SRC=orc_rust_b200/csrc
SAN="-fsanitize=address,-fsanitize=undefined,-fno-sanitize-recover=undefined,-fno-omit-frame-pointer"
pids=()
for s in k_int.cu k_streams.cu k_strings.cu k_decompress.cu meta.cc tz.cc schema.cc plan.cc job.cc export.cc selection.cc predicate.cc c_api.cc; do
  if [ ! -f $B/obj/$s.o ] || [ $SRC/$s -nt $B/obj/$s.o ] || [ -n "$(find $SRC -name '*.h' -newer $B/obj/$s.o)" ]; then
    nvcc -gencode arch=compute_100a,code=sm_100a -O1 -g -std=c++17 -Xcompiler "-fPIC,$SAN" -x cu -c $SRC/$s -o $B/obj/$s.o &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
g++ -O1 -g -std=c++17 ${SAN//,/ } -c tools/fuzz_host.cc -o $B/obj/fuzz_host.o
nvcc -o $B/fuzz_host $B/obj/*.o -gencode arch=compute_100a,code=sm_100a -cudart static -Xlinker -lasan -Xlinker -lubsan
[ -n "$(ls $B/seeds/gen_*.orc 2>/dev/null)" ] || python tools/fuzz_host_seeds.py $B/seeds
FILES=$(find tests/golden -name '*.orc' -size -400k | sort)
ASAN_OPTIONS=detect_leaks=${LEAKS:-0}:allocator_may_return_null=1:max_allocation_size_mb=4096 UBSAN_OPTIONS=print_stacktrace=1 \
  $B/fuzz_host $ITERS $SEED $FILES $B/seeds/*.orc $B/seeds/*.sec
