#!/bin/bash
# A/B: resident CTAs per SM of the inflate / LZO kernels
out=gpurun_out/ab12; mkdir -p $out
for v in 10 12 16; do
  ORCB_NVCC_DEFS="-DORCB_INF_CTAS=$v" python -m orc_rust_b200.build --force > $out/build_$v.log 2>&1
  echo "== ORCB_INF_CTAS=$v"
  bash tools/gpu_codec.sh ab12/v$v zlib lzo 2>&1 | grep -E "^(zlib|lzo) ms"
done
python -m orc_rust_b200.build --force > $out/build_final.log 2>&1
