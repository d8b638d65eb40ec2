#!/bin/bash
# A/B: resident CTAs per SM of k_decompress_bits
out=gpurun_out/ab9; mkdir -p $out
for v in 4 5 6; do
  ORCB_NVCC_DEFS="-DORCB_BITS_CTAS=$v" python -m orc_rust_b200.build --force > $out/build_$v.log 2>&1
  echo "== ORCB_BITS_CTAS=$v"
  bash tools/gpu_codec.sh ab9/v$v zstd lzo 2>&1 | grep -E "^(zstd|lzo) ms"
done
python -m orc_rust_b200.build --force > $out/build_final.log 2>&1
bash tools/gpu_codec.sh ab9/tile snappy lz4 2>&1 | grep -E "^(snappy|lz4) ms|k_decompress"
python - <<'PY'
import json
b=json.loads(open('gpurun_out/ab9/tile/b_snappy.json').read().strip().splitlines()[-1])
print({k['name'][:14]:round(k['ms'],2) for k in b['roofline']['kernels_alone']})
PY
