#!/bin/bash
# GPU box: per-kernel times of every library variant under lib/variants/ (and the product build) on the given
# workloads; a quick tensor-path parity run per variant first.  usage: bash scripts/gpu_variants.sh <tag> [workloads...]
TAG=${1:-var}; shift
WL=${@:-cfg2}
mkdir -p gpurun_out
OUT=gpurun_out/variants_$TAG.log
: > $OUT
for lib in modulated-deform-conv_b200/lib/libmdconv_b200.so modulated-deform-conv_b200/lib/variants/*.so; do
  [ -f "$lib" ] || continue
  echo "#### $lib $(cat ${lib%.so}.defs 2>/dev/null)" | tee -a $OUT
  export MDCONV_B200_LIB=$PWD/$lib
  timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tensor_path_matches_oracle and f16" 2>&1 | tail -1 | tee -a $OUT
  for w in $WL; do
    echo "== $w" | tee -a $OUT
    timeout 120 python scripts/kernel_times.py $w 10 2>&1 | tail -16 | tee -a $OUT
  done
done
unset MDCONV_B200_LIB
