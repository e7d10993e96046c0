#!/bin/bash
# GPU box: per-kernel times (fwd+bwd) of the product library and of every variant under lib/variants/.
# usage: gpu_bwd_variants.sh [workload] [grep-pattern]
W=${1:-cfg2}; PAT=${2:-tc_grad_sample|col2im|tc_grad_input|sum}
for lib in modulated-deform-conv_b200/lib/libmdconv_b200.so modulated-deform-conv_b200/lib/variants/*.so; do
  [ -f "$lib" ] || continue
  export MDCONV_B200_LIB=$PWD/$lib
  echo "## $(basename $lib) $(cat ${lib%.so}.defs 2>/dev/null)"
  python scripts/kernel_times.py $W 10 2>&1 | grep -E "$PAT"
done
