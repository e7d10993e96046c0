#!/bin/bash
# GPU box: forward-only kernel time of the product library and of every variant under lib/variants/ (timing
# experiments; variants built with MDCONV_EXP_* compute wrong results on purpose).  usage: gpu_fwd_variants.sh [workload]
W=${1:-cfg2}
for lib in modulated-deform-conv_b200/lib/libmdconv_b200.so modulated-deform-conv_b200/lib/variants/*.so; do
  [ -f "$lib" ] || continue
  export MDCONV_B200_LIB=$PWD/$lib
  echo "## $(basename $lib) $(cat ${lib%.so}.defs 2>/dev/null)"
  python scripts/kernel_times.py $W 10 fwd_only 2>&1 | grep -E "tc_forward"
done
