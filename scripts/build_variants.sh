#!/bin/bash
# Builds tuning variants of the library next to the product build (CPU box; ~1 min each):
#   modulated-deform-conv_b200/lib/variants/libmdconv_b200_<name>.so
# usage: bash scripts/build_variants.sh name1="-DFOO=1 -DBAR=2" name2="..."
# scripts/gpu_variants.sh then times every variant found there on the GPU box.
set -e
cd "$(dirname "$0")/../modulated-deform-conv_b200/csrc"
for spec in "$@"; do
  name="${spec%%=*}"; defs="${spec#*=}"
  echo "== variant $name: $defs"
  make -j8 VARIANT=_$name DEFS="$defs" > /dev/null
  echo "$defs" > ../lib/variants/libmdconv_b200_$name.defs
done
ls -la ../lib/variants/
