#!/bin/bash
# A/B timing of libfra builds under build_variants/ (tools only; the product loads its own libfra.so).
# usage: tools/ab_variants.sh "<prof_step args>" A B C ...
ARGS=$1; shift
PKG=fpga_real_time_fft_analyzer_b200
cp $PKG/libfra.so /tmp/libfra_keep.so
for v in "$@"; do
  cp build_variants/libfra_$v.so $PKG/libfra.so
  for a in $ARGS; do :; done
  echo "== variant $v"
  IFS='|' read -ra RUNS <<< "$ARGS"
  for r in "${RUNS[@]}"; do python tools/prof_step.py $r 2>&1 | grep -E "K1|ms/step" ; done
done
cp /tmp/libfra_keep.so $PKG/libfra.so
