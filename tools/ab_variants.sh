#!/bin/bash
# A/B of engine variants built by tools/build_variant.py (GPU box).  usage: bash tools/ab_variants.sh TAG "S1 S2 .." NAME...
TAG=$1; SIZES=$2; shift 2
mkdir -p gpurun_out
V=nonstationary_multivariate_gaussian_process_b200/variants
for name in "$@"; do
  for S in $SIZES; do
    if [ "$name" = "main" ]; then unset NMGP_B200_LIB; else export NMGP_B200_LIB=$V/libnmgp_b200_$name.so; fi
    echo "== variant $name S=$S" | tee -a gpurun_out/ab_$TAG.txt
    timeout 300 python tools/run_config.py nonseparable 100 6 $S 3 2>&1 | grep "^{" | tee -a gpurun_out/ab_$TAG.txt
  done
done
unset NMGP_B200_LIB
