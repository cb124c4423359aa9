#!/bin/bash
O=gpurun_out/r2e; mkdir -p $O
D=$PWD/selfconcordantsmoothoptimization.jl_b200/libscs_b200.so
V=$PWD/gpurun_variants/lib_r2_c2_s64.so
for n in 8192 16384 32768 65536 131072 262144; do
  for cfg in "c4_1cta $D 0" "c2_2cta $V 1"; do
    set -- $cfg
    SCS_B200_LIB=$2 SCS_I8_2CTA=$3 timeout 100 python tools/time_gram.py $n 4096 20 > $O/g_$1_$n.log 2>&1
    echo "$1 n=$n: $(grep '^gram ' $O/g_$1_$n.log)"
  done
done
