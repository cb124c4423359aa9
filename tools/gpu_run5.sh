#!/bin/bash
O=gpurun_out/r2f; mkdir -p $O
D=$PWD/selfconcordantsmoothoptimization.jl_b200/libscs_b200.so
V=$PWD/gpurun_variants
for n in 16384 262144 1000000; do
  reps=20; [ $n -ge 1000000 ] && reps=6
  for cfg in "c4_1cta $D 0 0" "c4_1cta_nolock $D 0 1" "c2_1cta $V/lib_r2_c2.so 0 0" "c2_2cta $V/lib_r2_c2.so 1 0" "c1_1cta $V/lib_r2_c1.so 0 0" "c1_1cta_nolock $V/lib_r2_c1.so 0 1"; do
    set -- $cfg
    SCS_B200_LIB=$2 SCS_I8_2CTA=$3 SCS_I8_NOLOCK=$4 timeout 100 python tools/time_gram.py $n 4096 $reps > $O/g_$1_$n.log 2>&1
    echo "$1 n=$n: $(grep '^gram ' $O/g_$1_$n.log) $(grep -c Error $O/g_$1_$n.log)"
  done
done
