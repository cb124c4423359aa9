#!/bin/bash
O=gpurun_out/r2g; mkdir -p $O
V=$PWD/gpurun_variants
n=1000000
for cfg in "c2s64_slack0 $V/lib_c2_s64.so 0" "c2s64_slack1 $V/lib_c2_s64.so 1" "c2s32_slack0 $V/lib_c2_s32.so 0" "c2s32_slack1 $V/lib_c2_s32.so 1" "c2s32_slack2 $V/lib_c2_s32.so 2" "c2s16_slack1 $V/lib_c2_s16.so 1" "c2s16_slack2 $V/lib_c2_s16.so 2" "c2s16_slack4 $V/lib_c2_s16.so 4"; do
  set -- $cfg
  SCS_B200_LIB=$2 SCS_I8_2CTA=1 SCS_I8_SLACK=$3 timeout 100 python tools/time_gram.py $n 4096 8 > $O/g_$1.log 2>&1
  echo "$1: $(grep '^gram ' $O/g_$1.log) $(grep -c Error $O/g_$1.log)"
done
