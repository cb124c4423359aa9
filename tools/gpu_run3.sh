#!/bin/bash
# int8 SYRK variants: cluster 4 / 2, 1-CTA / 2-CTA UMMA, lock-step on / off, segment 64 / 128 — full C2 size, no profiler
O=gpurun_out/r2c; mkdir -p $O
(timeout 600 python -m pytest tests/test_gpu_solvers.py -q --timeout 400 -k "wide or sparse" > $O/pytest_wide_sparse.log 2>&1; echo "rc=$?" >> $O/pytest_wide_sparse.log)
run() { # name lib 2cta nolock
  SCS_B200_LIB=$2 SCS_I8_2CTA=$3 SCS_I8_NOLOCK=$4 timeout 200 python tools/time_gram.py 1000000 4096 6 > $O/gram_$1.log 2>&1
  echo "$1: $(grep '^gram ' $O/gram_$1.log)"
}
D=$PWD/selfconcordantsmoothoptimization.jl_b200/libscs_b200.so
V=$PWD/gpurun_variants
run c4_1cta $D 0 0
run c4_1cta_nolock $D 0 1
run c2_2cta $V/lib_r2_c2_s64.so 1 0
run c2_2cta_nolock $V/lib_r2_c2_s64.so 1 1
run c2_1cta $V/lib_r2_c2_s64.so 0 0
run c2_2cta_s128 $V/lib_r2_c2_s128.so 1 0
run c4_2cta_nolock $D 1 1
SCS_B200_LIB=$V/lib_r2_c2_s64.so SCS_I8_2CTA=1 timeout 300 ncu --set full --clock-control none -k regex:k_i8syrk -c 1 -o $O/ncu_i8syrk_c2_2cta -f python tools/time_gram.py 262144 4096 1 > $O/ncu_c2_2cta.log 2>&1
tail -3 $O/pytest_wide_sparse.log
