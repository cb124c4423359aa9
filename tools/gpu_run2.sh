#!/bin/bash
# round-2 GPU session: full GPU test-suite, solve timing, int8 SYRK experiment (1-CTA vs 2-CTA UMMA), ncu captures, bench
O=gpurun_out/r2b; mkdir -p $O
(timeout 1200 python -m pytest tests -m gpu -q --timeout 400 --deselect tests/test_gpu_multirank.py > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log)
for m in 4096 8192; do timeout 120 python tools/time_solve.py $m > $O/solve_$m.log 2>&1; done
smi() { nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap --format=csv,noheader,nounits -lms 100 > $1 & echo $!; }
for v in 0 1; do
  pid=$(smi $O/smi_gram_2cta$v.csv)
  SCS_I8_2CTA=$v timeout 200 python tools/time_gram.py 1000000 4096 6 > $O/gram_2cta$v.log 2>&1
  kill $pid
done
for v in 0 1; do
  SCS_I8_2CTA=$v timeout 300 ncu --set full --clock-control none -k regex:k_i8syrk -c 1 -o $O/ncu_i8syrk_2cta$v -f python tools/time_gram.py 262144 4096 1 > $O/ncu_2cta$v.log 2>&1
done
(timeout 400 python bench.py --steps 10 --warmup 3 > $O/bench_c2.json 2> $O/bench_c2.err; echo "rc=$?" >> $O/bench_c2.err)
tail -4 $O/pytest.log; cat $O/solve_4096.log $O/solve_8192.log $O/gram_2cta0.log $O/gram_2cta1.log | grep -v "^$"; ls -la $O
