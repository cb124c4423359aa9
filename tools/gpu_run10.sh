#!/bin/bash
# paired trailing update of the look-ahead Cholesky: correctness + isolated timing at m = 4096 / 8192, both settings
mkdir -p gpurun_out/r2x
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -k "linear_solve" > gpurun_out/r2x/pytest_solve.log 2>&1
tail -3 gpurun_out/r2x/pytest_solve.log
for m in 4096 8192; do for pr in 0 1; do
  SCS_SOLVE_PAIR=$pr timeout 200 python tools/time_solve.py $m > gpurun_out/r2x/time_solve_${m}_pair${pr}.txt 2>&1
  echo "m=$m pair=$pr"; tail -3 gpurun_out/r2x/time_solve_${m}_pair${pr}.txt
done; done
