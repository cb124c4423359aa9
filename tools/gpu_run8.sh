#!/bin/bash
O=gpurun_out/r2i; mkdir -p $O
(timeout 600 python -m pytest tests/test_gpu_solvers.py -q --timeout 400 -k "sparse" > $O/pytest_sparse.log 2>&1; echo "rc=$?" >> $O/pytest_sparse.log)
tail -5 $O/pytest_sparse.log
for w in sparse sparse_dense; do
  (timeout 600 python bench.py --workload $w --steps 6 --warmup 3 --no-cpu-baseline > $O/bench_$w.json 2> $O/bench_$w.err; echo "rc=$?" >> $O/bench_$w.err)
  tail -3 $O/bench_$w.err; head -c 700 $O/bench_$w.json; echo
done
