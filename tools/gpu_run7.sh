#!/bin/bash
# 2-GPU session: multi-rank parity (incl. packed slab-pipelined Gram all-reduce, wide branch across ranks), bench C2/C3 at N=2
O=gpurun_out/r2h; mkdir -p $O
(timeout 900 python -m pytest tests/test_gpu_multirank.py -q --timeout 900 > $O/pytest_mr.log 2>&1; echo "rc=$?" >> $O/pytest_mr.log)
tail -15 $O/pytest_mr.log
for w in c2 c3; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 3 --workload $w > $O/bench_${w}_n2.json 2> $O/bench_${w}_n2.err
  echo "bench $w rc=$?"; tail -c 600 $O/bench_${w}_n2.json
done
