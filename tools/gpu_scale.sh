#!/bin/bash
# 8-GPU session: multi-rank parity, then every named shape at the GPU counts its shard (+ int8 planes) fits
O=gpurun_out/r2s; mkdir -p $O
(timeout 900 python -m pytest tests/test_gpu_multirank.py -q --timeout 900 > $O/pytest_mr.log 2>&1; echo "rc=$?" >> $O/pytest_mr.log); tail -3 $O/pytest_mr.log
run() { # workload N steps warmup
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $2 --steps $3 --warmup $4 --workload $1 --no-cpu-baseline > $O/bench_$1_n$2.json 2> $O/bench_$1_n$2.err
  echo "$1 N=$2 rc=$? $(python -c "import json;d=json.load(open('$O/bench_$1_n$2.json'));print(round(d['value'],3),'it/s',round(d['ms_per_step'],3),'ms', {k:round(v,2) for k,v in d['stages_ms_per_step'].items()})" 2>&1 | tail -1)"
}
for n in 8 4 2; do run c2 $n 20 5; done
for n in 8 4 2; do run c3 $n 20 5; done
for n in 8 4; do run c4 $n 10 3; run c5 $n 10 3; done
