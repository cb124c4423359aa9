#!/bin/bash
# C4 (2,000,000 x 8192) on 4 GPUs with the paired Cholesky update and the 1.5-pass objective + gradient
O=gpurun_out/r2aa; mkdir -p $O
run() { # tag workload N steps warmup
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $3 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $3 --steps $4 --warmup $5 --workload $2 --no-cpu-baseline > $O/bench_$1.json 2> $O/bench_$1.err
  echo "$1 rc=$? $(python -c "import json;d=json.load(open('$O/bench_$1.json'));print(round(d['value'],3),'it/s',round(d['ms_per_step'],3),'ms', {k:round(v,2) for k,v in d['stages_ms_per_step'].items()}, (d.get('parity_at_scale') or {}).get('ok'))" 2>&1 | tail -1)"
}
run c4_n4 c4 4 10 3
