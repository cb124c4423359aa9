#!/bin/bash
# 8-GPU tuning: slab count of the packed Gram all-reduce (SCS_GRAM_SLABS), and C4 with the 1.5-pass gradient
O=gpurun_out/r2t; mkdir -p $O
run() { # tag workload N steps warmup
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $3 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $3 --steps $4 --warmup $5 --workload $2 --no-cpu-baseline --no-parity > $O/bench_$1.json 2> $O/bench_$1.err
  echo "$1 rc=$? $(python -c "import json;d=json.load(open('$O/bench_$1.json'));print(round(d['value'],3),'it/s',round(d['ms_per_step'],3),'ms', {k:round(v,2) for k,v in d['stages_ms_per_step'].items()})" 2>&1 | tail -1)"
}
for s in 1 2 4; do export SCS_GRAM_SLABS=$s; run c2_n8_slabs$s c2 8 20 5; done
export SCS_GRAM_SLABS=2; run c2_n4_slabs2 c2 4 20 5; export SCS_GRAM_SLABS=1; run c2_n4_slabs1 c2 4 20 5
unset SCS_GRAM_SLABS
run c4_n8 c4 8 10 3
