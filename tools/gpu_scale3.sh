#!/bin/bash
# 8-GPU check of the peer-memory Gram exchange (SCS_P2P=1 default) against NCCL (SCS_P2P=0)
O=gpurun_out/r2v; mkdir -p $O
run() { # tag workload N steps warmup
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $3 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $3 --steps $4 --warmup $5 --workload $2 --no-cpu-baseline > $O/bench_$1.json 2> $O/bench_$1.err
  echo "$1 rc=$? $(python -c "import json;d=json.load(open('$O/bench_$1.json'));print(round(d['value'],3),'it/s',round(d['ms_per_step'],3),'ms', {k:round(v,2) for k,v in d['stages_ms_per_step'].items()}, (d.get('parity_at_scale') or {}).get('ok'))" 2>&1 | tail -1)"; grep -i "peer-memory" $O/bench_$1.err | head -2
}
export SCS_P2P=1; run c2_n8_p2p c2 8 20 5
export SCS_P2P=0; run c2_n8_nccl c2 8 20 5
export SCS_P2P=1; run c4_n8_p2p c4 8 10 3
run c5_n8_p2p c5 8 10 3
