#!/bin/bash
O=gpurun_out/r2u; mkdir -p $O
(timeout 600 python -m pytest tests/test_gpu_multirank.py -q --timeout 600 > $O/pytest_mr.log 2>&1; echo "rc=$?" >> $O/pytest_mr.log); tail -4 $O/pytest_mr.log | cut -c1-300
for v in 1 0; do
  SCS_P2P=$v timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_c2_n2_p2p$v.json 2> $O/bench_c2_n2_p2p$v.err
  echo "p2p=$v rc=$? $(python -c "import json;d=json.load(open('$O/bench_c2_n2_p2p$v.json'));print(round(d['value'],3),'it/s',round(d['ms_per_step'],3),'ms', {k:round(v,2) for k,v in d['stages_ms_per_step'].items()}, d['parity_at_scale']['ok'], d['multirank_parity'])" 2>&1 | tail -1)"; grep -i "peer-memory\|error" $O/bench_c2_n2_p2p$v.err | head -3
done
