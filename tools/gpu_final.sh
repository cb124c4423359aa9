#!/bin/bash
# what the driver runs at round end, on one B200: GPU test-suite, smoke, default bench
O=gpurun_out/r2z; mkdir -p $O
(timeout 1200 python -m pytest tests -m gpu -q --timeout 400 > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log); tail -4 $O/pytest.log
(timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log); tail -4 $O/smoke.log
(timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_c2_n1.json 2> $O/bench_c2_n1.err; echo "bench rc=$?"); python -c "
import json
d=json.load(open('$O/bench_c2_n1.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['stages_ms_per_step'], d['roofline']['frac'], d['parity_at_scale']['ok'])"
