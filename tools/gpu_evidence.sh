#!/bin/bash
# round-2 evidence on one B200: bench lines (no profiler), then the ncu launch lists and --set full captures of the same commands
O=gpurun_out/r2k; mkdir -p $O
(timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_c2_n1.json 2> $O/bench_c2_n1.err; echo "c2 rc=$?")
(timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $O/bench_c2_reference.json 2> $O/bench_c2_reference.err; echo "ref rc=$?")
(timeout 600 python bench.py --workload c3 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_c3_n1.json 2> $O/bench_c3_n1.err; echo "c3 rc=$?")
(timeout 600 python bench.py --workload sparse --steps 20 --warmup 5 > $O/bench_sparse_n1.json 2> $O/bench_sparse_n1.err; echo "sparse rc=$?")
Q="--no-peak --no-cpu-baseline --no-parity"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/ncu_launches_c2.csv python bench.py --steps 2 --warmup 1 $Q > $O/ncu_launches_c2.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/ncu_launches_c3.csv python bench.py --workload c3 --rows 2000000 --steps 3 --warmup 1 $Q > $O/ncu_launches_c3.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/ncu_launches_smoke.csv python -c "import __graft_entry__ as g; g.smoke()" > $O/ncu_launches_smoke.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_i8syrk|k_residues|k_fused_grad|k_crt|k_wstat|k_colscale" -s 9 -c 7 -f -o $O/ncu_full_c2 python bench.py --steps 1 --warmup 1 $Q > $O/ncu_full_c2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_lqn_head|k_lqn_update|k_fused_grad" -s 4 -c 6 -f -o $O/ncu_full_c3 python bench.py --workload c3 --rows 2000000 --steps 3 --warmup 1 $Q > $O/ncu_full_c3.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_sp_" -s 4 -c 5 -f -o $O/ncu_full_sparse python bench.py --workload sparse --steps 1 --warmup 1 $Q > $O/ncu_full_sparse.log 2>&1
# text summaries on the box; the reports themselves are too large to travel back (64 MiB limit)
for r in c2 c3 sparse; do python tools/ncu_summary.py $O/ncu_full_$r.ncu-rep > $O/ncu_full_summary_$r.txt 2>&1; done
ncu -i $O/ncu_full_c2.ncu-rep --page raw --csv > $O/ncu_full_c2_raw.csv 2>/dev/null
rm -f $O/ncu_full_c3.ncu-rep $O/ncu_full_sparse.ncu-rep $O/ncu_full_c2.ncu-rep
gzip -f $O/ncu_full_c2_raw.csv
ls -la $O
