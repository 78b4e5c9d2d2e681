#!/bin/bash
# Round 2, GPU call 33 (the same as call 28, after the last kernel changes) (1 GPU): whole GPU suite, the default bench line + reference arm, ncu capture of the sparse kernels as shipped
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -rs 2>&1 | tail -8
echo "== default bench line"
timeout 600 python bench.py > $O/r02_bench_n1_final2.log 2> $O/r02_bench_n1_final2.err; echo "rc=$?"; cut -c1-600 $O/r02_bench_n1_final2.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_bench_ref_final2.log 2> $O/r02_bench_ref_final2.err; echo "rc=$?"; cut -c1-400 $O/r02_bench_ref_final2.log
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
echo "== ncu: sparse radix, config 4 at 1/5 scale (6 sp_ kernels per call: skip the warm-up call)"
CMD="python bench.py --workload config4 --reads 20000000 --sparse-algo auto --steps 1 --warmup 1"
$CMD > $O/r02_plain_sp3.log 2> $O/r02_plain_sp3.err &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sp_ -s 6 -c 6 -o $O/r02_prof_sp3 $CMD > $O/r02_ncu_sp3.log 2>&1
echo "rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_launches_sp3.csv $CMD > $O/r02_ncu_sp3_list.log 2>&1
echo "rc=$?"
ls -la $O/r02_prof_sp3.ncu-rep
