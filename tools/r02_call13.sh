#!/bin/bash
# Round 2, GPU call 13 (1 GPU): whole GPU test suite, TMA-vs-direct scan experiment, ncu captures of the shipped kernels.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -rs 2>&1 | tail -12
echo "== TMA staging experiment"
timeout 120 tools/microbench4 > $O/r02_microbench4.txt 2>&1; cat $O/r02_microbench4.txt
echo "== ncu: dense k=12 default (scatter7v2 + count7)"
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-probe"
$CMD > $O/r02_plain_k12.log 2> $O/r02_plain_k12.err &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:part_ -s 6 -c 2 -o $O/r02_prof_k12 $CMD > $O/r02_ncu_k12.log 2>&1
echo "rc=$?"
echo "== ncu: config 2 (k=8 default = checksum bins)"
CMD="python bench.py --workload config2 --steps 2 --warmup 3 --no-e2e --no-cpu --no-probe"
$CMD > $O/r02_plain_k8.log 2> $O/r02_plain_k8.err &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:smem16 -s 9 -c 3 -o $O/r02_prof_k8 $CMD > $O/r02_ncu_k8.log 2>&1
echo "rc=$?"
echo "== ncu: sparse radix, config 4 at 1/5 scale"
CMD="python bench.py --workload config4 --reads 20000000 --sparse-algo auto --steps 1 --warmup 1"
$CMD > $O/r02_plain_sp.log 2> $O/r02_plain_sp.err &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sp_ -s 8 -c 4 -o $O/r02_prof_sp $CMD > $O/r02_ncu_sp.log 2>&1
echo "rc=$?"
ls -la $O/*.ncu-rep | tail -5
