#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "config2_full or config4_full or config5_quarter" 2>&1 | tail -4
CMD="python bench.py --workload config4 --sparse-algo auto --steps 1 --warmup 1"
$CMD > $O/r02_plain_spfull.log 2> $O/r02_plain_spfull.err &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sp_leaf -s 3 -c 1 -o $O/r02_prof_spfull $CMD > $O/r02_ncu_spfull.log 2>&1
echo "rc=$?"
