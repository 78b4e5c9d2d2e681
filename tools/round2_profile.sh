#!/bin/bash
# ncu evidence for the dense k = 12 step of the next round (B200_PROFILING.md recipe):
#   gpurun --timeout 900 -- 'ALGO=7 bash tools/round2_profile.sh'
# ALGO = the --algo value to profile (0 = what KC_DENSE_AUTO picks).  The plain run comes first and
# must exit 0; ncu only runs behind it.  Outputs: gpurun_out/r02_launches_a$ALGO.csv (every launch with
# its device time: compare SHARES) and gpurun_out/r02_prof_a$ALGO.ncu-rep (--set full of the partition
# kernels; read here with `ncu -i ... --page raw --csv` / `--page source --csv --print-source cuda,sass`).
set -u
mkdir -p gpurun_out
A=${ALGO:-0}
CMD="python bench.py --algo $A --steps 2 --warmup 3 --no-e2e --no-cpu"
$CMD > gpurun_out/r02_plain_a$A.log 2> gpurun_out/r02_plain_a$A.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02_launches_a$A.csv $CMD > gpurun_out/r02_ncu_l_a$A.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/r02_plain2_a$A.log 2> gpurun_out/r02_plain2_a$A.err &&
ncu --set full --clock-control none --import-source on -k regex:part_ -s 8 -c 4 -o gpurun_out/r02_prof_a$A $CMD > gpurun_out/r02_ncu_f_a$A.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out/r02_prof_a$A.ncu-rep 2>/dev/null
