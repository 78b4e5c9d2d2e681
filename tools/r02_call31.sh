#!/bin/bash
# ncu capture of perseq_kernel (row a1, the reference's own kernel shape) at k = 3 and k = 6
set -u
mkdir -p gpurun_out
O=gpurun_out
python tools/prof_perseq.py > $O/r02_plain_perseq.log 2>&1 && cat $O/r02_plain_perseq.log &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:perseq -s 2 -c 1 -o $O/r02_prof_perseq_k3 python tools/prof_perseq.py > $O/r02_ncu_perseq3.log 2>&1
echo rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:perseq -s 5 -c 1 -o $O/r02_prof_perseq_k6 python tools/prof_perseq.py > $O/r02_ncu_perseq6.log 2>&1
echo rc=$?
ls -la $O/r02_prof_perseq*.ncu-rep
