#!/bin/bash
# Round 2, GPU call 3: first B200 run of the second-generation scatter (algo 10) — parity, timing, ncu.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_zzz_first_gpu_run.py -m gpu -q -x 2>&1 | tail -15
for A in 6 7 10; do
  timeout 200 python bench.py --algo $A --steps 20 --warmup 3 --no-e2e --no-cpu --no-probe > $O/r02_c3_a$A.log 2> $O/r02_c3_a$A.err
  python - <<PY
import json
try:
    d=json.load(open("$O/r02_c3_a$A.log"))
    print("algo=$A ms/step %.4f kernels %s fp %s" % (d["ms_per_step"], d["roofline"]["kernel_ms"], d["config"]["table_fingerprint"]))
except Exception as e:
    print("algo=$A failed:", e); print(open("$O/r02_c3_a$A.err").read()[-1500:])
PY
done
CMD="python bench.py --algo 10 --steps 2 --warmup 3 --no-e2e --no-cpu --no-probe"
$CMD > $O/r02_plain_a10.log 2> $O/r02_plain_a10.err &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:part_ -s 6 -c 2 -o $O/r02_prof_a10 $CMD > $O/r02_ncu_f_a10.log 2>&1
echo "full capture rc=$?"; ls -la $O/r02_prof_a10.ncu-rep 2>/dev/null
