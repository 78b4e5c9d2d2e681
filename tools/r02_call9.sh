#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 200 python tools/debug_w2.py 2>&1 | grep "differing" | head
timeout 600 python -m pytest tests/test_zzz_first_gpu_run.py -m gpu -q -x 2>&1 | tail -3
for T in 1024; do
  KC_W2_THREADS=$T timeout 200 python bench.py --algo 10 --steps 20 --warmup 3 --no-e2e --no-cpu --no-probe > $O/r02_c10_t$T.log 2> $O/r02_c10_t$T.err
  python - <<PY
import json
try:
    d=json.load(open("$O/r02_c10_t$T.log"))
    print("threads=$T ms/step %.4f kernels %s fp %s" % (d["ms_per_step"], d["roofline"]["kernel_ms"], d["config"]["table_fingerprint"]))
except Exception as e:
    print("threads=$T failed:", e); print(open("$O/r02_c10_t$T.err").read()[-1500:])
PY
done
CMD="python bench.py --algo 10 --steps 2 --warmup 3 --no-e2e --no-cpu --no-probe"
$CMD > $O/r02_plain_a10f.log 2> $O/r02_plain_a10f.err &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:part_scatter -s 3 -c 1 -o $O/r02_prof_a10f $CMD > $O/r02_ncu_f_a10f.log 2>&1
echo "full capture rc=$?"
