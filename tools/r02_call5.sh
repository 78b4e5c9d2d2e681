#!/bin/bash
# Round 2, GPU call 5: second-generation scatter with the self-service flush, 1024 vs 512 threads.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_zzz_first_gpu_run.py -m gpu -q -x 2>&1 | tail -5
for T in 1024 512; do
  KC_W2_THREADS=$T timeout 200 python bench.py --algo 10 --steps 20 --warmup 3 --no-e2e --no-cpu --no-probe > $O/r02_c5_t$T.log 2> $O/r02_c5_t$T.err
  python - <<PY
import json
try:
    d=json.load(open("$O/r02_c5_t$T.log"))
    print("threads=$T ms/step %.4f kernels %s fp %s" % (d["ms_per_step"], d["roofline"]["kernel_ms"], d["config"]["table_fingerprint"]))
except Exception as e:
    print("threads=$T failed:", e); print(open("$O/r02_c5_t$T.err").read()[-1500:])
PY
done
for T in 1024 512; do
CMD="python bench.py --algo 10 --steps 2 --warmup 3 --no-e2e --no-cpu --no-probe"
KC_W2_THREADS=$T $CMD > $O/r02_plain_a10c.log 2> $O/r02_plain_a10c.err &&
KC_W2_THREADS=$T timeout 900 ncu --set full --clock-control none --import-source on -k regex:part_scatter -s 3 -c 1 -o $O/r02_prof_a10c_t$T $CMD > $O/r02_ncu_f_a10c_t$T.log 2>&1
echo "full capture T=$T rc=$?"
done
