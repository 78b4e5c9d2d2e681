#!/bin/bash
# Round 2, GPU call 14: TMA bulk flush in the second-generation scatter (KC_W2_TMA=1) vs the lane flush; config 2 after the sliced reduce.
set -u
mkdir -p gpurun_out
O=gpurun_out
for T in 0 1; do
  if [ $T = 1 ]; then export KC_W2_TMA=1; else unset KC_W2_TMA; fi
  timeout 120 python tools/debug_w2.py 2>&1 | grep "differing" | head -2
  timeout 200 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --no-probe > $O/r02_c14_tma$T.log 2> $O/r02_c14_tma$T.err
  python - <<PY
import json
try:
    d=json.load(open("$O/r02_c14_tma$T.log"))
    print("tma=$T ms/step %.4f kernels %s fp %s" % (d["ms_per_step"], d["roofline"]["kernel_ms"], d["config"]["table_fingerprint"]))
except Exception as e:
    print("tma=$T failed:", e); print(open("$O/r02_c14_tma$T.err").read()[-1500:])
PY
done
unset KC_W2_TMA
timeout 200 python bench.py --workload config2 --steps 50 --warmup 5 --no-e2e --no-cpu --no-probe > $O/r02_c14_config2.log 2> $O/r02_c14_config2.err
python - <<PY
import json
d=json.load(open("$O/r02_c14_config2.log")); print("config2 ms/step %.4f kernels %s" % (d["ms_per_step"], d["roofline"]["kernel_ms"]))
PY
timeout 200 python bench.py --workload config4 --sparse-algo auto --steps 3 --warmup 1 > $O/r02_c14_c4.log 2>/dev/null; python -c "
import json; d=json.load(open('$O/r02_c14_c4.log')); print('config4 auto ms', d['ms_per_step'], d['config']['self_check']['ok'])"
