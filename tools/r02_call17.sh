#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_zzz_first_gpu_run.py -m gpu -q -x 2>&1 | tail -4
timeout 600 python bench.py --no-cpu > $O/r02_c17_bench.log 2> $O/r02_c17_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("$O/r02_c17_bench.log"))
print(d["ms_per_step"], json.dumps(d["e2e"])[:1800])
PY
