#!/bin/bash
# vectorised round filter in sp_scatter_kernel + clz table size: parity tests, then timings of the multi-round inputs
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity_variants.py tests/test_zzz_first_gpu_run.py tests/test_gpu_parity.py -m gpu -q -x -k "sparse or radix or config4 or config5 or fingerprint" 2>&1 | grep -v "^$" | tail -15 | cut -c1-250
for spec in "config4 0 auto" "config5 100000000 auto" "config5 0 auto"; do
  set -- $spec
  KC_TRACE=1 timeout 400 python bench.py --workload $1 --reads $2 --sparse-algo $3 --steps 3 --warmup 1 > $O/r02_c29_sp_$1_$2_$3.log 2> $O/r02_c29_sp_$1_$2_$3.err
  echo "$spec rc=$?"; python - <<PY
import json
try:
    d=json.load(open("$O/r02_c29_sp_$1_$2_$3.log"))
    print("  ms/step", d.get("ms_per_step"), "distinct", d["config"].get("distinct_kmers"), "self_check", d["config"].get("self_check",{}).get("ok"), d.get("error"))
except Exception as e:
    print("  failed:", e)
PY
  grep -c "radix: scatter" $O/r02_c29_sp_$1_$2_$3.err
  grep "kc_trace" $O/r02_c29_sp_$1_$2_$3.err | tail -7
done
