#!/bin/bash
# __syncwarp after the divergent bodies (kc_warp_scan, leaf phase A / B): whole GPU suite, then every bench line that uses them
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
for W in config3 config2; do
  timeout 300 python bench.py --workload $W --steps 20 --warmup 3 --no-e2e --no-cpu --no-probe > $O/r02_c32_$W.log 2> $O/r02_c32_$W.err
  python - <<PY
import json
try:
    d=json.load(open("$O/r02_c32_$W.log")); print("$W ms/step %.4f kernels %s" % (d["ms_per_step"], d["roofline"].get("kernel_ms")))
except Exception as e:
    print("$W failed:", e)
PY
done
for spec in "config4 0 auto" "config4 0 hash" "config5 50000000 auto" "config5 0 auto"; do
  set -- $spec
  KC_TRACE=1 timeout 400 python bench.py --workload $1 --reads $2 --sparse-algo $3 --steps 3 --warmup 1 > $O/r02_c32_sp_$1_$2_$3.log 2> $O/r02_c32_sp_$1_$2_$3.err
  echo "$spec rc=$?"; python - <<PY
import json
try:
    d=json.load(open("$O/r02_c32_sp_$1_$2_$3.log"))
    print("  ms/step", d.get("ms_per_step"), "distinct", d["config"].get("distinct_kmers"), "self_check", d["config"].get("self_check",{}).get("ok"), d.get("error"))
except Exception as e:
    print("  failed:", e)
PY
  grep "kc_trace" $O/r02_c32_sp_$1_$2_$3.err | tail -7 | grep "scatter\|leaf"
done
timeout 300 python tools/measure_aux.py > $O/r02_aux_rows2.json 2> $O/r02_aux_rows2.err; python -c "
import json
d=json.load(open('$O/r02_aux_rows2.json'))
for r in d['rows']: print({k:(round(v,3) if isinstance(v,float) else v) for k,v in r.items()})
"
