#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
for spec in "config5 0 auto" "config4 0 auto" "config5 100000000 auto"; do
  set -- $spec
  KC_TRACE=1 timeout 300 python bench.py --workload $1 --reads $2 --sparse-algo $3 --steps 2 --warmup 1 > $O/r02_c16_sp_$1_$2_$3.log 2> $O/r02_c16_sp_$1_$2_$3.err
  echo "$spec rc=$?"; python - <<PY
import json
try:
    d=json.load(open("$O/r02_c16_sp_$1_$2_$3.log"))
    print("  ms/step", d.get("ms_per_step"), "distinct", d["config"].get("distinct_kmers"), "self_check", d["config"].get("self_check",{}).get("ok"), d.get("error"))
except Exception as e:
    print("  failed:", e)
PY
  echo "  scatter calls in the last 5 sparse calls: $(grep -c 'radix: scatter' $O/r02_c16_sp_$1_$2_$3.err)"
done
echo "== default bench line (e2e + cpu baseline) and the reference arm"
timeout 600 python bench.py > $O/r02_bench_n1.log 2> $O/r02_bench_n1.err; echo "bench rc=$?"; cut -c1-900 $O/r02_bench_n1.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_bench_ref.log 2> $O/r02_bench_ref.err; echo "reference arm rc=$?"; cut -c1-300 $O/r02_bench_ref.log
