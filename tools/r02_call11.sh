#!/bin/bash
# Round 2, GPU call 11 (1 GPU): new defaults (dense k=12 -> WIDE2, sparse -> auto/radix), pooled allocations, self-checks.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_zzz_first_gpu_run.py -m gpu -q -x 2>&1 | tail -4
echo "== dense k=12, library default vs round 1's default, full size and a 1/8 shard"
for L in 0 387500000; do for R in new r01; do
  if [ $R = r01 ]; then export KC_DENSE_AUTO_R01=1; else unset KC_DENSE_AUTO_R01; fi
  LA=""; [ $L != 0 ] && LA="--length $L"
  timeout 200 python bench.py $LA --steps 20 --warmup 3 --no-e2e --no-cpu --no-probe > $O/r02_c11_d_${L}_$R.log 2> $O/r02_c11_d_${L}_$R.err
  python - <<PY
import json
try:
    d=json.load(open("$O/r02_c11_d_${L}_$R.log"))
    print("L=$L default=$R ms/step %.4f kernels %s fp %s frac %.4f" % (d["ms_per_step"], d["roofline"]["kernel_ms"], d["config"]["table_fingerprint"], d["roofline"]["step"]["frac"]))
except Exception as e:
    print("L=$L $R failed:", e); print(open("$O/r02_c11_d_${L}_$R.err").read()[-1500:])
PY
done; done
unset KC_DENSE_AUTO_R01
echo "== sparse, auto vs hash, self-check in the line"
for spec in "config4 20000000 auto" "config4 0 auto" "config4 0 hash" "config5 50000000 auto" "config5 100000000 auto"; do
  set -- $spec
  KC_TRACE=1 timeout 300 python bench.py --workload $1 --reads $2 --sparse-algo $3 --steps 3 --warmup 1 > $O/r02_c11_sp_$1_$2_$3.log 2> $O/r02_c11_sp_$1_$2_$3.err
  echo "$spec rc=$?"; python - <<PY
import json
try:
    d=json.load(open("$O/r02_c11_sp_$1_$2_$3.log"))
    print("  ms/step", d.get("ms_per_step"), "distinct", d["config"].get("distinct_kmers"), "self_check", d["config"].get("self_check",{}).get("ok"), d.get("error"), d["config"].get("timing"))
except Exception as e:
    print("  failed:", e)
PY
  grep kc_trace $O/r02_c11_sp_$1_$2_$3.err | tail -9
done
