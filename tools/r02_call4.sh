#!/bin/bash
# Round 2, GPU call 4: compact second-generation scatter (algo 10) + collapsed leaf sort of the sparse radix path.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_zzz_first_gpu_run.py -m gpu -q -x 2>&1 | tail -5
for A in 10; do
  timeout 200 python bench.py --algo $A --steps 20 --warmup 3 --no-e2e --no-cpu --no-probe > $O/r02_c4_a$A.log 2> $O/r02_c4_a$A.err
  python - <<PY
import json
try:
    d=json.load(open("$O/r02_c4_a$A.log"))
    print("algo=$A ms/step %.4f kernels %s fp %s" % (d["ms_per_step"], d["roofline"]["kernel_ms"], d["config"]["table_fingerprint"]))
except Exception as e:
    print("algo=$A failed:", e); print(open("$O/r02_c4_a$A.err").read()[-1500:])
PY
done
echo "== sparse radix with collapsed leaf sub-buckets"
timeout 300 python -m pytest tests/test_gpu_parity_variants.py -m gpu -q -x -k "sparse_radix" 2>&1 | tail -3
for spec in "config4 20000000" "config4 0" "config5 20000000" "config5 50000000"; do
  set -- $spec
  KC_TRACE=1 timeout 300 python bench.py --workload $1 --reads $2 --sparse-algo radix --steps 2 --warmup 1 > $O/r02_c4_sp_$1_$2.log 2> $O/r02_c4_sp_$1_$2.err
  echo "$spec rc=$?"; python -c "
import json;d=json.load(open('$O/r02_c4_sp_$1_$2.log'));print(d['ms_per_step'],d['config']['distinct_kmers'])"; grep kc_trace $O/r02_c4_sp_$1_$2.err | tail -9
done
CMD="python bench.py --algo 10 --steps 2 --warmup 3 --no-e2e --no-cpu --no-probe"
$CMD > $O/r02_plain_a10b.log 2> $O/r02_plain_a10b.err &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:part_scatter -s 3 -c 1 -o $O/r02_prof_a10b $CMD > $O/r02_ncu_f_a10b.log 2>&1
echo "full capture rc=$?"
