#!/bin/bash
# Round 2 multi-GPU call:  gpurun --gpus N --timeout 1500 -- 'N=<N> bash tools/r02_multi.sh'
# (TAG=<suffix> keeps the files of an earlier run)
# BASELINE configs 3, 4, 5 at full scale on N B200; every run has its own time limit (8 GPUs x 15 min were lost in
# round 1 to one process that would not exit).
set -u
mkdir -p gpurun_out
O=gpurun_out
N=${N:-8}
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== config 3 (3.1 Gbp, k=12, dense, ncclReduce) N=$N"
[ "${DENSE:-1}" = 1 ] && timeout 240 $R --master-port 29751 bench.py --gpus $N --steps 20 --warmup 3 --no-probe > $O/r02_m${N}${TAG:-}_dense.log 2> $O/r02_m${N}${TAG:-}_dense.err; echo "rc=$? (124 = hung)"
[ "${DENSE:-1}" = 1 ] && python - <<PY
import json
try:
    d=json.load(open("$O/r02_m${N}${TAG:-}_dense.log")); print("dense N=$N ms/step %.4f kernels %s launch %s fp %s e2e %s" % (d["ms_per_step"], d["roofline"]["kernel_ms"], d["config"]["launch"], d["config"]["table_fingerprint"], (d.get("e2e") or {}).get("ms_per_step")))
except Exception as e:
    print("failed", e); print(open("$O/r02_m${N}${TAG:-}_dense.err").read()[-1500:])
PY
SPECS=${SPECS:-config4:0:auto config4:0:hash config5:0:auto config5:0:hash}
for spec in $SPECS; do
  IFS=: read W RD A <<< "$spec"
  timeout 240 $R --master-port 29752 bench.py --gpus $N --workload $W --reads $RD --sparse-algo $A --steps ${STEPS:-2} --warmup ${WARM:-1} > $O/r02_m${N}${TAG:-}_${W}_${RD}_$A.log 2> $O/r02_m${N}${TAG:-}_${W}_${RD}_$A.err
  echo "$W reads=$RD $A rc=$?"; python - <<PY
import json
try:
    d=json.load(open("$O/r02_m${N}${TAG:-}_${W}_${RD}_$A.log"))
    print("  ms/step", d.get("ms_per_step"), "distinct", d["config"].get("distinct_kmers"), "self_check", d["config"].get("self_check",{}).get("ok"), d["config"].get("self_check",{}).get("fingerprint_out"), d.get("error"))
except Exception as e:
    print("  failed:", e); print(open("$O/r02_m${N}${TAG:-}_${W}_${RD}_$A.err").read()[-1000:])
PY
done
