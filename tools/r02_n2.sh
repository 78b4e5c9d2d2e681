#!/bin/bash
# Round 2, 2-GPU call: NCCL parity tests, dense N=2, sparse configs 4/5 at N=2 with self-checks.  Every run is bounded.
set -u
mkdir -p gpurun_out
O=gpurun_out
N=${N:-2}
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== NCCL parity tests"
timeout 600 python -m pytest tests -m gpu -q -rs -k "nccl or multi_gpu" 2>&1 | tail -8
echo "== dense N=$N"
timeout 240 $R --master-port 29741 bench.py --gpus $N --steps 20 --warmup 3 --no-probe > $O/r02_n${N}_dense.log 2> $O/r02_n${N}_dense.err; echo "rc=$? (124 = hung)"
python - <<PY
import json
try:
    d=json.load(open("$O/r02_n${N}_dense.log")); print("dense N=$N ms/step %.4f kernels %s fp %s e2e %s" % (d["ms_per_step"], d["roofline"]["kernel_ms"], d["config"]["table_fingerprint"], (d.get("e2e") or {}).get("ms_per_step")))
except Exception as e:
    print("failed", e); print(open("$O/r02_n${N}_dense.err").read()[-1200:])
PY
echo "== sparse N=$N"
for spec in "config4 0 auto" "config4 0 hash" "config5 50000000 auto" "config5 50000000 hash" "config5 100000000 auto"; do
  set -- $spec
  timeout 240 $R --master-port 29742 bench.py --gpus $N --workload $1 --reads $2 --sparse-algo $3 --steps 2 --warmup 1 > $O/r02_n${N}_sp_$1_$2_$3.log 2> $O/r02_n${N}_sp_$1_$2_$3.err
  echo "$spec rc=$?"; python - <<PY
import json
try:
    d=json.load(open("$O/r02_n${N}_sp_$1_$2_$3.log"))
    print("  ms/step", d.get("ms_per_step"), "distinct", d["config"].get("distinct_kmers"), "self_check", d["config"].get("self_check",{}).get("ok"), d["config"].get("self_check",{}).get("fingerprint_out"), d.get("error"))
except Exception as e:
    print("  failed:", e); print(open("$O/r02_n${N}_sp_$1_$2_$3.err").read()[-800:])
PY
done
