#!/bin/bash
# Second GPU call of the next round, on 2 GPUs:  gpurun --gpus 2 --timeout 900 -- 'bash tools/round2_multi_gpu.sh'
# (N=2 first: the N=8 run of round 1 printed its result and then hung in library teardown for the whole
# time limit, which is what bench.py's leave() now avoids — check that cheaply before spending 8 GPUs.)
set -u
mkdir -p gpurun_out
O=gpurun_out
N=${N:-2}
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== 1. NCCL parity tests (verified hash-sharded path, then the first-run-pending range-sharded radix path)"
timeout 600 python -m pytest tests -m gpu -q -rxX -k "nccl" > $O/r02_pytest_nccl.log 2>&1; echo "pytest rc=$?"; tail -6 $O/r02_pytest_nccl.log
echo "== 2. dense bench at N=$N: must print its line AND exit"
timeout 600 $R --master-port 29741 bench.py --gpus $N --steps 20 --warmup 3 > $O/r02_dense_n$N.log 2> $O/r02_dense_n$N.err; echo "dense N=$N rc=$? (124 = hung)"; cut -c1-300 $O/r02_dense_n$N.log
echo "== 3. config 4 at 1/5 scale, N=$N: hash-sharded vs range-sharded radix"
for A in hash radix; do
  timeout 600 $R --master-port 29742 bench.py --gpus $N --workload config4 --reads 20000000 --sparse-algo $A --steps 1 --warmup 1 > $O/r02_c4_n${N}_$A.log 2> $O/r02_c4_n${N}_$A.err
  echo "config4 N=$N $A rc=$?"; cut -c1-400 $O/r02_c4_n${N}_$A.log; tail -2 $O/r02_c4_n${N}_$A.err
done
echo "== 4. dense bench at N=$N with the packed host path in e2e (opt-in until seen green)"
KC_BENCH_E2E_PACKED_N=1 timeout 300 $R --master-port 29743 bench.py --gpus $N --steps 20 --warmup 3 > $O/r02_dense_n${N}_packed.log 2> $O/r02_dense_n${N}_packed.err; echo "rc=$?"
python - <<PY
import json
try:
    d = json.load(open("$O/r02_dense_n${N}_packed.log"))
    print("e2e:", json.dumps(d["e2e"])[:500])
except Exception as e:
    print("no line:", e)
PY
