#!/bin/bash
# Round 2, GPU call 1 (one GPU): everything round 1 wrote but never timed.
#   gpurun --timeout 1500 -- 'bash tools/r02_call1.sh'
# Every step has its own timeout; outputs land in gpurun_out/r02_*.
set -u
mkdir -p gpurun_out
O=gpurun_out
nproc; cat /sys/fs/cgroup/cpu.max 2>/dev/null; nvidia-smi --query-gpu=name,memory.total --format=csv
echo "== 1. k=8 (config 2): shipped 16-bit bins (--algo 0) vs checksum variant (--algo 3)"
for A in 0 3; do
  timeout 200 python bench.py --workload config2 --algo $A --steps 50 --warmup 5 --no-e2e --no-cpu --no-probe > $O/r02_config2_a$A.log 2> $O/r02_config2_a$A.err
  python - <<PY
import json
try:
    d=json.load(open("$O/r02_config2_a$A.log"))
    print("config2 algo=$A ms/step %.4f kernels %s checksum %s" % (d["ms_per_step"], d["roofline"].get("kernel_ms"), d["config"]["table_checksum"]))
except Exception as e:
    print("config2 algo=$A failed:", e)
PY
done
echo "== 2. sparse config 4 at 1/5 scale and full scale: hash vs radix"
for R in 20000000 0; do for A in hash radix; do
  timeout 300 python bench.py --workload config4 --reads $R --sparse-algo $A --steps 2 --warmup 1 > $O/r02_c4_${A}_$R.log 2> $O/r02_c4_${A}_$R.err
  echo "config4 reads=$R algo=$A rc=$?"; cut -c1-700 $O/r02_c4_${A}_$R.log; tail -2 $O/r02_c4_${A}_$R.err
done; done
echo "== 3. sparse config 5 at 1/10 and 1/4 scale: hash vs radix"
for R in 20000000 50000000; do for A in hash radix; do
  timeout 300 python bench.py --workload config5 --reads $R --sparse-algo $A --steps 2 --warmup 1 > $O/r02_c5_${A}_$R.log 2> $O/r02_c5_${A}_$R.err
  echo "config5 reads=$R algo=$A rc=$?"; cut -c1-700 $O/r02_c5_${A}_$R.log; tail -2 $O/r02_c5_${A}_$R.err
done; done
echo "== 4. primitive rates (tools/microbench3)"
timeout 120 tools/microbench3 > $O/r02_microbench3.txt 2>&1; cat $O/r02_microbench3.txt
echo "== 5. host -> GPU through the packed form: packer-thread sweep"
for T in 8 16 32; do
  KC_HOSTPACK_THREADS=$T timeout 200 python bench.py --probe --probe-e2e --steps 5 > $O/r02_e2e_packed_t$T.log 2> $O/r02_e2e_packed_t$T.err
  echo "threads=$T $(cut -c1-300 $O/r02_e2e_packed_t$T.log)"
done
lscpu | head -25; numactl -H 2>/dev/null | head
