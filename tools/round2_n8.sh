#!/bin/bash
# Third GPU call of the next round, on 8 GPUs (only after tools/round2_multi_gpu.sh showed at N=2 that
# bench.py exits after its result line):
#   gpurun --gpus 8 --timeout 900 -- 'bash tools/round2_n8.sh'
# BASELINE configs 3, 4, 5 at full scale on 8 B200; every run has its own timeout (8 GPUs x 15 min were
# lost in round 1 to one process that would not exit).
set -u
mkdir -p gpurun_out
O=gpurun_out
R="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
echo "== config 3 (3.1 Gbp, k=12, dense, ncclReduce; rank 0 first probes the variants at N = 1 size: ~2 min on a fresh box)"
timeout 600 $R --master-port 29751 bench.py --gpus 8 --steps 20 --warmup 3 ${DENSE_ALGO:+--algo $DENSE_ALGO} > $O/r02_n8_dense.log 2> $O/r02_n8_dense.err; echo "rc=$?"; cut -c1-400 $O/r02_n8_dense.log
for W in config4 config5; do for A in hash radix; do
  echo "== $W (full scale) $A"
  timeout 300 $R --master-port 29752 bench.py --gpus 8 --workload $W --sparse-algo $A --steps 1 --warmup 1 > $O/r02_n8_${W}_$A.log 2> $O/r02_n8_${W}_$A.err
  echo "rc=$?"; cut -c1-500 $O/r02_n8_${W}_$A.log; tail -2 $O/r02_n8_${W}_$A.err
done; done
