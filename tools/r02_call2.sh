#!/bin/bash
# Round 2, GPU call 2: per-kernel launch lists of the sparse radix path (ncu duration only; SHARES matter).
set -u
mkdir -p gpurun_out
O=gpurun_out
for spec in "config4 0" "config4 20000000" "config5 20000000" "config5 50000000"; do
  set -- $spec
  CMD="python bench.py --workload $1 --reads $2 --sparse-algo radix --steps 1 --warmup 1"
  timeout 300 $CMD > $O/r02_l_$1_$2.log 2> $O/r02_l_$1_$2.err &&
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02_launches_$1_$2.csv $CMD > $O/r02_ncu_$1_$2.log 2>&1
  echo "$spec rc=$?"
  python - <<PY
import csv,collections
rows=[r for r in csv.reader(open("$O/r02_launches_$1_$2.csv")) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value"); ui=hdr.index("Metric Unit")
for r in rows[1:]:
    print(r[ki][:60], r[vi], r[ui])
PY
done
