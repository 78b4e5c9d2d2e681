#!/bin/bash
# First GPU call of the next round: everything that was written while the GPU budget of round 1
# was spent (sparse radix path, deferred-retry scatter, bench exit path) in ONE gpurun call.
#   gpurun --timeout 2700 -- 'bash tools/round2_first_call.sh'
# Every step is bounded by its own timeout; outputs land in gpurun_out/r02_*.
set -u
mkdir -p gpurun_out
O=gpurun_out
echo "== 1. GPU parity tests (the first-run-pending file is last, xfail(strict=False), every case in its own time-limited process)"
timeout 1500 python -m pytest tests -m gpu -q -x -rxX > $O/r02_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $O/r02_pytest.log
echo "== 2. dense k=12: shipped (--algo 0) vs deferred-retry scatter (--algo 4) vs paired count (--algo 5) vs 14-mer + 13-mer count (--algo 6) vs seven windows per record (--algo 7) vs the combinations 4+5 (--algo 8) and 4+6 (--algo 9); KC_PART_PAIR=1|2 combines a count variant with any scatter variant"
for A in 0 4 5 6 7 8 9; do
  timeout 300 python bench.py --algo $A --steps 20 --warmup 3 --no-e2e --no-cpu > $O/r02_dense_abl$A.log 2> $O/r02_dense_abl$A.err
  python - <<PY
import json
d=json.load(open("$O/r02_dense_abl$A.log"))
print("algo=$A ms/step %.4f kernels %s checksum %s" % (d["ms_per_step"], d["roofline"]["kernel_ms"], d["config"]["table_checksum"]))
PY
done
echo "== 2a. dense k=12 scatter flush variants (env KC_PART_ABLATE: 3 deferred retry, 4 TMA bulk flush, 5 256-bit stores)"
for A in 3 4 5; do
  KC_PART_ABLATE=$A timeout 300 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu > $O/r02_dense_env$A.log 2> $O/r02_dense_env$A.err
  python - <<PY
import json
try:
    d=json.load(open("$O/r02_dense_env$A.log"))
    print("KC_PART_ABLATE=$A ms/step %.4f kernels %s checksum %s (must equal the shipped path's)" % (d["ms_per_step"], d["roofline"]["kernel_ms"], d["config"]["table_checksum"]))
except Exception as e:
    print("KC_PART_ABLATE=$A failed:", e)
PY
done
echo "== 2a'. deferred-retry scatter + two-increment count together"
KC_PART_ABLATE=3 KC_PART_PAIR=2 timeout 300 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu > $O/r02_dense_env3_pair.log 2> $O/r02_dense_env3_pair.err
python - <<PY
import json
try:
    d=json.load(open("$O/r02_dense_env3_pair.log"))
    print("KC_PART_ABLATE=3 KC_PART_PAIR=2 ms/step %.4f kernels %s checksum %s" % (d["ms_per_step"], d["roofline"]["kernel_ms"], d["config"]["table_checksum"]))
except Exception as e:
    print("combined run failed:", e)
PY
echo "== 2b. k=8 (config 2 and the 3.1 Gbp genome): shipped 16-bit bins vs checksum variant (--algo 3)"
for W in config2 genome_k8; do for A in 0 3; do
  timeout 300 python bench.py --workload $W --algo $A --steps 20 --warmup 3 --no-e2e --no-cpu > $O/r02_${W}_a$A.log 2> $O/r02_${W}_a$A.err
  python - <<PY
import json
d=json.load(open("$O/r02_${W}_a$A.log"))
print("$W algo=$A ms/step %.4f kernels %s checksum %s" % (d["ms_per_step"], d["roofline"]["kernel_ms"], d["config"]["table_checksum"]))
PY
done; done
echo "== 3. sparse config 4 at 1/5 scale and full scale: hash vs radix"
for R in 20000000 0; do for A in hash radix; do
  timeout 600 python bench.py --workload config4 --reads $R --sparse-algo $A --steps 1 --warmup 1 > $O/r02_c4_${A}_$R.log 2> $O/r02_c4_${A}_$R.err
  echo "config4 reads=$R algo=$A rc=$?"; cut -c1-400 $O/r02_c4_${A}_$R.log; tail -2 $O/r02_c4_${A}_$R.err
done; done
echo "== 3a. sparse radix with 256-bit flush stores (KC_RADIX_FLUSH=1), config 4 at 1/5 scale"
KC_RADIX_FLUSH=1 timeout 600 python bench.py --workload config4 --reads 20000000 --sparse-algo radix --steps 1 --warmup 1 > $O/r02_c4_radix256.log 2> $O/r02_c4_radix256.err
echo "rc=$?"; cut -c1-300 $O/r02_c4_radix256.log
echo "== 3b. primitive rates (tools/microbench3): shared increments by pattern, flush variants"
make -s -C tools microbench3 2>/dev/null; timeout 120 tools/microbench3 > $O/r02_microbench3.txt 2>&1; cat $O/r02_microbench3.txt
echo "== 4. default bench line (with e2e + cpu baseline)"
timeout 600 python bench.py > $O/r02_bench_n1.log 2> $O/r02_bench_n1.err; echo "bench rc=$?"; cut -c1-600 $O/r02_bench_n1.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_bench_ref.log 2> $O/r02_bench_ref.err; echo "reference arm rc=$?"; cut -c1-300 $O/r02_bench_ref.log
echo "== 5. host -> GPU through the packed form: packer-thread sweep (the default bench line above already chose between plain and packed)"
for T in 8 16 32 64 128; do
  KC_HOSTPACK_THREADS=$T timeout 300 python bench.py --probe --probe-e2e --steps 5 > $O/r02_e2e_packed_t$T.log 2> $O/r02_e2e_packed_t$T.err
  echo "threads=$T $(cut -c1-200 $O/r02_e2e_packed_t$T.log)"
done
nproc; cat /sys/fs/cgroup/cpu.max 2>/dev/null; lscpu | head -20
