"""bench.py contract checks that need no GPU: the CPU reference arm (`--impl reference`) prints
one JSON line with the keys the driver reads, alone and under torchrun (rank 0 only)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "bench.py")
KEYS = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
        "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"}


def _check_line(out, n_gpus):
    lines = [ln for ln in out.splitlines() if ln.strip()]
    assert len(lines) == 1, out
    d = json.loads(lines[0])
    assert KEYS <= set(d), sorted(KEYS - set(d))
    assert d["impl"] == "reference" and d["metric"] == "bases/sec" and d["n_gpus"] == n_gpus
    assert d["value"] > 0 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and "sample" in d["cpu_baseline"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert "workload" in d["config"]
    return d


def test_reference_arm_single():
    r = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--workload", "config2", "--steps", "2", "--warmup", "1",
                        "--ref-sample", "4000000"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    d = _check_line(r.stdout, 1)
    assert d["steps"] == 2 and d["warmup"] == 1 and d["config"]["k"] == 8


def test_reference_arm_under_torchrun_prints_once():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29663", BENCH, "--impl", "reference", "--gpus", "2", "--workload", "config2", "--steps", "1",
           "--warmup", "0", "--ref-sample", "2000000"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, OMP_NUM_THREADS="1"))
    assert r.returncode == 0, r.stderr[-2000:]
    _check_line(r.stdout, 2)


def test_reference_arm_time_box_shrinks_the_sample():
    r = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--workload", "config2", "--steps", "3", "--warmup", "1",
                        "--ref-sample", "268435456", "--ref-budget-s", "0.5"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    d = _check_line(r.stdout, 1)
    assert d["config"]["bases_per_step"] < 268435456  # the calibration pass shrank it
