"""examen_b200 — the reference's main() flow as a CLI ("next" row f3) — on a FASTA file,
against the distances the reference's own sequentialKmerCount2 produced (golden)."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "dna-kmeres-parallel_b200", "examen_b200")


def _build():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "dna-kmeres-parallel_b200"), "examen_b200"], check=True)


def test_driver_builds_and_reports_usage():
    _build()
    r = subprocess.run([EXE], capture_output=True, text=True)
    assert r.returncode == 2 and "usage:" in r.stderr
    r = subprocess.run([EXE, "/nonexistent.fasta"], capture_output=True, text=True)
    assert r.returncode == 1 and "Error opening" in r.stderr  # the reference exits 0 here (main.cu:477-480)


@pytest.mark.gpu
def test_driver_matches_reference_distances(golden, oracle, tmp_path):
    _build()
    for case in golden["distance"]:
        k, seqs = case["k"], case["seqs"]
        if any("\n" in s or "\r" in s or s == "" for s in seqs):
            continue
        fasta = tmp_path / "in.fasta"
        fasta.write_bytes("".join(">s%d\n%s\n\n" % (i, s) for i, s in enumerate(seqs)).encode("latin-1"))
        out, sums = tmp_path / "parallel_results.csv", tmp_path / "sums.txt"
        r = subprocess.run([EXE, str(fasta), "-k", str(k), "--out", str(out), "--sums", str(sums)],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        assert "sequences read ." in r.stdout and "Elapsed parallel timer step 1" in r.stdout
        assert out.read_text() == "".join("%f\n" % d for d in np.array(case["dist"], dtype=np.float32))
        data = b"".join(s.encode("latin-1") + b"\0" for s in seqs)
        offs = np.cumsum([0] + [len(s) + 1 for s in seqs])
        want, _ = oracle.count_per_seq(data, offs, k)
        assert sums.read_bytes() == oracle.dump_counts(want, k, len(seqs))


@pytest.mark.gpu
def test_driver_gpu_parse_matches_host_parse(golden, tmp_path):
    """--gpu-parse: the same distances and sums as the host-parsed run"""
    _build()
    rng = np.random.default_rng(3)
    seqs = ["".join("ACGT"[c] for c in rng.integers(0, 4, int(rng.integers(30, 4000)))) for _ in range(12)]
    fasta = tmp_path / "in.fasta"
    fasta.write_bytes("".join(">s%d\n%s\n\n" % (i, "\n".join(s[j:j + 60] for j in range(0, len(s), 60))) for i, s in enumerate(seqs)).encode())
    outs = []
    for flag in ([], ["--gpu-parse"]):
        out, sums = tmp_path / ("r%d.csv" % len(flag)), tmp_path / ("s%d.txt" % len(flag))
        r = subprocess.run([EXE, str(fasta), "-k", "4", "--out", str(out), "--sums", str(sums)] + flag, capture_output=True, text=True,
                           timeout=180)  # a never-run kernel may hang: bounded, and the run is its own process
        assert r.returncode == 0, r.stderr
        outs.append((out.read_text(), sums.read_text()))
    assert outs[0] == outs[1]
