#!/usr/bin/env python
"""Generate tests/golden/*.json from the REFERENCE's own host code.

Run in the authoring container (needs /root/reference and oracle/_ref, built by
`make -C oracle ref`).  Every expected value below is produced by the reference's
permutation() / permutationsCountAll() / importSeqs() / importSeqsNoNL() /
sequentialKmerCount2(), compiled unmodified (oracle/ref_harness.cu); nothing is
computed by this repo's oracle or engine.  The JSON files are committed so the
tests can run where /root/reference does not exist (the GPU box).
"""
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, HERE)
import oracle as O  # noqa: E402
from mt64 import mt19937_64  # noqa: E402


def dirty_sequence(rng, n):
    """random ACGT with N runs, lower case, CR and other junk sprinkled in"""
    s = rng.choice(list(b"ACGT"), size=n).astype(np.uint8)
    for _ in range(max(1, n // 200)):
        p = int(rng.integers(0, n))
        ln = int(rng.integers(1, 12))
        s[p:p + ln] = ord("N")
    junk = b"acgtn\r-*RYKM.X"
    for _ in range(max(1, n // 150)):
        s[int(rng.integers(0, n))] = junk[int(rng.integers(0, len(junk)))]
    return bytes(s)


def main():
    assert O.ref(3) is not None, "build oracle/_ref first: make -C oracle ref"
    out = {}

    # ---- enumeration order, k = 3..6 (utils.h:21-50) -----------------------
    out["permutation"] = {str(k): O.ref(k).permutation() for k in (3, 4)}
    out["permutation_fnv"] = {}
    for k in (3, 4, 5, 6):
        perms = O.ref(k).permutation()
        out["permutation_fnv"][str(k)] = "%016x" % O.fnv1a64(np.frombuffer("".join(perms).encode(), dtype=np.uint8))

    # ---- known-answer strings at k=3 ------------------------
    kat = []
    for s in ["ACGTNACGTacgtAAAA", "AAAA", "AC", "ACG", "ACGTACGTAC", "NNNACGNNN", "ACGT\rACGT", "", "A",
              "TTTTTTTTTT", "ACGTACGTACGTACGTACGTACGTACGTACGTACGT", "NNNNNNNN", "acgtacgt", "ACG|ACG", "GATTACA"]:
        c = O.ref(3).count_all(s)
        kat.append({"seq": s, "k": 3, "counts": {str(i): int(c[i]) for i in np.nonzero(c)[0]}})
    out["kat_k3"] = kat

    # ---- random dirty sequences, k = 3..6 ----------------------------------
    rng = np.random.default_rng(20261018)
    dirty = []
    for k in (3, 4, 5, 6):
        for n in (1, k - 1, k, k + 1, 31, 257, 1500):
            s = dirty_sequence(rng, n) if n >= 8 else bytes(rng.choice(list(b"ACGT"), size=n).astype(np.uint8))
            c = O.ref(k).count_all(s)
            dirty.append({"seq": s.decode("latin-1"), "k": k,
                          "counts": {str(i): int(c[i]) for i in np.nonzero(c)[0]}})
    out["dirty"] = dirty

    # ---- 1 Mbp std::mt19937_64(1234) vector -----------------
    g = mt19937_64(1234)
    seq = bytes(b"ACGT"[g.next() & 3] for _ in range(1000000))
    c = O.ref(3).count_all(seq)
    out["mt19937_64_1mbp_k3"] = {"seed": 1234, "n": 1000000, "head": seq[:32].decode(),
                                 "counts": [int(x) for x in c],
                                 "seq_fnv": "%016x" % O.fnv1a64(np.frombuffer(seq, dtype=np.uint8))}

    # ---- BASELINE config 1: 1 Mbp splitmix sequence, k=3, via the reference --
    seed1 = 0xB2000001
    seq1 = O.gen_bases(seed1, 0, 1000000).tobytes()
    c = O.ref(3).count_all(seq1)
    out["config1_k3"] = {"seed": seed1, "n": 1000000, "head": seq1[:32].decode(), "counts": [int(x) for x in c]}
    # the same sequence at k = 4..6 (still the reference's code, other K builds)
    for k in (4, 5, 6):
        c = O.ref(k).count_all(seq1[:200000])
        out["config1_prefix200k_k%d" % k] = {"seed": seed1, "n": 200000, "invalid": int(c[0]),
                                             "table_fnv": "%016x" % O.fnv1a64(c[1:].astype("<i4"))}

    # ---- loader fixtures: importSeqs / importSeqsNoNL -----------------------
    fastas = {
        "blank_separated": ">s1 first\nACGTAC\nGGT\n\n>s2\nTTTT\n\n>s3 last no newline\nACGNNAC",
        "crlf": ">s1\r\nACGT\r\nAC\r\n\r\n>s2\r\nGG\r\n",
        "no_blank_lines": ">s1\nACGT\nAC\n>s2\nGGGG\n>s3\nTT\nTT\n",
        "trailing_blank": ">s1\nACGT\n\n>s2\nGGCC\n\n",
        "stray_and_pipe": "stray line\n>s1\nAC|GT\nAA\n\nmore stray\n>s2\nCCCC\n",
        "single_line_records": ">a\nACGTACGT\n\n>b\nTTTTAAAA\n\n>c\nGGGG",
        "empty": "",
        "header_only": ">s1\n",
    }
    loader = []
    for name, text in fastas.items():
        for mode in (0, 1):
            with tempfile.NamedTemporaryFile("wb", suffix=".fasta", delete=False) as f:
                f.write(text.encode("latin-1"))
                path = f.name
            r = O.ref(3).import_seqs(path, mode)
            os.unlink(path)
            loader.append({"name": name, "mode": mode, "fasta": text, "num_seqs": r["num_seqs"], "ids": r["ids"],
                           "seqs": [s.decode("latin-1") for s in r["seqs"]],
                           "offsets": [int(x) for x in r["offsets"]],
                           "data": r["data"].decode("latin-1")})
    out["loader"] = loader
    # MAX_SEQS = 100 behaviour (main.cu:30,514,524): 130 three-line records
    many = "".join(">r%d\nACGT\nGGCC\nTTAA\n\n" % i for i in range(130))
    with tempfile.NamedTemporaryFile("wb", suffix=".fasta", delete=False) as f:
        f.write(many.encode())
        path = f.name
    r = O.ref(3).import_seqs(path, 0)
    os.unlink(path)
    out["loader_max_seqs"] = {"records": 130, "max_seqs": 100, "num_seqs": r["num_seqs"],
                              "last_seq": r["seqs"][-1].decode(), "offsets_tail": [int(x) for x in r["offsets"][-3:]],
                              "num_offsets": int(len(r["offsets"]))}

    # ---- distance step (sequentialKmerCount2, main.cu:587-621) --------------
    dist = []
    for k, seqs in [(3, ["ACGTACGTAC", "ACGTTTTTAC", "GGGGGGGG", "ACGTACGTACGTNNACGT"]),
                    (3, [dirty_sequence(rng, 300).decode("latin-1") for _ in range(6)]),
                    (4, [dirty_sequence(rng, 500).decode("latin-1") for _ in range(5)])]:
        d = O.ref(k).distance(seqs)
        dist.append({"k": k, "seqs": seqs, "dist": [float(x) for x in d], "dist_hex": [x.tobytes().hex() for x in d]})
    out["distance"] = dist
    out["triangular_index"] = [[i, j, n, O.ref(3).triangular_index(i, j, n)]
                               for n in (2, 5, 100) for i in (1, 2, n - 1) for j in (1, 2, 3) if i <= n - 1]

    path = os.path.join(HERE, "reference_golden.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
