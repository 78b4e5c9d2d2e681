#!/usr/bin/env python
"""Per-kernel fingerprint of the SASS instruction stream of dna-kmeres-parallel_b200/build/*.o
(opcode + operands, addresses and encodings stripped).  Used to prove that a refactor of a
.cu file left a GPU-verified kernel bit-identical when no GPU is at hand:

    python tools/sass_fingerprint.py > /tmp/now.txt && diff profiles/r01_sass_fingerprints.txt /tmp/now.txt
"""
import glob
import hashlib
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    objs = sorted(glob.glob(os.path.join(ROOT, "dna-kmeres-parallel_b200", "build", "*.o")))
    for o in objs:
        txt = subprocess.run(["cuobjdump", "-sass", o], capture_output=True, text=True, stdin=subprocess.DEVNULL).stdout
        name, h, n = None, None, 0
        out = []
        for line in txt.splitlines():
            m = re.search(r"Function : (\S+)", line)
            if m:
                if name:
                    out.append((name, n, h.hexdigest()[:16]))
                name, h, n = m.group(1), hashlib.md5(), 0
                continue
            m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?)\s*/\*", line)
            if m and name:
                h.update(m.group(1).encode())
                n += 1
        if name:
            out.append((name, n, h.hexdigest()[:16]))
        if not out:
            continue  # no kernels in this object (c++filt without names would wait on stdin)
        demangled = subprocess.run(["c++filt"] + [x[0] for x in out], capture_output=True, text=True,
                                   stdin=subprocess.DEVNULL).stdout.splitlines()
        for (nm, n, d), dm in sorted(zip(out, demangled), key=lambda t: t[1]):
            print("%-10s %5d %s  %s" % (os.path.basename(o), n, d, dm[:150]))


if __name__ == "__main__":
    sys.exit(main())
