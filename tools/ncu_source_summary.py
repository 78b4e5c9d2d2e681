#!/usr/bin/env python
"""Per-source-line warp instructions and stall samples from
`ncu -i X.ncu-rep --page source --csv --print-source cuda,sass [-k regex:...]`.
usage: ncu_source_summary.py cs.csv [min_pct]"""
import csv
import os
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    minpct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.4
    out = []
    fname, hdr = None, None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = os.path.basename(r[1])
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r
            ie = hdr.index("Instructions Executed")
            isamp = hdr.index("# Samples")
            continue
        if hdr is None or len(r) < len(hdr):
            continue
        if r[2] != "-":  # a SASS row under its source line
            continue
        try:
            out.append((fname, int(r[0]), int(r[ie] or 0), int(r[isamp] or 0), r[1].strip()))
        except ValueError:
            pass
    tot = sum(o[2] for o in out) or 1
    tots = sum(o[3] for o in out) or 1
    print("total warp-instructions %d, stall samples %d" % (tot, tots))
    for f, ln, n, s, src in out:
        if 100.0 * n / tot >= minpct or 100.0 * s / tots >= minpct:
            print("%-14s %5d %11d %5.1f%%  samp %6d %5.1f%% | %s" % (f, ln, n, 100.0 * n / tot, s, 100.0 * s / tots, src[:110]))


if __name__ == "__main__":
    main()
