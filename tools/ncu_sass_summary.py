#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page source --csv --print-source sass [-k regex:...]` output:
stall reasons, opcode mix and the hottest SASS lines of every kernel in the file.
usage: ncu_sass_summary.py sass.csv [top_n]"""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    topn = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    i = 0
    seen = set()
    while i < len(rows):
        if rows[i] and rows[i][0] == "Kernel Name":
            name = rows[i][1]
            hdr = rows[i + 1]
            ix = {h: j for j, h in enumerate(hdr)}
            st = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
            tot, ops, opsamp, opwf = collections.Counter(), collections.Counter(), collections.Counter(), collections.Counter()
            inst = 0
            lines = []
            i += 2
            while i < len(rows) and not (rows[i] and rows[i][0] == "Kernel Name"):
                r = rows[i]
                i += 1
                if len(r) < len(hdr):
                    continue
                for s in st:
                    tot[s] += int(r[ix[s]] or 0)
                n = int(r[ix["Instructions Executed"]] or 0)
                inst += n
                toks = r[ix["Source"]].split()
                op = toks[1] if toks[0].startswith("@") else toks[0]
                op = op.split(".")[0]
                ops[op] += n
                opsamp[op] += int(r[ix["# Samples"]] or 0)
                if "L1 Wavefronts Shared" in ix:
                    opwf[op] += int(r[ix["L1 Wavefronts Shared"]] or 0)
                lines.append((int(r[ix["# Samples"]] or 0), n, r[ix["Source"]]))
            if name in seen:  # one launch per kernel is enough
                continue
            seen.add(name)
            print("====", name[:110])
            print("SASS instructions %d, warp-instructions executed %d" % (len(lines), inst))
            S = sum(tot.values()) or 1
            print("stalls:", ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / S) for k, v in tot.most_common(9)))
            for k, v in ops.most_common(18):
                print("   %-8s exec %11d (%4.1f%%)  samples %7d  smem wf %d" % (k, v, 100.0 * v / max(inst, 1), opsamp[k], opwf[k]))
            if topn:
                for smp, n, src in sorted(lines, reverse=True)[:topn]:
                    print("   hot: samples %6d exec %10d  %s" % (smp, n, src))
        else:
            i += 1


if __name__ == "__main__":
    main()
