"""Collect the sparse (config 4 / 5) bench lines of gpurun_out/r02_*.log into profiles/r02_sparse_runs.json.
Each entry keeps the file name, N, workload, algorithm, ms per call and the full-scale self-check of that run."""
import glob, json, os, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def entry(path):
    txt = open(path).read().strip()
    line = next((l for l in txt.splitlines() if l.startswith("{")), None)
    e = {"file": os.path.basename(path)}
    if line is None:
        e.update(error=(txt.splitlines() or ["no output"])[-1][:300], ms_per_step=None)
        return e
    d = json.loads(line)
    c = d.get("config", {})
    if "sparse" not in c.get("workload", ""):
        return None
    e.update(n_gpus=d.get("n_gpus"), workload=c.get("workload"), k=c.get("k"), reads=c.get("reads"), algo=c.get("algo"),
             ms_per_step=d.get("ms_per_step"), bases_per_s=d.get("value"), distinct_kmers=c.get("distinct_kmers"),
             roofline_frac=(d.get("roofline") or {}).get("frac"), self_check=c.get("self_check"),
             phases=c.get("phases"), error=d.get("error"), timing=c.get("timing"))
    return e


def code_state(name):
    """which leaf pass / planning the run had (the file tags are the GPU calls of round 2, in order)"""
    import re
    m = re.match(r"r02_(c|m|n)(\d+)([b-z]?)_", name)
    if not m:
        return None
    kind, num, b = m.group(1), int(m.group(2)), m.group(3) == "b"
    if kind == "c" and num == 23:
        return "A/B of call 23: distinct-first leaf pass" if "dedupe" in name else "A/B of call 23: record-sort leaf pass"
    if kind in ("m", "n") and re.match(r"r02_[mn]\d+[c-z]_", name):
        return "distinct-first leaf pass, warp converged after divergent bodies" + (
            ", rounds appended through kc_sparse_radix_count_round_append" if re.match(r"r02_[mn]\d+[e-z]_", name) else "")
    if kind == "c" and num >= 32:
        return "distinct-first leaf pass, warp converged after divergent bodies, rounds appended on one GPU"
    if (kind == "c" and num >= 24) or b:
        s = "distinct-first leaf pass"
        if kind == "c" and num in (24, 25):
            s += " (plan / append logic of that call still being fixed: config 5 lines show allocation stalls or a retried plan)"
        return s
    return "record-sort leaf pass (round 2's first form)"


def main():
    out = []
    for p in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", "r02_*.log"))):
        try:
            e = entry(p)
        except Exception as ex:  # a truncated line from a timed-out run
            e = {"file": os.path.basename(p), "error": "unparsable: %s" % ex, "ms_per_step": None}
            if "_sp_" not in p and "config4" not in p and "config5" not in p:
                e = None
        # runs under ncu are never bench values; lines without the self-check predate csrc/check.cu (early round 2)
        if e:
            e["code"] = code_state(e["file"])
        if e and "_ncu_" not in e["file"] and (e.get("self_check") or e.get("ms_per_step") is None) \
                and ("config4" in e["file"] or "config5" in e["file"]):
            out.append(e)
    json.dump(out, open(os.path.join(ROOT, "profiles", "r02_sparse_runs.json"), "w"), indent=1)
    for e in out:
        sc = e.get("self_check") or {}
        print("%-46s N=%s %-5s ms=%s ok=%s" % (e["file"], e.get("n_gpus"), e.get("algo"), e.get("ms_per_step"), sc.get("ok")))


if __name__ == "__main__":
    main()
