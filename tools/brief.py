"""print the interesting fields of a bench.py JSON line (headline + workloads)"""
import json
import sys


def brief(x):
    keep = {"roofline": ("frac", "achieved", "kernel_ms", "unit"), "e2e": ("value", "pageable", "error"), "cpu_baseline": ("value", "cores")}
    out = {}
    for k in ("value", "ms_per_step", "steps", "roofline", "e2e", "cpu_baseline", "parity", "error", "trace", "bench_seconds", "verdict", "blocks", "clocks", "device_seam_matches_host_merge", "gpu_launches"):
        if k in x:
            out[k] = {kk: vv for kk, vv in (x[k] or {}).items() if kk in keep[k]} if k in keep else x[k]
    return out


d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("HEAD", json.dumps(brief(d)))
for w, x in d.get("workloads", {}).items():
    print(w.upper(), json.dumps(brief(x)))
print("total_s", d.get("bench_seconds_total"))
