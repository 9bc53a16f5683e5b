#!/usr/bin/env python
"""Summarise ncu evidence (gpurun_out/*.ncu-rep, launch lists) into profiles/<name>.md (run in the build container)."""
import csv
import io
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_issued.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed_pipe_lsu.sum",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg.per_second", "launch__cluster_size", "launch__shared_mem_per_block_dynamic"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    # some section metrics carry a "UNIT.Section." prefix (e.g. TPC.TriageCompute.sm__pipe_tensor...): index them by bare name too
    bare = [h.split(".", 2)[2] if h.count(".") >= 2 and h.split(".")[0].isupper() else h for h in hdr]
    recs = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        for b, v in zip(bare, r):
            d.setdefault(b, v)
        recs.append(d)
    u = dict(zip(hdr, units))
    for b, v in zip(bare, units):
        u.setdefault(b, v)
    return recs, u


def stalls(rep, top=12):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        return []
    hdr = rows[1]
    try:
        ia, isrc, isamp = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples")
    except ValueError:
        return []
    data = [(int(r[isamp] or 0), r[isrc]) for r in rows[2:] if len(r) > isamp]
    tot = sum(d[0] for d in data) or 1
    return [(n, 100.0 * n / tot, src) for n, src in sorted(data, key=lambda x: -x[0])[:top]]


def launches(path):
    rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
    hdr = rows[0]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = {}
    for r in rows[1:]:
        if len(r) <= iv:
            continue
        name = r[ik].split("(")[0][-60:]
        v = float(r[iv].replace(",", ""))
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
    return agg


def main():
    import glob
    import json
    out_dir = os.path.join(ROOT, "profiles")
    os.makedirs(out_dir, exist_ok=True)
    src = os.path.join(ROOT, "gpurun_out")
    prefix = sys.argv[1] if len(sys.argv) > 1 else "r02"
    SF = 600_037_902
    algo = {"q6": 28.0 * SF, "q1": 38.0 * SF, "sum": 8.0 * 10_000_000}
    # bench.py roofline.traffic keys: "<workload>:<rows_per_gpu>[:<queries>]"
    traffic_key = {"q6": "q6:%d" % SF, "q1": "q1:%d" % SF, "sum": "sum:10000000", "tc_ivf": "ivf:10000000:10000", "tc_bruteforce": "bruteforce:1000000:10000"}
    traffic = {}
    lines = ["# ncu evidence, %s" % prefix, "",
             "Captured with `tools/profile_%s.sh` under gpurun (`ncu --set full --clock-control none --import-source on`); the `.ncu-rep`" % prefix,
             "files stay in gpurun_out/ (scratch).  Durations under ncu are cold-cache and serialised: compare shares, not absolutes.", ""]
    names = sorted(os.path.basename(f)[len(prefix) + 1:-len(".ncu-rep")] for f in glob.glob(os.path.join(src, "%s_*.ncu-rep" % prefix)))
    for name in names:
        rep = os.path.join(src, "%s_%s.ncu-rep" % (prefix, name))
        recs, units = raw(rep)
        for rec in recs[:1]:
            lines += ["## %s -- `%s`" % (name, rec.get("Kernel Name", "?")[:110]), "", "| metric | value | unit |", "|---|---|---|"]
            for k in KEYS:
                if k in rec and rec[k] != "":
                    lines.append("| %s | %s | %s |" % (k, rec[k], units.get(k, "")))
            try:
                rd = float(rec["dram__bytes_read.sum"]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[units["dram__bytes_read.sum"]]
                wr = float(rec["dram__bytes_write.sum"]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[units["dram__bytes_write.sum"]]
                dur = float(rec["gpu__time_duration.sum"]) * {"ms": 1e-3, "us": 1e-6, "s": 1, "ns": 1e-9}[units["gpu__time_duration.sum"]]
                lines += ["", "DRAM traffic per launch = %.4f GB (read %.4f + write %.4f); duration under ncu %.3f ms => %.0f GB/s of DRAM traffic." % ((rd + wr) / 1e9, rd / 1e9, wr / 1e9, dur * 1e3, (rd + wr) / dur / 1e9)]
                if algo.get(name):
                    lines.append("Algorithmic bytes per launch = %.4f GB => traffic / algorithmic = %.3f." % (algo[name] / 1e9, (rd + wr) / algo[name]))
                if name in traffic_key:
                    traffic[traffic_key[name]] = {"bytes_per_launch": rd + wr, "kernel": rec.get("Kernel Name", "?")[:80],
                                                  "source": "profiles/%s_ncu_summary.md section '%s' (dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of this bench command)" % (prefix, name)}
            except Exception as ex:
                lines.append("(traffic summary unavailable: %s)" % ex)
            st = stalls(rep)
            if st:
                lines += ["", "Top sampled instructions (warp-stall samples):", "", "| samples | % | SASS |", "|---|---|---|"]
                lines += ["| %d | %.1f | `%s` |" % (n, pct, s_[:90]) for n, pct, s_ in st]
            lines.append("")
    for name in ("q6", "q1", "sum", "bruteforce", "ivf"):
        p = os.path.join(src, "%s_launches_%s.csv" % (prefix, name))
        if os.path.exists(p):
            agg = launches(p)
            tot = sum(v[1] for v in agg.values()) or 1
            lines += ["## launch list: `python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --workload %s` (set-up kernels included)" % name, "",
                      "| kernel | launches | total ns | share |", "|---|---|---|---|"]
            for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                lines.append("| `%s` | %d | %.0f | %.1f %% |" % (k, c, v, 100 * v / tot))
            lines.append("")
    open(os.path.join(out_dir, "%s_ncu_summary.md" % prefix), "w").write("\n".join(lines) + "\n")
    if traffic:
        open(os.path.join(out_dir, "ncu_traffic.json"), "w").write(json.dumps(traffic, indent=1) + "\n")
    print("\n".join(lines[:60]))


if __name__ == "__main__":
    main()
