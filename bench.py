#!/usr/bin/env python
"""bench.py -- the measurement contract.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload q6|q1|sum|bruteforce]

Default workload = BASELINE.json configs[1]: TPC-H Q6 (3-predicate filter + SUM, fp64) over SF100 synthetic lineitem
columns (600 037 902 rows, 28 B/row = 16.8 GB) on one B200.  A "step" is one pass of the fused scan->filter->agg over
the rank's columns (+ the NCCL exchange of partial aggregates when N > 1).  One JSON line is printed by rank 0.

  value        rows/s, whole job, columns RESIDENT in HBM when the timed region starts (CUDA events, max over ranks)
  e2e          the same metric through the C-ABI call with HOST (pinned) column buffers: H2D copies inside the timed region
  roofline     algorithmic bytes per launch / CUDA-event duration of the dominant kernel, against MEASURED_PEAKS.json
  cpu_baseline the oracle port (C restatement of the Go operator chain) on the host cores, bounded sample
  --impl reference   times that CPU implementation alone (rank 0), same metric / config / unit

Multi-GPU: weak scaling -- every rank owns an SF100-sized, disjoint row range of an N x SF100 table (block ranges, the
reference's buildScanParallelRun split); partial (sum, count) records are exchanged with NCCL all_gather and merged.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SF100_ROWS = 600_037_902
WORKLOADS = {
    "q6": dict(name="tpch_q6_sf100_fp64", metric="scan+filter+agg rows/sec (TPC-H Q6, fp64)", bytes_per_row=28.0, rows=SF100_ROWS),
    "q1": dict(name="tpch_q1_sf100_fp64_packed_keys", metric="scan+filter+group-agg rows/sec (TPC-H Q1, fp64)", bytes_per_row=38.0, rows=SF100_ROWS),
    "sum": dict(name="int64_sum_10m_rows", metric="int64 SUM rows/sec", bytes_per_row=8.0, rows=10_000_000),
    "bruteforce": dict(name="bruteforce_l2_top10_1Mx768_10k_queries", metric="ANN top-k qps (768-d, brute-force L2 top-10)", bytes_per_row=None, rows=1_000_000),
    "ivf": dict(name="ivfflat_l2_top10_10Mx768_nlist1024_nprobe32_10k_queries", metric="ANN top-k qps (768-d, IVF-flat nlist=1024 nprobe=32 top-10)", bytes_per_row=None, rows=10_000_000),
}


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler:
    """nvidia-smi clocks during the timed region (B200_PROFILING.md recipe)"""
    Q = "timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self, t_begin=None, t_end=None):
        """t_begin / t_end: wall-clock (time.time()) bounds of the timed region; samples outside are dropped"""
        import datetime
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                if t_begin is not None:
                    ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                    if ts < t_begin - 0.02 or ts > t_end + 0.02:
                        continue
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(section, default_size):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed `ncu --set full` capture
    of this very command (profiles/r01_ncu_summary.md, written by tools/summarize_ncu.py).  Only meaningful at the size the
    capture was taken at, so None for any other --rows."""
    if not default_size:
        return None
    try:
        txt = open(os.path.join(ROOT, "profiles", "r01_ncu_summary.md")).read()
        sec = txt.split("## %s -- " % section, 1)[1].split("\n## ", 1)[0]
        gb = float(sec.split("DRAM traffic per launch = ", 1)[1].split(" GB", 1)[0])
        return gb * 1e9
    except Exception:
        return None


# =====================================================================================================================
# CPU reference arm / cpu_baseline: the oracle port on the host cores
# =====================================================================================================================
def cpu_reference(workload, sample_rows, steps, warmup, threads, dataset=None, queries=None):
    import oracle_lib as O
    from matrixone_b200 import datagen
    if workload == "bruteforce":
        # GoBruteForceIndex.Search on the FULL dataset (per-query work must not shrink), one query per host thread per step
        dim = 768
        ds = dataset if dataset is not None else datagen.vectors_f32(20, 0, sample_rows, dim)
        qs = queries if queries is not None else datagen.vectors_f32(21, 0, threads, dim)
        nq = qs.shape[0]
        for _ in range(warmup):
            O.bruteforce(ds, qs[:max(1, nq // 4)], 10, 0, threads)
        t0 = time.perf_counter()
        for _ in range(steps):
            O.bruteforce(ds, qs, 10, 0, threads)
        dt = time.perf_counter() - t0
        return nq * steps / dt, dt / steps
    if workload == "ivf":
        # IvfflatSearchIndex.Search (findCentroids + scan of the nprobe probed lists + heap) on a full per-GPU shard: nlist 1024,
        # nprobe 32, top-10; `queries` per step spread over the host threads.  dataset = (entries, list id per entry, centroids)
        dim, nlist, nprobe, k = 768, 1024, 32, 10
        if dataset is None:
            centers = datagen.vectors_f32(30, 0, nlist, dim) * 4
            ents = datagen.vectors_f32(31, 0, sample_rows, dim, centers, 1.0)
            assign = datagen.vector_components(31, 0, sample_rows, nlist)   # the generating component: a valid list assignment
        else:
            ents, assign, centers = dataset
        qs = queries if queries is not None else datagen.vectors_f32(32, 0, 8 * threads, dim, centers, 1.0)
        nq = qs.shape[0]
        keys = np.zeros(nq * k, dtype=np.int64); dists = np.zeros(nq * k)
        run = lambda m: O.go().og_ivf_search_f32(O.p(ents), O.p(assign), ents.shape[0], dim, O.p(centers), nlist, O.p(qs), m, nprobe, k, 0, 0, threads, O.p(keys), O.p(dists))
        for _ in range(warmup):
            run(max(1, nq // 4))
        t0 = time.perf_counter()
        for _ in range(steps):
            run(nq)
        dt = time.perf_counter() - t0
        return nq * steps / dt, dt / steps
    if workload == "q6":
        cols = datagen.lineitem(10, 0, sample_rows)
        P = datagen.q6_params()
        fn = lambda: O.q6(cols, sample_rows, P, nthreads=threads)
    elif workload == "q1":
        cols = datagen.lineitem(10, 0, sample_rows)
        fn = lambda: O.q1(cols, sample_rows, datagen.Q1_CUTOFF, nthreads=threads)
    elif workload == "sum":
        v, _ = datagen.int64_column(1, 0, sample_rows)
        s = np.zeros(1, dtype=np.int64); nul = np.zeros(1, dtype=np.uint8)
        fn = lambda: O.go().og_sum_int64_mt(O.p(v), None, sample_rows, threads, O.p(s), O.p(nul))
    else:
        raise SystemExit("cpu reference for workload %s: use --workload q6|q1|sum" % workload)
    for _ in range(warmup):
        fn()
    t0 = time.perf_counter()
    for _ in range(steps):
        fn()
    dt = time.perf_counter() - t0
    return sample_rows * steps / dt, dt / steps


def run_reference_arm(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return 0
    wl = WORKLOADS[args.workload]
    threads = os.cpu_count() or 1
    sample = min(wl["rows"], 1 << 25)
    steps = max(1, args.steps)
    warmup = max(1, min(args.warmup, 2))
    if args.workload in ("bruteforce", "ivf"):
        steps, warmup = min(steps, 3), 1
    if args.workload == "ivf":
        sample = 1_250_000   # one GPU's shard of the 10 M-row index
    value, sec_per_step = cpu_reference(args.workload, sample, steps, warmup, threads)
    unit = "queries/s" if args.workload in ("bruteforce", "ivf") else "rows/s"
    if args.workload == "ivf":
        line = {"impl": "reference", "metric": wl["metric"], "value": value, "unit": unit, "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
                "ms_per_step": sec_per_step * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": wl["name"], "rows": sample, "queries_per_step": 8 * threads},
                "cpu_baseline": {"value": value, "unit": unit, "cores": threads, "kind": "port",
                                 "sample": "1.25 M x 768 shard (nlist 1024, nprobe 32, top-10), %d queries per step over %d host threads; oracle/oracle_go.c og_ivf_search_f32 (IvfflatSearchIndex.Search)" % (8 * threads, threads)},
                "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(line))
        return 0
    if args.workload == "bruteforce":
        line = {"impl": "reference", "metric": wl["metric"], "value": value, "unit": unit, "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
                "ms_per_step": sec_per_step * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": wl["name"], "rows": sample, "queries_per_step": threads},
                "cpu_baseline": {"value": value, "unit": unit, "cores": threads, "kind": "port",
                                 "sample": "full 1 M x 768 dataset, %d queries per step (one per host thread); oracle/oracle_go.c GoBruteForceIndex.Search (metric.L2DistanceSq + FastMaxHeap)" % threads},
                "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(line))
        return 0
    line = {
        "impl": "reference", "metric": wl["metric"], "value": value, "unit": "rows/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": sec_per_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64" if args.workload != "sum" else "int64",
        "data": "synthetic", "config": {"workload": wl["name"], "rows_per_step": sample},
        "cpu_baseline": {"value": value, "unit": "rows/s", "cores": threads, "kind": "port",
                         "sample": "first %d rows of the workload per step; oracle/oracle_go.c operator chain (filter conjunct by conjunct + Shrink, projection, aggexec fill) on %d pthreads over 8192-row blocks" % (sample, threads)},
        "e2e": {"value": value, "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# =====================================================================================================================
# our arm
# =====================================================================================================================
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="q6", choices=sorted(WORKLOADS))
    ap.add_argument("--rows", type=int, default=0, help="override rows per GPU (testing only; the default is the BASELINE config)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--queries", type=int, default=10_000)
    ap.add_argument("--metric", default="l2", choices=["l2", "ip", "cos"], help="bruteforce workload only (experiments; BASELINE config 4 is l2)")
    ap.add_argument("--tune", action="append", default=[], metavar="NAME=VALUE",
                    help="kernel-variant knob passed to MoB200_SetTuning (experiments only; the default run uses none)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    os.environ.setdefault("MO_B200_DEVICE", str(local))
    from matrixone_b200 import capi, datagen, ops, shard
    from matrixone_b200.vector import DeviceBuffer, PinnedArray
    lib = capi.load_library()
    capi.check(lib.MoB200_Init(local), lib)
    for kv in args.tune:
        name, _, val = kv.partition("=")
        lib.MoB200_SetTuning(name.encode(), int(val))

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier_sync():
        capi.check(lib.MoB200_Sync(), lib)
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    wl = WORKLOADS[args.workload]
    n = args.rows or wl["rows"]
    W = max(3, args.warmup)
    K = max(1, args.steps)
    row0 = rank * n                      # this rank's disjoint block range of the N x SF100 table
    result = {}

    # ---------------------------------------------------------------------------------------------- workload set-up
    if args.workload in ("q6", "q1"):
        names = ["shipdate", "quantity", "extendedprice", "discount"] + (["tax", "returnflag", "linestatus"] if args.workload == "q1" else [])
        size = {"shipdate": 4, "returnflag": 1, "linestatus": 1}
        bufs = {k: DeviceBuffer(size.get(k, 8) * n, lib) for k in names}
        ptr = lambda k: bufs[k].ptr if k in bufs else None
        capi.check(lib.MoB200_GenLineitem(10, row0, n, ptr("shipdate"), ptr("quantity"), ptr("extendedprice"), ptr("discount"), ptr("tax"),
                                          ptr("returnflag"), ptr("linestatus")), lib)
        P = datagen.q6_params()
        if args.workload == "q6":
            def step(cols=bufs):
                return ops.q6_filter_sum(cols["shipdate"], cols["discount"], cols["quantity"], cols["extendedprice"], n, *P)
            rec_bytes = shard.Q6_REC_BYTES
            pack = lambda res: shard.pack_q6(res[0], res[1])
            merge = lambda buf: shard.merge_q6(buf, world)
        else:
            def step(cols=bufs):
                return ops.q1_group_agg(cols["shipdate"], cols["quantity"], cols["extendedprice"], cols["discount"], cols["tax"],
                                        cols["returnflag"], cols["linestatus"], n, datagen.Q1_CUTOFF)
            rec_bytes = shard.Q1_REC_BYTES
            pack = lambda res: shard.pack_q1(res, row0)
            merge = lambda buf: shard.merge_q1(buf, world)
        units = n
        unit_name = "rows/s"
        alg_bytes = wl["bytes_per_row"] * n
        h2d_bytes = int(alg_bytes)
    elif args.workload == "sum":
        dv = DeviceBuffer(8 * n, lib)
        capi.check(lib.MoB200_GenInt64(1, (row0 // 64) * 64, n, dv.ptr, None, 0), lib)
        bufs = {"col": dv}
        def step(cols=bufs):
            return ops.agg_sum(capi.T_INT64, cols["col"], None, n)
        rec_bytes = 16
        pack = lambda res: np.asarray([res[1], 0], dtype=np.int64).tobytes()
        merge = lambda buf: int(np.frombuffer(bytes(buf), dtype=np.int64).reshape(world, 2)[:, 0].sum())
        units, unit_name, alg_bytes, h2d_bytes = n, "rows/s", 8.0 * n, 8 * n
    elif args.workload == "ivf":
        # BASELINE config 5: entries = mixture of 1024 Gaussians, centroids = the generating means, assignment by argmin L2sq
        # (Productl2).  Lists are sharded across ranks (every rank holds whole lists of a contiguous row slice of the table and
        # the full centroid table); each rank answers every query over its lists; per-rank top-k are gathered and merged.
        dim, nq, k, nlist, nprobe = 768, args.queries, 10, 1024, 32
        n_local = n // world
        centers = datagen.vectors_f32(30, 0, nlist, dim) * np.float32(4.0)
        dcent = DeviceBuffer.from_numpy(centers, lib)
        raw = DeviceBuffer(4 * n_local * dim, lib)
        capi.check(lib.MoB200_GenVectorsF32(31, rank * n_local, n_local, dim, raw.ptr, dcent.ptr, nlist, 1.0), lib)
        ivf = ops.IvfflatSearchIndex.build(raw, n_local, centers, capi.METRIC_L2, lib)
        ivf.row_ids += rank * n_local                      # global primary keys
        ivf.d_ids.free(); ivf.d_ids = DeviceBuffer.from_numpy(ivf.row_ids, lib)
        raw.free(); dcent.free()
        dq = DeviceBuffer(4 * nq * dim, lib)
        capi.check(lib.MoB200_GenVectorsF32(32, 0, nq, dim, dq.ptr, ivf.d_cent.ptr, nlist, 1.0), lib)
        bufs = {"queries": dq}
        def step(cols=bufs):
            return ivf.search(cols["queries"], k, nprobe)
        rec_bytes = nq * k * 16
        pack = lambda res: shard.pack_topk(res[0], res[1])
        merge = lambda buf: ops.topk_merge(*shard.unpack_topk(buf, world, nq, k), nq, k)
        units, unit_name = nq, "queries/s"
        alg_bytes = None
        h2d_bytes = 4 * nq * dim
        n = n_local
    else:  # bruteforce: dataset rows sharded across ranks, every rank sees all queries
        dim, nq, k = 768, args.queries, 10
        n_local = n // world if world > 1 else n
        ds = DeviceBuffer(4 * n_local * dim, lib)
        capi.check(lib.MoB200_GenVectorsF32(20, rank * n_local, n_local, dim, ds.ptr, None, 0, 1.0), lib)
        dq = DeviceBuffer(4 * nq * dim, lib)
        capi.check(lib.MoB200_GenVectorsF32(21, 0, nq, dim, dq.ptr, None, 0, 1.0), lib)
        idx = ops.BruteForceIndex(ds, dim, {"l2": capi.METRIC_L2, "ip": capi.METRIC_IP, "cos": capi.METRIC_COS}[args.metric], key_base=rank * n_local, lib=lib)
        bufs = {"queries": dq}
        def step(cols=bufs):
            return idx.search(cols["queries"], k)
        rec_bytes = nq * k * 16
        pack = lambda res: shard.pack_topk(res[0], res[1])
        merge = lambda buf: ops.topk_merge(*shard.unpack_topk(buf, world, nq, k), nq, k)   # k-way merge kernel on every rank
        units, unit_name = nq, "queries/s"
        alg_bytes = None
        h2d_bytes = 4 * nq * dim
        n = n_local

    gather_buf = None
    search_dev = None
    if dist is not None:
        import torch
        send = torch.zeros(rec_bytes, dtype=torch.uint8, device="cuda")
        gather_buf = torch.zeros(rec_bytes * world, dtype=torch.uint8, device="cuda")
        if args.workload in ("bruteforce", "ivf"):
            # MergeTop seam kept on the device: every rank's (keys, distances) go from the search kernels' output buffers through
            # NCCL into the merge kernel; the host sees only the final top-k
            search_dev = {"k": torch.empty(nq * k, dtype=torch.int64, device="cuda"), "d": torch.empty(nq * k, dtype=torch.float64, device="cuda"),
                          "gk": torch.empty(world * nq * k, dtype=torch.int64, device="cuda"), "gd": torch.empty(world * nq * k, dtype=torch.float64, device="cuda"),
                          "ok": torch.empty(nq * k, dtype=torch.int64, device="cuda"), "od": torch.empty(nq * k, dtype=torch.float64, device="cuda")}
            host_step = step
            def step(cols=bufs):
                if cols is not bufs:          # e2e arm: host queries in, host results out
                    return host_step(cols)
                out = (search_dev["k"].data_ptr(), search_dev["d"].data_ptr())
                if args.workload == "ivf":
                    ivf.search(cols["queries"], k, nprobe, out=out)
                else:
                    idx.search(cols["queries"], k, out=out)
                return None

    def exchange(res):
        """the reduce seam (MergeGroup / MergeTop): all ranks receive every rank's partial record"""
        if dist is None:
            return
        import torch
        if search_dev is not None and res is None:
            dist.all_gather_into_tensor(search_dev["gk"], search_dev["k"])
            dist.all_gather_into_tensor(search_dev["gd"], search_dev["d"])
            torch.cuda.current_stream().synchronize()     # NCCL ran on torch's stream, the merge kernel runs on the library's
            ops.topk_merge_device(search_dev["gk"].data_ptr(), search_dev["gd"].data_ptr(), world, nq, k, search_dev["ok"].data_ptr(), search_dev["od"].data_ptr())
            return search_dev["ok"].cpu().numpy(), search_dev["od"].cpu().numpy()
        send.copy_(torch.frombuffer(bytearray(pack(res)), dtype=torch.uint8))
        dist.all_gather_into_tensor(gather_buf, send)
        return merge(gather_buf.cpu().numpy().tobytes())

    # ---------------------------------------------------------------------------------------------- resident timing
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()          # started before the warm-up: nvidia-smi needs ~100 ms to deliver its first sample
    for wi in range(W):
        merged = exchange(step())
        if wi == 0 and search_dev is not None:
            # the device-side MergeTop seam must give what the host-side one (shard.py, covered by the gloo CPU test) gives
            import torch
            hres = host_step(bufs)
            send.copy_(torch.frombuffer(bytearray(pack(hres)), dtype=torch.uint8))
            dist.all_gather_into_tensor(gather_buf, send)
            hk, hd = merge(gather_buf.cpu().numpy().tobytes())
            if not (np.array_equal(hk, merged[0]) and np.array_equal(hd, merged[1])):
                raise SystemExit("device-side top-k exchange disagrees with the host-side merge")
    if rank == 0:
        time.sleep(0.25)
    barrier_sync()
    t_region0 = time.time()
    launches0 = lib.MoB200_KernelLaunchCount()
    kernel_ms = []
    kms = C.c_float()
    capi.check(lib.MoB200_TimerStart(), lib)
    t_wall0 = time.perf_counter()
    for _ in range(K):
        res = step()
        lib.MoB200_LastKernelMs(C.byref(kms)); kernel_ms.append(kms.value)
        exchange(res)
    ms = C.c_float()
    capi.check(lib.MoB200_TimerStop(C.byref(ms)), lib)
    barrier_sync()
    wall_ms = (time.perf_counter() - t_wall0) * 1e3
    launches = lib.MoB200_KernelLaunchCount() - launches0
    clocks = sampler.stop(t_region0, time.time()) if rank == 0 else None
    total_ms = max_over_ranks(max(ms.value, 0.0))
    ms_per_step = total_ms / K
    # weak workloads: every rank processes `units` rows of its own; strong (search): all ranks work on the SAME `units` queries
    value = units * (1 if args.workload in ("bruteforce", "ivf") else world) * K / (total_ms * 1e-3)
    kern_ms = statistics.mean(kernel_ms)

    # ---------------------------------------------------------------------------------------------- e2e: host buffers
    e2e = None
    if not args.no_e2e and args.workload in ("q6", "q1", "sum"):
        try:
            avail = 0
            for ln in open("/proc/meminfo"):
                if ln.startswith("MemAvailable"):
                    avail = int(ln.split()[1]) * 1024
            need = sum(b.nbytes for b in bufs.values())
            n_e2e = n
            if avail and need * world * 1.5 > avail:
                n_e2e = max(1 << 20, int(n * (avail / (need * world * 1.5))) // 8192 * 8192)
            host = {}
            dts = {"shipdate": np.int32, "returnflag": np.uint8, "linestatus": np.uint8, "col": np.int64}
            for kname, b in bufs.items():
                dt = np.dtype(dts.get(kname, np.float64))
                pa = PinnedArray((n_e2e,), dt, lib)
                capi.check(lib.MoB200_Download(pa.ptr, b.ptr, n_e2e * dt.itemsize), lib)
                host[kname] = pa
            n_saved = n
            ke = min(K, 5)
            hcols = {kname: pa.array for kname, pa in host.items()}
            n = n_e2e     # step() closes over n
            for _ in range(1):
                exchange(step(hcols))
            barrier_sync()
            t0 = time.perf_counter()
            for _ in range(ke):
                exchange(step(hcols))
            barrier_sync()
            dt_e = max_over_ranks(time.perf_counter() - t0)
            n = n_saved
            per_row = need / n_saved
            e2e = {"value": n_e2e * world * ke / dt_e, "unit": unit_name, "h2d_bytes_per_step": int(per_row * n_e2e), "d2h_bytes_per_step": rec_bytes,
                   "steps": ke, "rows_per_gpu": n_e2e, "host_memory": "pinned (MoB200_HostAlloc)", "timer": "wall clock around the C-ABI calls, max over ranks"}
            for pa in host.values():
                pa.free()
        except Exception as ex:  # report, never fake
            e2e = {"value": None, "unit": unit_name, "error": str(ex)[:200]}
    elif args.workload in ("bruteforce", "ivf"):
        qhost = bufs["queries"].to_numpy(np.float32)
        t0 = time.perf_counter()
        ke = min(K, 3)
        for _ in range(ke):
            exchange(step({"queries": qhost}))
        barrier_sync()
        dt_e = max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": units * ke / dt_e, "unit": unit_name, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": rec_bytes, "steps": ke,
               "note": "dataset resident (index built once, as the reference keeps it in memory); queries from host, keys+distances to host"}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    # ---------------------------------------------------------------------------------------------- roofline + cpu baseline
    peak, peak_src = measured_peaks()
    if args.workload == "ivf":
        # dominant kernel = tc_candidates_kernel over (list, 128-query tile) units: 2 * K' flop per (query, probed row) pair,
        # K' = 3 * dim (hi/lo operand split); pairs = queries * nprobe * mean list length
        pairs = float(args.queries) * 32 * (n / 1024.0)
        kused = int(lib.MoB200_SetTuning(b"get_tc_kused", 0)) or 3 * 768
        flop = 2.0 * pairs * kused
        try:
            mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            tpeak, tsrc = float(mp["bf16_tflops"]), "measured (MEASURED_PEAKS.json bf16_tflops, burst)"
        except Exception:
            tpeak, tsrc = 1590.0, "fallback (B200_PROFILING.md 1.59 PFLOP/s)"
        ach = flop / (kern_ms * 1e-3) / 1e12
        roofline = {"bound": "tensor", "achieved": ach, "peak": tpeak, "unit": "TFLOP/s", "frac": ach / tpeak,
                    "traffic": ncu_traffic("tc_ivf", n == 1_250_000 and args.queries == 10_000),
                    "kernel": "tc_candidates_kernel (tcgen05 bf16, K = %d per pair) over (list, query-tile) units" % kused, "kernel_ms": kern_ms, "algorithmic_flop_per_launch": flop,
                    "peak_source": tsrc, "traffic_source": "profiles/r01_ncu_summary.md (ncu --set full capture of this command; null at other sizes)",
                    "tc_refined_queries": int(lib.MoB200_SetTuning(b"get_tc_refined", 0)),
                    "tc_fallback_queries": int(lib.MoB200_SetTuning(b"get_tc_fallbacks", 0)),
                    "note": "useful flop only: tiles are padded to 128 queries x 256 rows, so the tensor pipe does more work than counted"}
    elif alg_bytes is not None:
        achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": ncu_traffic({"q6": "q6", "q1": "q1", "sum": None}[args.workload], args.rows == 0) if args.workload != "sum" else None,
                    "kernel": {"q6": "q6_kernel", "q1": "q1_kernel", "sum": "agg_kernel"}[args.workload], "kernel_ms": kern_ms,
                    "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src}
    else:
        # dominant kernel = tc_candidates_kernel: one bf16 GEMM with K' = 3 * dim (hi/lo operand split), 2 * Q * N * K' flop
        # K elements the timed candidate pass multiplies per (query, row): 768 = hi.hi only (one-term level), 2304 = three-term product
        kused = int(lib.MoB200_SetTuning(b"get_tc_kused", 0)) or 3 * 768
        flop = 2.0 * args.queries * n * kused
        tpeak, tsrc = 1709.9, "fallback"
        try:
            mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            tpeak, tsrc = float(mp["bf16_tflops"]), "measured (MEASURED_PEAKS.json bf16_tflops, burst: the kernel is timed alone)"
        except Exception:
            tpeak, tsrc = 1590.0, "fallback (B200_PROFILING.md 1.59 PFLOP/s)"
        ach = flop / (kern_ms * 1e-3) / 1e12
        roofline = {"bound": "tensor", "achieved": ach, "peak": tpeak, "unit": "TFLOP/s", "frac": ach / tpeak,
                    "traffic": ncu_traffic("tc_bruteforce", args.rows == 0 and args.queries == 10_000 and args.metric == "l2" and world == 1),
                    "kernel": "tc_candidates_kernel (tcgen05 bf16, K = %d per pair)" % kused, "kernel_ms": kern_ms, "algorithmic_flop_per_launch": flop, "peak_source": tsrc,
                    "tc_refined_queries": int(lib.MoB200_SetTuning(b"get_tc_refined", 0)),
                    "tc_fallback_queries": int(lib.MoB200_SetTuning(b"get_tc_fallbacks", 0))}
    cpu_baseline = None
    if not args.no_cpu and world == 1 and args.workload == "bruteforce":
        threads = os.cpu_count() or 1
        hds = ds.to_numpy(np.float32).reshape(n, 768)
        hqs = bufs["queries"].to_numpy(np.float32).reshape(-1, 768)[:threads]
        v, sec = cpu_reference("bruteforce", n, 1, 1, threads, hds, hqs)
        cpu_baseline = {"value": v, "unit": "queries/s", "cores": threads, "kind": "port",
                        "sample": "full dataset, %d queries (one per host thread), 1 pass after a quarter-size warm-up; oracle/oracle_go.c GoBruteForceIndex.Search" % hqs.shape[0]}
    if not args.no_cpu and world == 1 and args.workload == "ivf":
        threads = os.cpu_count() or 1
        ents = ivf.d_data.to_numpy(np.float32).reshape(ivf.n, 768)                  # list-ordered entries
        assign = np.repeat(np.arange(1024, dtype=np.int32), np.diff(ivf.offsets))     # their list ids
        cents = ivf.d_cent.to_numpy(np.float32).reshape(1024, 768)
        hqs = bufs["queries"].to_numpy(np.float32).reshape(-1, 768)[:8 * threads]
        v, sec = cpu_reference("ivf", ivf.n, 1, 1, threads, (ents, assign, cents), hqs)
        cpu_baseline = {"value": v, "unit": "queries/s", "cores": threads, "kind": "port",
                        "sample": "the GPU's own shard (%d x 768, nlist 1024, nprobe 32), %d queries over %d host threads, 1 pass after a quarter-size warm-up; oracle/oracle_go.c og_ivf_search_f32" % (ivf.n, hqs.shape[0], threads)}
    if not args.no_cpu and world == 1 and args.workload in ("q6", "q1", "sum"):
        threads = os.cpu_count() or 1
        sample = min(n, 1 << 25)
        v, sec = cpu_reference(args.workload, sample, 3, 1, threads)
        cpu_baseline = {"value": v, "unit": "rows/s", "cores": threads, "kind": "port",
                        "sample": "first %d rows, 3 passes after 1 warm-up; oracle/oracle_go.c operator chain on %d pthreads" % (sample, threads)}

    line = {
        "metric": wl["metric"], "value": value, "unit": unit_name, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak" if args.workload not in ("bruteforce", "ivf") else "strong", "vs_baseline": None,
        "dtype": {"q6": "f64", "q1": "f64", "sum": "int64", "bruteforce": "f32", "ivf": "f32"}[args.workload], "data": "synthetic",
        "config": {"workload": wl["name"], "rows_per_gpu": n, "l2": "inputs (%.1f GB per GPU) are larger than the 126 MB L2; no flush needed" % ((alg_bytes or 4.0 * n * 768) / 1e9),
                   "parallelism": "block-range shards x%d, NCCL all_gather of partial aggregates" % world if world > 1 else "1 GPU",
                   "timer": "CUDA events on the library stream (MoB200_TimerStart/Stop), max over ranks", "wall_ms_rank0": wall_ms},
        "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
    }
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    # stdout carries exactly ONE JSON line: library chatter (NCCL version banners, torchrun notices) is sent to stderr by
    # pointing fd 1 at fd 2 for the run and printing the result line to the saved descriptor.
    sys.stdout.flush()
    _real_stdout = os.dup(1)
    os.dup2(2, 1)
    _buf = []
    _print = print

    def print(*a, **k):  # noqa: A001  (only the JSON line goes through here)
        _buf.append(" ".join(str(x) for x in a))

    rc = main()
    sys.stdout.flush()
    os.dup2(_real_stdout, 1)
    for line in _buf:
        os.write(1, (line + "\n").encode())
    sys.exit(rc)
