"""Concurrent callers: the reference calls the cgo surface from one pipeline goroutine per core, i.e. from arbitrary OS threads at the
same time (pkg/sql/compile/scope.go:442-499; cgo pins the goroutine to its thread for the call).  16 OS threads hammer a mix of entry
points for a few seconds -- every thread owns its stream, arena and staging buffer inside the library (csrc/runtime.cu) -- and every
single result is compared with the oracle answer computed up front.  ctypes releases the GIL for the duration of each foreign call, so
the calls really overlap inside libmo_b200.so."""
import threading
import time

import numpy as np
import pytest

import oracle_lib as O
from matrixone_b200 import capi, datagen, ops
from matrixone_b200.vector import Vector, bitmap_from_bools, xcall

pytestmark = pytest.mark.gpu
NTHREADS = 16
SECONDS = 4.0


def _work_items():
    """(name, callable -> result, expected) triples; inputs are host memory, shared read-only between the threads"""
    rng = np.random.default_rng(7)
    items = []
    n = 8192                                                      # one block, the unit the pipeline hands over
    a = rng.integers(-10 ** 9, 10 ** 9, n).astype(np.int64); b = rng.integers(-10 ** 9, 10 ** 9, n).astype(np.int64)
    nulls = bitmap_from_bools(rng.random(n) < 0.1)

    def go_add():
        r = np.zeros(n, dtype=np.int64); rn = np.zeros(n // 64, dtype=np.uint64)
        prm = np.zeros(2, dtype=np.int64); prm[1] = -1
        xcall(capi.XCALL_GO_ARITH(0, capi.T_INT64), [Vector(data=r, nulls=rn, length=n), Vector(data=a, nulls=nulls, length=n), Vector(data=b, length=n), Vector(data=prm.view(np.uint8), length=n)], n)
        return r, rn
    r0 = np.zeros(n, dtype=np.int64); rn0 = np.zeros(n // 64, dtype=np.uint64)
    O.go().og_arith(0, capi.T_INT64, O.p(r0), O.p(a), O.p(b), n, 0, 0, O.p(nulls), None, O.p(rn0), 0, None)
    live = ~np.unpackbits(rn0.view(np.uint8), bitorder="little")[:n].astype(bool)
    items.append(("go_arith", go_add, lambda res: np.array_equal(res[1], rn0) and np.array_equal(res[0][live], r0[live])))

    def go_lt():
        r = np.zeros(n, dtype=np.uint8); rn = np.zeros(n // 64, dtype=np.uint64)
        xcall(capi.XCALL_GO_COMPARE(4, capi.T_INT64), [Vector(data=r, nulls=rn, length=n), Vector(data=a, length=n), Vector(data=b, length=n)], n)
        return r
    items.append(("go_compare", go_lt, lambda res: np.array_equal(res, (a < b).astype(np.uint8))))

    col = rng.integers(-1000, 1000, 300_000).astype(np.int32)
    want_sum = int(col.astype(np.int64).sum())
    items.append(("agg_sum", lambda: ops.agg_sum(capi.T_INT32, col), lambda res: res == (0, want_sum, False)))

    cols = datagen.lineitem(10, 0, 200_000)
    P = datagen.q6_params()
    want6 = O.q6(cols, 200_000, P)
    items.append(("q6", lambda: ops.q6_filter_sum(cols["shipdate"], cols["discount"], cols["quantity"], cols["extendedprice"], 200_000, *P),
                  lambda res: res[1] == want6[1] and abs(res[0] - want6[0]) <= 1e-11 * abs(want6[0])))
    want1 = O.q1(cols, 200_000, datagen.Q1_CUTOFF)
    q1cols = [cols[k] for k in ("shipdate", "quantity", "extendedprice", "discount", "tax", "returnflag", "linestatus")]
    items.append(("q1", lambda: ops.q1_group_agg(*q1cols, 200_000, datagen.Q1_CUTOFF),
                  lambda res: [(g["returnflag"], g["count_order"]) for g in res] == [(g["returnflag"], g["count_order"]) for g in want1]))

    v = (rng.random(100_000) < 0.3).astype(np.uint8)
    want_sels = np.flatnonzero(v)
    items.append(("filter_sels", lambda: ops.filter_sels(v), lambda res: np.array_equal(res, want_sels)))

    ds = rng.standard_normal((3000, 64)).astype(np.float32); qs = rng.standard_normal((32, 64)).astype(np.float32)
    wk, wd = O.bruteforce(ds, qs, 5)
    def bf():
        idx = ops.BruteForceIndex(ds, 64)
        try:
            return idx.search(qs, 5)
        finally:
            idx.destroy()
    items.append(("bruteforce", bf, lambda res: np.array_equal(res[1], wd) and np.array_equal(res[0], wk)))

    key = rng.integers(0, 50, 100_000).astype(np.int32); x = rng.standard_normal(100_000)
    p = ops.FusedPlan([capi.T_INT32, capi.T_FLOAT64]).group_by(0).agg(capi.AGG_COUNT, -1).agg(capi.AGG_MAX, 1)
    cnt = np.bincount(key, minlength=50); mx = np.full(50, -np.inf); np.maximum.at(mx, key, x)
    def plan():
        return {g["key"]: (g["rows"], g["aggs"][1][0]) for g in p.run([key, x], 100_000, max_groups=64)}
    items.append(("plan", plan, lambda res: all(res[k] == (cnt[k], mx[k]) for k in range(50))))
    return items


def test_sixteen_threads_mixed_xcalls_match_oracle(gpu):
    items = _work_items()
    for name, fn, ok in items:      # single-threaded sanity first
        assert ok(fn()), name
    errors, counts = [], [0] * NTHREADS
    deadline = time.time() + SECONDS

    def worker(tid):
        rng = np.random.default_rng(tid)
        try:
            while time.time() < deadline:
                name, fn, ok = items[int(rng.integers(0, len(items)))]
                if not ok(fn()):
                    errors.append("thread %d: %s returned a wrong result" % (tid, name))
                    return
                counts[tid] += 1
        except Exception as ex:  # noqa: BLE001
            errors.append("thread %d: %r" % (tid, ex))

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(NTHREADS)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[:3]
    assert min(counts) >= 3 and sum(counts) >= 20 * NTHREADS, counts      # every thread made progress
