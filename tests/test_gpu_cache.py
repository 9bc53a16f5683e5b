"""Device column cache (MoB200_ColumnPin): a host range declared immutable is uploaded once; afterwards every entry point that is handed a host
pointer inside it reads the device copy ('columns must be cached/resident on device across calls, or the numbers
are H2D-bound').  Results must be identical with and without the cache; a new generation replaces the copy; eviction is LRU."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as O
from matrixone_b200 import capi, datagen, ops

pytestmark = pytest.mark.gpu


def stats(lib):
    h, m, b = C.c_uint64(), C.c_uint64(), C.c_uint64()
    lib.MoB200_ColumnCacheStats(C.byref(h), C.byref(m), C.byref(b))
    return h.value, m.value, b.value


def test_pinned_columns_are_used_by_host_pointer_calls(gpu):
    n = 3_000_000
    cols = datagen.lineitem(10, 0, n)
    P = datagen.q6_params()
    names = ("shipdate", "discount", "quantity", "extendedprice")
    want = ops.q6_filter_sum(*[cols[k] for k in names], n, *P)
    try:
        capi.check(gpu.MoB200_ColumnCacheConfigure(1 << 30))
        for i, k in enumerate(names):
            capi.check(gpu.MoB200_ColumnPin(cols[k].ctypes.data, cols[k].nbytes, 1))
        h0, m0, b0 = stats(gpu)
        assert b0 == sum(cols[k].nbytes for k in names)
        assert ops.q6_filter_sum(*[cols[k] for k in names], n, *P) == want
        h1, m1, _ = stats(gpu)
        assert h1 - h0 == 4                                            # all four columns came from the cache
        # a sub-range of a pinned column (a block inside it) hits too
        sub = cols["quantity"][8192:8192 * 3]
        assert ops.agg_sum(capi.T_FLOAT64, sub)[1] == float(sub.sum())
        assert stats(gpu)[0] - h1 == 1
        # the host buffer changes, the caller announces a new generation
        cols["discount"][:] = 0.03
        capi.check(gpu.MoB200_ColumnPin(cols["discount"].ctypes.data, cols["discount"].nbytes, 2))
        got = ops.q6_filter_sum(*[cols[k] for k in names], n, *P)
        o = O.q6(cols, n, P)
        assert got[1] == o[1] and abs(got[0] - o[0]) <= 1e-11 * abs(o[0]) and got != want
        # unpin: back to staging, same result
        for k in names:
            capi.check(gpu.MoB200_ColumnUnpin(cols[k].ctypes.data))
        assert stats(gpu)[2] == 0
        assert ops.q6_filter_sum(*[cols[k] for k in names], n, *P) == got
    finally:
        gpu.MoB200_ColumnCacheConfigure(0)


def test_cache_lru_eviction_and_capacity(gpu):
    a = [np.full(1 << 18, i, dtype=np.int64) for i in range(4)]       # 2 MiB each
    try:
        capi.check(gpu.MoB200_ColumnCacheConfigure(5 << 20))           # room for two
        for i in range(3):
            capi.check(gpu.MoB200_ColumnPin(a[i].ctypes.data, a[i].nbytes, 0))
        assert stats(gpu)[2] == 4 << 20                                # the oldest was evicted
        h0 = stats(gpu)[0]
        assert ops.agg_sum(capi.T_INT64, a[0])[1] == 0                 # staged, not cached
        assert ops.agg_sum(capi.T_INT64, a[2])[1] == 2 << 18
        assert stats(gpu)[0] - h0 == 1
        capi.check(gpu.MoB200_ColumnPin(a[3].ctypes.data, 64 << 20, 0) if False else 0)
        big = np.zeros(1 << 20, dtype=np.int64)                        # 8 MiB > capacity: not cached, calls keep working
        capi.check(gpu.MoB200_ColumnPin(big.ctypes.data, big.nbytes, 0))
        assert stats(gpu)[2] == 4 << 20
        assert ops.agg_sum(capi.T_INT64, big)[1] == 0
    finally:
        gpu.MoB200_ColumnCacheConfigure(0)
