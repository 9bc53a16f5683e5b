"""GPU parity: single-column aggregates (config 1 of BASELINE.json and the aggexec family) against the oracle."""
import numpy as np
import pytest

import oracle_lib as O
from matrixone_b200 import capi, datagen, ops
from matrixone_b200.vector import DeviceBuffer, bitmap_from_bools

pytestmark = pytest.mark.gpu

I64MAX = np.iinfo(np.int64).max
I64MIN = np.iinfo(np.int64).min


def _oracle_sum_i64(T, v, nulls):
    s = np.zeros(1, dtype=np.int64); nul = np.ones(1, dtype=np.uint8); c = np.zeros(1, dtype=np.int64)
    rc = O.go().og_sum_int64(T, O.p(v), O.p(nulls), 0, None, len(v), O.p(s), O.p(nul), O.p(c), None)
    return rc, int(s[0]), bool(nul[0]), int(c[0])


@pytest.mark.parametrize("n", [0, 1, 5, 63, 64, 65, 8191, 8192, 1_000_003])
def test_sum_signed_bit_exact(gpu, n):
    rng = np.random.default_rng(n + 1)
    for T, dt in ((capi.T_INT8, np.int8), (capi.T_INT16, np.int16), (capi.T_INT32, np.int32), (capi.T_INT64, np.int64)):
        info = np.iinfo(dt)
        v = rng.integers(max(info.min, -2**31), min(info.max, 2**31 - 1), size=n, dtype=dt, endpoint=True)
        for nulls in (None, bitmap_from_bools(rng.random(n) < 0.05)):
            rc0, s0, nul0, c0 = _oracle_sum_i64(T, v, nulls)
            rc, s, nul = ops.agg_sum(T, v, nulls, n)
            assert (rc, nul) == (rc0, nul0) and (nul or s == s0), (T, n)
            assert ops.agg_count(T, v, nulls, n) == c0
            rca, avg, nula = ops.agg_avg(T, v, nulls, n)
            assert nula == nul0 and (nula or avg == float(s0) / float(c0))


def test_config1_int64_sum_10m_rows_device_resident(gpu):
    """BASELINE config 1: int64 SUM over a 10 M-row vector via XCall; values uniform in [-2^31, 2^31), +/- 5 % nulls"""
    n = 10_000_000
    v, nm = datagen.int64_column(1, 0, n, 50)
    nulls = bitmap_from_bools(nm)
    dv = DeviceBuffer(8 * n); dn = DeviceBuffer(8 * ((n + 63) // 64))
    capi.check(gpu.MoB200_GenInt64(1, 0, n, dv.ptr, dn.ptr, 50))
    assert (dv.to_numpy(np.int64) == v).all() and (dn.to_numpy(np.uint64) == nulls).all()      # device generator == numpy twin
    for nl_host, nl_dev in ((None, None), (nulls, dn)):
        rc0, s0, nul0, c0 = _oracle_sum_i64(capi.T_INT64, v, nl_host)
        assert ops.agg_sum(capi.T_INT64, dv, nl_dev, n) == (rc0, s0, nul0)       # resident path
        assert ops.agg_sum(capi.T_INT64, v, nl_host, n) == (rc0, s0, nul0)       # host path (staged)
    dv.free(); dn.free()


def test_sum_int64_overflow_follows_serial_prefix_rule(gpu):
    """int64OfCheck fires on the RUNNING sum in row order (sumavg2.go:89-94,139-164)"""
    cases = [
        np.asarray([I64MAX, 1, -5], dtype=np.int64),            # prefix overflows, total fits  -> error
        np.asarray([I64MAX, -5, 1], dtype=np.int64),            # same multiset, safe order     -> ok
        np.asarray([I64MIN, -1, 10], dtype=np.int64),
        np.asarray([I64MIN, 10, -1], dtype=np.int64),
        np.asarray([I64MAX, I64MAX, I64MIN, I64MIN], dtype=np.int64),
        np.asarray([I64MAX, I64MIN, I64MAX, I64MIN], dtype=np.int64),
        np.asarray([I64MIN, I64MIN], dtype=np.int64),           # wraps to 0: "sum >= 0" branch
    ]
    rng = np.random.default_rng(8)
    big = rng.integers(-2**62, 2**62, size=300_000, dtype=np.int64)
    cases += [big, np.sort(big), np.sort(big)[::-1].copy()]
    for v in cases:
        for nulls in (None, bitmap_from_bools(np.arange(len(v)) % 7 == 3)):
            rc0, s0, nul0, _ = _oracle_sum_i64(capi.T_INT64, v, nulls)
            rc, s, nul = ops.agg_sum(capi.T_INT64, v, nulls)
            assert rc == rc0, (v[:4], rc, rc0)
            if rc0 == 0:
                assert s == s0 and nul == nul0


def test_sum_unsigned_and_overflow(gpu):
    rng = np.random.default_rng(2)
    for T, dt in ((capi.T_UINT8, np.uint8), (capi.T_UINT16, np.uint16), (capi.T_UINT32, np.uint32), (capi.T_UINT64, np.uint64)):
        v = rng.integers(0, min(np.iinfo(dt).max, 2**40), size=100_001, dtype=dt, endpoint=True)
        nulls = bitmap_from_bools(rng.random(len(v)) < 0.1)
        s = np.zeros(1, dtype=np.uint64); nul = np.ones(1, dtype=np.uint8)
        rc0 = O.go().og_sum_uint64(T, O.p(v), O.p(nulls), 0, None, len(v), O.p(s), O.p(nul), None, None)
        assert ops.agg_sum(T, v, nulls) == (rc0, int(s[0]), bool(nul[0]))
    v = np.asarray([2**63, 2**63, 5], dtype=np.uint64)
    assert ops.agg_sum(capi.T_UINT64, v)[0] == capi.RC_OUT_OF_RANGE


@pytest.mark.parametrize("n", [1, 100, 8192, 777_777])
def test_sum_float_within_1e5_and_deterministic(gpu, n):
    rng = np.random.default_rng(n)
    for T, dt in ((capi.T_FLOAT32, np.float32), (capi.T_FLOAT64, np.float64)):
        v = (rng.standard_normal(n) * 1e3).astype(dt) + dt(500)
        nulls = bitmap_from_bools(rng.random(n) < 0.05)
        s = np.zeros(1, dtype=np.float64); nul = np.ones(1, dtype=np.uint8); c = np.zeros(1, dtype=np.int64)
        O.go().og_sum_float64(T, O.p(v), O.p(nulls), 0, None, n, O.p(s), O.p(nul), O.p(c))
        rc, got, isnull = ops.agg_sum(T, v, nulls)
        assert rc == 0 and isnull == bool(nul[0])
        if not isnull:
            assert abs(got - s[0]) <= 1e-5 * abs(s[0]) + 1e-9      # tolerance stated by north_star: 1e-5 relative
            live = np.ascontiguousarray(v.astype(np.float64)[~np.unpackbits(nulls.view(np.uint8), bitorder="little")[:n].astype(bool)])
            kahan = O.go().og_kahan_sum(O.p(live), int(c[0]))
            assert abs(got - kahan) <= 1e-9 * abs(kahan) + 1e-9    # and far tighter against a compensated sum
            assert ops.agg_sum(T, v, nulls)[1] == got              # bitwise run-to-run determinism
            rca, avg, _ = ops.agg_avg(T, v, nulls)
            assert abs(avg - s[0] / c[0]) <= 1e-5 * abs(avg)


def test_all_null_and_empty(gpu):
    v = np.arange(100, dtype=np.int64)
    nulls = bitmap_from_bools(np.ones(100, dtype=bool))
    assert ops.agg_sum(capi.T_INT64, v, nulls) == (0, 0, True)       # SUM of an all-null group is NULL (sumavg2.go:160-162)
    assert ops.agg_count(capi.T_INT64, v, nulls) == 0
    assert ops.agg_min(capi.T_INT64, v, nulls)[1] is True
    assert ops.agg_avg(capi.T_FLOAT64, v.astype(np.float64), nulls)[2] is True


MM_TYPES = [(capi.T_INT8, np.int8), (capi.T_INT16, np.int16), (capi.T_INT32, np.int32), (capi.T_INT64, np.int64),
            (capi.T_UINT8, np.uint8), (capi.T_UINT16, np.uint16), (capi.T_UINT32, np.uint32), (capi.T_UINT64, np.uint64),
            (capi.T_FLOAT32, np.float32), (capi.T_FLOAT64, np.float64), (capi.T_DATE, np.int32), (capi.T_TIMESTAMP, np.int64)]


@pytest.mark.parametrize("n", [1, 1000, 300_007])
def test_min_max_bit_exact(gpu, n):
    rng = np.random.default_rng(n + 5)
    for T, dt in MM_TYPES:
        if np.issubdtype(dt, np.floating):
            v = rng.standard_normal(n).astype(dt)
        else:
            info = np.iinfo(dt)
            v = rng.integers(info.min, info.max, size=n, dtype=dt, endpoint=True)
        nulls = bitmap_from_bools(rng.random(n) < 0.2)
        for is_max, fn in ((0, ops.agg_min), (1, ops.agg_max)):
            for nl in (None, nulls):
                ag = np.zeros(1, dtype=dt); nul = np.ones(1, dtype=np.uint8)
                O.go().og_minmax(is_max, T, O.p(v), O.p(nl), 0, None, n, O.p(ag), O.p(nul))
                got, isnull = fn(T, v, nl)
                assert isnull == bool(nul[0]) and (isnull or got == ag[0]), (T, is_max)


def test_min_with_nan_follows_go_first_value_rule(gpu):
    """minmax2.go:69-75: the first non-null value initialises and `value < agg` is false against NaN"""
    for v in (np.asarray([np.nan, 1.0, -2.0]), np.asarray([3.0, np.nan, -2.0, np.nan]), np.asarray([np.nan, np.nan])):
        ag = np.zeros(1); nul = np.ones(1, dtype=np.uint8)
        O.go().og_minmax(0, capi.T_FLOAT64, O.p(v), None, 0, None, len(v), O.p(ag), O.p(nul))
        got, _ = ops.agg_min(capi.T_FLOAT64, v)
        assert (np.isnan(got) and np.isnan(ag[0])) or got == ag[0]
