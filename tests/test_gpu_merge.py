"""GPU parity of the device-resident (asynchronous) results and of the MergeGroup kernels (mergeGroup.go:132-247): the partial
record a scan writes to caller-owned device memory, NCCL-gatherable, and the on-device fold of such records.  Checked against the
synchronous API (itself checked against the oracle in test_gpu_tpch.py / test_gpu_agg.py), against the host-side merge in
matrixone_b200/shard.py, and against the oracle's multi-worker pipelines (whose partial states are merged in worker order)."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as O
from matrixone_b200 import capi, datagen, ops, shard
from matrixone_b200.vector import DeviceBuffer

pytestmark = pytest.mark.gpu


def dev_lineitem(lib, seed, row0, n):
    names = ("shipdate", "quantity", "extendedprice", "discount", "tax", "returnflag", "linestatus")
    size = {"shipdate": 4, "returnflag": 1, "linestatus": 1}
    bufs = {k: DeviceBuffer(size.get(k, 8) * max(n, 1)) for k in names}
    capi.check(lib.MoB200_GenLineitem(seed, row0, n, *[bufs[k].ptr for k in names]))
    return bufs


def free(bufs):
    for b in bufs.values():
        b.free()


@pytest.mark.parametrize("n", [1, 1000, 8193, 2_000_001])
def test_q6_device_result_equals_synchronous_result(gpu, n):
    bufs = dev_lineitem(gpu, 10, 0, n)
    P = datagen.q6_params()
    want = ops.q6_filter_sum(bufs["shipdate"], bufs["discount"], bufs["quantity"], bufs["extendedprice"], n, *P)
    out = DeviceBuffer(16); nul = DeviceBuffer(8)
    ops.q6_filter_sum_device(bufs["shipdate"], bufs["discount"], bufs["quantity"], bufs["extendedprice"], n, *P, out_ptr=out.ptr, out_nulls_ptr=nul.ptr)
    raw = out.to_numpy(np.float64)           # MoB200_Download is stream-ordered behind the kernel
    assert raw[0] == want[0] and int(raw.view(np.int64)[1]) == want[1]
    assert bool(nul.to_numpy(np.uint64)[0] & np.uint64(1)) == want[2]
    free(bufs); out.free(); nul.free()


def test_q6_merge_kernel_equals_host_merge_and_global_scan(gpu):
    n, world = 1_500_007, 4
    P = datagen.q6_params()
    parts_dev = DeviceBuffer(16 * world)
    parts = []
    for r in range(world):
        r0, r1 = shard.block_range(r, world, n)
        b = dev_lineitem(gpu, 10, r0, r1 - r0)
        ops.q6_filter_sum_device(b["shipdate"], b["discount"], b["quantity"], b["extendedprice"], r1 - r0, *P, out_ptr=parts_dev.ptr + 16 * r)
        parts.append(ops.q6_filter_sum(b["shipdate"], b["discount"], b["quantity"], b["extendedprice"], r1 - r0, *P))
        free(b)
    fin = DeviceBuffer(16)
    ops.q6_merge_device(parts_dev.ptr, world, fin.ptr)
    raw = fin.to_numpy(np.float64)
    got = (float(raw[0]), int(raw.view(np.int64)[1]))
    packed = b"".join(shard.pack_q6(s, c) for s, c, _ in parts)
    hs, hc, hnull = shard.merge_q6(packed, world)
    assert got == (hs, hc)                                            # device fold == host fold, bit for bit (same order)
    assert ops.q6_merge([(s, c) for s, c, _ in parts]) == (hs, hc, hnull)      # host-pointer form of the same XCall
    cols = datagen.lineitem(10, 0, n)
    want, ns, _ = O.q6(cols, n, P, nthreads=world)                    # the oracle's own block-range workers + MergeGroup
    assert hc == ns and abs(hs - want) <= 1e-11 * abs(want)
    # an empty partial is NULL and skipped
    assert ops.q6_merge([(0.0, 0), (2.5, 3), (0.0, 0)]) == (2.5, 3, False)
    assert ops.q6_merge([(0.0, 0)])[2] is True
    parts_dev.free(); fin.free()


@pytest.mark.parametrize("n", [1, 64, 100_003, 3_000_001])
def test_q1_device_result_equals_synchronous_result(gpu, n):
    bufs = dev_lineitem(gpu, 10, 0, n)
    cut = datagen.Q1_CUTOFF
    args = [bufs[k] for k in ("shipdate", "quantity", "extendedprice", "discount", "tax", "returnflag", "linestatus")]
    want = ops.q1_group_agg(*args, n, cut, row_base=77)
    out = DeviceBuffer(ops.Q1_RESULT_BYTES)
    ops.q1_group_agg_device(*args, n, cut, out.ptr, row_base=77)
    got = ops.q1_result_from_bytes(out.to_numpy(np.uint8).tobytes())
    assert got == want
    assert all(g["first_row"] >= 77 for g in got)
    free(bufs); out.free()


def test_q1_device_result_wide_retry_and_too_many_groups(gpu):
    """5..8 distinct keys take the gated 8-slot retry; more than 8 report ngroups = -1 on the device (MoError on the host form)"""
    n = 50_000
    rng = np.random.default_rng(3)
    cols = datagen.lineitem(10, 0, n)
    for nkeys in (6, 8, 9):
        rf = (65 + rng.integers(0, nkeys, n)).astype(np.uint8)
        ls = np.full(n, ord("F"), dtype=np.uint8)
        dev = {k: DeviceBuffer.from_numpy(v) for k, v in cols.items()}
        dev["returnflag"].free(); dev["returnflag"] = DeviceBuffer.from_numpy(rf)
        dev["linestatus"].free(); dev["linestatus"] = DeviceBuffer.from_numpy(ls)
        args = [dev[k] for k in ("shipdate", "quantity", "extendedprice", "discount", "tax", "returnflag", "linestatus")]
        out = DeviceBuffer(ops.Q1_RESULT_BYTES)
        ops.q1_group_agg_device(*args, n, datagen.Q1_CUTOFF, out.ptr)
        raw = out.to_numpy(np.uint8).tobytes()
        if nkeys <= 8:
            got = ops.q1_result_from_bytes(raw)
            want = ops.q1_group_agg(*args, n, datagen.Q1_CUTOFF)
            assert got == want and len(got) == nkeys
            hc = dict(cols); hc["returnflag"] = rf; hc["linestatus"] = ls
            ores = O.q1(hc, n, datagen.Q1_CUTOFF)
            assert [(g["returnflag"], g["count_order"], g["first_row"]) for g in got] == [(g["returnflag"], g["count_order"], g["first_row"]) for g in ores]
        else:
            assert capi.Q1Result.from_buffer_copy(raw).ngroups == -1
            with pytest.raises(capi.MoError):
                ops.q1_group_agg(*args, n, datagen.Q1_CUTOFF)
        free(dev); out.free()


def test_q1_merge_kernel_equals_host_merge_and_global_scan(gpu):
    n, world = 2_000_003, 8
    cut = datagen.Q1_CUTOFF
    RB = ops.Q1_RESULT_BYTES
    parts_dev = DeviceBuffer(RB * world)
    host_parts = []
    for r in range(world):
        r0, r1 = shard.block_range(r, world, n)
        b = dev_lineitem(gpu, 10, r0, r1 - r0)
        args = [b[k] for k in ("shipdate", "quantity", "extendedprice", "discount", "tax", "returnflag", "linestatus")]
        ops.q1_group_agg_device(*args, r1 - r0, cut, parts_dev.ptr + RB * r, row_base=r0)
        host_parts.append((ops.q1_group_agg(*args, r1 - r0, cut), r0))
        free(b)
    fin = DeviceBuffer(RB)
    ops.q1_merge_device(parts_dev.ptr, world, fin.ptr)
    got = ops.q1_result_from_bytes(fin.to_numpy(np.uint8).tobytes())
    hm = shard.merge_q1(b"".join(shard.pack_q1(g, r0) for g, r0 in host_parts), world)
    assert len(got) == len(hm)
    for a, b in zip(got, hm):
        for f in ("returnflag", "linestatus", "count_order", "first_row", "sum_qty", "sum_base_price", "sum_disc_price", "sum_charge", "sum_disc", "avg_qty", "avg_price", "avg_disc"):
            assert a[f] == b[f], f                                     # same association (rank order) => bitwise equal
    assert ops.q1_merge(parts_dev.to_numpy(np.uint8).tobytes(), world) == got
    cols = datagen.lineitem(10, 0, n)
    ores = O.q1(cols, n, cut, nthreads=world)
    assert [(g["returnflag"], g["linestatus"], g["count_order"], g["first_row"]) for g in got] == [(g["returnflag"], g["linestatus"], g["count_order"], g["first_row"]) for g in ores]
    for a, b in zip(got, ores):
        for f in ("sum_qty", "sum_base_price", "sum_disc_price", "sum_charge", "avg_qty", "avg_price", "avg_disc"):
            assert abs(a[f] - b[f]) <= 1e-11 * abs(b[f])
    parts_dev.free(); fin.free()


TYPES = [capi.T_INT8, capi.T_INT32, capi.T_INT64, capi.T_UINT16, capi.T_UINT64, capi.T_FLOAT32, capi.T_FLOAT64]


@pytest.mark.parametrize("T", TYPES)
@pytest.mark.parametrize("op", [capi.AGG_SUM, capi.AGG_AVG, capi.AGG_COUNT, capi.AGG_MIN, capi.AGG_MAX])
def test_agg_state_and_merge_equal_whole_column(gpu, T, op):
    """split a column in 5 ragged ranges: device-resident partial states folded by the merge kernel == the aggregate of the whole column"""
    rng = np.random.default_rng(T * 7 + op)
    dt = np.dtype(capi.NP_OF_T[T])
    n = 70_001
    if dt.kind == "f":
        col = rng.standard_normal(n).astype(dt) * 100
    else:
        info = np.iinfo(dt)
        col = rng.integers(max(info.min, -1000), min(info.max, 1000), n).astype(dt)
    nullmask = rng.random(n) < 0.1
    nullmask[:13000] = True            # the first range is all NULL: its state must be skipped by the merge
    from matrixone_b200.vector import bitmap_from_bools
    cuts = [0, 13000, 13001, 40000, 40064, n]
    states_dev = DeviceBuffer(24 * 5)
    states_host = []
    for i in range(5):
        a, b = cuts[i], cuts[i + 1]
        seg = np.ascontiguousarray(col[a:b]); segn = bitmap_from_bools(nullmask[a:b])
        dc, dn = DeviceBuffer.from_numpy(seg), DeviceBuffer.from_numpy(segn)
        ops.agg_state_device(op, T, dc, dn, b - a, states_dev.ptr + 24 * i)
        states_host.append(ops.agg_state(op, T, seg, segn, b - a))
        dc.free(); dn.free()
    sd = states_dev.to_numpy(np.uint64).reshape(5, 3)
    assert np.array_equal(sd, np.stack(states_host))                    # asynchronous state == synchronous state
    fin = DeviceBuffer(16); fnul = DeviceBuffer(8)
    ops.agg_merge_device(op, T, states_dev.ptr, 5, fin.ptr, out_bytes=16, out_nulls_ptr=fnul.ptr)
    fv = fin.to_numpy(np.uint64)
    rc, bits, cnt, isnull = ops.agg_merge(op, T, sd)
    assert fv[0] == bits and int(fv[1]) == cnt and bool(fnul.to_numpy(np.uint64)[0] & np.uint64(1)) == isnull and rc == 0
    whole_n = bitmap_from_bools(nullmask)
    if op == capi.AGG_COUNT:
        assert int(bits) == ops.agg_count(T, col, whole_n) == int((~nullmask).sum())
    elif op in (capi.AGG_MIN, capi.AGG_MAX):
        want, wn = (ops.agg_min if op == capi.AGG_MIN else ops.agg_max)(T, col, whole_n)
        assert np.asarray([bits], dtype=np.uint64).view(np.uint8)[:dt.itemsize].view(dt)[0] == want and not isnull
    elif op == capi.AGG_SUM:
        rcw, want, wn = ops.agg_sum(T, col, whole_n)
        if dt.kind == "f":
            got = np.asarray([bits], dtype=np.uint64).view(np.float64)[0]
            assert abs(got - want) <= 1e-9 * max(1.0, abs(want))
        elif dt.kind == "u":
            assert int(bits) == want
        else:
            assert int(np.asarray([bits], dtype=np.uint64).view(np.int64)[0]) == want
    else:
        rcw, want, wn = ops.agg_avg(T, col, whole_n)
        got = np.asarray([bits], dtype=np.uint64).view(np.float64)[0]
        assert abs(got - want) <= 1e-9 * max(1.0, abs(want))
    states_dev.free(); fin.free(); fnul.free()


def test_agg_merge_overflow_rules(gpu):
    """BatchMerge keeps int64OfCheck / uint64OfCheck (sumavg2.go:229-234)"""
    big = np.uint64(2 ** 62)
    st = np.asarray([[big, 1, 0], [big, 1, 0]], dtype=np.uint64)
    rc, bits, cnt, isnull = ops.agg_merge(capi.AGG_SUM, capi.T_INT64, st)
    assert rc == capi.RC_OUT_OF_RANGE
    rc, bits, cnt, isnull = ops.agg_merge(capi.AGG_SUM, capi.T_UINT64, st)
    assert rc == 0 and int(bits) == 2 ** 63
    st2 = np.asarray([[np.uint64(2 ** 63), 1, 0], [np.uint64(2 ** 63), 1, 0]], dtype=np.uint64)
    assert ops.agg_merge(capi.AGG_SUM, capi.T_UINT64, st2)[0] == capi.RC_OUT_OF_RANGE
    # a partial that already failed keeps failing the merge
    st3 = np.asarray([[1, 1, capi.RC_OUT_OF_RANGE], [1, 1, 0]], dtype=np.uint64)
    assert ops.agg_merge(capi.AGG_SUM, capi.T_INT64, st3)[0] == capi.RC_OUT_OF_RANGE
    # all partials NULL -> NULL
    assert ops.agg_merge(capi.AGG_SUM, capi.T_INT64, np.zeros((3, 3), dtype=np.uint64))[3] is True
    assert ops.agg_merge(capi.AGG_COUNT, capi.T_INT64, np.zeros((3, 3), dtype=np.uint64))[3] is False


def test_sum_signed_prefix_overflow_async_and_sync_agree(gpu):
    """the exact serial-order prefix check (slow path) behind the device gate: total fits int64 but a prefix does not"""
    n = 300_000
    col = np.zeros(n, dtype=np.int64)
    col[0] = np.iinfo(np.int64).max; col[1] = 1; col[2] = -5          # prefix overflows at row 1
    rc, v, isnull = ops.agg_sum(capi.T_INT64, col)
    assert rc == capi.RC_OUT_OF_RANGE
    dc = DeviceBuffer.from_numpy(col); st = DeviceBuffer(24)
    ops.agg_state_device(capi.AGG_SUM, capi.T_INT64, dc, None, n, st.ptr)
    s = st.to_numpy(np.uint64)
    assert int(s[2]) == capi.RC_OUT_OF_RANGE
    col2 = col.copy(); col2[1] = -1; col2[2] = 1                       # magnitudes exceed int64 but no prefix leaves it
    rc, v, isnull = ops.agg_sum(capi.T_INT64, col2)
    assert rc == 0 and v == np.iinfo(np.int64).max
    dc2 = DeviceBuffer.from_numpy(col2)
    ops.agg_state_device(capi.AGG_SUM, capi.T_INT64, dc2, None, n, st.ptr)
    s = st.to_numpy(np.uint64)
    assert int(s[2]) == 0 and int(s.view(np.int64)[0]) == np.iinfo(np.int64).max
    dc.free(); dc2.free(); st.free()
