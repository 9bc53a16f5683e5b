"""GPU parity of the standalone column operators (csrc/colops.cu) against the oracle restatement of the reference operators:
og_filter_sels (filter.go:125-141), og_shuffle_fixed (shuffle.go:21-26), og_nulls_filter (nulls.go:264-280), og_group_ids
(first-seen group ids of IntHashMap insert), og_sum_* / og_count / og_minmax (aggexec BatchFill).  Integer / index work: bit-exact;
float64 group sums: 1e-5 relative as north_star states (measured ~1e-13: atomics change the association order only)."""
import numpy as np
import pytest

import oracle_lib as O
from matrixone_b200 import capi, ops
from matrixone_b200.vector import DeviceBuffer, bitmap_from_bools, bitmap_to_bools

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [0, 1, 15, 16, 17, 2047, 2048, 2049, 100_003, 3_000_000])
@pytest.mark.parametrize("density", [0.0, 0.02, 0.5, 1.0])
def test_filter_sels_matches_oracle(gpu, n, density):
    rng = np.random.default_rng(n + int(density * 100))
    v = (rng.random(n) < density).astype(np.uint8)
    v[v != 0] = rng.integers(1, 256, int((v != 0).sum())).astype(np.uint8)      # any non-zero byte is true
    for with_nulls in (False, True):
        nulls = bitmap_from_bools(rng.random(n) < 0.1) if with_nulls and n else None
        want = np.zeros(max(n, 1), dtype=np.int64)
        k = O.go().og_filter_sels(O.p(v), O.p(nulls), n, O.p(want))
        got = ops.filter_sels(v, nulls, n)
        assert got.shape[0] == k and np.array_equal(got, want[:k])


def test_filter_sels_device_form_and_unaligned_input(gpu):
    rng = np.random.default_rng(5)
    n = 1_000_003
    v = (rng.random(n + 3) < 0.3).astype(np.uint8)
    dv = DeviceBuffer.from_numpy(v)
    view = DeviceBuffer.__new__(DeviceBuffer); view.lib, view.nbytes, view.ptr = dv.lib, n, dv.ptr + 3      # start not 16-byte aligned
    dsels, dcnt = DeviceBuffer(8 * n), DeviceBuffer(8)
    ops.filter_sels_device(view, None, n, dsels.ptr, dcnt.ptr)
    cnt = int(dcnt.to_numpy(np.int64)[0])
    want = np.flatnonzero(v[3:3 + n])
    assert cnt == want.shape[0] and np.array_equal(dsels.to_numpy(np.int64, cnt), want)
    view.ptr = None
    dv.free(); dsels.free(); dcnt.free()


@pytest.mark.parametrize("dt", [np.uint8, np.int16, np.int32, np.float64, np.complex128])
def test_shuffle_with_nulls_filter_matches_oracle(gpu, dt):
    rng = np.random.default_rng(np.dtype(dt).itemsize)
    n = 50_021
    src = rng.integers(0, 250, n).astype(dt)
    snulls = bitmap_from_bools(rng.random(n) < 0.2)
    for m in (0, 1, 63, 64, 65, 20_000):
        sels = np.sort(rng.choice(n, m, replace=False)).astype(np.int64) if m else np.zeros(0, dtype=np.int64)
        if m > 10:
            sels[5] = sels[4]                      # Union may repeat rows
        want = np.zeros(m, dtype=dt); wn = np.zeros((m + 63) // 64, dtype=np.uint64)
        O.go().og_shuffle_fixed(O.p(want), O.p(src), O.p(sels), m, np.dtype(dt).itemsize)
        O.go().og_nulls_filter(O.p(snulls), n, O.p(sels), m, O.p(wn))
        got, gn = ops.shuffle(src, sels, snulls, want_nulls=True)
        assert np.array_equal(got.view(np.uint8), want.view(np.uint8)) and np.array_equal(gn, wn)
        assert np.array_equal(ops.shuffle(src, sels).view(np.uint8), want.view(np.uint8))     # no nulls requested


def test_shuffle_varlena_cells(gpu):
    """24-byte varlena cells move as opaque cells (Shrink of a varchar / vecf32 column keeps the area)"""
    rng = np.random.default_rng(9)
    cells = rng.integers(0, 256, (1000, 24)).astype(np.uint8)
    sels = np.sort(rng.choice(1000, 300, replace=False)).astype(np.int64)
    dst = np.zeros((300, 24), dtype=np.uint8)
    from matrixone_b200.vector import Vector, xcall
    xcall(capi.XCALL_SHUFFLE(24), [Vector(data=dst.reshape(-1), length=300), Vector(data=cells.reshape(-1), length=1000), Vector(data=sels, length=300)], 300)
    assert np.array_equal(dst, cells[sels])


def test_pack_keys_layout(gpu):
    """fillKeys byte layout (inthashmap.go:113-180): marker byte then value bytes per column; a NULL contributes its marker only"""
    rf = np.frombuffer(b"ANRA", dtype=np.uint8).copy(); ls = np.frombuffer(b"FOFO", dtype=np.uint8).copy()
    keys, _ = ops.pack_keys([rf, ls], has_null=True)
    assert [int(k) for k in keys] == [(int(c1) << 8) | (int(c2) << 24) for c1, c2 in zip(rf, ls)]
    nul = [bitmap_from_bools([False, True, False, False]), None]
    keys, _ = ops.pack_keys([rf, ls], nulls=nul, has_null=True)
    assert int(keys[1]) == 1 | (0 << 8) | (int(ls[1]) << 16)              # marker 1, then column 2 at offset 1
    keys, skip = ops.pack_keys([rf, ls], nulls=nul, has_null=False)
    assert [int(k) for k in keys[[0, 2, 3]]] == [int(rf[i]) | (int(ls[i]) << 8) for i in (0, 2, 3)]
    assert list(bitmap_to_bools(skip, 4)) == [False, True, False, False]
    k32, _ = ops.pack_keys([np.asarray([7, 8], dtype=np.int32), np.asarray([1, 2], dtype=np.int16)], has_null=True)
    assert [int(k) for k in k32] == [(7 << 8) | (1 << 48), (8 << 8) | (2 << 48)]
    with pytest.raises(capi.MoError):
        ops.pack_keys([np.zeros(4, dtype=np.int64), np.zeros(4, dtype=np.int8)], has_null=True)    # 11 bytes: StrHashMap territory


@pytest.mark.parametrize("card", [1, 4, 1000, 1_000_000])
def test_group_ids_first_seen_order_matches_oracle(gpu, card):
    rng = np.random.default_rng(card)
    n = 400_000 if card < 1_000_000 else 1_200_000
    universe = rng.integers(0, 2 ** 63, card, dtype=np.uint64)
    universe[0] = np.uint64(0xFFFFFFFFFFFFFFFF)                            # the table's empty-slot sentinel is a legal key
    keys = universe[rng.integers(0, card, n)]
    cap = min(n, card) + 8
    table = ops.GroupTable(cap)
    # two batches: the table persists, ids of batch 2 continue the numbering (exec2.go:325-362)
    half = n // 2
    g1 = table.insert(keys[:half]); ng1 = int(table.ngroups[0])
    g2 = table.insert(keys[half:]); ng2 = int(table.ngroups[0])
    # oracle: any exact map reproduces first-seen ids; vectorised restatement for speed (og_group_ids is O(n * groups))
    _, first_idx, inv = np.unique(keys, return_index=True, return_inverse=True)
    order = np.argsort(first_idx, kind="stable")
    rank = np.empty_like(order); rank[order] = np.arange(order.shape[0])
    want = (rank[inv] + 1).astype(np.uint64)
    assert np.array_equal(np.concatenate([g1, g2]), want)
    assert ng2 == order.shape[0] and ng1 == int(want[:half].max())
    assert np.array_equal(table.keys[:ng2], keys[np.sort(first_idx)])      # keys appended in id order (GetBinaryInsertList)
    if card <= 1000:
        tk = np.zeros(cap, dtype=np.uint64); og = np.zeros(n, dtype=np.uint64)
        assert O.go().og_group_ids(O.p(keys), n, O.p(og), O.p(tk), 0, cap) == ng2
        assert np.array_equal(og, want) and np.array_equal(tk[:ng2], table.keys[:ng2])


def test_group_ids_skip_rows_and_overflow(gpu):
    keys = np.asarray([5, 5, 9, 7, 9, 5], dtype=np.uint64)
    skip = bitmap_from_bools([False, True, False, False, True, False])
    t = ops.GroupTable(8)
    assert list(t.insert(keys, skip)) == [1, 0, 2, 3, 0, 1]
    small = ops.GroupTable(2)
    with pytest.raises(capi.MoError):
        small.insert(keys)


SUM_TYPES = [capi.T_INT8, capi.T_INT32, capi.T_INT64, capi.T_UINT8, capi.T_UINT64, capi.T_FLOAT32, capi.T_FLOAT64]


@pytest.mark.parametrize("T", SUM_TYPES)
@pytest.mark.parametrize("ngroups", [1, 6, 33, 5000])
def test_group_sum_avg_count_match_oracle(gpu, T, ngroups):
    rng = np.random.default_rng(T * 31 + ngroups)
    dt = np.dtype(capi.NP_OF_T[T])
    n = 200_003
    col = (rng.standard_normal(n) * 50).astype(dt) if dt.kind == "f" else rng.integers(max(np.iinfo(dt).min, -100), min(np.iinfo(dt).max, 100), n).astype(dt)
    nulls = bitmap_from_bools(rng.random(n) < 0.1)
    groups = rng.integers(0, ngroups + 1, n).astype(np.uint64)            # 0 = GroupNotMatched
    if ngroups > 2:
        groups[groups == 2] = 3                                           # group 2 stays empty -> NULL
    st_dt = np.float64 if dt.kind == "f" else (np.uint64 if dt.kind == "u" else np.int64)
    for two_batches in (False, True):
        want = np.zeros(ngroups, dtype=st_dt); wnull = np.ones(ngroups, dtype=np.uint8); wcnt = np.zeros(ngroups, dtype=np.int64)
        fn = O.go().og_sum_float64 if dt.kind == "f" else (O.go().og_sum_uint64 if dt.kind == "u" else O.go().og_sum_int64)
        args = [T, O.p(col), O.p(nulls), 0, O.p(groups), n, O.p(want), O.p(wnull), O.p(wcnt)] + ([] if dt.kind == "f" else [None])
        assert fn(*args) == 0
        state = np.zeros(ngroups, dtype=st_dt); snull = bitmap_from_bools(np.ones(ngroups, dtype=bool)); cnt = np.zeros(ngroups, dtype=np.int64)
        if two_batches:      # BatchFill twice into the same state == once over the concatenation
            h = n // 2 // 64 * 64
            assert ops.group_agg(capi.AGG_AVG, T, state, snull, cnt, groups[:h], col[:h], nulls[:h // 64], h) == 0
            assert ops.group_agg(capi.AGG_AVG, T, state, snull, cnt, groups[h:], col[h:], nulls[h // 64:], n - h) == 0
        else:
            assert ops.group_agg(capi.AGG_AVG, T, state, snull, cnt, groups, col, nulls) == 0
        assert np.array_equal(cnt, wcnt)
        assert np.array_equal(bitmap_to_bools(snull, ngroups), wnull.astype(bool))
        live = wnull == 0
        if dt.kind == "f":
            np.testing.assert_allclose(state[live], want[live], rtol=1e-5, atol=1e-9)
            np.testing.assert_allclose(state[live], want[live], rtol=1e-11, atol=1e-9)       # what we actually achieve
        else:
            assert np.array_equal(state[live], want[live])
    # COUNT(col) and COUNT(*)
    for star in (0, 1):
        wc = np.zeros(ngroups, dtype=np.int64)
        O.go().og_count(star, None if star else O.p(nulls), 0, O.p(groups), n, O.p(wc))
        cs = np.zeros(ngroups, dtype=np.int64)
        assert ops.group_agg(capi.AGG_COUNT, T, cs, None, None, groups, None if star else col, None if star else nulls, n) == 0
        assert np.array_equal(cs, wc)


MM_TYPES = [capi.T_INT8, capi.T_INT64, capi.T_UINT16, capi.T_UINT64, capi.T_FLOAT32, capi.T_FLOAT64, capi.T_DATE]


@pytest.mark.parametrize("T", MM_TYPES)
@pytest.mark.parametrize("is_max", [0, 1])
def test_group_minmax_matches_oracle_incl_nan_rule(gpu, T, is_max):
    rng = np.random.default_rng(T * 3 + is_max)
    dt = np.dtype(capi.NP_OF_T[T])
    n, ngroups = 120_001, 40
    if dt.kind == "f":
        col = (rng.standard_normal(n) * 1000).astype(dt)
        col[rng.random(n) < 0.01] = np.nan
    else:
        info = np.iinfo(dt)
        col = rng.integers(info.min, info.max, n, dtype=np.int64 if dt.kind == "i" else np.uint64).astype(dt)
    nulls = bitmap_from_bools(rng.random(n) < 0.1)
    groups = rng.integers(0, ngroups + 1, n).astype(np.uint64)
    groups[groups == 7] = 8
    if dt.kind == "f":       # group 3's first non-null value is NaN: the Go loop never replaces it (minmax2.go:69-75)
        first3 = np.flatnonzero((groups == 3) & ~bitmap_to_bools(nulls, n))[0]
        col[first3] = np.nan
    want = np.zeros(ngroups, dtype=dt); wnull = np.ones(ngroups, dtype=np.uint8)
    assert O.go().og_minmax(is_max, T, O.p(col), O.p(nulls), 0, O.p(groups), n, O.p(want), O.p(wnull)) == 0
    state = np.zeros(ngroups, dtype=np.uint64); snull = bitmap_from_bools(np.ones(ngroups, dtype=bool))
    h = n // 3 // 64 * 64
    op = capi.AGG_MAX if is_max else capi.AGG_MIN
    assert ops.group_agg(op, T, state, snull, None, groups[:h], col[:h], nulls[:h // 64], h) == 0
    assert ops.group_agg(op, T, state, snull, None, groups[h:], col[h:], nulls[h // 64:], n - h) == 0
    assert np.array_equal(bitmap_to_bools(snull, ngroups), wnull.astype(bool))
    got = state.view(np.uint8).reshape(ngroups, 8)[:, :dt.itemsize].copy().view(dt).reshape(-1)
    live = wnull == 0
    if dt.kind == "f":
        assert np.array_equal(np.isnan(got[live]), np.isnan(want[live])) and np.isnan(got[2])
        ok = ~np.isnan(want) & live
        assert np.array_equal(got[ok], want[ok])
    else:
        assert np.array_equal(got[live], want[live])


def test_group_sum_overflow_rules(gpu):
    """int64OfCheck on the running sum of each group (sumavg2.go:89-94): certain overflow, pathological no-overflow, uint64"""
    mx = np.iinfo(np.int64).max
    groups = np.asarray([1, 2, 1, 2, 1], dtype=np.uint64)
    col = np.asarray([mx, 5, 1, 6, -3], dtype=np.int64)                    # group 1: prefix mx + 1 overflows
    st = np.zeros(2, dtype=np.int64); sn = bitmap_from_bools([True, True])
    assert ops.group_agg(capi.AGG_SUM, capi.T_INT64, st, sn, None, groups, col) == capi.RC_OUT_OF_RANGE
    col2 = np.asarray([mx, 5, -3, 6, 1], dtype=np.int64)                   # magnitudes exceed int64, no prefix does: exact serial check passes
    st = np.zeros(2, dtype=np.int64); sn = bitmap_from_bools([True, True])
    assert ops.group_agg(capi.AGG_SUM, capi.T_INT64, st, sn, None, groups, col2) == 0
    assert list(st) == [mx - 2, 11]
    w = np.zeros(2, dtype=np.int64); wn = np.ones(2, dtype=np.uint8)
    assert O.go().og_sum_int64(capi.T_INT64, O.p(col2), None, 0, O.p(groups), 5, O.p(w), O.p(wn), None, None) == 0 and list(w) == list(st)
    ucol = np.asarray([2 ** 63, 1, 2 ** 63, 1, 0], dtype=np.uint64)
    ust = np.zeros(2, dtype=np.uint64); sn = bitmap_from_bools([True, True])
    assert ops.group_agg(capi.AGG_SUM, capi.T_UINT64, ust, sn, None, groups, ucol) == capi.RC_OUT_OF_RANGE
    with pytest.raises(capi.MoError):                                      # group id beyond the state
        ops.group_agg(capi.AGG_SUM, capi.T_INT64, np.zeros(1, dtype=np.int64), bitmap_from_bools([True]), None, groups, col2)


def test_filter_shrink_group_agg_chain_reproduces_q1_on_reference_lineitem(gpu):
    """the reference's operator chain, operator by operator on the GPU (compare -> sels -> Shrink -> fillKeys -> group ids -> BatchFill), on the
    reference's own lineitem: identical groups / counts to the fused Q1 kernel and q1.result"""
    import golden_tables as G
    from matrixone_b200 import datagen
    from matrixone_b200.vector import Vector, xcall
    cols, ints, expected = G.tpch_fixture()
    n = len(cols["shipdate"])
    r = np.zeros(n, dtype=np.uint8); rn = np.zeros((n + 63) // 64, dtype=np.uint64)
    cut = np.asarray([datagen.Q1_CUTOFF], dtype=np.int32)
    xcall(capi.XCALL_GO_COMPARE(5, capi.T_DATE), [Vector(data=r, nulls=rn, length=n), Vector(data=cols["shipdate"], length=n), Vector(data=cut, length=n)], n)
    sels = ops.filter_sels(r, rn, n)
    shr = {k: ops.shuffle(cols[k], sels) for k in ("quantity", "extendedprice", "discount", "tax", "returnflag", "linestatus")}
    keys, _ = ops.pack_keys([shr["returnflag"], shr["linestatus"]], has_null=False)
    table = ops.GroupTable(16)
    groups = table.insert(keys)
    ng = int(table.ngroups[0])
    assert ng == 4
    disc_price = shr["extendedprice"] * (1.0 - shr["discount"])
    charge = disc_price * (1.0 + shr["tax"])
    out = []
    for g in range(ng):
        out.append({"returnflag": int(table.keys[g]) & 0xff, "linestatus": (int(table.keys[g]) >> 8) & 0xff})
    def agg(op, col):
        st = np.zeros(ng, dtype=np.float64); sn = bitmap_from_bools(np.ones(ng, dtype=bool)); cnt = np.zeros(ng, dtype=np.int64)
        assert ops.group_agg(op, capi.T_FLOAT64, st, sn, cnt, groups, col) == 0
        return st, cnt
    for name, col in (("sum_qty", shr["quantity"]), ("sum_base_price", shr["extendedprice"]), ("sum_disc_price", disc_price), ("sum_charge", charge)):
        st, cnt = agg(capi.AGG_SUM, col)
        for g in range(ng):
            out[g][name] = st[g]; out[g]["count_order"] = int(cnt[g])
    for name, col in (("avg_qty", shr["quantity"]), ("avg_price", shr["extendedprice"]), ("avg_disc", shr["discount"])):
        st, cnt = agg(capi.AGG_AVG, col)
        for g in range(ng):
            out[g][name] = st[g] / cnt[g]
    G.check_q1_result(out, expected)
