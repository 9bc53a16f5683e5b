"""The generic fused operator (csrc/plan.cu, MO_XCALL_PLAN) against the oracle's operator chains: Q6 and Q1 expressed as plan descriptors
must give what og_q6 / og_q1 (the reference's conjunct-by-conjunct filter + Shrink, projection vectors, hash group, BatchFill passes) give --
on synthetic lineitem, on the reference's own lineitem (vs q6.result / q1.result), with nulls, and for a random plan with up to 1 M groups
against a numpy restatement assembled from the oracle's primitive operators."""
import numpy as np
import pytest

import golden_tables as G
import oracle_lib as O
from matrixone_b200 import capi, datagen, ops
from matrixone_b200.vector import DeviceBuffer, bitmap_from_bools

pytestmark = pytest.mark.gpu
Q1COLS = ("shipdate", "quantity", "extendedprice", "discount", "tax", "returnflag", "linestatus")


@pytest.fixture(params=[1, 0], ids=["specialised", "interpreter"])
def mode(request, gpu):
    """recognised Q6 / Q1 shapes dispatch to the hand-specialised kernels; plan_specialise = 0 forces the generic interpreter"""
    assert gpu.MoB200_SetTuning(b"plan_specialise", request.param) == 0
    yield request.param
    gpu.MoB200_SetTuning(b"plan_specialise", 1)


def q1_groups(res):
    out = []
    for g in res:
        a = g["aggs"]
        out.append({"returnflag": g["key"] & 0xff, "linestatus": (g["key"] >> 8) & 0xff, "first_row": g["first_row"], "sum_qty": a[0][0], "sum_base_price": a[1][0],
                    "sum_disc_price": a[2][0], "sum_charge": a[3][0], "avg_qty": a[4][0], "avg_price": a[5][0], "avg_disc": a[6][0], "count_order": a[7][1]})
    return out


@pytest.mark.parametrize("n", [1, 255, 256, 257, 100_003, 2_000_001])
def test_q6_plan_matches_oracle_and_specialised_kernel(gpu, mode, n):
    cols = datagen.lineitem(10, 0, n)
    P = datagen.q6_params()
    want, ns, nul = O.q6(cols, n, P, nthreads=1)
    res = ops.q6_plan().run([cols["shipdate"], cols["discount"], cols["quantity"], cols["extendedprice"]], n)
    if ns == 0:
        assert res == []
        return
    assert len(res) == 1 and res[0]["rows"] == ns and res[0]["aggs"][0][1] == ns
    assert abs(res[0]["aggs"][0][0] - want) <= 1e-5 * abs(want)
    assert abs(res[0]["aggs"][0][0] - want) <= 1e-11 * abs(want)
    spec = ops.q6_filter_sum(cols["shipdate"], cols["discount"], cols["quantity"], cols["extendedprice"], n, *P)
    assert spec[1] == res[0]["rows"]


@pytest.mark.parametrize("n", [64, 100_003, 3_000_001])
def test_q1_plan_matches_oracle(gpu, mode, n):
    cols = datagen.lineitem(10, 0, n)
    want = O.q1(cols, n, datagen.Q1_CUTOFF)
    res = q1_groups(ops.q1_plan(datagen.Q1_CUTOFF, row_base=5).run([cols[k] for k in Q1COLS], n))
    assert [(g["returnflag"], g["linestatus"], g["count_order"], g["first_row"] - 5) for g in res] == [(g["returnflag"], g["linestatus"], g["count_order"], g["first_row"]) for g in want]
    for a, b in zip(res, want):
        for f in ("sum_qty", "sum_base_price", "sum_disc_price", "sum_charge", "avg_qty", "avg_price", "avg_disc"):
            assert abs(a[f] - b[f]) <= 1e-11 * abs(b[f]), f


def test_plans_reproduce_reference_results_on_reference_lineitem(gpu, mode):
    cols, ints, expected = G.tpch_fixture()
    n = len(cols["shipdate"])
    r6 = ops.q6_plan().run([cols["shipdate"], cols["discount"], cols["quantity"], cols["extendedprice"]], n)
    G.check_q6_result(r6[0]["aggs"][0][0], expected)
    G.check_q1_result(q1_groups(ops.q1_plan(datagen.Q1_CUTOFF).run([cols[k] for k in Q1COLS], n)), expected)
    # resident columns, device-resident result (asynchronous form)
    dev = [DeviceBuffer.from_numpy(cols[k]) for k in Q1COLS]
    p = ops.q1_plan(datagen.Q1_CUTOFF)
    out = DeviceBuffer(p.result_bytes(16))
    p.run(dev, n, max_groups=16, out_ptr=out.ptr)
    G.check_q1_result(q1_groups(p.parse(out.to_numpy(np.uint8).tobytes())), expected)
    for b in dev + [out]:
        b.free()


def test_q6_plan_with_nulls_matches_operator_chain(gpu):
    """nullable inputs: a NULL predicate operand rejects the row (filter.go:125-141), a NULL product is skipped by SUM (sumavg2.go:139-164).
    Oracle = the reference's chain assembled from its primitives: og_compare / og_between -> og_filter_sels -> og_shuffle_fixed (+ og_nulls_filter)
    -> og_arith (mul) -> og_sum_float64."""
    n = 200_000
    rng = np.random.default_rng(4)
    cols = datagen.lineitem(10, 0, n)
    names = ("shipdate", "discount", "quantity", "extendedprice")
    nulls = {k: bitmap_from_bools(rng.random(n) < 0.07) for k in names}
    P = datagen.q6_params()
    lib = O.go()
    live = np.arange(n, dtype=np.int64)
    cur = {k: cols[k].copy() for k in names}; curn = {k: nulls[k].copy() for k in names}

    def conjunct(make):
        nonlocal live, cur, curn
        m = live.shape[0]
        r = np.zeros(max(m, 1), dtype=np.uint8); rn = np.zeros((m + 63) // 64 + 1, dtype=np.uint64)
        make(r, rn, m)
        sels = np.zeros(max(m, 1), dtype=np.int64)
        k = lib.og_filter_sels(O.p(r), O.p(rn), m, O.p(sels))
        sels = sels[:k]
        nxt, nxtn = {}, {}
        for kk in names:
            d = np.zeros(max(k, 1), dtype=cur[kk].dtype); dn = np.zeros((k + 63) // 64 + 1, dtype=np.uint64)
            lib.og_shuffle_fixed(O.p(d), O.p(cur[kk]), O.p(sels), k, cur[kk].dtype.itemsize)
            lib.og_nulls_filter(O.p(curn[kk]), m, O.p(sels), k, O.p(dn))
            nxt[kk], nxtn[kk] = d[:max(k, 1)], dn
        cur, curn, live = nxt, nxtn, live[sels]

    def cmp_i32(op, c):
        cv = np.asarray([c], dtype=np.int32)
        return lambda r, rn, m: lib.og_compare(op, capi.T_DATE, O.p(r), O.p(cur["shipdate"]), O.p(cv), m, 0, 1, O.p(curn["shipdate"]), None, O.p(rn))
    conjunct(cmp_i32(3, P[0]))
    conjunct(cmp_i32(4, P[1]))
    lo = np.asarray([P[2]]); hi = np.asarray([P[3]])
    conjunct(lambda r, rn, m: lib.og_between(capi.T_FLOAT64, O.p(r), O.p(cur["discount"]), O.p(lo), O.p(hi), m, O.p(curn["discount"]), O.p(rn)))
    qh = np.asarray([P[4]])
    conjunct(lambda r, rn, m: lib.og_compare(4, capi.T_FLOAT64, O.p(r), O.p(cur["quantity"]), O.p(qh), m, 0, 1, O.p(curn["quantity"]), None, O.p(rn)))
    m = live.shape[0]
    prod = np.zeros(m); pn = np.zeros((m + 63) // 64 + 1, dtype=np.uint64)
    assert lib.og_arith(2, capi.T_FLOAT64, O.p(prod), O.p(cur["extendedprice"]), O.p(cur["discount"]), m, 0, 0, O.p(curn["extendedprice"]), O.p(curn["discount"]), O.p(pn), 0, None) == 0
    s = np.zeros(1); sn = np.ones(1, dtype=np.uint8); sc = np.zeros(1, dtype=np.int64)
    lib.og_sum_float64(capi.T_FLOAT64, O.p(prod), O.p(pn), 0, None, m, O.p(s), O.p(sn), O.p(sc))
    res = ops.q6_plan().run([cols[k] for k in names], n, nulls=[nulls[k] for k in names])
    assert len(res) == 1 and res[0]["rows"] == m
    assert res[0]["aggs"][0][1] == int(sc[0]) and abs(res[0]["aggs"][0][0] - s[0]) <= 1e-11 * abs(s[0])
    assert int(sc[0]) < m     # some qualifying rows carry a NULL price: counted by rows, skipped by SUM


@pytest.mark.parametrize("card", [1, 7, 300, 20_000, 1_000_000])
def test_random_plan_group_by_cardinalities(gpu, card):
    """group by an int32 key of `card` distinct values, nullable value column; SUM / AVG / COUNT / COUNT(*) / MIN / MAX of an expression"""
    rng = np.random.default_rng(card)
    n = 1_500_000
    key = rng.integers(0, card, n).astype(np.int32) * 3 - 7
    x = (rng.standard_normal(n) * 10).astype(np.float64); y = rng.integers(-50, 50, n).astype(np.int64)
    xn = rng.random(n) < 0.1
    flt = rng.integers(0, 100, n).astype(np.int16)
    p = ops.FusedPlan([capi.T_INT32, capi.T_FLOAT64, capi.T_INT64, capi.T_INT16])
    p.where(3, ">=", 10).where(3, "!=", 50)
    e = p.add(p.mul(p.col(1), p.const(2.5)), p.col(2))          # x * 2.5 + y
    p.group_by(0)
    for kind in (capi.AGG_SUM, capi.AGG_AVG, capi.AGG_COUNT, capi.AGG_MIN, capi.AGG_MAX):
        p.agg(kind, e)
    p.agg(capi.AGG_COUNT, -1)
    res = p.run([key, x, y, flt], n, nulls=[None, bitmap_from_bools(xn), None, None], max_groups=card + 8)
    sel = (flt >= 10) & (flt != 50)
    ev = x * 2.5 + y
    ks = key[sel]; es = ev[sel]; ens = xn[sel]
    # first-seen group ids from the oracle's restatement of the hash map (vectorised: any exact map gives the same ids)
    _, first_idx, inv = np.unique(ks, return_index=True, return_inverse=True)
    order = np.argsort(first_idx, kind="stable")
    rank = np.empty_like(order); rank[order] = np.arange(order.shape[0])
    gid = rank[inv]
    ng = order.shape[0]
    assert len(res) == ng
    rows = np.bincount(gid, minlength=ng)
    live = ~ens
    cnt = np.bincount(gid[live], minlength=ng)
    sums = np.bincount(gid[live], weights=es[live], minlength=ng)
    mins = np.full(ng, np.inf); np.minimum.at(mins, gid[live], es[live])
    maxs = np.full(ng, -np.inf); np.maximum.at(maxs, gid[live], es[live])
    first_rows = np.flatnonzero(sel)[np.sort(first_idx)]
    if ng <= 8192:
        assert [g["first_row"] for g in res] == list(first_rows)
    for i, g in enumerate(res):
        raw = int(ks[np.sort(first_idx)[i]]) & 0xffffffff
        assert g["key"] == raw and g["rows"] == rows[i]
        a = g["aggs"]
        assert a[5][1] == rows[i] and a[2][1] == cnt[i] and a[0][1] == cnt[i]
        if cnt[i]:
            assert abs(a[0][0] - sums[i]) <= 1e-9 * max(1.0, abs(sums[i]))
            assert abs(a[1][0] - sums[i] / cnt[i]) <= 1e-9 * max(1.0, abs(sums[i] / cnt[i]))
            assert a[3][0] == mins[i] and a[4][0] == maxs[i]
        if i > 2000:
            break


def test_plan_too_many_groups_is_loud(gpu):
    key = np.arange(5000, dtype=np.int32)
    p = ops.FusedPlan([capi.T_INT32]).group_by(0).agg(capi.AGG_COUNT, -1)
    with pytest.raises(capi.MoError):
        p.run([key], 5000, max_groups=100)
    p2 = ops.FusedPlan([capi.T_INT64, capi.T_INT32]).group_by(0, 1).agg(capi.AGG_COUNT, -1)      # 12 key bytes: StrHashMap territory
    with pytest.raises(capi.MoError):
        p2.run([np.zeros(4, dtype=np.int64), np.zeros(4, dtype=np.int32)], 4)
