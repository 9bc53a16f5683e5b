"""Pins the oracle's decimal restatement (og_d64_* / og_d128_* / og_sum_d64) to what the reference holds: the Decimal64 / Decimal128
FunctionTestCase tables of arithmetic_{plus,minus,multi}_test.go (tests/golden/decimal_kat.json) and -- end to end, digit for digit --
03_QUERIES/q6.result and the SUM / COUNT columns of q1.result computed in DECIMAL(15,2) arithmetic over the reference's own lineitem."""
import json
import os

import numpy as np
import pytest

import golden_tables as G
import oracle_lib as O

CASES = json.load(open(os.path.join(G.HERE, "decimal_kat.json")))["cases"]


def run_case_oracle(c):
    lib = O.go()
    op = {"add": 0, "sub": 1, "mul": 2}[c["op"]]
    w = 64 if c["inputs"][0]["type"] == "decimal64" else 128
    vals = [[int(v) for v in i["values"]] for i in c["inputs"]]
    n = len(vals[0])
    n1, n2 = (G.bitmap(i["nulls"], n) if i["nulls"] else None for i in c["inputs"])
    rn = np.zeros((n + 63) // 64, dtype=np.uint64); er = np.full(1, -1, dtype=np.int64)
    if w == 64:
        a, b = (np.asarray(v, dtype=np.int64) for v in vals)
        if op == 2:
            r = np.zeros((n, 2), dtype=np.uint64)
            rc = lib.og_d64_mul(O.p(r), O.p(a), O.p(b), n, 0, 0, 0, 0, O.p(n1), O.p(n2), O.p(rn))
            out = O.d128_to_int(r)
        else:
            r = np.zeros(n, dtype=np.int64)
            rc = lib.og_d64_addsub(op, O.p(r), O.p(a), O.p(b), n, 0, 0, 0, 0, O.p(n1), O.p(n2), O.p(rn), O.p(er))
            out = [int(x) for x in r]
    else:
        a, b = (O.int_to_d128(v) for v in vals)
        r = np.zeros((n, 2), dtype=np.uint64)
        if op == 2:
            rc = lib.og_d128_mul(O.p(r), O.p(a), O.p(b), n, 0, 0, 0, 0, O.p(n1), O.p(n2), O.p(rn), O.p(er))
        else:
            rc = lib.og_d128_addsub(op, O.p(r), O.p(a), O.p(b), n, 0, 0, 0, 0, O.p(n1), O.p(n2), O.p(rn), O.p(er))
        out = O.d128_to_int(r)
    return rc, out, [bool((int(rn[i >> 6]) >> (i & 63)) & 1) for i in range(n)]


@pytest.mark.parametrize("c", CASES, ids=["%s:%d" % (c["file"], c["line"]) for c in CASES])
def test_decimal_tables(c):
    rc, out, nulls = run_case_oracle(c)
    assert (rc != 0) == c["expect"]["want_err"]
    assert nulls == list(c["expect"]["nulls"])
    for i, v in enumerate(c["expect"]["values"]):
        if not nulls[i]:
            assert out[i] == int(v)


def decimal_q6_q1_oracle(threads=1):
    """TPC-H Q6 / Q1 on the DECIMAL(15,2) schema with the oracle's operators (what the reference's chain does on decimal64 columns):
    compare on the unscaled values, 1 - l_discount (d64 sub, scale 0 vs 2), * (d64 mul -> d128 scale 4), * (1 + l_tax) (d128 mul, scale 4 x 2 -> 6),
    SUM into Decimal128 per group."""
    from matrixone_b200 import datagen
    lib = O.go()
    cols, ints, expected = G.tpch_fixture()
    n = len(cols["shipdate"])
    qty = ints["quantity"] * 100; price = ints["extendedprice_cents"]; disc = ints["discount_pct"]; tax = ints["tax_pct"]     # DECIMAL(15,2) unscaled
    P = datagen.q6_params()
    # ---- Q6: predicates on unscaled values (0.03 -/+ 0.01 at scale 2 = 2..4 ; 24 = 2400), revenue = sum(price * disc) at scale 4
    m = (cols["shipdate"] >= P[0]) & (cols["shipdate"] < P[1]) & (disc >= 2) & (disc <= 4) & (qty < 2400)
    k = int(m.sum())
    prod = np.zeros((k, 2), dtype=np.uint64); rn = np.zeros((k + 63) // 64, dtype=np.uint64)
    assert lib.og_d64_mul(O.p(prod), O.p(np.ascontiguousarray(price[m])), O.p(np.ascontiguousarray(disc[m])), k, 0, 0, 2, 2, None, None, O.p(rn)) == 0
    s = np.zeros((1, 2), dtype=np.uint64); c = np.zeros(1, dtype=np.int64)
    lib.og_sum_d128(O.p(prod), None, 0, None, k, O.p(s), O.p(c))
    q6 = O.decimal_str(O.d128_to_int(s)[0], lib.og_mul_result_scale(2, 2))
    # ---- Q1
    sel = cols["shipdate"] <= datagen.Q1_CUTOFF
    k = int(sel.sum())
    q_, p_, d_, t_ = (np.ascontiguousarray(x[sel]) for x in (qty, price, disc, tax))
    keys = (cols["returnflag"][sel].astype(np.uint64) | (cols["linestatus"][sel].astype(np.uint64) << np.uint64(8)))
    groups = np.zeros(k, dtype=np.uint64); tk = np.zeros(16, dtype=np.uint64)
    ng = lib.og_group_ids(O.p(keys), k, O.p(groups), O.p(tk), 0, 16)
    one = np.asarray([1], dtype=np.int64)
    rn = np.zeros((k + 63) // 64, dtype=np.uint64); er = np.full(1, -1, dtype=np.int64)
    t1 = np.zeros(k, dtype=np.int64)
    assert lib.og_d64_addsub(1, O.p(t1), O.p(one), O.p(d_), k, 1, 0, 0, 2, None, None, O.p(rn), O.p(er)) == 0          # 1 - l_discount  (scale 2)
    t2 = np.zeros((k, 2), dtype=np.uint64)
    assert lib.og_d64_mul(O.p(t2), O.p(p_), O.p(t1), k, 0, 0, 2, 2, None, None, O.p(rn)) == 0                           # price * (1 - disc): d128 scale 4
    t3 = np.zeros(k, dtype=np.int64)
    assert lib.og_d64_addsub(0, O.p(t3), O.p(one), O.p(t_), k, 1, 0, 0, 2, None, None, O.p(rn), O.p(er)) == 0          # 1 + l_tax  (scale 2)
    t3w = O.int_to_d128([int(v) for v in t3])                                                                              # cast to Decimal128 for the d128 multiply
    t4 = np.zeros((k, 2), dtype=np.uint64)
    assert lib.og_d128_mul(O.p(t4), O.p(t2), O.p(t3w), k, 0, 0, 4, 2, None, None, O.p(rn), O.p(er)) == 0               # ... * (1 + tax): scale 6
    out = {}
    def dsum(col, width):
        s = np.zeros((ng, 2), dtype=np.uint64); c = np.zeros(ng, dtype=np.int64)
        (lib.og_sum_d64 if width == 64 else lib.og_sum_d128)(O.p(col), None, 0, O.p(groups), k, O.p(s), O.p(c))
        return O.d128_to_int(s), c
    sq, cnt = dsum(q_, 64); sp, _ = dsum(p_, 64); sdp, _ = dsum(t2, 128); sch, _ = dsum(t4, 128)
    for g in range(ng):
        out[(chr(int(tk[g]) & 0xff), chr((int(tk[g]) >> 8) & 0xff))] = {"sum_qty": O.decimal_str(sq[g], 2), "sum_base_price": O.decimal_str(sp[g], 2),
                                                                         "sum_disc_price": O.decimal_str(sdp[g], 4), "sum_charge": O.decimal_str(sch[g], 6), "count_order": str(int(cnt[g]))}
    return q6, out, expected


def test_decimal_q6_q1_digit_for_digit():
    q6, q1, expected = decimal_q6_q1_oracle()
    assert q6 == expected["q6_revenue"]                       # "43092.5479"
    assert set(q1) == set(expected["q1"])
    for k, g in q1.items():
        for f, v in g.items():
            assert v == expected["q1"][k][f], (k, f, v, expected["q1"][k][f])


def test_decimal_scale_rules_and_overflow():
    lib = O.go()
    assert [lib.og_mul_result_scale(a, b) for a, b in ((2, 2), (4, 2), (6, 6), (10, 10), (14, 3), (0, 0))] == [4, 6, 12, 12, 14, 0]
    # scale-down rounds half up on the magnitude (d128DivPow10Once): 1.5e-12 -> 2e-12 ; -1.5e-12 -> -2e-12 ; 1.4 -> 1
    a = np.asarray([15, -15, 14, 25], dtype=np.int64); b = np.asarray([10 ** 13] * 4, dtype=np.int64)      # scales 13 + 13 = 26 -> 13: divide by 10^13
    r = np.zeros((4, 2), dtype=np.uint64); rn = np.zeros(1, dtype=np.uint64)
    assert lib.og_d64_mul(O.p(r), O.p(a), O.p(b), 4, 0, 0, 13, 13, None, None, O.p(rn)) == 0
    assert O.d128_to_int(r) == [15, -15, 14, 25]
    b2 = np.asarray([1] * 4, dtype=np.int64)
    assert lib.og_d64_mul(O.p(r), O.p(a), O.p(b2), 4, 0, 0, 13, 13, None, None, O.p(rn)) == 0              # 15e-26 -> scale 13: 0
    assert O.d128_to_int(r) == [0, 0, 0, 0]
    c = np.asarray([15, -15, 14, 5], dtype=np.int64); d = np.asarray([10 ** 12] * 4, dtype=np.int64)
    assert lib.og_d64_mul(O.p(r), O.p(c), O.p(d), 4, 0, 0, 13, 13, None, None, O.p(rn)) == 0               # x * 10^12 / 10^13 = x / 10: 1.5 -> 2, -1.5 -> -2, 1.4 -> 1, 0.5 -> 1
    assert O.d128_to_int(r) == [2, -2, 1, 1]
    # add: first offending row fails the call, rows before it hold their results
    mx = np.iinfo(np.int64).max
    x = np.asarray([1, mx, 5], dtype=np.int64); y = np.asarray([2, 1, 6], dtype=np.int64); z = np.zeros(3, dtype=np.int64); er = np.full(1, -1, dtype=np.int64); rn[:] = 0
    assert lib.og_d64_addsub(0, O.p(z), O.p(x), O.p(y), 3, 0, 0, 2, 2, None, None, O.p(rn), O.p(er)) == 20203 and er[0] == 1 and z[0] == 3
    # scale overflow: |x| * 10^diff >= 2^63
    x = np.asarray([10 ** 17], dtype=np.int64); y = np.asarray([1], dtype=np.int64); er[:] = -1
    assert lib.og_d64_addsub(0, O.p(z), O.p(x), O.p(y), 1, 0, 0, 0, 2, None, None, O.p(rn), O.p(er)) == 20203 and er[0] == 0
