"""GPU parity of Decimal64 / Decimal128 arithmetic and decimal SUM (csrc/decimal.cu) -- integer work, bit-exact: the reference's
FunctionTestCase tables, random vectors vs the oracle restatement (results, nulls, return code, first offending row), and TPC-H Q6 / Q1 on the
DECIMAL(15,2) schema over the reference's lineitem, digit for digit against q6.result / q1.result, operator by operator on the GPU."""
import numpy as np
import pytest

import golden_tables as G
import oracle_lib as O
from matrixone_b200 import capi, datagen, ops
from test_oracle_decimal import CASES

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("c", CASES, ids=["%s:%d" % (c["file"], c["line"]) for c in CASES])
def test_gpu_decimal_tables(gpu, c):
    op = {"add": 0, "sub": 1, "mul": 2}[c["op"]]
    w = 64 if c["inputs"][0]["type"] == "decimal64" else 128
    vals = [[int(v) for v in i["values"]] for i in c["inputs"]]
    n = len(vals[0])
    n1, n2 = (G.bitmap(i["nulls"], n) if i["nulls"] else None for i in c["inputs"])
    a, b = (np.asarray(v, dtype=np.int64) for v in vals) if w == 64 else (O.int_to_d128(v) for v in vals)
    rc, r, rn, er = ops.dec_arith(op, w, a, b, 0, 0, n, n1, n2)
    assert (rc != 0) == c["expect"]["want_err"]
    nulls = [bool((int(rn[i >> 6]) >> (i & 63)) & 1) for i in range(n)]
    assert nulls == list(c["expect"]["nulls"])
    out = O.d128_to_int(r) if r.ndim == 2 else [int(x) for x in r]
    assert all(out[i] == int(v) for i, v in enumerate(c["expect"]["values"]) if not nulls[i])


def _rand_nulls(rng, n, p):
    b = rng.random(((n + 63) // 64) * 64) < p
    b[n:] = False
    return np.packbits(b, bitorder="little").view(np.uint64).copy()


@pytest.mark.parametrize("op", [0, 1, 2])
@pytest.mark.parametrize("scales", [(2, 2), (0, 2), (4, 2), (6, 0), (10, 10), (13, 13)])
@pytest.mark.parametrize("shape", ["vv", "vc", "cv"])
def test_gpu_d64_arith_matches_oracle(gpu, op, scales, shape):
    rng = np.random.default_rng(op * 100 + scales[0] * 7 + scales[1] + len(shape))
    n = 30_011
    lim = 10 ** 15 if op != 2 else 2 ** 62
    a = rng.integers(-lim, lim, n).astype(np.int64); b = rng.integers(-lim, lim, n).astype(np.int64)
    c1, c2 = shape == "cv", shape == "vc"
    aa = a[:1].copy() if c1 else a; bb = b[:1].copy() if c2 else b
    n1 = None if c1 else _rand_nulls(rng, n, 0.05); n2 = None if c2 else _rand_nulls(rng, n, 0.05)
    pre = _rand_nulls(rng, n, 0.1)
    lib = O.go()
    rn0 = pre.copy(); er0 = np.full(1, -1, dtype=np.int64)
    if op == 2:
        r0 = np.zeros((n, 2), dtype=np.uint64)
        rc0 = lib.og_d64_mul(O.p(r0), O.p(aa), O.p(bb), n, int(c1), int(c2), scales[0], scales[1], O.p(n1), O.p(n2), O.p(rn0))
    else:
        r0 = np.zeros(n, dtype=np.int64)
        rc0 = lib.og_d64_addsub(op, O.p(r0), O.p(aa), O.p(bb), n, int(c1), int(c2), scales[0], scales[1], O.p(n1), O.p(n2), O.p(rn0), O.p(er0))
    rc1, r1, rn1, er1 = ops.dec_arith(op, 64, aa, bb, scales[0], scales[1], n, n1, n2, pre)
    assert rc1 == rc0
    if rc0 == 0:
        assert np.array_equal(rn1, rn0)
        live = ~np.unpackbits(rn0.view(np.uint8), bitorder="little")[:n].astype(bool)
        assert np.array_equal(r1[live], r0[live])
    else:
        assert er1 == er0[0]


@pytest.mark.parametrize("op", [0, 1, 2])
@pytest.mark.parametrize("scales", [(4, 2), (2, 4), (12, 12), (6, 6)])
def test_gpu_d128_arith_matches_oracle(gpu, op, scales):
    rng = np.random.default_rng(op * 10 + scales[0])
    n = 20_003
    big = [int(x) * int(y) for x, y in zip(rng.integers(-2 ** 62, 2 ** 62, n), rng.integers(1, 2 ** 30 if op != 2 else 4, n))]
    small = [int(x) for x in rng.integers(-10 ** 9, 10 ** 9, n)]
    a, b = O.int_to_d128(big), O.int_to_d128(small)
    n1 = _rand_nulls(rng, n, 0.05)
    lib = O.go()
    r0 = np.zeros((n, 2), dtype=np.uint64); rn0 = np.zeros((n + 63) // 64, dtype=np.uint64); er0 = np.full(1, -1, dtype=np.int64)
    if op == 2:
        rc0 = lib.og_d128_mul(O.p(r0), O.p(a), O.p(b), n, 0, 0, scales[0], scales[1], O.p(n1), None, O.p(rn0), O.p(er0))
    else:
        rc0 = lib.og_d128_addsub(op, O.p(r0), O.p(a), O.p(b), n, 0, 0, scales[0], scales[1], O.p(n1), None, O.p(rn0), O.p(er0))
    rc1, r1, rn1, er1 = ops.dec_arith(op, 128, a, b, scales[0], scales[1], n, n1, None)
    assert rc1 == rc0
    if rc0 == 0:
        assert np.array_equal(rn1, rn0)
        live = ~np.unpackbits(rn0.view(np.uint8), bitorder="little")[:n].astype(bool)
        assert np.array_equal(r1[live], r0[live])
    else:
        assert er1 == er0[0]


def test_gpu_decimal_overflow_first_offender(gpu):
    mx = np.iinfo(np.int64).max
    x = np.asarray([1, mx, 5, mx], dtype=np.int64); y = np.asarray([2, 1, 6, 1], dtype=np.int64)
    rc, r, rn, er = ops.dec_arith(0, 64, x, y, 2, 2, 4)
    assert rc == capi.RC_INVALID_ARGUMENT and er == 1 and r[0] == 3
    rc, r, rn, er = ops.dec_arith(0, 64, np.asarray([10 ** 17], dtype=np.int64), np.asarray([1], dtype=np.int64), 0, 2, 1)
    assert rc == capi.RC_INVALID_ARGUMENT and er == 0                                  # scale overflow
    huge = O.int_to_d128([2 ** 100, 3]); rc, r, rn, er = ops.dec_arith(2, 128, huge, O.int_to_d128([2 ** 40, 3]), 0, 0, 2)
    assert rc == capi.RC_INVALID_ARGUMENT and er == 0                                  # Decimal128 Mul overflow


@pytest.mark.parametrize("ngroups", [1, 4, 700, 100_000])
def test_gpu_decimal_sum_matches_oracle(gpu, ngroups):
    rng = np.random.default_rng(ngroups)
    n = 300_001
    col = rng.integers(-2 ** 62, 2 ** 62, n).astype(np.int64)          # sums leave int64: the Decimal128 state must carry
    nulls = _rand_nulls(rng, n, 0.1)
    groups = rng.integers(0, ngroups + 1, n).astype(np.uint64)
    lib = O.go()
    s0 = np.zeros((ngroups, 2), dtype=np.uint64); c0 = np.zeros(ngroups, dtype=np.int64)
    lib.og_sum_d64(O.p(col), O.p(nulls), 0, O.p(groups), n, O.p(s0), O.p(c0))
    s1 = np.zeros((ngroups, 2), dtype=np.uint64); c1 = np.zeros(ngroups, dtype=np.int64)
    h = n // 2 // 64 * 64
    ops.dec_sum(64, col[:h], s1, c1, groups[:h], nulls[:h // 64], h)                  # two batches into the same state
    ops.dec_sum(64, col[h:], s1, c1, groups[h:], nulls[h // 64:], n - h)
    assert np.array_equal(s1, s0) and np.array_equal(c1, c0)
    wide = O.int_to_d128([int(v) * 3 for v in col[:50_000]])
    s0[:] = 0; c0[:] = 0; s1[:] = 0; c1[:] = 0
    lib.og_sum_d128(O.p(wide), None, 0, O.p(groups), 50_000, O.p(s0), O.p(c0))
    ops.dec_sum(128, wide, s1, c1, groups[:50_000], None, 50_000)
    assert np.array_equal(s1, s0) and np.array_equal(c1, c0)
    if ngroups == 1:                                                                   # no group-by: groups vector omitted
        s2 = np.zeros((1, 2), dtype=np.uint64); c2 = np.zeros(1, dtype=np.int64)
        ops.dec_sum(128, wide, s2, c2, None, None, 50_000)
        assert O.d128_to_int(s2)[0] == sum(int(v) * 3 for v in col[:50_000]) and c2[0] == 50_000


def test_gpu_decimal_q6_q1_digit_for_digit_on_reference_lineitem(gpu):
    """the DECIMAL(15,2) operator chain on the GPU: compare -> sels -> Shrink -> d64 sub / mul -> d128 mul -> fillKeys -> group ids -> decimal SUM"""
    from matrixone_b200.vector import Vector, xcall
    cols, ints, expected = G.tpch_fixture()
    n = len(cols["shipdate"])
    qty = ints["quantity"] * 100; price = ints["extendedprice_cents"]; disc = ints["discount_pct"]; tax = ints["tax_pct"]
    P = datagen.q6_params()

    def cmp(op, T, col, const):
        r = np.zeros(n, dtype=np.uint8); rn = np.zeros((n + 63) // 64, dtype=np.uint64)
        xcall(capi.XCALL_GO_COMPARE(op, T), [Vector(data=r, nulls=rn, length=n), Vector(data=col, length=n), Vector(data=np.asarray([const], dtype=col.dtype), length=n)], n)
        return r
    # Decimal64 columns of equal scale compare as their int64 unscaled values (types.Decimal64.Compare)
    m = cmp(3, capi.T_DATE, cols["shipdate"], P[0]) & cmp(4, capi.T_DATE, cols["shipdate"], P[1]) & cmp(3, capi.T_INT64, disc, 2) & cmp(5, capi.T_INT64, disc, 4) & cmp(4, capi.T_INT64, qty, 2400)
    sels = ops.filter_sels(m)
    rc, prod, _, _ = ops.dec_arith(2, 64, ops.shuffle(price, sels), ops.shuffle(disc, sels), 2, 2, len(sels))
    assert rc == 0
    s = np.zeros((1, 2), dtype=np.uint64); c = np.zeros(1, dtype=np.int64)
    ops.dec_sum(128, prod, s, c, None, None, len(sels))
    assert O.decimal_str(O.d128_to_int(s)[0], 4) == expected["q6_revenue"]
    # ---- Q1
    sels = ops.filter_sels(cmp(5, capi.T_DATE, cols["shipdate"], datagen.Q1_CUTOFF))
    k = len(sels)
    q_, p_, d_, t_ = (ops.shuffle(x, sels) for x in (qty, price, disc, tax))
    keys, _ = ops.pack_keys([ops.shuffle(cols["returnflag"], sels), ops.shuffle(cols["linestatus"], sels)], has_null=False)
    table = ops.GroupTable(16)
    groups = table.insert(keys)
    ng = int(table.ngroups[0])
    one = np.asarray([1], dtype=np.int64)
    rc, t1, _, _ = ops.dec_arith(1, 64, one, d_, 0, 2, k); assert rc == 0
    rc, t2, _, _ = ops.dec_arith(2, 64, p_, t1, 2, 2, k); assert rc == 0
    rc, t3, _, _ = ops.dec_arith(0, 64, one, t_, 0, 2, k); assert rc == 0
    rc, t4, _, _ = ops.dec_arith(2, 128, t2, O.int_to_d128([int(v) for v in t3]), 4, 2, k); assert rc == 0
    def dsum(col, width):
        s = np.zeros((ng, 2), dtype=np.uint64); c = np.zeros(ng, dtype=np.int64)
        ops.dec_sum(width, col, s, c, groups, None, k)
        return O.d128_to_int(s), c
    sq, cnt = dsum(q_, 64); sp, _ = dsum(p_, 64); sdp, _ = dsum(t2, 128); sch, _ = dsum(t4, 128)
    assert ng == len(expected["q1"])
    for g in range(ng):
        e = expected["q1"][(chr(int(table.keys[g]) & 0xff), chr((int(table.keys[g]) >> 8) & 0xff))]
        assert (O.decimal_str(sq[g], 2), O.decimal_str(sp[g], 2), O.decimal_str(sdp[g], 4), O.decimal_str(sch[g], 6), str(int(cnt[g]))) == \
               (e["sum_qty"], e["sum_base_price"], e["sum_disc_price"], e["sum_charge"], e["count_order"])
