"""Pins the oracle's elementwise family and its Q6 / Q1 operator chains to what the REFERENCE ITSELF holds:

  * the FunctionTestCase tables of pkg/sql/plan/function/{arithmetic_plus,arithmetic_minus,arithmetic_multi,arithmetic_div_mod,
    arithmetic_div_zero,func_compare_logic,operatorSet}_test.go (inputs, null lists, expected values, expected error), transcribed by
    tests/golden/extract_goldens.py into tests/golden/function_kat.json;
  * the 6005-row lineitem of test/distributed/cases/benchmark/tpch/02_LOAD/03_insert_lineitem.sql with the expected
    03_QUERIES/q6.result (revenue 43092.5479) and q1.result.
CPU only; the same tables go through the GPU in tests/test_gpu_reference_tables.py."""
import ctypes as C

import numpy as np
import pytest

import golden_tables as G
import oracle_lib as O

def same_value(ty, got, want):
    """the comparison FunctionTestCase.Run applies (func_testcase.go:390-432): exact for every type but float64, which uses
    assertx.InEpsilonF64 (|want - got| < 1e-9, pkg/common/assertx/float64.go:22,64)"""
    if ty == "float64":
        return got == want or abs(float(got) - float(want)) < 1e-9 or (got != got and want != want)
    return got == want or (got != got and want != want)


ARITH = list(G.function_cases(G.ARITH_OP))
CMP = list(G.function_cases(G.CMP_OP))
LOGIC = list(G.function_cases({"and", "or", "xor"}))


def test_tables_are_not_empty():
    assert len(ARITH) >= 55 and len(CMP) >= 10 and len(LOGIC) == 3
    assert sum(c["want_err"] for c in ARITH) >= 5


@pytest.mark.parametrize("c", ARITH, ids=[c["id"] for c in ARITH])
def test_og_arith_reproduces_reference_table(c):
    n, dt = c["n"], G.NP[c["type"]]
    r = np.zeros(n, dtype=dt); rn = np.zeros((n + 63) // 64, dtype=np.uint64); row = np.full(1, -1, dtype=np.int64)
    # unit tests run with no statement profile: division by zero yields NULL (checkDivisionByZeroBehavior, baseTemplate.go:1369-1400)
    rc = O.go().og_arith(G.ARITH_OP[c["op"]], G.TID[c["type"]], O.p(r), O.p(c["a"]), O.p(c["b"]), n, int(c["c1"]), int(c["c2"]),
                         O.p(c["n1"]), O.p(c["n2"]), O.p(rn), 1, O.p(row))
    if c["want_err"]:
        assert rc != 0 and row[0] >= 0
        return
    assert rc == 0
    nulls = [bool((int(rn[i >> 6]) >> (i & 63)) & 1) for i in range(n)]
    assert nulls == c["want_nulls"]
    for i in range(n):
        if not nulls[i]:
            assert same_value(c["type"], r[i], c["want"][i]), (i, r[i], c["want"][i])


@pytest.mark.parametrize("c", CMP, ids=[c["id"] for c in CMP])
def test_og_compare_reproduces_reference_table(c):
    n = c["n"]
    r = np.zeros(n, dtype=np.uint8); rn = np.zeros((n + 63) // 64, dtype=np.uint64)
    rc = O.go().og_compare(G.CMP_OP[c["op"]], G.TID[c["type"]], O.p(r), O.p(c["a"]), O.p(c["b"]), n, int(c["c1"]), int(c["c2"]), O.p(c["n1"]), O.p(c["n2"]), O.p(rn))
    assert rc == 0 and not c["want_err"]
    nulls = [bool((int(rn[i >> 6]) >> (i & 63)) & 1) for i in range(n)]
    assert nulls == c["want_nulls"]
    for i in range(n):
        if not nulls[i]:
            assert bool(r[i]) == bool(c["want"][i])


@pytest.mark.parametrize("c", LOGIC, ids=[c["id"] for c in LOGIC])
def test_og_logic_reproduces_reference_table(c):
    n = c["n"]
    if c["op"] == "xor":   # xorFn: plain a != b on non-null rows (logicalOperator.go:23-28) -- the oracle's restatement is og_compare NE on bools
        r = np.zeros(n, dtype=np.uint8); rn = np.zeros(1, dtype=np.uint64)
        assert O.go().og_compare(1, 10, O.p(r), O.p(c["a"]), O.p(c["b"]), n, 0, 0, O.p(c["n1"]), O.p(c["n2"]), O.p(rn)) == 0
    else:
        r = np.zeros(n, dtype=np.uint8); rn = np.zeros(1, dtype=np.uint64)
        cols = (C.c_void_p * 2)(O.p(c["a"]), O.p(c["b"])); nulls = (C.c_void_p * 2)(O.p(c["n1"]), O.p(c["n2"])); kind = (C.c_int32 * 2)(0, 0)
        assert O.go().og_multi_logic(1 if c["op"] == "or" else 0, O.p(r), O.p(rn), 2, cols, nulls, kind, n) == 0
    nulls = [bool((int(rn[0]) >> i) & 1) for i in range(n)]
    assert nulls == c["want_nulls"]
    for i in range(n):
        if not nulls[i]:
            assert bool(r[i]) == bool(c["want"][i])


def test_compare_f32_scale_rounds_both_sides_first():
    """func_compare.go:725-734: float32 columns declared with scale > 0 compare after rounding to `scale` decimals"""
    a = np.asarray([1.234, 1.235, -2.675, 0.1, 3.0, 1e10], dtype=np.float32)
    b = np.asarray([1.2341, 1.2449, -2.67, 0.1000001, 2.9951, 1e10], dtype=np.float32)
    for scale in (1, 2, 3):
        pw = 10.0 ** scale
        # math.Round = half away from zero
        rnd = lambda v: np.float32(np.copysign(np.floor(abs(float(v)) * pw + 0.5), float(v)) / pw)
        for op, f in ((0, np.equal), (1, np.not_equal), (2, np.greater), (3, np.greater_equal), (4, np.less), (5, np.less_equal)):
            r = np.zeros(6, dtype=np.uint8); rn = np.zeros(1, dtype=np.uint64)
            assert O.go().og_compare_f32_scale(op, scale, O.p(r), O.p(a), O.p(b), 6, 0, 0, None, None, O.p(rn)) == 0
            want = [bool(f(rnd(x), rnd(y))) for x, y in zip(a, b)]
            assert [bool(v) for v in r] == want, (scale, op)
    # scale makes a difference: 1.234 vs 1.2341 differ as float32, are equal at 2 and 3 decimals
    r = np.zeros(6, dtype=np.uint8); rn = np.zeros(1, dtype=np.uint64)
    O.go().og_compare(0, 30, O.p(r), O.p(a), O.p(b), 6, 0, 0, None, None, O.p(rn))
    assert not r[0]
    O.go().og_compare_f32_scale(0, 2, O.p(r), O.p(a), O.p(b), 6, 0, 0, None, None, O.p(rn))
    assert r[0]


def test_q6_q1_chains_reproduce_reference_results_on_reference_lineitem():
    from matrixone_b200 import datagen
    cols, ints, expected = G.tpch_fixture()
    n = len(cols["shipdate"])
    assert n == 6005
    for threads in (1, 3):
        s, ns, nul = O.q6(cols, n, datagen.q6_params(), nthreads=threads)
        G.check_q6_result(s, expected)
        assert not nul and ns > 0
        G.check_q1_result(O.q1(cols, n, datagen.Q1_CUTOFF, nthreads=threads), expected)
    # the DECIMAL(15,2) result is exact in integers: sum(price_cents * disc_pct) / 10^4 -- the fixture pins the predicate set too
    P = datagen.q6_params()
    m = (cols["shipdate"] >= P[0]) & (cols["shipdate"] < P[1]) & (ints["discount_pct"] >= 2) & (ints["discount_pct"] <= 4) & (ints["quantity"] < 24)
    exact = int((ints["extendedprice_cents"][m] * ints["discount_pct"][m]).sum())
    assert "%d.%04d" % (exact // 10000, exact % 10000) == expected["q6_revenue"]
