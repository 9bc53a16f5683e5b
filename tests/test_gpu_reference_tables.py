"""The reference-held tables and fixtures of tests/test_oracle_reference_tables.py, through the GPU (C-ABI):

  * FunctionTestCase tables of pkg/sql/plan/function/*_test.go -> MO_XCALL_GO_ARITH / GO_COMPARE / GO_MULTI_AND / GO_MULTI_OR / Logic_VecXor;
  * the reference's 6005-row lineitem -> fused Q6 / Q1 kernels (host and resident, synchronous and device-result forms) vs
    03_QUERIES/q6.result, q1.result;
  * float32 compare with scale > 0 (func_compare.go:725-734) vs the oracle restatement."""
import numpy as np
import pytest

import golden_tables as G
import oracle_lib as O
from matrixone_b200 import capi, datagen, ops
from matrixone_b200.vector import DeviceBuffer, Vector, xcall
from test_oracle_reference_tables import ARITH, CMP, LOGIC, same_value

pytestmark = pytest.mark.gpu


def _nulls_list(rn, n):
    return [bool((int(rn[i >> 6]) >> (i & 63)) & 1) for i in range(n)]


@pytest.mark.parametrize("c", ARITH, ids=[c["id"] for c in ARITH])
def test_gpu_arith_reproduces_reference_table(gpu, c):
    n, dt = c["n"], G.NP[c["type"]]
    r = np.zeros(n, dtype=dt); rn = np.zeros((n + 63) // 64, dtype=np.uint64)
    params = np.zeros(2, dtype=np.int64); params.view(np.int32)[0] = 1; params[1] = -1      # div0 -> NULL (SELECT behaviour)
    rc, msg = xcall(capi.XCALL_GO_ARITH(G.ARITH_OP[c["op"]], G.TID[c["type"]]),
                    [Vector(data=r, nulls=rn, length=n), Vector(data=c["a"], nulls=c["n1"], length=n), Vector(data=c["b"], nulls=c["n2"], length=n),
                     Vector(data=params.view(np.uint8), length=n)], n, raise_on_error=False)
    if c["want_err"]:
        assert rc in (capi.RC_OUT_OF_RANGE, capi.RC_DIVISION_BY_ZERO) and params[1] >= 0, (rc, msg)
        return
    assert rc == 0, msg
    nulls = _nulls_list(rn, n)
    assert nulls == c["want_nulls"]
    for i in range(n):
        if not nulls[i]:
            assert same_value(c["type"], r[i], c["want"][i]), (i, r[i], c["want"][i])


@pytest.mark.parametrize("c", CMP, ids=[c["id"] for c in CMP])
def test_gpu_compare_reproduces_reference_table(gpu, c):
    n = c["n"]
    r = np.zeros(n, dtype=np.uint8); rn = np.zeros((n + 63) // 64, dtype=np.uint64)
    xcall(capi.XCALL_GO_COMPARE(G.CMP_OP[c["op"]], G.TID[c["type"]]),
          [Vector(data=r, nulls=rn, length=n), Vector(data=c["a"], nulls=c["n1"], length=n), Vector(data=c["b"], nulls=c["n2"], length=n)], n)
    nulls = _nulls_list(rn, n)
    assert nulls == c["want_nulls"]
    assert all(bool(r[i]) == bool(c["want"][i]) for i in range(n) if not nulls[i])


@pytest.mark.parametrize("c", LOGIC, ids=[c["id"] for c in LOGIC])
def test_gpu_logic_reproduces_reference_table(gpu, c):
    n = c["n"]
    r = np.zeros(n, dtype=np.uint8)
    if c["op"] == "xor":
        # xorFn (logicalOperator.go:23-28) through the mo.h entry point: the caller pre-ORs the nulls (Logic_VecXor skips null rows)
        rn = (c["n1"] if c["n1"] is not None else np.zeros(1, dtype=np.uint64)) | (c["n2"] if c["n2"] is not None else np.zeros(1, dtype=np.uint64))
        assert gpu.Logic_VecXor(r.ctypes.data, c["a"].ctypes.data, c["b"].ctypes.data, n, rn.ctypes.data, 0) == 0
    else:
        rn = np.zeros(1, dtype=np.uint64)
        cnt = np.array([2], dtype=np.int32)
        xcall(capi.XCALL_GO_MULTI_OR if c["op"] == "or" else capi.XCALL_GO_MULTI_AND,
              [Vector(data=r, nulls=rn, length=n), Vector(data=cnt.view(np.uint8), length=n), Vector(data=c["a"], nulls=c["n1"], length=n), Vector(data=c["b"], nulls=c["n2"], length=n)], n)
    nulls = _nulls_list(rn, n)
    assert nulls == c["want_nulls"]
    assert all(bool(r[i]) == bool(c["want"][i]) for i in range(n) if not nulls[i])


@pytest.mark.parametrize("scale", [1, 2, 3, 6])
@pytest.mark.parametrize("op", range(6))
def test_gpu_compare_f32_scale_matches_oracle(gpu, op, scale):
    rng = np.random.default_rng(op * 10 + scale)
    n = 20_011
    a = (rng.integers(-3000, 3000, n) / 1000.0).astype(np.float32)
    b = (a + rng.choice([0.0, 1e-4, -1e-4, 5e-3, -5e-3, 0.5], n)).astype(np.float32)
    a[:4] = [np.inf, -np.inf, np.nan, 1e30]; b[:4] = [np.inf, 1.0, 1.0, 1e30]
    for shape in ("vv", "vc", "cv"):
        c1, c2 = shape == "cv", shape == "vc"
        aa = a[:1].copy() if c1 else a; bb = b[:1].copy() if c2 else b
        n1 = None if c1 else np.packbits(rng.random(((n + 63) // 64) * 64) < 0.05, bitorder="little").view(np.uint64).copy()
        if n1 is not None and n & 63:
            n1[-1] &= np.uint64((1 << (n & 63)) - 1)
        r0 = np.full(n, 7, dtype=np.uint8); rn0 = np.zeros((n + 63) // 64, dtype=np.uint64)
        assert O.go().og_compare_f32_scale(op, scale, O.p(r0), O.p(aa), O.p(bb), n, int(c1), int(c2), O.p(n1), None, O.p(rn0)) == 0
        r1 = np.full(n, 7, dtype=np.uint8); rn1 = np.zeros((n + 63) // 64, dtype=np.uint64)
        xcall(capi.XCALL_GO_COMPARE_F32_SCALE(op, scale), [Vector(data=r1, nulls=rn1, length=n), Vector(data=aa, nulls=n1, length=n), Vector(data=bb, length=n)], n)
        assert (rn1 == rn0).all() and (r1 == r0).all(), shape


def test_gpu_q6_q1_reproduce_reference_results_on_reference_lineitem(gpu):
    cols, ints, expected = G.tpch_fixture()
    n = len(cols["shipdate"])
    P = datagen.q6_params()
    # host columns (the drop-in call)
    s, rows, nul = ops.q6_filter_sum(cols["shipdate"], cols["discount"], cols["quantity"], cols["extendedprice"], n, *P)
    G.check_q6_result(s, expected)
    os_, orows, _ = O.q6(cols, n, P)
    assert rows == orows and abs(s - os_) <= 1e-12 * abs(os_)
    q1 = ops.q1_group_agg(cols["shipdate"], cols["quantity"], cols["extendedprice"], cols["discount"], cols["tax"], cols["returnflag"], cols["linestatus"], n, datagen.Q1_CUTOFF)
    G.check_q1_result(q1, expected)
    # resident columns, device-resident results
    dev = {k: DeviceBuffer.from_numpy(v) for k, v in cols.items()}
    out6 = DeviceBuffer(16)
    ops.q6_filter_sum_device(dev["shipdate"], dev["discount"], dev["quantity"], dev["extendedprice"], n, *P, out_ptr=out6.ptr)
    G.check_q6_result(float(out6.to_numpy(np.float64)[0]), expected)
    out1 = DeviceBuffer(ops.Q1_RESULT_BYTES)
    ops.q1_group_agg_device(dev["shipdate"], dev["quantity"], dev["extendedprice"], dev["discount"], dev["tax"], dev["returnflag"], dev["linestatus"], n, datagen.Q1_CUTOFF, out1.ptr)
    G.check_q1_result(ops.q1_result_from_bytes(out1.to_numpy(np.uint8).tobytes()), expected)
    for b in list(dev.values()) + [out6, out1]:
        b.free()
