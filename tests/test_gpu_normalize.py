"""normalize_l2 (vector-valued XCall ids 112/113) against the oracle restatement of metric.NormalizeL2
(pkg/vectorindex/metric/distance_func.go:411-434 == pkg/vectorize/moarray/external.go:262-285) and the reference's own
TestNormalizeL2 table (external_test.go:400-480, float32 compared with reflect.DeepEqual = bit-exact)."""
import json
import os

import numpy as np
import pytest

import oracle_lib as O
from matrixone_b200 import capi, ops
from matrixone_b200.vector import bitmap_from_bools, varlena_column

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _decode(cells, area, i, dtype):
    c = cells[24 * i:24 * i + 24]
    if c[0] <= 23:
        return c[1:1 + c[0]].copy().view(dtype)
    off, ln = c.view(np.uint32)[1], c.view(np.uint32)[2]
    return area[off:off + ln].copy().view(dtype)


def _oracle(v, dtype):
    out = np.empty_like(v)
    fn = O.go().og_normalize_l2_f32 if dtype == np.float32 else O.go().og_normalize_l2_f64
    assert fn(O.p(v), O.p(out), v.shape[0]) == 0
    return out


def test_reference_table_bit_exact(gpu):
    k = json.load(open(os.path.join(GOLD, "moarray_kat.json")))
    for c in k["normalize_l2"]:
        dt = np.float32 if c["dtype"] == "f32" else np.float64
        v = np.array(c["v"], dtype=dt)
        cells, area = varlena_column([v], dtype=dt)
        oc, oa = ops.normalize_l2(cells, area, 1, dt)
        got = _decode(oc, oa, 0, dt)
        want = np.array(c["want"], dtype=dt)
        if dt == np.float32:
            assert got.tobytes() == want.tobytes(), (c, got)
        else:
            np.testing.assert_allclose(got, want, rtol=1e-15, atol=0)
        assert got.tobytes() == _oracle(v, dt).tobytes()


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("dim", [1, 3, 5, 8, 33, 128, 131, 768])
def test_matches_oracle_with_nulls_and_zero_rows(gpu, dtype, dim):
    rng = np.random.default_rng(dim * 7 + (1 if dtype == np.float64 else 0))
    n = 1000
    mat = (rng.standard_normal((n, dim)) * rng.choice([1e-3, 1.0, 1e3], size=(n, 1))).astype(dtype)
    mat[5] = 0            # zero vector: copied
    mat[17] = 0
    null = rng.random(n) < 0.1
    null[5] = False
    cells, area = varlena_column([mat[i] for i in range(n)], dtype=dtype)
    oc, oa = ops.normalize_l2(cells, area, n, dtype, nulls=bitmap_from_bools(null))
    for i in range(n):
        if null[i]:
            assert not oc[24 * i:24 * i + 24].any()
            continue
        assert _decode(oc, oa, i, dtype).tobytes() == _oracle(mat[i], dtype).tobytes(), i
    # the result mirrors the argument's layout
    live = ~null
    assert (oc.reshape(n, 24)[live][:, 0] == cells.reshape(n, 24)[live][:, 0]).all()


def test_const_argument_and_unaligned_rows(gpu):
    rng = np.random.default_rng(3)
    v = rng.standard_normal(100).astype(np.float32)
    cells, area = varlena_column([v], dtype=np.float32)
    oc, oa = ops.normalize_l2(cells, area, 7, np.float32)     # dataSz == 24, len 7: a const argument
    want = _oracle(v, np.float32).tobytes()
    for i in range(7):
        assert _decode(oc, oa, i, np.float32).tobytes() == want
    # ragged rows: offsets that are not multiples of 16 bytes
    rows = [rng.standard_normal(d).astype(np.float32) for d in (7, 9, 6, 31, 64, 10, 100)]
    cells, area = varlena_column(rows, dtype=np.float32)
    oc, oa = ops.normalize_l2(cells, area, len(rows), np.float32)
    for i, r in enumerate(rows):
        assert _decode(oc, oa, i, np.float32).tobytes() == _oracle(r, np.float32).tobytes()


def test_empty_vector_is_an_error(gpu):
    cells, area = varlena_column([np.ones(4, np.float32), np.zeros(0, np.float32)], dtype=np.float32)
    with pytest.raises(capi.MoError) as e:
        ops.normalize_l2(cells, area, 2, np.float32)
    assert "cannot normalize empty vector" in str(e.value)
