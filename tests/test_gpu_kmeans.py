"""Elkan k-means on the GPU (csrc/kmeans.cu) against the oracle restatement of ElkanClusterer (oracle_go.c og_km_cluster_*, pinned to the reference's
step tables by tests/test_oracle_kmeans.py): final centroids BIT FOR BIT, assignments and iteration count equal."""
import json
import os

import numpy as np
import pytest

import oracle_lib as O
from matrixone_b200 import capi, ops

pytestmark = pytest.mark.gpu
K = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "kmeans_kat.json")))


def _oracle(v, init, max_iter, rnd=None):
    n, dim = v.shape; k = len(init)
    cent = np.ascontiguousarray(init, dtype=v.dtype).copy(); assign = np.zeros(n, np.int64)
    fn = O.go().og_km_cluster_f32 if v.dtype == np.float32 else O.go().og_km_cluster_f64
    r = np.ascontiguousarray(rnd, dtype=np.float32) if rnd is not None else None
    iters = fn(O.p(v), n, dim, O.p(cent), k, max_iter, O.p(r), 0 if r is None else len(r), O.p(assign))
    return cent, assign, iters


def _check(v, init, max_iter=500, rnd=None):
    oc, oa, oi = _oracle(v, init, max_iter, rnd)
    gc, ga, gi = ops.kmeans_elkan(v, init, max_iter, rnd)
    assert oi > 0 and gi == oi, (gi, oi)
    assert (ga == oa).all(), np.flatnonzero(ga != oa)[:10]
    assert gc.tobytes() == oc.tobytes(), np.abs(gc - oc).max()
    return gc, ga, gi


def test_reference_table_vectors(gpu):
    v = np.array(K["init_bounds"]["vectors"]); init = np.array(K["init_bounds"]["centroids"])
    c, a, it = _check(v, init)
    assert list(a) == [0, 0, 0, 1, 1]
    np.testing.assert_allclose(c, K["recalc"]["centroids"], rtol=1e-12)          # the converged means are the recalculateCentroids table's


def test_reference_test_cluster_expectation_with_random_init(gpu):
    """Test_Cluster (clusterer_test.go:441-470) end to end: kmeans.Random initialisation (Go's PCG restated on the host side) + the GPU loop"""
    v = np.array([[1, 2, 3, 4], [1, 2, 4, 5], [1, 2, 4, 5], [1, 2, 3, 4], [1, 2, 4, 5], [1, 2, 4, 5],
                  [10, 2, 4, 5], [10, 3, 4, 5], [10, 5, 4, 5], [10, 2, 4, 5], [10, 3, 4, 5], [10, 5, 4, 5]], dtype=np.float64)
    c, a, it = ops.kmeans_elkan(v, 2)
    np.testing.assert_allclose(c, [[10, 3.333333333333333, 4, 5], [1, 2, 3.6666666666666665, 4.666666666666666]], rtol=1e-12)
    assert list(a) == [1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0]


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("n,k,dim", [(3000, 16, 64), (2000, 7, 13), (5000, 64, 128), (400, 3, 1), (20_000, 128, 96)])
def test_matches_oracle_bit_for_bit(gpu, dtype, n, k, dim):
    rng = np.random.default_rng(n + k + dim)
    centers = rng.standard_normal((k, dim)) * 3
    v = (centers[rng.integers(0, k, n)] + rng.standard_normal((n, dim))).astype(dtype)
    init = v[rng.choice(n, k, replace=False)]
    rnd = rng.random(dim * k * 8).astype(np.float32)      # two initial centroids drawn from one blob leave a cluster empty: it re-seeds from this stream
    _check(v, init, max_iter=30, rnd=rnd)


def test_max_iterations_and_empty_clusters(gpu):
    rng = np.random.default_rng(9)
    v = rng.standard_normal((1000, 8)).astype(np.float32)
    init = v[rng.choice(1000, 5, replace=False)].copy()
    init[3] = 1e6                      # nobody is near it: the cluster is empty after the first assignment and re-seeds from the rnd stream
    init[4] = -1e6
    rnd = rng.random(64).astype(np.float32)
    _check(v, init, max_iter=3, rnd=rnd)
    _check(v, init, max_iter=1, rnd=rnd)
    with pytest.raises(capi.MoError):
        ops.kmeans_elkan(v, init, 5, rnd[:4])       # the stream runs dry
