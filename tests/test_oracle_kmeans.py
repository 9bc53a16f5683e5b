"""Pins the Elkan k-means restatement (oracle_go.c og_km_*) to the reference's own per-step tables (clusterer_test.go:501-939, transcribed by
tests/golden/extract_goldens.py --kmeans into tests/golden/kmeans_kat.json) and checks the whole loop against a plain Lloyd iteration in
numpy on well-separated data (Elkan's bounds only skip distance computations; assignments and means agree).  No GPU."""
import json
import os

import numpy as np

import oracle_lib as O

K = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "kmeans_kat.json")))


def _close(a, b):   # assertx.InEpsilonF64
    return np.allclose(np.asarray(a, np.float64), np.asarray(b, np.float64), rtol=1e-9, atol=1e-9)


def test_init_bounds_table():
    c = K["init_bounds"]
    v = np.array(c["vectors"]); cent = np.array(c["centroids"]); n, dim = v.shape; k = len(cent)
    lower = np.zeros((n, k)); upper = np.zeros(n); assign = np.zeros(n, np.int64)
    O.go().og_km_init_bounds_f64(O.p(v), n, dim, O.p(cent), k, O.p(lower), O.p(upper), O.p(assign))
    assert list(assign) == c["assignment"]
    for i, m in enumerate(c["metas"]):
        assert _close(lower[i], m["lower"]) and _close(upper[i], m["upper"])


def test_centroid_distances_table():
    c = K["centroid_dists"]
    cent = np.array(c["centroids"]); k, dim = cent.shape
    half = np.zeros((k, k)); minhalf = np.zeros(k)
    O.go().og_km_centroid_dists_f64(O.p(cent), k, dim, O.p(half), O.p(minhalf))
    assert _close(half, c["half"]) and list(minhalf) == c["minhalf"]          # the reference also requires DeepEqual on minHalf


def test_recalculate_centroids_table():
    c = K["recalc"]
    v = np.array(c["vectors"]); n, dim = v.shape; k = len(c["centroids"])
    assign = np.array(c["assignments"], np.int64); newc = np.zeros((k, dim)); members = np.zeros(k, np.int64); used = np.zeros(1, np.int64)
    assert O.go().og_km_recalc_f64(O.p(v), n, dim, O.p(assign), k, O.p(newc), O.p(members), None, 0, O.p(used)) == 0
    assert _close(newc, c["centroids"]) and list(members) == [3, 2]


def test_update_bounds_table():
    c = K["update_bounds"]
    cent = np.array(c["centroids"]); newc = np.array(c["new_centroids"]); k, dim = cent.shape; n = len(c["metas"])
    lower = np.array([m["lower"] for m in c["metas"]]); upper = np.array([m["upper"] for m in c["metas"]])
    rec = np.zeros(n, np.uint8); assign = np.zeros(n, np.int64); shift = np.zeros(k)       # the test leaves km.assignments zero
    O.go().og_km_update_bounds_f64(O.p(cent), O.p(newc), k, dim, n, O.p(assign), O.p(lower), O.p(upper), O.p(rec), O.p(shift))
    for i, m in enumerate(c["want"]):
        assert _close(lower[i], m["lower"]) and _close(upper[i], m["upper"]) and bool(rec[i]) == m["recompute"]


def test_reference_whole_run_expectations_pin_the_generator_and_the_loop():
    """TestRandom_InitCentroids (initializer_test.go:36-59) and Test_Cluster (clusterer_test.go:441-470): kmeans.Random initialisation on 12 vectors,
    k = 2.  The expected initial centroids are vectors 7 and 1 -- reached only if Go's PCG(1, 0) is restated correctly (gorand.py) -- and the expected
    final centroids come out of the restated Elkan loop started from them."""
    from matrixone_b200 import gorand
    v = np.array([[1, 2, 3, 4], [1, 2, 4, 5], [1, 2, 4, 5], [1, 2, 3, 4], [1, 2, 4, 5], [1, 2, 4, 5],
                  [10, 2, 4, 5], [10, 3, 4, 5], [10, 5, 4, 5], [10, 2, 4, 5], [10, 3, 4, 5], [10, 5, 4, 5]], dtype=np.float64)
    rows = gorand.random_init_rows(12, 2)
    assert rows == [7, 1]
    assert v[rows].tolist() == [[10, 3, 4, 5], [1, 2, 4, 5]]                       # wantCentroids of TestRandom_InitCentroids
    cent = v[rows].copy(); assign = np.zeros(12, np.int64)
    iters = O.go().og_km_cluster_f64(O.p(v), 12, 4, O.p(cent), 2, 500, None, 0, O.p(assign))
    assert iters > 0
    assert _close(cent, [[10, 3.333333333333333, 4, 5], [1, 2, 3.6666666666666665, 4.666666666666666]])       # want of Test_Cluster
    sse = sum(float(np.sqrt(((v[i] - cent[assign[i]]) ** 2).sum())) ** 2 for i in range(12))
    assert abs(sse - 12) < 1e-9                                                     # wantSSE


def test_whole_loop_agrees_with_lloyd_on_separated_clusters():
    rng = np.random.default_rng(0)
    k, dim, per = 8, 16, 200
    centers = rng.standard_normal((k, dim)) * 20
    v = (centers[:, None, :] + rng.standard_normal((k, per, dim))).reshape(-1, dim).astype(np.float32)
    rng.shuffle(v)
    n = len(v)
    init = v[rng.choice(n, k, replace=False)].copy()
    cent = init.copy(); assign = np.zeros(n, np.int64)
    iters = O.go().og_km_cluster_f32(O.p(v), n, dim, O.p(cent), k, 50, None, 0, O.p(assign))
    assert 1 < iters <= 51
    c = init.astype(np.float64)
    for _ in range(60):
        a = ((v[:, None, :].astype(np.float64) - c[None]) ** 2).sum(-1).argmin(1)
        nc = np.stack([v[a == j].mean(0) if (a == j).any() else c[j] for j in range(k)])
        if np.allclose(nc, c):
            break
        c = nc
    assert (a == assign).mean() > 0.999
    assert np.allclose(cent, c, rtol=1e-4, atol=1e-4)
