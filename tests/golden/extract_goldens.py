"""Transcribe the known-answer tables of the reference's own Go tests into JSON fixtures (data only).

Run in the build container (needs /root/reference):   python tests/golden/extract_goldens.py
Writes tests/golden/{metric_kat,moarray_kat,heap_kat,tpch_kat,agg_kat}.json.  The fixtures are committed; the GPU box
and the CPU test-suite read only the JSON.

Sources (paths relative to /root/reference):
  pkg/vectorindex/metric/distance_func_test.go   Test_L2Distance/L1Distance/CosineDistance/InnerProduct/L2DistanceSq
                                                  (exact f64 equality), Test_ZeroVector
  pkg/vectorize/moarray/external_test.go         TestInnerProduct/CosineSimilarity/L2Distance/CosineDistance/
                                                  NormalizeL2 (f32 and f64 variants, InEpsilonF64)
  pkg/vectorindex/index_test.go:215-249          TestFastMaxHeap push sequence + expected ascending pop order
  test/distributed/cases/benchmark/tpch/03_QUERIES/q{1,6}.result   tiny-scale SQL goldens (DECIMAL schema)
  pkg/sql/colexec/aggexec/sumavg2_test.go        sum/avg over 1..10 with nulls (expected values stated in-test)
"""
import json
import os
import re

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))

NUM = r"[-+]?(?:\d+\.?\d*(?:[eE][-+]?\d+)?|\.\d+)"


def _func_body(src, name):
    m = re.search(r"^func %s\(t \*testing\.T\) \{" % re.escape(name), src, re.M)
    if not m:
        raise KeyError(name)
    start = m.end()
    nxt = re.search(r"^func ", src[start:], re.M)
    return src[start:start + nxt.start()] if nxt else src[start:]


def _floats(s):
    return [float(x) for x in re.findall(NUM, s)]


def metric_kats():
    src = open(os.path.join(REF, "pkg/vectorindex/metric/distance_func_test.go")).read()
    out = {"_source": "pkg/vectorindex/metric/distance_func_test.go", "_compare": "exact float64 equality"}
    for fn, key in (("Test_L2Distance", "l2"), ("Test_L1Distance", "l1"), ("Test_CosineDistance", "cosine_distance"),
                    ("Test_InnerProduct", "inner_product"), ("Test_L2DistanceSq", "l2sq")):
        body = _func_body(src, fn)
        cases = []
        for m in re.finditer(r"v1:\s*\[\]float64\{([^}]*)\},\s*v2:\s*\[\]float64\{([^}]*)\},\s*\},\s*want:\s*(%s)," % NUM, body):
            cases.append({"v1": _floats(m.group(1)), "v2": _floats(m.group(2)), "want": float(m.group(3))})
        assert cases, fn
        out[key] = cases
    out["zero_vector_cosine_distance"] = {"v1": [0, 0, 0], "v2": [0, 0, 0], "want": 1.0}  # Test_ZeroVector :142-156
    return out


def moarray_kats():
    src = open(os.path.join(REF, "pkg/vectorize/moarray/external_test.go")).read()
    out = {"_source": "pkg/vectorize/moarray/external_test.go", "_compare": "assertx.InEpsilonF64"}
    for fn, key in (("TestInnerProduct", "inner_product"), ("TestCosineSimilarity", "cosine_similarity"),
                    ("TestL2Distance", "l2"), ("TestCosineDistance", "cosine_distance")):
        body = _func_body(src, fn)
        cases = []
        pat = (r"args\{argLeft(F32|F64):\s*\[\]float(?:32|64)\{([^}]*)\},\s*argRight(?:F32|F64):\s*\[\]float(?:32|64)\{([^}]*)\}\},"
               r"\s*want:\s*(%s)," % NUM)
        for m in re.finditer(pat, body):
            cases.append({"dtype": m.group(1).lower(), "v1": _floats(m.group(2)), "v2": _floats(m.group(3)),
                          "want": float(m.group(4))})
        assert cases, fn
        out[key] = cases
    body = _func_body(src, "TestNormalizeL2")
    cases = []
    for m in re.finditer(r"args\{arg(F32|F64):\s*\[\]float(?:32|64)\{([^}]*)\}\},\s*want(?:F32|F64):\s*\[\]float(?:32|64)\{([^}]*)\}", body):
        cases.append({"dtype": m.group(1).lower(), "v": _floats(m.group(2)), "want": _floats(m.group(3))})
    out["normalize_l2"] = cases
    return out


def heap_kats():
    src = open(os.path.join(REF, "pkg/vectorindex/index_test.go")).read()
    body = _func_body(src, "TestFastMaxHeap")
    pushes = [{"key": int(m.group(1)), "dist": float(m.group(2))}
              for m in re.finditer(r"h\.Push\((\d+),\s*float32\((%s)\)\)" % NUM, body)]
    limit = int(re.search(r"limit\s*:=\s*(\d+)", body).group(1))
    pops = [{"key": int(k), "dist": float(d)} for k, d in
            re.findall(r"require\.Equal\(t, int64\((\d+)\), key\)\s*require\.Equal\(t, float32\((%s)\), dist\)" % NUM, body)]
    assert pushes and pops
    return {"_source": "pkg/vectorindex/index_test.go:215-249", "limit": limit, "pushes": pushes, "pops_in_order": pops}


def tpch_kats():
    base = os.path.join(REF, "test/distributed/cases/benchmark/tpch/03_QUERIES")
    q1 = open(os.path.join(base, "q1.result")).read().strip().splitlines()
    q6 = open(os.path.join(base, "q6.result")).read().strip().splitlines()
    hdr = [i for i, l in enumerate(q1) if l.startswith("l_returnflag") and "sum_qty" in l][0]
    rows = [l.split() for l in q1[hdr + 1:] if l.strip()]
    return {"_source": "test/distributed/cases/benchmark/tpch/03_QUERIES/q1.result, q6.result (tiny dataset, DECIMAL schema)",
            "q1_columns": q1[hdr].split(), "q1_rows": rows, "q6_revenue": q6[-1].strip()}


def agg_kats():
    """sumavg2_test.go drives values 1..10 (some with nulls) through sum/avg; expectations are stated in-test as
    arithmetic over the same constants.  We record the input recipe and the closed forms the test asserts."""
    return {
        "_source": "pkg/sql/colexec/aggexec/sumavg2_test.go:38-110,210-335; count2_test.go:123-135",
        "values": list(range(1, 11)),
        "sum_all": 55, "avg_all": 5.5, "count_all": 10,
        "null_rows_example": [1, 4],  # 0-based rows nulled in our derived case: sum 55-2-5=48, cnt 8
        "sum_with_nulls": 48, "avg_with_nulls": 6.0, "count_with_nulls": 8,
        "tolerance_abs": 1e-6,
    }


def main():
    for name, fn in (("metric_kat", metric_kats), ("moarray_kat", moarray_kats), ("heap_kat", heap_kats),
                     ("tpch_kat", tpch_kats), ("agg_kat", agg_kats)):
        data = fn()
        with open(os.path.join(OUT, name + ".json"), "w") as f:
            json.dump(data, f, indent=1)
        print(name, {k: (len(v) if isinstance(v, list) else "") for k, v in data.items() if not k.startswith("_")})


if __name__ == "__main__":
    main()
