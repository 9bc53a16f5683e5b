"""Loaders for the reference-held fixtures under tests/golden/ (written by tests/golden/extract_goldens.py --round2).

function_kat.json   FunctionTestCase tables of pkg/sql/plan/function/{arithmetic_*,func_compare*,operatorSet}_test.go
tpch_lineitem.json  the 6005-row lineitem of test/distributed/cases/benchmark/tpch/02_LOAD/03_insert_lineitem.sql
tpch_kat.json       03_QUERIES/q1.result, q6.result
"""
import json
import os

import numpy as np

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NP = {"int8": np.int8, "int16": np.int16, "int32": np.int32, "int64": np.int64, "uint8": np.uint8, "uint16": np.uint16,
      "uint32": np.uint32, "uint64": np.uint64, "float32": np.float32, "float64": np.float64, "bool": np.uint8}
TID = {"bool": 10, "int8": 20, "int16": 21, "int32": 22, "int64": 23, "uint8": 25, "uint16": 26, "uint32": 27, "uint64": 28, "float32": 30, "float64": 31}
ARITH_OP = {"add": 0, "sub": 1, "mul": 2, "div": 3, "mod": 4}
CMP_OP = {"eq": 0, "ne": 1, "gt": 2, "ge": 3, "lt": 4, "le": 5}


def _dec(v):
    return {"inf": float("inf"), "-inf": float("-inf"), "nan": float("nan")}.get(v, v) if isinstance(v, str) else v


def _arr(ty, values):
    vals = [_dec(v) for v in values]
    if ty in ("uint64",):
        return np.asarray([int(v) for v in vals], dtype=np.uint64)
    if ty.startswith(("int", "uint")):
        return np.asarray([int(v) for v in vals], dtype=np.int64 if ty != "uint64" else np.uint64).astype(NP[ty])
    if ty == "bool":
        return np.asarray([1 if v else 0 for v in vals], dtype=np.uint8)
    with np.errstate(over="ignore"):
        return np.asarray(vals, dtype=np.float64).astype(NP[ty])


def bitmap(flags, n):
    w = np.zeros((n + 63) // 64, dtype=np.uint64)
    for i, f in enumerate(flags or []):
        if f:
            w[i >> 6] |= np.uint64(1) << np.uint64(i & 63)
    return w


def function_cases(ops):
    """yields dicts: op, type, n, a, b (numpy), c1, c2 (const flags), n1, n2 (bitmaps or None), want (numpy), want_nulls (bool list), want_err, id"""
    d = json.load(open(os.path.join(HERE, "function_kat.json")))
    for c in d["cases"]:
        if c["op"] not in ops or len(c["inputs"]) != 2:
            continue
        i0, i1 = c["inputs"]
        if i0["type"] != i1["type"]:
            continue
        n = max(len(i0["values"]), len(i1["values"]))
        ex = c["expect"]
        yield {"op": c["op"], "type": i0["type"], "n": n, "a": _arr(i0["type"], i0["values"]), "b": _arr(i1["type"], i1["values"]),
               "c1": bool(i0["const"]) or (len(i0["values"]) == 1 and n > 1), "c2": bool(i1["const"]) or (len(i1["values"]) == 1 and n > 1),
               "n1": bitmap(i0["nulls"], len(i0["values"])) if i0["nulls"] else None, "n2": bitmap(i1["nulls"], len(i1["values"])) if i1["nulls"] else None,
               "want": _arr(ex["type"], ex["values"]) if ex["values"] is not None else None, "want_type": ex["type"],
               "want_nulls": list(ex["nulls"]) if ex["nulls"] else [False] * n, "want_err": ex["want_err"],
               "id": "%s:%d" % (c["file"], c["line"])}


def tpch_fixture():
    """the reference's tiny lineitem as the fp64 columns the fused kernels read (+ exact integer forms) and the expected results"""
    d = json.load(open(os.path.join(HERE, "tpch_lineitem.json")))
    k = json.load(open(os.path.join(HERE, "tpch_kat.json")))
    cols = {"shipdate": np.asarray(d["shipdate_days"], dtype=np.int32), "quantity": np.asarray(d["quantity"], dtype=np.float64),
            "extendedprice": np.asarray(d["extendedprice_cents"], dtype=np.float64) / 100.0, "discount": np.asarray(d["discount_pct"], dtype=np.float64) / 100.0,
            "tax": np.asarray(d["tax_pct"], dtype=np.float64) / 100.0, "returnflag": np.frombuffer(d["returnflag"].encode(), dtype=np.uint8).copy(),
            "linestatus": np.frombuffer(d["linestatus"].encode(), dtype=np.uint8).copy()}
    ints = {"quantity": np.asarray(d["quantity"], dtype=np.int64), "extendedprice_cents": np.asarray(d["extendedprice_cents"], dtype=np.int64),
            "discount_pct": np.asarray(d["discount_pct"], dtype=np.int64), "tax_pct": np.asarray(d["tax_pct"], dtype=np.int64)}
    q1 = {}
    for row in k["q1_rows"]:
        q1[(row[0], row[1])] = dict(zip(k["q1_columns"][2:], row[2:]))
    return cols, ints, {"q6_revenue": k["q6_revenue"], "q1": q1}


def check_q6_result(got_sum, expected):
    assert abs(got_sum - float(expected["q6_revenue"])) <= 1e-5 * float(expected["q6_revenue"]), (got_sum, expected["q6_revenue"])


def check_q1_result(groups, expected):
    """groups: list of dicts (returnflag, linestatus as byte values).  The .result file prints DECIMAL results: sums with the scale of
    the expression, averages rounded to 2 decimals -- sums are compared at 1e-5 relative, averages within half a unit of the last
    printed digit."""
    assert len(groups) == len(expected["q1"])
    for g in groups:
        e = expected["q1"][(chr(g["returnflag"]), chr(g["linestatus"]))]
        assert g["count_order"] == int(e["count_order"])
        for f in ("sum_qty", "sum_base_price", "sum_disc_price", "sum_charge"):
            assert abs(g[f] - float(e[f])) <= 1e-5 * float(e[f]), (f, g[f], e[f])
        for f in ("avg_qty", "avg_price", "avg_disc"):
            assert abs(g[f] - float(e[f])) <= 0.005 + 1e-9, (f, g[f], e[f])
