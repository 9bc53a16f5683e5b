"""numpy twin of csrc/datagen.cu: the same counter-based columns, bit for bit, on the host.

Used to feed the CPU oracle the exact rows the GPU generated for itself (tests, bench cpu_baseline) without moving
multi-GB columns over PCIe.  Pure data synthesis -- no operator logic lives here.
"""
import numpy as np

M1 = np.uint64(0x9E3779B97F4A7C15)
M2 = np.uint64(0xBF58476D1CE4E5B9)
M3 = np.uint64(0x94D049BB133111EB)
STREAM_MUL = np.uint64(0xD6E8FEB86659FD93)

DATE_1992_01_02 = 8036   # days since 1970-01-01
DATE_1995_06_17 = 9298
DATE_1994_01_01 = 8766
DATE_1995_01_01 = 9131
DATE_1998_12_01 = 10561
Q1_CUTOFF = DATE_1998_12_01 - 112   # q1.sql:16: date '1998-12-01' - interval '112' day = 1998-08-11


def mix64(x):
    with np.errstate(over="ignore"):
        x = (x + M1).astype(np.uint64)
        x = ((x ^ (x >> np.uint64(30))) * M2).astype(np.uint64)
        x = ((x ^ (x >> np.uint64(27))) * M3).astype(np.uint64)
        return x ^ (x >> np.uint64(31))


def hash3(seed, stream, rows):
    with np.errstate(over="ignore"):
        s = mix64(np.asarray([np.uint64(seed) ^ (np.uint64(stream) * STREAM_MUL)], dtype=np.uint64))[0]
        return mix64((rows.astype(np.uint64) + s).astype(np.uint64))


def lineitem(seed, row0, n):
    """dict of columns for rows [row0, row0+n): identical to MoB200_GenLineitem."""
    r = np.arange(row0, row0 + n, dtype=np.uint64)
    sd = (DATE_1992_01_02 + (hash3(seed, 1, r) % np.uint64(2526)).astype(np.int64)).astype(np.int32)
    q = np.uint64(1) + hash3(seed, 2, r) % np.uint64(50)
    cents = q * (np.uint64(90000) + hash3(seed, 3, r) % np.uint64(120001))
    h6 = hash3(seed, 6, r)
    receipt = sd.astype(np.int64) + 1 + (h6 % np.uint64(30)).astype(np.int64)
    rf = np.where(receipt <= DATE_1995_06_17, np.where(((h6 >> np.uint64(32)) & np.uint64(1)) == 1, ord("R"), ord("A")), ord("N")).astype(np.uint8)
    ls = np.where(sd <= DATE_1995_06_17, ord("F"), ord("O")).astype(np.uint8)
    return {
        "shipdate": sd,
        "quantity": q.astype(np.float64),
        "extendedprice": cents.astype(np.float64) / 100.0,
        "discount": (hash3(seed, 4, r) % np.uint64(11)).astype(np.float64) / 100.0,
        "tax": (hash3(seed, 5, r) % np.uint64(9)).astype(np.float64) / 100.0,
        "returnflag": rf,
        "linestatus": ls,
    }


def int64_column(seed, row0, n, null_per_mille=0):
    r = np.arange(row0, row0 + n, dtype=np.uint64)
    vals = (hash3(seed, 1, r) & np.uint64(0xFFFFFFFF)).astype(np.uint32).view(np.int32).astype(np.int64)
    nulls = None
    if null_per_mille:
        nulls = (hash3(seed, 2, r) % np.uint64(1000)) < np.uint64(null_per_mille)
    return vals, nulls


def vectors_f32(seed, row0, n, dim, centers=None, sigma=1.0):
    r = np.arange(row0, row0 + n, dtype=np.uint64)
    out = np.empty((n, dim), dtype=np.float32)
    scale = np.float32(2.6428965e-05)
    cidx = None
    if centers is not None:
        cidx = (hash3(seed, 7, r) % np.uint64(centers.shape[0])).astype(np.int64)
    for j in range(dim):
        h = hash3(seed, 16 + j, r)
        s = ((h & np.uint64(0xFFFF)).astype(np.int64) + ((h >> np.uint64(16)) & np.uint64(0xFFFF)).astype(np.int64)
             + ((h >> np.uint64(32)) & np.uint64(0xFFFF)).astype(np.int64) + ((h >> np.uint64(48)) & np.uint64(0xFFFF)).astype(np.int64) - 131070)
        z = s.astype(np.float32) * scale
        if centers is not None:
            z = centers[cidx, j].astype(np.float32) + np.float32(sigma) * z
        out[:, j] = z
    return out


def vector_components(seed, row0, n, ncenters):
    """mixture component every row of vectors_f32(seed, row0, n, dim, centers) was drawn from"""
    r = np.arange(row0, row0 + n, dtype=np.uint64)
    return (hash3(seed, 7, r) % np.uint64(ncenters)).astype(np.int32)


def q6_params():
    """(date_lo, date_hi, disc_lo, disc_hi, qty_hi) with the constants folded in float64 like q6.sql:58-61."""
    return DATE_1994_01_01, DATE_1995_01_01, 0.03 - 0.01, 0.03 + 0.01, 24.0
