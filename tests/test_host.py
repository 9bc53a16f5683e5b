"""Host-side layout helpers (CPU): bitmap words, varlena cells, synthetic-column twin."""
import numpy as np

from matrixone_b200 import datagen
from matrixone_b200.vector import (bitmap_from_bools, bitmap_to_bools, bitmap_words, varlena_char1_column, varlena_column,
                                   varlena_column_from_matrix, Vector)


def test_bitmap_roundtrip_lsb_first():
    m = np.zeros(130, dtype=bool); m[[0, 63, 64, 129]] = True
    w = bitmap_from_bools(m)
    assert w.dtype == np.uint64 and len(w) == bitmap_words(130) == 3
    assert w[0] == (1 | (1 << 63)) and w[1] == 1 and w[2] == 2      # bit i at words[i>>6] & (1 << (i & 63)), bitmap.go:196-229
    assert (bitmap_to_bools(w, 130) == m).all()


def test_varlena_inline_and_area_cells():
    rows = [np.arange(3, dtype=np.float32), np.arange(768, dtype=np.float32), np.zeros(0, dtype=np.float32)]
    cells, area = varlena_column(rows)
    c = cells.reshape(3, 24)
    assert c[0, 0] == 12 and (c[0, 1:13].view(np.float32) == rows[0]).all()        # inline: bs[0] = len <= 23
    u = c[1].view(np.uint32)
    assert u[0] == 0xFFFFFFFF and u[1] == 0 and u[2] == 3072 and len(area) == 3072    # big: offset, len (varlena.h:63-105)
    assert c[2, 0] == 0
    m = np.arange(20, dtype=np.float32).reshape(2, 10)
    cells2, area2 = varlena_column_from_matrix(m)
    assert (cells2 == varlena_column(list(m))[0]).all() and (area2 == m.view(np.uint8).reshape(-1)).all()
    k = varlena_char1_column(np.frombuffer(b"ANR", dtype=np.uint8)).reshape(3, 24)
    assert (k[:, 0] == 1).all() and bytes(k[:, 1]) == b"ANR"


def test_fill_raw_ptr_len_const_vector_is_24_bytes():
    cells, area = varlena_column_from_matrix(np.ones((1, 16), dtype=np.float32))
    a = Vector(data=cells, area=area, length=8192, const=True).fill_raw_ptr_len()
    assert a.dataSz == 24 and a.areaSz == 64        # const detection is dataSz == 24, xcall.c:38-39


def test_lineitem_twin_is_deterministic_and_shaped_like_dbgen():
    a = datagen.lineitem(10, 1000, 50_000)
    b = datagen.lineitem(10, 0, 51_000)
    for k in a:
        assert (a[k] == b[k][1000:]).all()           # pure function of (seed, row)
    assert a["shipdate"].min() >= datagen.DATE_1992_01_02 and a["shipdate"].max() < datagen.DATE_1992_01_02 + 2526
    assert set(np.unique(a["quantity"])) <= set(range(1, 51))
    assert np.allclose(np.unique(a["discount"]), np.arange(11) / 100.0)
    assert set(np.unique(a["returnflag"])) == {ord("A"), ord("N"), ord("R")} and set(np.unique(a["linestatus"])) == {ord("F"), ord("O")}
    pairs = set(zip(a["returnflag"].tolist(), a["linestatus"].tolist()))
    assert (ord("N"), ord("F")) in pairs and (ord("R"), ord("O")) not in pairs   # the narrow N/F band exists, R/O cannot
    assert abs((a["extendedprice"] * 100).round() - a["extendedprice"] * 100).max() < 1e-6


def test_vector_twin_moments():
    v = datagen.vectors_f32(20, 0, 4000, 16)
    assert abs(v.mean()) < 0.02 and abs(v.std() - 1.0) < 0.02
    assert (datagen.vectors_f32(20, 100, 10, 16) == v[100:110]).all()
