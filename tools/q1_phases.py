#!/usr/bin/env python
"""Per-CTA phase timeline of the Q1 kernel (globaltimer stamps): where does the time go?"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from matrixone_b200 import capi, datagen, ops
from matrixone_b200.vector import DeviceBuffer
lib = capi.load_library(); capi.check(lib.MoB200_Init(0), lib)
n = int(os.environ.get("TUNE_ROWS", 200_000_000))
names = ["shipdate", "quantity", "extendedprice", "discount", "tax", "returnflag", "linestatus"]
size = {"shipdate": 4, "returnflag": 1, "linestatus": 1}
b = {k: DeviceBuffer(size.get(k, 8) * n, lib) for k in names}
capi.check(lib.MoB200_GenLineitem(10, 0, n, *[b[k].ptr for k in ["shipdate", "quantity", "extendedprice", "discount", "tax", "returnflag", "linestatus"]]), lib)
lib.MoB200_SetTuning(b"q1_debug", 1)
for _ in range(3):
    ops.q1_group_agg(b["shipdate"], b["quantity"], b["extendedprice"], b["discount"], b["tax"], b["returnflag"], b["linestatus"], n, datagen.Q1_CUTOFF)
for rep in range(3):
    capi.check(lib.MoB200_Memset(lib.MoB200_DebugBuffer(), 0, 8 * 16 * 1024), lib)
    ops.q1_group_agg(b["shipdate"], b["quantity"], b["extendedprice"], b["discount"], b["tax"], b["returnflag"], b["linestatus"], n, datagen.Q1_CUTOFF)
    dbg = np.zeros(16 * 1024, dtype=np.uint64)
    capi.check(lib.MoB200_Download(dbg.ctypes.data, lib.MoB200_DebugBuffer(), dbg.nbytes), lib)
    d = dbg.reshape(1024, 16)[:296].astype(np.int64)
    t0 = d[:, 0].min()
    loop = (d[:, 8:16].max(axis=1) - t0) / 1e3
    order = np.argsort(loop)
    print("rep", rep, "slowest-warp loop end per CTA: median %.1f us max %.1f us" % (np.median(loop), loop.max()))
    for c in order[-4:]:
        print("   cta %3d smid %3d slowpath %5d  warp ends(us):" % (c, d[c, 5], d[c, 7]), np.round((d[c, 8:16] - t0) / 1e3, 1), " ticket %.1f" % ((d[c, 3] - t0) / 1e3))
    print("   slow-path entries: total %d max/cta %d ; CTAs sharing the slow smid:" % (d[:, 7].sum(), d[:, 7].max()), np.flatnonzero(d[:, 5] == d[order[-1], 5]))
