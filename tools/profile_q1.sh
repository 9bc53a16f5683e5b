#!/bin/bash
mkdir -p gpurun_out
for v in 0 1; do
cat > /tmp/q1v.py <<PY
import os, sys
sys.path.insert(0, os.getcwd())
from matrixone_b200 import capi, datagen, ops
from matrixone_b200.vector import DeviceBuffer
lib = capi.load_library(); capi.check(lib.MoB200_Init(0), lib)
n = 100_000_000
names = ["shipdate", "quantity", "extendedprice", "discount", "tax", "returnflag", "linestatus"]
size = {"shipdate": 4, "returnflag": 1, "linestatus": 1}
b = {k: DeviceBuffer(size.get(k, 8) * n, lib) for k in names}
capi.check(lib.MoB200_GenLineitem(10, 0, n, *[b[k].ptr for k in names[:1] + ["quantity", "extendedprice", "discount", "tax", "returnflag", "linestatus"]]), lib)
lib.MoB200_SetTuning(b"q1_variant", $v)
for _ in range(3):
    ops.q1_group_agg(b["shipdate"], b["quantity"], b["extendedprice"], b["discount"], b["tax"], b["returnflag"], b["linestatus"], n, datagen.Q1_CUTOFF)
PY
ncu --set full --clock-control none --import-source on -k regex:q1_ -s 2 -c 1 -f -o gpurun_out/r01b_q1_v$v python /tmp/q1v.py > gpurun_out/r01b_q1_v$v.out 2>&1
done
ls -la gpurun_out/r01b*
