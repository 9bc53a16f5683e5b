#!/bin/bash
# round 2: vectorised plan interpreter -- parity + micro-benchmark
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_plan.py tests/test_gpu_tpch.py -x -q -m gpu 2>&1 | tail -5
timeout 600 python tools/profile_ops.py > gpurun_out/r02_ops_microbench_v2.json 2> gpurun_out/r02_ops_microbench_v2.err
tail -3 gpurun_out/r02_ops_microbench_v2.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02_ops_microbench_v2.json'))
for k, v in d.items():
    if 'plan' in k: print(k, v)
PY
