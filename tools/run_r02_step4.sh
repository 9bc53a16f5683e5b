#!/bin/bash
# round 2, step 4: interpreter v4 (switches outside the row loops) + k-means tests with an rnd stream
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_plan.py tests/test_gpu_kmeans.py -q -m gpu 2>&1 | tail -12
timeout 300 python tools/profile_ops.py > gpurun_out/r02_ops_microbench_v4.json 2> gpurun_out/r02_ops_microbench_v4.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02_ops_microbench_v4.json'))
for k, v in d.items():
    if 'plan' in k: print(k, v)
PY
