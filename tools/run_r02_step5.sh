#!/bin/bash
# round 2, step 5: the whole GPU suite with everything new in (bloom, join, k-means, normalize, lz4, interpreter v4, dynamic unit feed) + microbench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -15
timeout 400 python tools/profile_ops.py > gpurun_out/r02_ops_microbench_v5.json 2> gpurun_out/r02_ops_microbench_v5.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02_ops_microbench_v5.json'))
for k, v in d.items():
    if 'plan' in k or 'bloom' in k or 'lz4' in k: print(k, v)
PY
