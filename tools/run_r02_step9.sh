#!/bin/bash
# round 2, step 9: microbench after the coalesced sels output / LZ4 sync trimming / bloom probe change (+ their tests)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_blockdecode.py tests/test_gpu_lz4.py tests/test_gpu_colops.py tests/test_gpu_bloom.py -q -m gpu 2>&1 | tail -3
timeout 300 python tools/profile_ops.py > gpurun_out/r02_ops_microbench_v9.json 2> gpurun_out/r02_ops_microbench_v9.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02_ops_microbench_v9.json'))
for k, v in d.items():
    if 'lz4' in k or 'filter_sels' in k or 'bloom' in k or 'kmeans' in k: print(k, v)
PY
