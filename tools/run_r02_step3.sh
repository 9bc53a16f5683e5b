#!/bin/bash
# round 2, step 3: grouped-load plan interpreter, normalize_l2, centroid assignment on the tensor cores, dynamic unit feed of the candidate kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kmeans.py tests/test_gpu_join.py tests/test_gpu_bloom.py tests/test_gpu_plan.py tests/test_gpu_normalize.py tests/test_gpu_search.py -q -m gpu 2>&1 | tail -25
timeout 300 python tools/profile_ops.py > gpurun_out/r02_ops_microbench_v3.json 2> gpurun_out/r02_ops_microbench_v3.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02_ops_microbench_v3.json'))
for k, v in d.items():
    if 'plan' in k or 'bloom' in k: print(k, v)
PY
for m in 0 1; do
  timeout 600 python bench.py --workload ivf --no-cpu --no-pageable --tune tc_sched=$m > gpurun_out/r02_ivf_sched$m.json 2> gpurun_out/r02_ivf_sched$m.err
  python tools/brief.py gpurun_out/r02_ivf_sched$m.json
done
timeout 300 python bench.py --workload bruteforce --no-cpu --no-pageable > gpurun_out/r02_bf_step3.json 2> gpurun_out/r02_bf_step3.err
python tools/brief.py gpurun_out/r02_bf_step3.json
