#!/bin/bash
# round 2, step 6: ncu evidence for the dynamic unit feed (IVF candidate pass) and the v4 interpreter; memcheck / racecheck over the new kernels
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_lz4.py tests/test_gpu_elementwise.py tests/test_gpu_search.py -q -m gpu 2>&1 | tail -4
timeout 300 python tools/profile_ops.py > gpurun_out/r02_ops_microbench_v6.json 2> gpurun_out/r02_ops_microbench_v6.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02_ops_microbench_v6.json'))
for k, v in d.items():
    if 'lz4' in k: print(k, v)
PY
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu"
ncu --set full --clock-control none --import-source on -k regex:tc_candidates_kernel -s 9 -c 1 -f -o gpurun_out/r02_tc_ivf_dyn $B --workload ivf > gpurun_out/r02_tc_ivf_dyn.out 2>&1
ncu --set full --clock-control none --import-source on -k regex:plan_kernel -s 2 -c 1 -f -o gpurun_out/r02_plan_q6_v4 python tools/profile_ops.py 100000000 > gpurun_out/r02_plan_q6_v4.out 2>&1
ncu --set full --clock-control none --import-source on -k regex:plan_kernel -s 8 -c 1 -f -o gpurun_out/r02_plan_q1_v4 python tools/profile_ops.py 100000000 > gpurun_out/r02_plan_q1_v4.out 2>&1
ncu --set full --clock-control none --import-source on -k regex:bloom_kernel -s 2 -c 1 -f -o gpurun_out/r02_bloom_test python tools/profile_ops.py 100000000 > gpurun_out/r02_bloom_test.out 2>&1
ls -la gpurun_out/*.ncu-rep | tail
for t in tests/test_gpu_bloom.py tests/test_gpu_lz4.py tests/test_gpu_normalize.py; do
  timeout 900 compute-sanitizer --tool memcheck --error-exitcode 77 python -m pytest $t -q -m gpu -x -k "not device_resident" 2>&1 | tail -4
done
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 77 python -m pytest tests/test_gpu_join.py -q -m gpu -x -k "1000 or 5000 or 70000 or heavy or empty" 2>&1 | tail -4
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 77 python -m pytest tests/test_gpu_kmeans.py -q -m gpu -x -k "table or 2000 or 400 or empty" 2>&1 | tail -4
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 77 python -m pytest tests/test_gpu_plan.py -q -m gpu -x -k "test_plans_reproduce or test_q6_plan_with_nulls or too_many or (cardinalities and 300)" 2>&1 | tail -4
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 78 python -m pytest tests/test_gpu_plan.py -q -m gpu -x -k "test_plans_reproduce" 2>&1 | tail -4
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 78 python -m pytest tests/test_gpu_join.py -q -m gpu -x -k "5000 and 37" 2>&1 | tail -4
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 77 python -m pytest tests/test_gpu_search.py -q -m gpu -x -k "ivf_tensor_core_scan_is_exact and not 4096" 2>&1 | tail -4
