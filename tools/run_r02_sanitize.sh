# memcheck over the round-2 kernels (small inputs; compute-sanitizer slows kernels ~10-50x)
for t in tests/test_gpu_colops.py tests/test_gpu_decimal.py tests/test_gpu_merge.py; do
  timeout 900 compute-sanitizer --tool memcheck --error-exitcode 77 python -m pytest $t -q -m gpu -x -k "not 1000000 and not 3000000 and not 2000001 and not 3_000_000 and not 100000" 2>&1 | tail -6
done
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 77 python -m pytest tests/test_gpu_plan.py -q -m gpu -x -k "test_plans_reproduce or test_q6_plan_with_nulls or too_many or (cardinalities and 300)" 2>&1 | tail -6
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 77 python -m pytest tests/test_gpu_tpch.py -q -m gpu -x -k "nullable or (variants and 10) or predicate_edges" 2>&1 | tail -6
# shared-memory race check of the kernels with shared-memory protocols (plan register file / tables, compaction scan, grouped aggregates)
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 78 python -m pytest tests/test_gpu_colops.py -q -m gpu -x -k "(filter_sels_matches and 2049) or (group_sum and 33 and T23) or layout" 2>&1 | tail -6
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 78 python -m pytest tests/test_gpu_plan.py -q -m gpu -x -k "test_plans_reproduce" 2>&1 | tail -6
