python -m pytest tests/test_gpu_tpch.py tests/test_gpu_agg.py tests/test_gpu_merge.py tests/test_gpu_cache.py tests/test_gpu_concurrency.py -q -m gpu 2>&1 | tail -8
for v in 0 8 9; do
  python bench.py --workload q1 --steps 20 --warmup 5 --no-e2e --no-cpu --tune q1_variant=$v > gpurun_out/r02_q1_variant_$v.json 2> gpurun_out/r02_q1_variant_$v.err
  echo "q1 variant $v:"; python tools/brief.py gpurun_out/r02_q1_variant_$v.json | head -1 | cut -c1-330
done
python bench.py --workload sum --steps 20 --warmup 5 > gpurun_out/r02_bench_sum.json 2> gpurun_out/r02_bench_sum.err; python tools/brief.py gpurun_out/r02_bench_sum.json | cut -c1-600
