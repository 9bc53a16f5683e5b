# reference-CUDA comparison, Q1 kernel variants, drop-in block costs
python tools/ref_cuda_compare.py > gpurun_out/r02_ref_cuda_compare.json 2> gpurun_out/r02_ref_cuda_compare.err; echo rc=$?; cat gpurun_out/r02_ref_cuda_compare.json; tail -c 400 gpurun_out/r02_ref_cuda_compare.err
for v in 0 5 6 7; do
  python bench.py --workload q1 --steps 20 --warmup 5 --no-e2e --no-cpu --tune q1_variant=$v > gpurun_out/r02_q1_variant_$v.json 2> gpurun_out/r02_q1_variant_$v.err
  echo "q1 variant $v:"; python tools/brief.py gpurun_out/r02_q1_variant_$v.json | head -1 | cut -c1-420
done
python bench.py --workload dropin > gpurun_out/r02_bench_dropin.json 2> gpurun_out/r02_bench_dropin.err; python tools/brief.py gpurun_out/r02_bench_dropin.json
python -m pytest tests/test_gpu_cache.py tests/test_gpu_concurrency.py tests/test_gpu_decimal.py tests/test_gpu_plan.py -q -m gpu 2>&1 | tail -15
