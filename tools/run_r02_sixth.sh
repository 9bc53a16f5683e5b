timeout 300 python -m pytest tests/test_gpu_tpch.py -q -m gpu -k "variants" 2>&1 | tail -6
for v in 0 10; do
  timeout 300 python bench.py --workload q1 --steps 20 --warmup 5 --no-e2e --no-cpu --tune q1_variant=$v > gpurun_out/r02_q1_variant_$v.json 2> gpurun_out/r02_q1_variant_$v.err
  echo "q1 variant $v:"; python tools/brief.py gpurun_out/r02_q1_variant_$v.json | head -1 | cut -c1-330
done
timeout 600 python tools/profile_ops.py > gpurun_out/r02_ops.json 2> gpurun_out/r02_ops.err; cat gpurun_out/r02_ops.json; tail -c 500 gpurun_out/r02_ops.err
