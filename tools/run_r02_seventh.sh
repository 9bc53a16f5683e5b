timeout 900 python -m pytest tests/test_gpu_plan.py tests/test_gpu_tpch.py tests/test_gpu_merge.py tests/test_gpu_concurrency.py -q -m gpu -x 2>&1 | tail -8
timeout 600 python tools/profile_ops.py > gpurun_out/r02_ops.json 2> gpurun_out/r02_ops.err; cat gpurun_out/r02_ops.json; tail -c 300 gpurun_out/r02_ops.err
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_all_n1.json 2> gpurun_out/r02_bench_all_n1.err; echo rc=$?; tail -c 600 gpurun_out/r02_bench_all_n1.err
python tools/brief.py gpurun_out/r02_bench_all_n1.json
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_ref_n1.json 2> gpurun_out/r02_bench_ref_n1.err; echo rc=$?; python tools/brief.py gpurun_out/r02_bench_ref_n1.json | cut -c1-400
