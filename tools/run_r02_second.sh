python -m pytest tests/test_gpu_merge.py tests/test_gpu_reference_tables.py tests/test_gpu_colops.py -q -m gpu -x 2>&1 | tail -25
python bench.py --workload q1 --steps 20 --warmup 5 > gpurun_out/r02_bench_q1.json 2> gpurun_out/r02_bench_q1.err; echo rc=$?; tail -c 600 gpurun_out/r02_bench_q1.err
python tools/brief.py gpurun_out/r02_bench_q1.json
python bench.py --workload sum --steps 20 --warmup 5 > gpurun_out/r02_bench_sum.json 2> gpurun_out/r02_bench_sum.err; python tools/brief.py gpurun_out/r02_bench_sum.json
