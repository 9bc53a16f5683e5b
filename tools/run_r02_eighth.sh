timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -6
bash tools/run_r02_sanitize.sh
timeout 300 python tools/profile_ops.py > gpurun_out/r02_ops.json 2> gpurun_out/r02_ops.err; cat gpurun_out/r02_ops.json | cut -c1-1800
timeout 200 python bench.py --workload sum --steps 20 --warmup 5 > gpurun_out/r02_bench_sum.json 2> gpurun_out/r02_bench_sum.err; python tools/brief.py gpurun_out/r02_bench_sum.json | cut -c1-500
