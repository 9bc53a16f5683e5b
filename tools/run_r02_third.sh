timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -15
python bench.py --workload ivf --steps 20 --warmup 5 > gpurun_out/r02_bench_ivf.json 2> gpurun_out/r02_bench_ivf.err; echo rc=$?; tail -c 600 gpurun_out/r02_bench_ivf.err
python tools/brief.py gpurun_out/r02_bench_ivf.json
