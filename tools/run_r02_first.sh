python -m pytest tests/test_gpu_merge.py -x -q -m gpu 2>&1 | tail -15
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_all_n1.json 2> gpurun_out/r02_bench_all_n1.err; echo rc=$?
tail -c 1500 gpurun_out/r02_bench_all_n1.err
python tools/brief.py gpurun_out/r02_bench_all_n1.json
