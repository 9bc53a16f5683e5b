#!/bin/bash
# round 2, final: what the driver runs at round end, in one call -- the whole GPU suite, smoke(), the default bench (all configs) and the reference arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -4
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r02_bench_ref_n1.json 2> gpurun_out/r02_bench_ref_n1.err; echo ref rc=$?
timeout 900 python bench.py > gpurun_out/r02_bench_all_n1.json 2> gpurun_out/r02_bench_all_n1.err; echo bench rc=$?
tail -c 600 gpurun_out/r02_bench_all_n1.err
python tools/brief.py gpurun_out/r02_bench_all_n1.json
python tools/brief.py gpurun_out/r02_bench_ref_n1.json
