#!/bin/bash
# Round-end measurement set (run under gpurun, 1 GPU): one JSON line per BASELINE config into gpurun_out/final_<workload>.json
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 python bench.py "$@" > gpurun_out/final_$name.json 2> gpurun_out/final_$name.err; tail -c 2500 gpurun_out/final_$name.json | head -c 2500; echo; }
run q6
run q1 --workload q1 --steps 30
run sum --workload sum --steps 200
run bruteforce --workload bruteforce --steps 5
run ivf --workload ivf --rows 1250000 --steps 5
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_q6_ref.json 2> gpurun_out/final_q6_ref.err
