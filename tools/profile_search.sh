#!/bin/bash
# ncu evidence for the vector-search kernels (run under gpurun, 1 GPU).  Outputs go to gpurun_out/; summaries are written by
# tools/summarize_ncu.py into profiles/.
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_launches_ivf.csv $B --workload ivf --rows 1250000 > gpurun_out/r01_launches_ivf.out 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_launches_bruteforce.csv $B --workload bruteforce > gpurun_out/r01_launches_bruteforce.out 2>&1
ncu --set full --clock-control none --import-source on -k regex:tc_candidates_kernel -s 3 -c 1 -f -o gpurun_out/r01_tc_bruteforce $B --workload bruteforce > gpurun_out/r01_tc_bruteforce.out 2>&1
# ivf: tensor-core launches per step = centroid probe, list pass, refine pass -> the list pass of the 4th step is launch #11
ncu --set full --clock-control none --import-source on -k regex:tc_candidates_kernel -s 10 -c 1 -f -o gpurun_out/r01_tc_ivf $B --workload ivf --rows 1250000 > gpurun_out/r01_tc_ivf.out 2>&1
ls -la gpurun_out | tail -12
