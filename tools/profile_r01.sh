#!/bin/bash
# ncu evidence for round 1 (run under gpurun, 1 GPU).  Outputs go to gpurun_out/; summaries are written by
# tools/summarize_ncu.py into profiles/.
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_q6.csv $B > gpurun_out/r01_launches_q6.out 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_q1.csv $B --workload q1 > gpurun_out/r01_launches_q1.out 2>&1
ncu --set full --clock-control none --import-source on -k regex:q6_kernel -s 3 -c 1 -f -o gpurun_out/r01_q6 $B > gpurun_out/r01_q6.out 2>&1
ncu --set full --clock-control none --import-source on -k regex:q1_ -s 3 -c 1 -f -o gpurun_out/r01_q1 $B --workload q1 > gpurun_out/r01_q1.out 2>&1
TUNE_ROWS=100000000 ncu --set full --clock-control none --import-source on -k regex:agg_kernel -s 8 -c 1 -f -o gpurun_out/r01_agg python tools/tune.py sum > gpurun_out/r01_agg.out 2>&1
ncu --set full --clock-control none --import-source on -k regex:bf_topk_kernel -s 1 -c 1 -f -o gpurun_out/r01_bf python tools/tune.py bf > gpurun_out/r01_bf.out 2>&1
ls -la gpurun_out | tail -20
