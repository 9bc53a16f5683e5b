#!/bin/bash
# ncu evidence for the vector-search kernels (run under gpurun, 1 GPU).  Outputs go to gpurun_out/; summaries are written by
# tools/summarize_ncu.py into profiles/.
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_launches_ivf.csv $B --workload ivf --rows 1250000 > gpurun_out/r01_launches_ivf.out 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_launches_bruteforce.csv $B --workload bruteforce > gpurun_out/r01_launches_bruteforce.out 2>&1
ncu --set full --clock-control none --import-source on -k regex:tc_candidates_kernel -s 3 -c 1 -f -o gpurun_out/r01_tc_bruteforce $B --workload bruteforce > gpurun_out/r01_tc_bruteforce.out 2>&1
# ivf: tensor-core launches per step = centroid probe, list pass (no refine pass on this data) -> the list pass of the 4th step is launch #8
ncu --set full --clock-control none --import-source on -k regex:tc_candidates_kernel -s 7 -c 1 -f -o gpurun_out/r01_tc_ivf $B --workload ivf --rows 1250000 > gpurun_out/r01_tc_ivf.out 2>&1
# row-wise distance kernels (tools/tune.py dist launches each id 7 times: xcall l2sq, go l2sq, go cosine)
TUNE_DIST_ROWS=2000000 ncu --set full --clock-control none --import-source on -k regex:rowdist_kernel -s 8 -c 1 -f -o gpurun_out/r01_rowdist_l2 python tools/tune.py dist > gpurun_out/r01_rowdist_l2.out 2>&1
TUNE_DIST_ROWS=2000000 ncu --set full --clock-control none --import-source on -k regex:rowdist_kernel -s 16 -c 1 -f -o gpurun_out/r01_rowdist_cos python tools/tune.py dist > gpurun_out/r01_rowdist_cos.out 2>&1
ls -la gpurun_out | tail -12
