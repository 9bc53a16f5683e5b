#!/bin/bash
# ncu evidence for round 2 (run under gpurun, 1 GPU).  Outputs go to gpurun_out/; tools/summarize_ncu.py r02 writes profiles/r02_ncu_summary.md
# and profiles/ncu_traffic.json from them.
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu"
for w in q6 q1 sum bruteforce ivf; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_$w.csv $B --workload $w > gpurun_out/r02_launches_$w.out 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:q6_kernel -s 3 -c 1 -f -o gpurun_out/r02_q6 $B --workload q6 > gpurun_out/r02_q6.out 2>&1
ncu --set full --clock-control none --import-source on -k regex:q1_ -s 3 -c 1 -f -o gpurun_out/r02_q1 $B --workload q1 > gpurun_out/r02_q1.out 2>&1
ncu --set full --clock-control none --import-source on -k regex:agg_kernel -s 6 -c 1 -f -o gpurun_out/r02_sum $B --workload sum > gpurun_out/r02_sum.out 2>&1
ncu --set full --clock-control none --import-source on -k regex:tc_candidates_kernel -s 9 -c 1 -f -o gpurun_out/r02_tc_ivf $B --workload ivf > gpurun_out/r02_tc_ivf.out 2>&1
ncu --set full --clock-control none --import-source on -k regex:split_residual_kernel -s 5 -c 1 -f -o gpurun_out/r02_split_residual $B --workload ivf > gpurun_out/r02_split_residual.out 2>&1
ncu --set full --clock-control none --import-source on -k regex:plan_kernel -s 2 -c 1 -f -o gpurun_out/r02_plan_q6 python tools/profile_ops.py 100000000 > gpurun_out/r02_plan_q6.out 2>&1
ncu --set full --clock-control none --import-source on -k regex:plan_kernel -s 8 -c 1 -f -o gpurun_out/r02_plan_q1 python tools/profile_ops.py 100000000 > gpurun_out/r02_plan_q1.out 2>&1
ncu --set full --clock-control none --import-source on -k regex:select_kernel -s 2 -c 1 -f -o gpurun_out/r02_select python tools/profile_ops.py 100000000 > gpurun_out/r02_select.out 2>&1
ncu --set full --clock-control none --import-source on -k regex:group_agg_kernel -s 5 -c 1 -f -o gpurun_out/r02_group_agg python tools/profile_ops.py 100000000 > gpurun_out/r02_group_agg.out 2>&1
ls -la gpurun_out | tail -30
