#!/bin/bash
# round 2, step 8: lz4 with a 4 KB ring / 32 warps per SM at scan scale; plan v4 (inline) re-check; ncu of the IVF candidate pass (dynamic unit feed)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_lz4.py tests/test_gpu_plan.py -q -m gpu 2>&1 | tail -3
timeout 300 python tools/profile_ops.py > gpurun_out/r02_ops_microbench_v8.json 2> gpurun_out/r02_ops_microbench_v8.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02_ops_microbench_v8.json'))
for k, v in d.items():
    if 'lz4' in k or 'plan_q' in k or 'kmeans' in k: print(k, v)
PY
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu"
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:tc_candidates_kernel<0, 32>" -s 3 -c 1 -f -o gpurun_out/r02_tc_ivf_dyn $B --workload ivf > gpurun_out/r02_tc_ivf_dyn.out 2>&1
tail -3 gpurun_out/r02_tc_ivf_dyn.out
ncu -i gpurun_out/r02_tc_ivf_dyn.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); hdr=rows[0]; vals=rows[-1]
for w in ['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_sector_hit_rate.pct','sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_tensor.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed']:
    for i,h in enumerate(hdr):
        if h==w: print(w,'=',vals[i])
"
