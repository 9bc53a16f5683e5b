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
# ncu evidence of the IVF candidate pass with the dynamic unit feed (numbers under a profiler are not bench values)
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu"
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:tc_candidates_kernel<\(bool\)0" -s 3 -c 1 -f -o gpurun_out/r02_tc_ivf_dyn $B --workload ivf > gpurun_out/r02_tc_ivf_dyn.out 2>&1
tail -2 gpurun_out/r02_tc_ivf_dyn.out | cut -c1-200
ncu -i gpurun_out/r02_tc_ivf_dyn.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); hdr=rows[0]; vals=rows[-1]
for w in ['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_sector_hit_rate.pct','sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_tensor.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed']:
    for i,h in enumerate(hdr):
        if h==w: print(w,'=',vals[i][:120])
"
