#!/bin/bash
# 2-GPU check of every workload (run under gpurun --gpus 2)
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 "$@" > gpurun_out/final_n2_$name.json 2> gpurun_out/final_n2_$name.err; tail -c 1800 gpurun_out/final_n2_$name.json; echo; }
run q6 --steps 30 --no-cpu
run q1 --workload q1 --steps 20 --no-cpu
run bruteforce --workload bruteforce --steps 3 --no-cpu
run ivf --workload ivf --rows 2500000 --steps 3 --no-cpu
