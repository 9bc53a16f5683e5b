# 2 GPUs: the multi-GPU form of the default bench (one process per GPU, NCCL), as the driver launches it
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_all_n2.json 2> gpurun_out/r02_bench_all_n2.err; echo rc=$?
tail -c 1500 gpurun_out/r02_bench_all_n2.err
python tools/brief.py gpurun_out/r02_bench_all_n2.json
