# 8 GPUs: the multi-GPU form of the default bench (one process per GPU, NCCL), as the driver launches it
mkdir -p gpurun_out
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_bench_all_n8.json 2> gpurun_out/r02_bench_all_n8.err; echo rc=$?
tail -c 800 gpurun_out/r02_bench_all_n8.err
python tools/brief.py gpurun_out/r02_bench_all_n8.json
