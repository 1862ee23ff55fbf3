#!/bin/bash
# 2-GPU sanity of the torchrun path + wide-tile experiment on rank-local GPU 0
set -x
mkdir -p gpurun_out
B200FFT_TILE1024=16 timeout 600 python bench.py --steps 3 --logs 19,20 --no-e2e --no-cpu > gpurun_out/bench_k_wide.json 2> gpurun_out/bench_k.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_k_2gpu.json 2>> gpurun_out/bench_k.err; echo "2gpu rc=$?" >> gpurun_out/bench_k.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/bench_k_2gpu_ref.json 2>> gpurun_out/bench_k.err; echo "2gpu ref rc=$?" >> gpurun_out/bench_k.err
ls -la gpurun_out
