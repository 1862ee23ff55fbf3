#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_h.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_h.log
timeout 600 python bench.py --steps 3 --no-e2e --no-cpu > gpurun_out/bench_h.json 2> gpurun_out/bench_h.err; echo "bench rc=$?" >> gpurun_out/bench_h.err
for mb in 32 48; do
B200FFT_CHUNK_MB=$mb timeout 300 python bench.py --steps 3 --logs 16,18,20 --no-e2e --no-cpu > gpurun_out/bench_h_chunk$mb.json 2>> gpurun_out/bench_h.err
done
ls -la gpurun_out
