#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_j.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_j.log
timeout 600 python bench.py --steps 3 --no-e2e --no-cpu > gpurun_out/bench_j.json 2> gpurun_out/bench_j.err; echo "bench rc=$?" >> gpurun_out/bench_j.err
for k in 3 4; do
B200FFT_STREAMS=$k timeout 600 python bench.py --steps 3 --no-e2e --no-cpu --no-extras > gpurun_out/bench_j_s$k.json 2>> gpurun_out/bench_j.err
done
ls -la gpurun_out
