#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_i.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_i.log
timeout 600 python bench.py --steps 3 --no-e2e --no-cpu > gpurun_out/bench_i.json 2> gpurun_out/bench_i.err; echo "bench rc=$?" >> gpurun_out/bench_i.err
B200FFT_OVERLAP=0 timeout 600 python bench.py --steps 3 --no-e2e --no-cpu > gpurun_out/bench_i_noov.json 2>> gpurun_out/bench_i.err
ls -la gpurun_out
