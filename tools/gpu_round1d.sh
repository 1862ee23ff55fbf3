#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_d.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_d.log
timeout 900 python bench.py --steps 3 --no-e2e --no-cpu > gpurun_out/bench_d.json 2> gpurun_out/bench_d.err; echo "bench rc=$?" >> gpurun_out/bench_d.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:run_kernel -s 20 -c 2 -o gpurun_out/prof_r1d_n20 python bench.py --profile --steps 1 --logs 20 > gpurun_out/ncu_full_n20d.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:run_kernel -s 3 -c 3 -o gpurun_out/prof_r1d_small python bench.py --profile --steps 1 --logs 10,12,14 > gpurun_out/ncu_full_smalld.log 2>&1
ls -la gpurun_out
