#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_c.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_c.log
timeout 900 python bench.py --steps 3 --no-e2e --no-cpu > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err; echo "bench rc=$?" >> gpurun_out/bench_c.err
B200FFT_TILE1024=16 timeout 300 python bench.py --steps 3 --logs 19,20 --no-e2e --no-cpu > gpurun_out/bench_c_wide.json 2>> gpurun_out/bench_c.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:run_kernel -s 20 -c 2 -o gpurun_out/prof_r1c_n20 python bench.py --profile --steps 1 --logs 20 > gpurun_out/ncu_full_n20c.log 2>&1
B200FFT_TILE1024=16 timeout 600 ncu --set full --clock-control none --import-source on -k regex:run_kernel -s 20 -c 2 -o gpurun_out/prof_r1c_n20wide python bench.py --profile --steps 1 --logs 20 > gpurun_out/ncu_full_n20cw.log 2>&1
ls -la gpurun_out
