#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_e.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_e.log
timeout 900 python bench.py --steps 3 --no-e2e --no-cpu > gpurun_out/bench_e.json 2> gpurun_out/bench_e.err; echo "bench rc=$?" >> gpurun_out/bench_e.err
B200FFT_PIPELINE=0 timeout 900 python bench.py --steps 3 --no-e2e --no-cpu > gpurun_out/bench_e_nopipe.json 2>> gpurun_out/bench_e.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:run_ -s 20 -c 2 -o gpurun_out/prof_r1e_n20 python bench.py --profile --steps 1 --logs 20 > gpurun_out/ncu_full_n20e.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:run_ -s 3 -c 3 -o gpurun_out/prof_r1e_small python bench.py --profile --steps 1 --logs 10,12,13 > gpurun_out/ncu_full_smalle.log 2>&1
ls -la gpurun_out
