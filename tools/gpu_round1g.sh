#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_g.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_g.log
timeout 900 python bench.py --steps 3 > gpurun_out/bench_g.json 2> gpurun_out/bench_g.err; echo "bench rc=$?" >> gpurun_out/bench_g.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:run_ -s 20 -c 2 -o gpurun_out/prof_r1g_n20 python bench.py --profile --steps 1 --logs 20 > gpurun_out/ncu_full_n20g.log 2>&1
ls -la gpurun_out
