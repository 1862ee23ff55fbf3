#!/bin/bash
# second GPU contact: full parity suite, bench with graph replay, chunk sweep, ncu on the 2^20 kernels
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_b.log
timeout 900 python bench.py --steps 3 > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err; echo "bench rc=$?" >> gpurun_out/bench_b.err
for mb in 32 96; do
  B200FFT_CHUNK_MB=$mb timeout 300 python bench.py --steps 3 --logs 15,16,18,20 --no-e2e --no-cpu > gpurun_out/bench_b_chunk$mb.json 2>> gpurun_out/bench_b.err
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:run_kernel -s 20 -c 4 -o gpurun_out/prof_r1b_n20 python bench.py --profile --steps 1 --logs 20 > gpurun_out/ncu_full_n20b.log 2>&1
ls -la gpurun_out
