#!/bin/bash
# first GPU contact: environment facts, smoke, parity tests, bench, chunk sweep, ncu launch list + one full capture
set -x
mkdir -p gpurun_out
{ nvidia-smi; nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,memory.total --format=csv; free -g; nproc; lscpu | grep "Model name"; } > gpurun_out/env.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
for mb in 16 64 96; do
  B200FFT_CHUNK_MB=$mb timeout 300 python bench.py --steps 3 --logs 14,16,18,20 --no-e2e --no-cpu > gpurun_out/bench_chunk$mb.json 2>> gpurun_out/bench.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1.csv python bench.py --profile --steps 1 --logs 10,12,14,16,18,20 > gpurun_out/ncu_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:run_kernel -c 12 -o gpurun_out/prof_r1_small python bench.py --profile --steps 1 --logs 10,11,12 > gpurun_out/ncu_full_small.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:run_kernel -s 40 -c 4 -o gpurun_out/prof_r1_n20 python bench.py --profile --steps 1 --logs 20 > gpurun_out/ncu_full_n20.log 2>&1
ls -la gpurun_out
