#!/bin/bash
set -x
OUT=gpurun_out/r1n
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log
timeout 900 python bench.py --steps 3 --no-cpu > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" >> $OUT/bench.err
for mb in 96 128; do
B200FFT_CHUNK_MB=$mb timeout 300 python bench.py --steps 3 --logs 16,18,20 --no-e2e --no-cpu > $OUT/bench_chunk$mb.json 2>> $OUT/bench.err
done
du -sh gpurun_out
