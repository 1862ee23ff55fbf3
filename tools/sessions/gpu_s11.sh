#!/bin/bash
# session 11: ring size at 2^20 in the bench's own setting (batch 4096, 32 GiB in + 32 GiB out), same box, alternating
OUT=gpurun_out/s11
mkdir -p $OUT
export PYTHONPATH=$PWD:$PWD/tests
for rep in 1 2; do
  for w in 8 10; do
    B200FFT_FUSED_W=$w timeout 300 python bench.py --logs 20 --steps 5 --warmup 3 --no-e2e --no-cpu > $OUT/bench_w${w}_$rep.json 2> $OUT/err.txt
    python -c "
import json,sys
d=json.load(open('$OUT/bench_w${w}_$rep.json')); print('W=$w rep $rep', d['roofline']['frac'], d['config']['per_size'][0]['plan'])"
  done
done
