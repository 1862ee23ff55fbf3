#!/bin/bash
OUT=gpurun_out/s10
mkdir -p $OUT
export PYTHONPATH=$PWD:$PWD/tests
for v in "" "B200FFT_FUSED_XFLAGS=2" "B200FFT_FUSED_LOOKAHEAD=800" "B200FFT_FUSED_HINTS=9"; do
  env $v timeout 200 python tools/ab_two_pass.py 15,16,17,18,19,20 >> $OUT/ab.log 2>&1
done
grep SUMMARY $OUT/ab.log
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -n 5 $OUT/pytest_gpu.log
timeout 600 python bench.py --no-cpu > $OUT/bench.json 2> $OUT/bench.err
tail -c 3000 $OUT/bench.json
