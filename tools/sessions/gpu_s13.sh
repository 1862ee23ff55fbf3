#!/bin/bash
# session 13: compute-sanitizer (memcheck, racecheck, synccheck) over the final build incl. the kernels added since session s2, then a
# parity subset through the re-split translation units
OUT=gpurun_out/s13
mkdir -p $OUT
export PYTHONPATH=$PWD:$PWD/tests
SANITIZE_DEFAULT_ONLY=1 bash tools/gpu_sanitize.sh $OUT/sanitize 200 2>&1 | tail -5
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "round2 or smooth or compiled or real_fft or accuracy_every_len" > $OUT/pytest_subset.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_subset.log; tail -3 $OUT/pytest_subset.log
