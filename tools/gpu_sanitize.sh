#!/bin/bash
# compute-sanitizer over one small exec of every kernel family (tests/sanitize_check.py); logs summarised under profiles/.
#   default build/switches: memcheck, racecheck, synccheck;  chunked two-pass (B200FFT_FUSED=0): racecheck;
#   dataflow kernel (B200FFT_FLOW=1): memcheck;  TMA-pipelined one-pass kernels (B200FFT_PIPELINE=1): racecheck
OUT=${1:-gpurun_out/sanitize}
LIMIT=${2:-170}
mkdir -p $OUT
export PYTHONPATH=$PWD:$PWD/tests
CS=/usr/local/cuda/bin/compute-sanitizer
run() {  # name tool env...
    local name=$1 tool=$2; shift 2
    ( export "$@" _X=1; timeout $LIMIT $CS --tool $tool --print-limit 20 --error-exitcode 86 python tests/sanitize_check.py > $OUT/$name.log 2>&1; echo "rc=$?" >> $OUT/$name.log )
    echo "== $name: $(grep -c '^ok' $OUT/$name.log) execs ok, $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|rc=' $OUT/$name.log | tr '\n' ' ')"
}
run default_memcheck memcheck
run default_racecheck racecheck
run default_synccheck synccheck
if [ -z "$SANITIZE_DEFAULT_ONLY" ]; then
run chunked_racecheck racecheck B200FFT_FUSED=0 SANITIZE_QUICK=1
run flow_memcheck memcheck B200FFT_FLOW=1 SANITIZE_QUICK=1
run pipelined_racecheck racecheck B200FFT_PIPELINE=1 SANITIZE_QUICK=1
fi
for f in $OUT/*.log; do echo "--- $f"; grep -E "^ok|SUMMARY|rc=|Error|error|Race|Hazard" $f | head -60; done > $OUT/summary.txt
