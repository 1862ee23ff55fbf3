#!/bin/bash
# session 7: smoke(), the two-pass memory-pipeline ceiling on the final build (butterflies skipped), both bench arms
OUT=gpurun_out/s7
mkdir -p $OUT
export PYTHONPATH=$PWD:$PWD/tests
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest_gpu.log; tail -3 $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -7 $OUT/smoke.log
timeout 300 python tools/ab_two_pass.py 15,16,17,18,19,20 > $OUT/ab_two_pass_default.txt 2>&1; tail -1 $OUT/ab_two_pass_default.txt
B200FFT_FUSED_NOCOMPUTE=1 timeout 300 python tools/ab_two_pass.py 15,16,17,18,19,20 > $OUT/ab_two_pass_nocompute.txt 2>&1; tail -1 $OUT/ab_two_pass_nocompute.txt
B200FFT_FUSED=0 timeout 300 python tools/ab_two_pass.py 15,16,17,18,19,20 > $OUT/ab_two_pass_chunked.txt 2>&1; tail -1 $OUT/ab_two_pass_chunked.txt
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_ref.err; echo "ref rc=$?"
timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/s7/bench.json'))
print("value",d["value"],"frac",d["roofline"]["frac"], "e2e", d["e2e"]["value"], "cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
r=json.load(open('gpurun_out/s7/bench_reference.json')); print("reference arm", r["value"], r["unit"], r["cpu_baseline"]["cores"])
PY
