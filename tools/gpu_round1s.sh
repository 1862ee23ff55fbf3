#!/bin/bash
# round-1 run s: dataflow four-step -- parity first (bounded), then A/B timings, then the bench line
set -x
OUT=gpurun_out/r1s
mkdir -p $OUT
export PYTHONPATH=$PWD:$PWD/tests
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.draw --format=csv > $OUT/env.txt 2>&1
timeout 400 python tests/variant_check.py > $OUT/variant_flow.log 2>&1; echo "rc=$?" >> $OUT/variant_flow.log
tail -3 $OUT/variant_flow.log
if ! grep -q VARIANT-OK $OUT/variant_flow.log; then echo "FLOW PARITY FAILED"; fi
for v in "" "B200FFT_FLOW=0" "B200FFT_FLOW_LOOKAHEAD=250" "B200FFT_FLOW_LOOKAHEAD=1000" "B200FFT_FLOW_LOOKAHEAD=2000"; do
  env $v timeout 200 python tools/ab_two_pass.py >> $OUT/ab.log 2>&1
done
grep SUMMARY $OUT/ab.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?" >> $OUT/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r1s/bench.json').read().strip().splitlines()[-1])
print({k:v for k,v in d.items() if k in('value','ms_per_step','roofline','e2e')})
for r in d['config']['per_size']: print(r)
PY
timeout 500 python -m pytest tests -m gpu -x -q -k "config or up_to_2_24 or device_path or ragged or shared_plan or flow or chunked" > $OUT/pytest_subset.log 2>&1; tail -5 $OUT/pytest_subset.log
